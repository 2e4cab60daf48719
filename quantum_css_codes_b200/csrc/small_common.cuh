// sm_100a kernels for small codes (n <= 32): K1 syndrome + K2 lookup decode + logical check +
// tally, with K3 (Philox sampler) fused in the SAMPLE instantiations.  Included by the
// small_*.cu translation units (split so nvcc can build them in parallel).
//
// Work decomposition: one thread owns one "unit" = VEC consecutive 32-shot words of every plane
// (VEC = 4 -> a 16-byte load per plane per thread, a warp reads 512 contiguous bytes of a plane).
// The grid is one full wave (SM count x resident CTAs) and grid-strides over units.
// Tallies: per-thread counters -> warp REDUX -> one 64-bit atomic per counter per CTA.
//
// Every kernel exists in two instantiations: FAST (tally only: both Pauli types, whole words, no
// output planes -- the Monte-Carlo hot path) and FULL (optional outputs, single-side calls, the
// ragged tail of a batch).  launch_split() sends the bulk of a tally-only batch through FAST and
// the last < VEC*32 shots through FULL.
#pragma once
#include <cuda_runtime.h>

#include "decode.cuh"
#include "launch.h"

namespace qcss {
namespace small {

constexpr int kThreads = 256;

struct GenericArgs {
    GenericSide x, z;
    DecodeIO io;
};

struct SideTables;

struct NamedArgs {
    const uint8_t* fm_x;
    const uint32_t* co_x;
    const uint32_t* e32_x;
    const uint8_t* fm_z;
    const uint32_t* co_z;
    const uint32_t* e32_z;
    DecodeIO io;
};

__device__ __forceinline__ void block_tally(const Counters& c, unsigned long long* tally) {
    __shared__ uint32_t part[kThreads / 32][5];
    uint32_t v[5] = {c.fail_x, c.fail_z, c.fail_any, c.miss_x, c.miss_z};
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = __reduce_add_sync(0xFFFFFFFFu, v[i]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 5; ++i) part[warp][i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < 5 && tally != nullptr) {
        unsigned long long sum = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) sum += part[w][threadIdx.x];
        if (sum != 0) atomicAdd(tally + 1 + threadIdx.x, sum);
    }
}

// Shared body: stage the lookup tables in shared memory, grid-stride over units, tally.
// FAST kernels stage the 32-bit-entry tables when the side has them (m <= kMaxE32M), otherwise
// (and in FULL kernels) the byte tables.
struct SideTables {
    const uint8_t* fm;
    const uint32_t* corr;
    const uint32_t* e32;
};

template <class P, bool FAST>
__device__ __forceinline__ SideLut stage_side(const P& pol, const SideTables& t, uint8_t*& cursor) {
    SideLut lut;
    lut.fm = nullptr;
    lut.corr = t.corr;
    lut.e32 = nullptr;
    if constexpr (!P::kSliced) {
        const int size = 1 << pol.m();
        bool e32 = false;
        if constexpr (FAST && P::kTallyM > 0) e32 = pol.use_e32() && t.e32 != nullptr;
        if (e32) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(cursor);
            for (int i = threadIdx.x; i < size; i += kThreads) dst[i] = t.e32[i];
            lut.e32 = cursor;
            cursor += (size_t)size * 4;
        } else if (t.fm != nullptr) {
            for (int i = threadIdx.x; i < size; i += kThreads) cursor[i] = t.fm[i];
            lut.fm = cursor;
            cursor += (size + 15) & ~15;
        }
    }
    return lut;
}

template <class PX, class PZ, int VEC, bool SAMPLE, bool FAST>
__device__ __forceinline__ void run_small(const PX& px, const PZ& pz, const DecodeIO& io,
                                          const SideTables& tx, const SideTables& tz) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* cursor = smem;
    const SideLut lut_x = stage_side<PX, FAST>(px, tx, cursor);
    const SideLut lut_z = stage_side<PZ, FAST>(pz, tz, cursor);
    if constexpr (!PX::kSliced || !PZ::kSliced) __syncthreads();

    // the gap sampler's table is read with per-lane indices on its rare path: shared memory, not the
    // kernel parameter block (whose address cannot be taken without a local copy)
    __shared__ GapTable s_gap;
    if constexpr (SAMPLE) {
        if (threadIdx.x < 32) s_gap.cdf[threadIdx.x] = io.gap.cdf[threadIdx.x];
        if (threadIdx.x == 32) s_gap.inv = io.gap.inv;
        __syncthreads();
    }

    Counters c = {0u, 0u, 0u, 0u, 0u};
    const int64_t units = (io.words + VEC - 1) / VEC;
    const int64_t step = (int64_t)gridDim.x * kThreads;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < units; u += step)
        process_unit<PX, PZ, VEC, SAMPLE, FAST>(px, pz, io, u, lut_x, lut_z, c, SAMPLE ? &s_gap : nullptr);
    block_tally(c, io.tally);
}

template <int NB, int MB, int VEC, bool SAMPLE, bool FAST>
__global__ void __launch_bounds__(kThreads, FAST ? 2 : 1)
k_small_generic(const __grid_constant__ GenericArgs a) {
    GenericPolicy<NB, MB> px{&a.x}, pz{&a.z};
    const SideTables tx{a.x.lut_fm, a.x.lut_corr, a.x.lut_e32}, tz{a.z.lut_fm, a.z.lut_corr, a.z.lut_e32};
    run_small<GenericPolicy<NB, MB>, GenericPolicy<NB, MB>, VEC, SAMPLE, FAST>(px, pz, a.io, tx, tz);
}

template <class DX, class DZ, int VEC, bool SAMPLE, bool FAST>
__global__ void __launch_bounds__(kThreads, FAST ? (SAMPLE ? 3 : 2) : 1)
k_small_named(const __grid_constant__ NamedArgs a) {
    StaticPolicy<DX> px;
    StaticPolicy<DZ> pz;
    const SideTables tx{a.fm_x, a.co_x, a.e32_x}, tz{a.fm_z, a.co_z, a.e32_z};
    run_small<StaticPolicy<DX>, StaticPolicy<DZ>, VEC, SAMPLE, FAST>(px, pz, a.io, tx, tz);
}

// ---- launch plumbing --------------------------------------------------------------------------
inline cudaError_t sm_count(int* out) {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    *out = cached;
    return cudaSuccess;
}

template <class Kernel, class Args>
cudaError_t launch_one(Kernel kernel, const Args& args, int64_t units, size_t smem, cudaStream_t stream) {
    int sms = 0;
    cudaError_t e = sm_count(&sms);
    if (e != cudaSuccess) return e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int64_t want = (units + kThreads - 1) / kThreads;
    int64_t wave = (int64_t)sms * per_sm;
    int64_t grid = want < wave ? want : wave;        // persistent: at most one full wave
    if (grid < 1) grid = 1;
    kernel<<<(unsigned)grid, kThreads, smem, stream>>>(args);
    return cudaGetLastError();
}

inline size_t side_smem(const GenericSide& s, bool lut, bool fast) {
    if (!lut || s.lut_fm == nullptr) return 0;
    const size_t size = (size_t)1 << s.m;
    if (fast && s.lut_e32 != nullptr) return size * 4;
    return (size + 15) & ~(size_t)15;
}

inline size_t lut_smem(const GenericSide& x, const GenericSide& z, bool x_lut, bool z_lut, bool fast) {
    return side_smem(x, x_lut, fast) + side_smem(z, z_lut, fast) + 16;
}

// A tally-only batch: whole units through the FAST kernel, the ragged rest through FULL.
inline bool fast_eligible(const SmallLaunch& l) {
    const DecodeIO& io = l.io;
    if (io.tally == nullptr) return false;
    if (l.x->mode == kModeNone || l.z->mode == kModeNone) return false;
    if (io.synd_x || io.synd_z || io.corr_x || io.corr_z || io.flip_x || io.flip_z || io.miss_x || io.miss_z)
        return false;
    if (l.sample) return io.ex_out == nullptr && io.ez_out == nullptr;
    return io.sides == 3;
}

template <int VEC, class KFast, class KFull, class Args>
cudaError_t launch_split(KFast kfast, KFull kfull, Args a, const SmallLaunch& l, size_t smem_fast,
                         size_t smem, cudaStream_t stream) {
    const DecodeIO io = l.io;
    int64_t fast_units = 0;
    if (fast_eligible(l)) {
        const int64_t whole_words = (io.tail_mask == 0xFFFFFFFFu) ? io.words : io.words - 1;
        fast_units = whole_words / VEC;
    }
    if (fast_units > 0) {
        a.io = io;
        a.io.words = fast_units * VEC;
        a.io.tail_mask = 0xFFFFFFFFu;
        cudaError_t e = launch_one(kfast, a, fast_units, smem_fast, stream);
        if (e != cudaSuccess) return e;
    }
    const int64_t done = fast_units * VEC;
    if (done < io.words) {
        a.io = io;
        a.io.words = io.words - done;
        a.io.first_word = io.first_word + (uint64_t)done;
        if (a.io.ex) a.io.ex += done;
        if (a.io.ez) a.io.ez += done;
        if (a.io.synd_x) a.io.synd_x += done;
        if (a.io.synd_z) a.io.synd_z += done;
        if (a.io.corr_x) a.io.corr_x += done;
        if (a.io.corr_z) a.io.corr_z += done;
        if (a.io.flip_x) a.io.flip_x += done;
        if (a.io.flip_z) a.io.flip_z += done;
        if (a.io.miss_x) a.io.miss_x += done;
        if (a.io.miss_z) a.io.miss_z += done;
        if (a.io.ex_out) a.io.ex_out += done;
        if (a.io.ez_out) a.io.ez_out += done;
        return launch_one(kfull, a, (a.io.words + VEC - 1) / VEC, smem, stream);
    }
    return cudaSuccess;
}

template <int NB, int MB, int VEC>
cudaError_t launch_generic(const SmallLaunch& l, cudaStream_t stream) {
    GenericArgs a;
    a.x = *l.x;
    a.z = *l.z;
    a.io = l.io;
    const bool lut = (MB != kSlicedM);
    const size_t smem = lut_smem(*l.x, *l.z, lut, lut, false), smem_fast = lut_smem(*l.x, *l.z, lut, lut, true);
    constexpr int SVEC = 1;                          // sampling kernels: one word per thread (see launch_named)
    if (l.sample)
        return launch_split<SVEC>(k_small_generic<NB, MB, SVEC, true, true>, k_small_generic<NB, MB, SVEC, true, false>,
                                  a, l, smem_fast, smem, stream);
    return launch_split<VEC>(k_small_generic<NB, MB, VEC, false, true>, k_small_generic<NB, MB, VEC, false, false>,
                             a, l, smem_fast, smem, stream);
}

template <class DX, class DZ>
cudaError_t launch_named(const SmallLaunch& l, cudaStream_t stream) {
    constexpr int VEC = 4;
    NamedArgs a;
    a.fm_x = l.x->lut_fm;
    a.co_x = l.x->lut_corr;
    a.e32_x = l.x->lut_e32;
    a.fm_z = l.z->lut_fm;
    a.co_z = l.z->lut_corr;
    a.e32_z = l.z->lut_e32;
    a.io = l.io;
    const size_t smem = lut_smem(*l.x, *l.z, !DX::kSliced, !DZ::kSliced, false);
    const size_t smem_fast = lut_smem(*l.x, *l.z, !DX::kSliced, !DZ::kSliced, true);
    // The sampling kernels load nothing, so the 16-byte-per-plane unit buys nothing there: one word per
    // thread keeps 4x fewer syndrome accumulators live and 4x fewer unrolled sampler sites (measured on
    // Golay-23 with VEC = 4: the first-error path inlined at 92 sites ran at 4.2e10 shots/s).
    constexpr int SVEC = 1;
    if (l.sample)
        return launch_split<SVEC>(k_small_named<DX, DZ, SVEC, true, true>, k_small_named<DX, DZ, SVEC, true, false>,
                                  a, l, smem_fast, smem, stream);
    return launch_split<VEC>(k_small_named<DX, DZ, VEC, false, true>, k_small_named<DX, DZ, VEC, false, false>,
                             a, l, smem_fast, smem, stream);
}

}  // namespace small

// one definition per translation unit (small_named_*.cu, small_generic*.cu)
cudaError_t launch_small_steane(const SmallLaunch& l, cudaStream_t stream);
cudaError_t launch_small_qrm15(const SmallLaunch& l, cudaStream_t stream);
cudaError_t launch_small_golay23(const SmallLaunch& l, cudaStream_t stream);
cudaError_t launch_small_generic16(const SmallLaunch& l, int mb, cudaStream_t stream);
cudaError_t launch_small_generic32(const SmallLaunch& l, int mb, cudaStream_t stream);

}  // namespace qcss
