// sm_100a kernels for small codes (n <= 32): K1 syndrome + K2 lookup decode + logical check +
// tally, with K3 (Philox sampler) fused in the SAMPLE instantiations.
//
// Work decomposition: one thread owns one "unit" = VEC consecutive 32-shot words of every plane
// (VEC = 4 -> a 16-byte load per plane per thread, a warp reads 512 contiguous bytes of a plane).
// The grid is a whole number of waves (SM count x resident CTAs) and grid-strides over units.
// Tallies: per-thread counters -> warp REDUX -> one 64-bit atomic per counter per CTA.
#include <cuda_runtime.h>

#include "decode.cuh"
#include "launch.h"
#include "named_codes.inc"

namespace qcss {

namespace {

constexpr int kThreads = 256;

struct GenericArgs {
    GenericSide x, z;
    DecodeIO io;
};

struct NamedArgs {
    const uint8_t* fm_x;
    const uint32_t* co_x;
    const uint8_t* fm_z;
    const uint32_t* co_z;
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

// Shared body: stage the flip/miss tables in shared memory, grid-stride over units, tally.
template <class PX, class PZ, int VEC, bool SAMPLE>
__device__ __forceinline__ void run_small(const PX& px, const PZ& pz, const DecodeIO& io,
                                          const uint8_t* g_fm_x, const uint32_t* g_co_x,
                                          const uint8_t* g_fm_z, const uint32_t* g_co_z) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* s_fm_x = smem;
    uint8_t* s_fm_z = smem;
    if constexpr (!PX::kSliced) {
        const int nx = (g_fm_x != nullptr) ? (1 << px.m()) : 0;
        for (int i = threadIdx.x; i < nx; i += kThreads) s_fm_x[i] = g_fm_x[i];
        s_fm_z = smem + ((nx + 15) & ~15);
    }
    if constexpr (!PZ::kSliced) {
        const int nz = (g_fm_z != nullptr) ? (1 << pz.m()) : 0;
        for (int i = threadIdx.x; i < nz; i += kThreads) s_fm_z[i] = g_fm_z[i];
    }
    if constexpr (!PX::kSliced || !PZ::kSliced) __syncthreads();

    auto fm_x = [s_fm_x](uint32_t k) { return (uint32_t)s_fm_x[k]; };
    auto fm_z = [s_fm_z](uint32_t k) { return (uint32_t)s_fm_z[k]; };
    auto co_x = [g_co_x](uint32_t k) { return __ldg(g_co_x + k); };
    auto co_z = [g_co_z](uint32_t k) { return __ldg(g_co_z + k); };

    Counters c = {0u, 0u, 0u, 0u, 0u};
    const int64_t units = (io.words + VEC - 1) / VEC;
    const int64_t step = (int64_t)gridDim.x * kThreads;
    for (int64_t u = (int64_t)blockIdx.x * kThreads + threadIdx.x; u < units; u += step)
        process_unit<PX, PZ, VEC, SAMPLE>(px, pz, io, u, fm_x, co_x, fm_z, co_z, c);
    block_tally(c, io.tally);
}

template <int NB, int MB, int VEC, bool SAMPLE>
__global__ void __launch_bounds__(kThreads)
k_small_generic(const __grid_constant__ GenericArgs a) {
    GenericPolicy<NB, MB> px{&a.x}, pz{&a.z};
    run_small<GenericPolicy<NB, MB>, GenericPolicy<NB, MB>, VEC, SAMPLE>(
        px, pz, a.io, a.x.lut_fm, a.x.lut_corr, a.z.lut_fm, a.z.lut_corr);
}

template <class DX, class DZ, int VEC, bool SAMPLE>
__global__ void __launch_bounds__(kThreads)
k_small_named(const __grid_constant__ NamedArgs a) {
    StaticPolicy<DX> px;
    StaticPolicy<DZ> pz;
    run_small<StaticPolicy<DX>, StaticPolicy<DZ>, VEC, SAMPLE>(px, pz, a.io, a.fm_x, a.co_x,
                                                                a.fm_z, a.co_z);
}

// ---- launch plumbing --------------------------------------------------------------------------
int g_sm_count = 0;

cudaError_t sm_count(int* out) {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
    }
    *out = g_sm_count;
    return cudaSuccess;
}

template <class Kernel, class Args>
cudaError_t launch(Kernel kernel, const Args& args, int64_t units, size_t smem, cudaStream_t stream) {
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

size_t lut_smem(const GenericSide& x, const GenericSide& z, bool x_lut, bool z_lut) {
    size_t nx = (x_lut && x.lut_fm != nullptr) ? ((size_t)1 << x.m) : 0;
    size_t nz = (z_lut && z.lut_fm != nullptr) ? ((size_t)1 << z.m) : 0;
    return ((nx + 15) & ~(size_t)15) + nz + 16;
}

template <int NB, int MB, int VEC>
cudaError_t launch_generic(const SmallLaunch& l, cudaStream_t stream) {
    GenericArgs a;
    a.x = *l.x;
    a.z = *l.z;
    a.io = l.io;
    const bool lut = (MB != kSlicedM);
    const size_t smem = lut_smem(*l.x, *l.z, lut, lut);
    const int64_t units = (l.io.words + VEC - 1) / VEC;
    if (l.sample) return launch(k_small_generic<NB, MB, VEC, true>, a, units, smem, stream);
    return launch(k_small_generic<NB, MB, VEC, false>, a, units, smem, stream);
}

template <class DX, class DZ>
cudaError_t launch_named(const SmallLaunch& l, cudaStream_t stream) {
    constexpr int VEC = 4;
    NamedArgs a;
    a.fm_x = l.x->lut_fm;
    a.co_x = l.x->lut_corr;
    a.fm_z = l.z->lut_fm;
    a.co_z = l.z->lut_corr;
    a.io = l.io;
    const size_t smem = lut_smem(*l.x, *l.z, !DX::kSliced, !DZ::kSliced);
    const int64_t units = (l.io.words + VEC - 1) / VEC;
    if (l.sample) return launch(k_small_named<DX, DZ, VEC, true>, a, units, smem, stream);
    return launch(k_small_named<DX, DZ, VEC, false>, a, units, smem, stream);
}

bool side_matches(const GenericSide& s, const uint32_t* rowmask, uint32_t lmask, const named::SideInfo& d) {
    if (s.n != d.n || s.m != d.m || s.mode == kModeNone) return false;
    if ((s.has_miss != 0) != d.has_miss || lmask != d.l) return false;
    for (int t = 0; t < d.m; ++t)
        if (rowmask[t] != d.rows[t]) return false;
    if (d.sliced) {
        if (s.tt_flip != d.tt_flip || s.tt_miss != d.tt_miss) return false;
        for (int j = 0; j < d.n; ++j)
            if (s.tt_corr[j] != d.tt_corr[j]) return false;
    }
    return true;
}

}  // namespace

int small_bucket_m(int mx, int mz) {
    int m = mx > mz ? mx : mz;
    if (m <= kSlicedM) return kSlicedM;
    if (m <= 8) return 8;
    return 16;
}

int match_named(const GenericSide& x, const uint32_t* rows_x, uint32_t lx, const GenericSide& z,
                const uint32_t* rows_z, uint32_t lz) {
    for (int i = 0; i < named::kNumNamed; ++i)
        if (side_matches(x, rows_x, lx, named::kNamed[i].x) && side_matches(z, rows_z, lz, named::kNamed[i].z))
            return i;
    return -1;
}

const char* named_name(int id) {
    return (id >= 0 && id < named::kNumNamed) ? named::kNamed[id].name : "generic";
}

cudaError_t launch_small(const SmallLaunch& l, cudaStream_t stream) {
#define QCSS_NAMED_CASE(ID, DX, DZ) \
    if (l.named_id == ID) return launch_named<named::DX, named::DZ>(l, stream);
    QCSS_FOR_EACH_NAMED(QCSS_NAMED_CASE)
#undef QCSS_NAMED_CASE
    const int n = l.x->n;
    const int mb = small_bucket_m(l.x->m, l.z->m);
    if (n <= 16) {
        if (mb == kSlicedM) return launch_generic<16, kSlicedM, 4>(l, stream);
        if (mb == 8) return launch_generic<16, 8, 4>(l, stream);
        return launch_generic<16, 16, 4>(l, stream);
    }
    if (mb == kSlicedM) return launch_generic<32, kSlicedM, 2>(l, stream);
    if (mb == 8) return launch_generic<32, 8, 2>(l, stream);
    return launch_generic<32, 16, 2>(l, stream);
}

}  // namespace qcss
