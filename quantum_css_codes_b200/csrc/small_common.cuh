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
#include "decode.cuh"
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>

#include "launch.h"
#endif

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

// kernel bodies of the static family, shared by the templates below and by the extern "C" entry points of a
// translation unit specialised for one code (nvcc: qcss_code_spec_source; NVRTC: spec_nvrtc.cu)
template <class DX, class DZ, int VEC, bool SAMPLE, bool FAST>
__device__ __forceinline__ void named_body(const NamedArgs& a) {
    StaticPolicy<DX> px;
    StaticPolicy<DZ> pz;
    const SideTables tx{a.fm_x, a.co_x, a.e32_x}, tz{a.fm_z, a.co_z, a.e32_z};
    run_small<StaticPolicy<DX>, StaticPolicy<DZ>, VEC, SAMPLE, FAST>(px, pz, a.io, tx, tz);
}

template <class DX, class DZ, int VEC, bool SAMPLE, bool FAST>
__global__ void __launch_bounds__(kThreads, FAST ? (SAMPLE ? 3 : 2) : 1)
k_small_named(const __grid_constant__ NamedArgs a) {
    named_body<DX, DZ, VEC, SAMPLE, FAST>(a);
}

// ---- fused gap sampler, CTA-wide two-phase form (Monte-Carlo tallies of the static kernels, p < 1/64) ----------
// k_small_named finishes every draw in place: thread = one 32-shot word, loop over the n qubits, and whenever ANY
// lane of the warp holds an error at that qubit (65 % of the warp-instructions at p = 1e-3, for one useful lane on
// average) the whole warp walks the gap logic -- ncu: 92 instructions per site-word of which the Philox rounds are
// 35 (profiles/r01_mc_fused_steane_gap_ncu_summary.txt).  Here a CTA iteration has three phases:
//   1  every thread computes the first-look Philox blocks of its site-words -- one block per EIGHT qubits (core.cuh) --
//      and pushes the few sites that may hold an error (3 %) to a shared-memory queue as (thread, qubit); nothing else
//      is kept;
//   2  the queue is handed out one item per lane: redo the block, finish the draw (same streams => same bits), and
//      XOR the error word into the owning thread's syndrome / logical accumulators in shared memory
//      (acc[row][thread]; rows of H and L are compile-time masks, the qubit index is the only runtime operand);
//   3  every thread reads its accumulators back, clears them, decodes and tallies as before.
// Bit-identical to the in-place kernel (tests compare both with the oracle); option "gapq" = 0 (qcss_set_option) selects the in-place one.
// Words per thread per CTA iteration: as many as keep the accumulators within ~48 KB (the three block barriers of
// an iteration are amortised over W * n site-words per thread: Steane 4, QRM-15 2, Golay-23 1).
template <class PX, class PZ>
struct GapqShape {
    static constexpr int kRows = PX::MB + PZ::MB + 2;
    static constexpr int kW = (kRows * 4 * kThreads * 4 <= 36 * 1024) ? 4 : ((kRows * 2 * kThreads * 4 <= 48 * 1024) ? 2 : 1);
    static constexpr size_t kSmem = (size_t)kRows * kW * kThreads * 4 + (size_t)kThreads * kW * PX::NB * 2;
};

template <class PX, class PZ>
__device__ __forceinline__ void run_small_gapq(const PX& px, const PZ& pz, const DecodeIO& io, const SideTables& tx,
                                               const SideTables& tz) {
    constexpr int MBX = PX::MB, MBZ = PZ::MB, ROWS = GapqShape<PX, PZ>::kRows, W = GapqShape<PX, PZ>::kW, N = PX::NB;
    static_assert(PX::NB == PZ::NB && N <= 32 && kThreads <= 256 && W <= 4, "queue items are (thread:8, word:2, qubit:5)");
    const int n = px.n();
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* cursor = smem;
    const SideLut lut_x = stage_side<PX, true>(px, tx, cursor);
    const SideLut lut_z = stage_side<PZ, true>(pz, tz, cursor);
    uint32_t* const acc = reinterpret_cast<uint32_t*>(cursor);                      // [ROWS][W][kThreads]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ GapTable s_gap;
    __shared__ int q_count[kThreads / 32 + 1][2];
    if (tid < 32) s_gap.cdf[tid] = io.gap.cdf[tid];
    if (tid == 32) s_gap.inv = io.gap.inv;
    if (tid <= kThreads / 32) q_count[tid][0] = q_count[tid][1] = 0;
    for (int i = tid; i < ROWS * W * kThreads; i += kThreads) acc[i] = 0u;
    __syncthreads();
    const uint32_t cdf31 = s_gap.cdf[31], look_hi = gap_look16(cdf31) << 16;
    Philox ph;
    ph.k0 = (uint32_t)io.seed;
    ph.k1 = (uint32_t)(io.seed >> 32);
    // Queue scope (uniform over the launch).  CTA-wide: one queue, block barriers between the phases, phase 2 runs in
    // as few warps as the items need.  Per warp: eight queues, warp barriers only -- every warp pays at least one
    // phase-2 pass, so it needs enough items per warp (expected 32 W n cdf31 / 2^32 >= 12) and only pays off where the
    // decode phase is uneven enough for block barriers to hurt (table-decoded sides: Golay-23 +11 %, QRM-15 +4 % at
    // p = 1e-3; Steane, both sides mux trees: equal at 1e-3, -15 % at 1e-4).
    const bool warpq = !(PX::kSliced && PZ::kSliced) && __umulhi(cdf31, (uint32_t)(32 * W * n)) >= 12u;
    uint16_t* const queue = reinterpret_cast<uint16_t*>(acc + ROWS * W * kThreads) + (warpq ? warp * (32 * W * N) : 0);
    int (*const qcs)[2] = warpq ? &q_count[warp] : &q_count[kThreads / 32];
    const int me = warpq ? lane : tid, team = warpq ? 32 : kThreads;
    auto phase_barrier = [warpq]() {
        if (warpq) __syncwarp();
        else __syncthreads();
    };

    Counters c = {0u, 0u, 0u, 0u, 0u};
    const int64_t units = io.words / W;                      // FAST: whole units only; thread-unit u = words W u .. W u + W - 1
    const int64_t step = (int64_t)gridDim.x * kThreads;
    const int64_t cta0 = (int64_t)blockIdx.x * kThreads;
    const int64_t iters = cta0 < units ? (units - cta0 + step - 1) / step : 0;      // uniform over the CTA
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t ubase = cta0 + it * step;
        const int64_t u = ubase + tid;
        const bool active = u < units;
        int* const qc = &(*qcs)[it & 1];
        // ---- 1: first looks (one block per EIGHT qubits, core.cuh); the hits collect in one mask per word, branch-free,
        //         and are pushed with one shared atomic per thread ----
        if (active) {
            uint32_t hit[W];
            int total = 0;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint64_t g = io.first_word + (uint64_t)(u * W + w);
                const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
                uint32_t m = 0u;
#pragma unroll
                for (int jo = 0; jo < (N + 7) / 8; ++jo) {
                    if (8 * jo < n) {
                        uint32_t hb[4];
                        gap_first8(ph, g_lo, g_hi, (uint32_t)jo, hb);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (8 * jo + k < N && gap_look(hb, k, look_hi)) m |= 1u << (8 * jo + k);
                    }
                }
                if (n < N) m &= (1u << n) - 1u;
                hit[w] = m;
                total += (int)popc32(m);
            }
            if (total != 0) {
                int at = atomicAdd(qc, total);
#pragma unroll
                for (int w = 0; w < W; ++w)
                    for (uint32_t m = hit[w]; m != 0u; m &= m - 1u)
                        queue[at++] = (uint16_t)((tid << 7) | (w << 5) | (int)ctz32(m));
            }
        }
        phase_barrier();
        // ---- 2: finish the queued draws, one per lane ----
        const int count = *qc;
        if (me == 0) (*qcs)[(it + 1) & 1] = 0;               // the other counter: untouched until the next phase 1
        for (int k = me; k < count; k += team) {
            const int item = queue[k], owner = item >> 7, w = (item >> 5) & 3, j = item & 31;
            uint32_t x, z;
            sample_site_word_gap(io.seed, io.first_word + (uint64_t)((ubase + owner) * W + w), (uint32_t)j, s_gap, cdf31, x, z);
            uint32_t* const mine = acc + w * kThreads + owner;                       // + row * W * kThreads
            if (x != 0u) {
#pragma unroll
                for (int t = 0; t < MBX; ++t)
                    if (px.rowbit(t, j)) atomicXor(mine + t * W * kThreads, x);
                if (px.lbit(j)) atomicXor(mine + MBX * W * kThreads, x);
            }
            if (z != 0u) {
#pragma unroll
                for (int t = 0; t < MBZ; ++t)
                    if (pz.rowbit(t, j)) atomicXor(mine + (MBX + 1 + t) * W * kThreads, z);
                if (pz.lbit(j)) atomicXor(mine + (MBX + 1 + MBZ) * W * kThreads, z);
            }
        }
        phase_barrier();
        // ---- 3: decode and tally ----
        if (active) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                uint32_t* const mine = acc + w * kThreads + tid;
                uint32_t sx[MBX], sz[MBZ];
#pragma unroll
                for (int t = 0; t < MBX; ++t) { sx[t] = mine[t * W * kThreads]; mine[t * W * kThreads] = 0u; }
                const uint32_t lex = mine[MBX * W * kThreads];
                mine[MBX * W * kThreads] = 0u;
#pragma unroll
                for (int t = 0; t < MBZ; ++t) { sz[t] = mine[(MBX + 1 + t) * W * kThreads]; mine[(MBX + 1 + t) * W * kThreads] = 0u; }
                const uint32_t lez = mine[(MBX + 1 + MBZ) * W * kThreads];
                mine[(MBX + 1 + MBZ) * W * kThreads] = 0u;
                const int64_t word = u * W + w;
                const WordOut ox = finish_side<true>(px, sx, lex, lut_x, nullptr, 0, nullptr, 0, nullptr, nullptr, word, 0xFFFFFFFFu);
                const WordOut oz = finish_side<true>(pz, sz, lez, lut_z, nullptr, 0, nullptr, 0, nullptr, nullptr, word, 0xFFFFFFFFu);
                c.fail_x += popc32(ox.flip);
                c.fail_z += popc32(oz.flip);
                c.fail_any += popc32(ox.flip | oz.flip);
                c.miss_x += popc32(ox.miss);
                c.miss_z += popc32(oz.miss);
            }
        }
    }
    block_tally(c, io.tally);
}

template <class DX, class DZ>
__device__ __forceinline__ void named_gapq_body(const NamedArgs& a) {
    StaticPolicy<DX> px;
    StaticPolicy<DZ> pz;
    run_small_gapq(px, pz, a.io, SideTables{a.fm_x, a.co_x, a.e32_x}, SideTables{a.fm_z, a.co_z, a.e32_z});
}

template <class DX, class DZ>
__global__ void __launch_bounds__(kThreads, 4)
k_small_named_gapq(const __grid_constant__ NamedArgs a) {
    named_gapq_body<DX, DZ>(a);
}

template <int NB, int MB>
__global__ void __launch_bounds__(kThreads, 2)
k_small_generic_gapq(const __grid_constant__ GenericArgs a) {
    GenericPolicy<NB, MB> px{&a.x}, pz{&a.z};
    run_small_gapq(px, pz, a.io, SideTables{a.x.lut_fm, a.x.lut_corr, a.x.lut_e32},
                   SideTables{a.z.lut_fm, a.z.lut_corr, a.z.lut_e32});
}

#if !defined(__CUDACC_RTC__)
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
    if (l.sample && l.io.use_gap && l.gapq) {
        using Shape = GapqShape<GenericPolicy<NB, MB>, GenericPolicy<NB, MB>>;
        return launch_split<Shape::kW>(k_small_generic_gapq<NB, MB>, k_small_generic<NB, MB, SVEC, true, false>, a, l,
                                       smem_fast + Shape::kSmem, smem, stream);
    }
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
    if (l.sample && l.io.use_gap && l.gapq) {
        // Monte-Carlo tallies below p = 1/64: the CTA-wide two-phase gap sampler for the whole words
        using Shape = GapqShape<StaticPolicy<DX>, StaticPolicy<DZ>>;
        return launch_split<Shape::kW>(k_small_named_gapq<DX, DZ>, k_small_named<DX, DZ, SVEC, true, false>, a, l,
                                       smem_fast + Shape::kSmem, smem, stream);
    }
    if (l.sample)
        return launch_split<SVEC>(k_small_named<DX, DZ, SVEC, true, true>, k_small_named<DX, DZ, SVEC, true, false>,
                                  a, l, smem_fast, smem, stream);
    return launch_split<VEC>(k_small_named<DX, DZ, VEC, false, true>, k_small_named<DX, DZ, VEC, false, false>,
                             a, l, smem_fast, smem, stream);
}

#endif  // !__CUDACC_RTC__

}  // namespace small

#if !defined(__CUDACC_RTC__)
// one definition per translation unit (small_named_*.cu, small_generic*.cu)
cudaError_t launch_small_steane(const SmallLaunch& l, cudaStream_t stream);
cudaError_t launch_small_qrm15(const SmallLaunch& l, cudaStream_t stream);
cudaError_t launch_small_golay23(const SmallLaunch& l, cudaStream_t stream);
cudaError_t launch_small_generic16(const SmallLaunch& l, int mb, cudaStream_t stream);
cudaError_t launch_small_generic32(const SmallLaunch& l, int mb, cudaStream_t stream);

#endif  // !__CUDACC_RTC__

}  // namespace qcss
