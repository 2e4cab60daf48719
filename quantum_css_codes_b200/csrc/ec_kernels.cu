// Kernels for the repeated-Steane-EC Pauli-frame Monte Carlo (ec_rounds.cuh).
// One thread owns one 32-shot word for all rounds: the state is 2 (m + 1) registers, nothing is read from
// or written to HBM except the six tallies, and the grid is one full wave striding over words.  Static
// instantiations for the three benchmark descriptors, generic (runtime H in the parameter block) otherwise.
#include <cstdlib>

#include "named_codes.inc"
#include "small_common.cuh"
#include "ec_rounds.cuh"

namespace qcss {

namespace {

using small::kThreads;
using small::SideTables;

struct EcGenericArgs {
    GenericSide x, z;
    EcParams ec;
};

struct EcNamedArgs {
    const uint8_t* fm_x;
    const uint32_t* co_x;
    const uint32_t* e32_x;
    const uint8_t* fm_z;
    const uint32_t* co_z;
    const uint32_t* e32_z;
    EcParams ec;
};

template <class PX, class PZ>
__device__ __forceinline__ void run_ec(const PX& px, const PZ& pz, const EcParams& ec, const SideTables& tx,
                                       const SideTables& tz) {
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* cursor = smem;
    const SideLut lut_x = small::stage_side<PX, true>(px, tx, cursor);
    const SideLut lut_z = small::stage_side<PZ, true>(pz, tz, cursor);
    __shared__ GapTable s_tab[2];                    // per-lane indexing on the samplers' rare path
    if (threadIdx.x < 32) {
        s_tab[0].cdf[threadIdx.x] = ec.tab_p.cdf[threadIdx.x];
        s_tab[1].cdf[threadIdx.x] = ec.tab_q.cdf[threadIdx.x];
    }
    if (threadIdx.x == 32) { s_tab[0].inv = ec.tab_p.inv; s_tab[1].inv = ec.tab_q.inv; }
    __syncthreads();

    Counters c = {0u, 0u, 0u, 0u, 0u};
    const int64_t step = (int64_t)gridDim.x * kThreads;
    for (int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x; w < ec.words; w += step)
        process_ec_word<PX, PZ>(px, pz, ec, s_tab[0], s_tab[1], w, lut_x, lut_z, c);
    small::block_tally(c, ec.tally);
}

template <int NB, int MB>
__global__ void __launch_bounds__(kThreads, 2)
k_ec_generic(const __grid_constant__ EcGenericArgs a) {
    GenericPolicy<NB, MB> px{&a.x}, pz{&a.z};
    const SideTables tx{a.x.lut_fm, a.x.lut_corr, a.x.lut_e32}, tz{a.z.lut_fm, a.z.lut_corr, a.z.lut_e32};
    run_ec(px, pz, a.ec, tx, tz);
}

// 3 CTAs/SM when both sides decode by mux tree (Steane: 4.45 -> 4.12 ms per 1e9 shot-rounds); with a table side
// the 80-register cap spills (QRM-15: 22.3 -> 33.3 ms), so those stay at 2.
template <class DX, class DZ>
__global__ void __launch_bounds__(kThreads, (DX::kSliced && DZ::kSliced) ? 3 : 2)
k_ec_named(const __grid_constant__ EcNamedArgs a) {
    StaticPolicy<DX> px;
    StaticPolicy<DZ> pz;
    const SideTables tx{a.fm_x, a.co_x, a.e32_x}, tz{a.fm_z, a.co_z, a.e32_z};
    run_ec(px, pz, a.ec, tx, tz);
}

// ---- CTA-wide two-phase form (both error rates below 1/64, static descriptors) ------------------------------
// Same idea as small_common.cuh::k_small_named_gapq: in the in-place kernel a warp walks the gap logic whenever any
// of its lanes holds an error, three draws per qubit per round.  Here, per round: (1) every thread computes the
// first-look Philox blocks of its 3 n site-words (one per eight sites) and queues those that may hold an error as (thread, stream, qubit); (2) the queue
// is handed out one item per lane, the draw finished and its error words XORed into the owner's DELTA rows in shared
// memory -- [dS_x | dl_x | dS_z | dl_z | H.a_x | H.b_x | L.b_x | H.b_z], the five places ec_rounds.cuh folds a
// draw into; (3) the owner applies the deltas to its register state (S, l), clears them and performs the two
// measurements exactly as process_ec_word does.  Two barriers per round (block or warp scope, see below).
template <class DX, class DZ>
struct EcqShape {
    static constexpr int kRows = EcDeltaRows<DX::MB, DZ::MB>::kRows;
    static constexpr size_t kSmem = (size_t)kRows * kThreads * 4 + (size_t)kThreads * 3 * DX::N * 2;
};

template <class DX, class DZ>
__global__ void __launch_bounds__(kThreads, (DX::kSliced && DZ::kSliced) ? 3 : 2)
k_ec_named_q(const __grid_constant__ EcNamedArgs a) {
    using PX = StaticPolicy<DX>;
    using PZ = StaticPolicy<DZ>;
    constexpr int MBX = PX::MB, MBZ = PZ::MB, N = DX::N, ROWS = EcqShape<DX, DZ>::kRows;
    PX px;
    PZ pz;
    const EcParams& ec = a.ec;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* cursor = smem;
    const SideLut lut_x = small::stage_side<PX, true>(px, SideTables{a.fm_x, a.co_x, a.e32_x}, cursor);
    const SideLut lut_z = small::stage_side<PZ, true>(pz, SideTables{a.fm_z, a.co_z, a.e32_z}, cursor);
    uint32_t* const acc = reinterpret_cast<uint32_t*>(cursor);                  // [ROWS][kThreads]
    uint16_t* const queue0 = reinterpret_cast<uint16_t*>(acc + ROWS * kThreads);  // [kThreads * 3 * N]
    __shared__ GapTable s_tab[2];
    __shared__ int q_count[kThreads / 32 + 1][2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) { s_tab[0].cdf[tid] = ec.tab_p.cdf[tid]; s_tab[1].cdf[tid] = ec.tab_q.cdf[tid]; }
    if (tid == 32) { s_tab[0].inv = ec.tab_p.inv; s_tab[1].inv = ec.tab_q.inv; }
    if (tid <= kThreads / 32) q_count[tid][0] = q_count[tid][1] = 0;
    for (int i = tid; i < ROWS * kThreads; i += kThreads) acc[i] = 0u;
    __syncthreads();
    const uint32_t look_p = gap_look16(s_tab[0].cdf[31]) << 16, look_q = gap_look16(s_tab[1].cdf[31]) << 16;
    Philox ph;
    ph.k0 = (uint32_t)ec.seed;
    ph.k1 = (uint32_t)(ec.seed >> 32);
    // queue scope as in small_common.cuh::run_small_gapq: per warp (warp barriers only) for the table-decoded codes when
    // a warp's 96 n site-words of a round yield enough items, CTA-wide otherwise
    const bool warpq = !(PX::kSliced && PZ::kSliced) &&
                       __umulhi(s_tab[0].cdf[31], (uint32_t)(32 * N)) + 2u * __umulhi(s_tab[1].cdf[31], (uint32_t)(32 * N)) >= 12u;
    uint16_t* const queue = queue0 + (warpq ? warp * (32 * 3 * N) : 0);
    int (*const qcs)[2] = warpq ? &q_count[warp] : &q_count[kThreads / 32];
    const int me = warpq ? lane : tid, team = warpq ? 32 : kThreads;
    auto phase_barrier = [warpq]() {
        if (warpq) __syncwarp();
        else __syncthreads();
    };

    Counters c = {0u, 0u, 0u, 0u, 0u};
    const int64_t step = (int64_t)gridDim.x * kThreads;
    const int64_t cta0 = (int64_t)blockIdx.x * kThreads;
    const int64_t iters = cta0 < ec.words ? (ec.words - cta0 + step - 1) / step : 0;     // uniform over the CTA
    int phase = 0;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t wbase = cta0 + it * step;
        const int64_t w = wbase + tid;
        const bool active = w < ec.words;
        const uint64_t g = ec.first_word + (uint64_t)w;
        const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
        uint32_t sx[MBX], sz[MBZ], lx = 0u, lz = 0u;
#pragma unroll
        for (int t = 0; t < MBX; ++t) sx[t] = 0u;
#pragma unroll
        for (int t = 0; t < MBZ; ++t) sz[t] = 0u;
#pragma unroll 1
        for (int r = 0; r < ec.rounds; ++r, phase ^= 1) {
            const uint32_t base = (uint32_t)(3 * r) << 5;
            int* const qc = &(*qcs)[phase];
            if (active) {
                uint32_t hit[3];
                int total = 0;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const uint32_t look_hi = k == 0 ? look_p : look_q;
                    uint32_t m = 0u;
#pragma unroll
                    for (int jo = 0; jo < (N + 7) / 8; ++jo) {   // eight sites share their first-look block (core.cuh)
                        uint32_t hb[4];
                        gap_first8(ph, g_lo, g_hi, ((base + 32u * k) >> 3) + (uint32_t)jo, hb);
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (8 * jo + c < N && gap_look(hb, c, look_hi)) m |= 1u << (8 * jo + c);
                    }
                    hit[k] = m;
                    total += (int)popc32(m);
                }
                if (total != 0) {                                // one shared atomic per thread
                    int at = atomicAdd(qc, total);
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        for (uint32_t m = hit[k]; m != 0u; m &= m - 1u)
                            queue[at++] = (uint16_t)((tid << 7) | (k << 5) | (int)ctz32(m));
                }
            }
            phase_barrier();
            const int count = *qc;
            if (me == 0) (*qcs)[phase ^ 1] = 0;
            for (int i = me; i < count; i += team) {
                const int item = queue[i], owner = item >> 7, k = (item >> 5) & 3, j = item & 31;
                const GapTable& tab = s_tab[k == 0 ? 0 : 1];
                uint32_t x, z;
                sample_site_word_gap(ec.seed, ec.first_word + (uint64_t)(wbase + owner), base + 32u * k + (uint32_t)j, tab,
                                     tab.cdf[31], x, z);
                uint32_t* const mine = acc + owner;
                ec_fold_draw(px, pz, k, j, x, z, [mine](int row, uint32_t v) { atomicXor(mine + row * kThreads, v); });
            }
            phase_barrier();
            if (active) {
                uint32_t* const mine = acc + tid;
                ec_apply_round(px, pz, sx, lx, sz, lz, lut_x, lut_z, w, [mine](int row) {
                    const uint32_t v = mine[row * kThreads];
                    mine[row * kThreads] = 0u;
                    return v;
                });
            }
        }
        if (active) {
            const uint32_t valid = (w == ec.words - 1) ? ec.tail_mask : 0xFFFFFFFFu;
            const WordOut ox = finish_side<true>(px, sx, lx, lut_x, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
            const WordOut oz = finish_side<true>(pz, sz, lz, lut_z, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
            c.fail_x += popc32(ox.flip & valid);
            c.fail_z += popc32(oz.flip & valid);
            c.fail_any += popc32((ox.flip | oz.flip) & valid);
            c.miss_x += popc32(ox.miss & valid);
            c.miss_z += popc32(oz.miss & valid);
        }
    }
    small::block_tally(c, ec.tally);
}

template <int NB, int MB>
cudaError_t launch_generic(const EcLaunch& l, cudaStream_t stream) {
    EcGenericArgs a;
    a.x = *l.x;
    a.z = *l.z;
    a.ec = l.ec;
    const bool lut = (MB != kSlicedM);
    return small::launch_one(k_ec_generic<NB, MB>, a, l.ec.words, small::lut_smem(*l.x, *l.z, lut, lut, true), stream);
}

template <class DX, class DZ>
cudaError_t launch_named(const EcLaunch& l, cudaStream_t stream) {
    EcNamedArgs a;
    a.fm_x = l.x->lut_fm;
    a.co_x = l.x->lut_corr;
    a.e32_x = l.x->lut_e32;
    a.fm_z = l.z->lut_fm;
    a.co_z = l.z->lut_corr;
    a.e32_z = l.z->lut_e32;
    a.ec = l.ec;
    const size_t lut = small::lut_smem(*l.x, *l.z, !DX::kSliced, !DZ::kSliced, true);
    const bool gapq_off = !l.gapq;
    if (l.ec.gap_p && l.ec.gap_q && !gapq_off && lut + EcqShape<DX, DZ>::kSmem <= 200 * 1024)
        return small::launch_one(k_ec_named_q<DX, DZ>, a, l.ec.words, lut + EcqShape<DX, DZ>::kSmem, stream);
    return small::launch_one(k_ec_named<DX, DZ>, a, l.ec.words, lut, stream);
}

}  // namespace

cudaError_t launch_ec_rounds(const EcLaunch& l, cudaStream_t stream) {
#define QCSS_EC_CASE(ID, DX, DZ) \
    if (l.named_id == ID) return launch_named<named::DX, named::DZ>(l, stream);
    QCSS_FOR_EACH_NAMED(QCSS_EC_CASE)
#undef QCSS_EC_CASE
    const int mb = small_bucket_m(l.x->m, l.z->m);
    if (l.x->n <= 16) {
        if (mb == kSlicedM) return launch_generic<16, kSlicedM>(l, stream);
        if (mb == 8) return launch_generic<16, 8>(l, stream);
        return launch_generic<16, 16>(l, stream);
    }
    if (mb == kSlicedM) return launch_generic<32, kSlicedM>(l, stream);
    if (mb == 8) return launch_generic<32, 8>(l, stream);
    return launch_generic<32, 16>(l, stream);
}

}  // namespace qcss
