// Kernels for the repeated-Steane-EC Pauli-frame Monte Carlo (ec_rounds.cuh).
// One thread owns one 32-shot word for all rounds: the state is 2 (m + 1) registers, nothing is read from
// or written to HBM except the six tallies, and the grid is one full wave striding over words.  Static
// instantiations for the three benchmark descriptors, generic (runtime H in the parameter block) otherwise.
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
    return small::launch_one(k_ec_named<DX, DZ>, a, l.ec.words,
                             small::lut_smem(*l.x, *l.z, !DX::kSliced, !DZ::kSliced, true), stream);
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
