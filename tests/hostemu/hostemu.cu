// TEST INFRASTRUCTURE ONLY -- never loaded by the product package.
// Runs the per-thread bodies of the CUDA kernels (csrc/core.cuh, csrc/decode.cuh: the same
// __host__ __device__ functions the sm_100a kernels call) sequentially on the host, so the bit
// logic (transposes, truth-table muxes, table walk, Philox sampler, static code descriptors)
// can be checked against the oracle in the GPU-less build container.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../quantum_css_codes_b200/csrc/decode.cuh"
#include "../../quantum_css_codes_b200/csrc/ec_rounds.cuh"
#include "../../quantum_css_codes_b200/csrc/host_compact.h"
#include "../../quantum_css_codes_b200/csrc/named_codes.inc"

using namespace qcss;

namespace {

bool g_fast = false;

template <class PX, class PZ, int VEC, bool SAMPLE>
void run(const PX& px, const PZ& pz, const DecodeIO& io, const GenericSide* gx, const GenericSide* gz,
         uint64_t* tally) {
    // FAST instantiation when the caller asks for tallies only over whole units (mirrors
    // small_common.cuh::launch_split for the part of the batch the FAST kernel would get)
    const bool fast = g_fast;
    const SideLut lut_x{gx->lut_fm, gx->lut_corr, (const uint8_t*)gx->lut_e32};
    const SideLut lut_z{gz->lut_fm, gz->lut_corr, (const uint8_t*)gz->lut_e32};
    const int64_t units = (io.words + VEC - 1) / VEC;
    for (int64_t u = 0; u < units; ++u) {
        Counters c = {0, 0, 0, 0, 0};
        if (fast) process_unit<PX, PZ, VEC, SAMPLE, true>(px, pz, io, u, lut_x, lut_z, c);
        else process_unit<PX, PZ, VEC, SAMPLE, false>(px, pz, io, u, lut_x, lut_z, c);
        tally[1] += c.fail_x; tally[2] += c.fail_z; tally[3] += c.fail_any;
        tally[4] += c.miss_x; tally[5] += c.miss_z;
    }
}

template <int NB, int MB, int VEC>
void run_generic(const GenericSide* x, const GenericSide* z, const DecodeIO& io, int sample, uint64_t* tally) {
    GenericPolicy<NB, MB> px{x}, pz{z};
    if (sample) run<GenericPolicy<NB, MB>, GenericPolicy<NB, MB>, VEC, true>(px, pz, io, x, z, tally);
    else run<GenericPolicy<NB, MB>, GenericPolicy<NB, MB>, VEC, false>(px, pz, io, x, z, tally);
}

template <class DX, class DZ>
void run_named(const GenericSide* x, const GenericSide* z, const DecodeIO& io, int sample, uint64_t* tally) {
    StaticPolicy<DX> px;
    StaticPolicy<DZ> pz;
    if (sample) run<StaticPolicy<DX>, StaticPolicy<DZ>, 4, true>(px, pz, io, x, z, tally);
    else run<StaticPolicy<DX>, StaticPolicy<DZ>, 4, false>(px, pz, io, x, z, tally);
}

template <class PX, class PZ>
void run_ec(const PX& px, const PZ& pz, const EcParams& ec, const GenericSide* gx, const GenericSide* gz,
            uint64_t* tally) {
    const SideLut lut_x{gx->lut_fm, gx->lut_corr, (const uint8_t*)gx->lut_e32};
    const SideLut lut_z{gz->lut_fm, gz->lut_corr, (const uint8_t*)gz->lut_e32};
    for (int64_t w = 0; w < ec.words; ++w) {
        Counters c = {0, 0, 0, 0, 0};
        process_ec_word<PX, PZ>(px, pz, ec, ec.tab_p, ec.tab_q, w, lut_x, lut_z, c);
        tally[1] += c.fail_x; tally[2] += c.fail_z; tally[3] += c.fail_any;
        tally[4] += c.miss_x; tally[5] += c.miss_z;
    }
}

// The CTA-wide two-phase gap sampler (small_common.cuh::run_small_gapq) replayed sequentially: the same phases
// over the same primitives (Philox::block, sample_site_word_gap, rowbit / lbit, finish_side), with a CTA's
// threads visited one after the other where the kernel has block barriers.  words must be a multiple of W.
template <class PX, class PZ, int W>
void run_gapq(const PX& px, const PZ& pz, const DecodeIO& io, const GenericSide* gx, const GenericSide* gz,
              uint64_t* tally) {
    constexpr int T = 256, MBX = PX::MB, MBZ = PZ::MB, ROWS = MBX + MBZ + 2, N = PX::NB;
    const SideLut lut_x{gx->lut_fm, gx->lut_corr, (const uint8_t*)gx->lut_e32};
    const SideLut lut_z{gz->lut_fm, gz->lut_corr, (const uint8_t*)gz->lut_e32};
    static uint32_t acc[ROWS * W * T];
    static uint16_t queue[T * W * N];
    memset(acc, 0, sizeof(acc));
    Philox ph;
    ph.k0 = (uint32_t)io.seed;
    ph.k1 = (uint32_t)(io.seed >> 32);
    const uint32_t cdf31 = io.gap.cdf[31], look_hi = gap_look16(cdf31) << 16;
    const int n = px.n();
    const int64_t units = io.words / W;
    for (int64_t ubase = 0; ubase < units; ubase += T) {
        int count = 0;
        for (int tid = 0; tid < T && ubase + tid < units; ++tid)                       // phase 1
            for (int w = 0; w < W; ++w) {
                const uint64_t g = io.first_word + (uint64_t)((ubase + tid) * W + w);
                for (int jo = 0; 8 * jo < n; ++jo) {                                  // eight sites share their first-look block
                    uint32_t hb[4];
                    gap_first8(ph, (uint32_t)g, (uint32_t)(g >> 32), (uint32_t)jo, hb);
                    for (int c = 0; c < 8; ++c) {
                        const int j = 8 * jo + c;
                        if (j < n && gap_look(hb, c, look_hi)) queue[count++] = (uint16_t)((tid << 7) | (w << 5) | j);
                    }
                }
            }
        for (int k = count - 1; k >= 0; --k) {                                         // phase 2, any order
            const int item = queue[k], owner = item >> 7, w = (item >> 5) & 3, j = item & 31;
            uint32_t x, z;
            sample_site_word_gap(io.seed, io.first_word + (uint64_t)((ubase + owner) * W + w), (uint32_t)j, io.gap, cdf31, x, z);
            uint32_t* mine = acc + w * T + owner;
            for (int t = 0; t < MBX; ++t)
                if (px.rowbit(t, j)) mine[t * W * T] ^= x;
            if (px.lbit(j)) mine[MBX * W * T] ^= x;
            for (int t = 0; t < MBZ; ++t)
                if (pz.rowbit(t, j)) mine[(MBX + 1 + t) * W * T] ^= z;
            if (pz.lbit(j)) mine[(MBX + 1 + MBZ) * W * T] ^= z;
        }
        for (int tid = 0; tid < T && ubase + tid < units; ++tid)                       // phase 3
            for (int w = 0; w < W; ++w) {
                uint32_t* mine = acc + w * T + tid;
                uint32_t sx[MBX], sz[MBZ];
                for (int t = 0; t < MBX; ++t) { sx[t] = mine[t * W * T]; mine[t * W * T] = 0u; }
                const uint32_t lex = mine[MBX * W * T];
                mine[MBX * W * T] = 0u;
                for (int t = 0; t < MBZ; ++t) { sz[t] = mine[(MBX + 1 + t) * W * T]; mine[(MBX + 1 + t) * W * T] = 0u; }
                const uint32_t lez = mine[(MBX + 1 + MBZ) * W * T];
                mine[(MBX + 1 + MBZ) * W * T] = 0u;
                const int64_t word = (ubase + tid) * W + w;
                const WordOut ox = finish_side<true>(px, sx, lex, lut_x, nullptr, 0, nullptr, 0, nullptr, nullptr, word, 0xFFFFFFFFu);
                const WordOut oz = finish_side<true>(pz, sz, lez, lut_z, nullptr, 0, nullptr, 0, nullptr, nullptr, word, 0xFFFFFFFFu);
                tally[1] += popc32(ox.flip); tally[2] += popc32(oz.flip); tally[3] += popc32(ox.flip | oz.flip);
                tally[4] += popc32(ox.miss); tally[5] += popc32(oz.miss);
            }
    }
}

template <int NB, int MB, int W>
void run_gapq_generic(const GenericSide* x, const GenericSide* z, const DecodeIO& io, uint64_t* tally) {
    GenericPolicy<NB, MB> px{x}, pz{z};
    run_gapq<GenericPolicy<NB, MB>, GenericPolicy<NB, MB>, W>(px, pz, io, x, z, tally);
}

// ec_kernels.cu::k_ec_named_q replayed sequentially (one CTA of 256 words at a time): first blocks -> queue ->
// ec_fold_draw into the owners' delta rows -> ec_apply_round, over the kernel's own helpers and row map.
template <class PX, class PZ>
void run_ecq(const PX& px, const PZ& pz, const EcParams& ec, const GenericSide* gx, const GenericSide* gz, uint64_t* tally) {
    constexpr int T = 256, MBX = PX::MB, MBZ = PZ::MB, ROWS = EcDeltaRows<MBX, MBZ>::kRows, N = PX::NB;
    const SideLut lut_x{gx->lut_fm, gx->lut_corr, (const uint8_t*)gx->lut_e32};
    const SideLut lut_z{gz->lut_fm, gz->lut_corr, (const uint8_t*)gz->lut_e32};
    static uint32_t acc[ROWS * T];
    static uint16_t queue[T * 3 * N];
    static uint32_t sx[T][MBX], sz[T][MBZ], lx[T], lz[T];
    memset(acc, 0, sizeof(acc));
    Philox ph;
    ph.k0 = (uint32_t)ec.seed;
    ph.k1 = (uint32_t)(ec.seed >> 32);
    const int n = px.n();
    for (int64_t wbase = 0; wbase < ec.words; wbase += T) {
        const int live = (int)((ec.words - wbase) < T ? (ec.words - wbase) : T);
        memset(sx, 0, sizeof(sx)); memset(sz, 0, sizeof(sz)); memset(lx, 0, sizeof(lx)); memset(lz, 0, sizeof(lz));
        for (int r = 0; r < ec.rounds; ++r) {
            const uint32_t base = (uint32_t)(3 * r) << 5;
            int count = 0;
            for (int tid = 0; tid < live; ++tid) {
                const uint64_t g = ec.first_word + (uint64_t)(wbase + tid);
                for (int k = 0; k < 3; ++k)
                    for (int jo = 0; 8 * jo < n; ++jo) {
                        uint32_t hb[4];
                        gap_first8(ph, (uint32_t)g, (uint32_t)(g >> 32), ((base + 32u * k) >> 3) + (uint32_t)jo, hb);
                        for (int c = 0; c < 8; ++c) {
                            const int j = 8 * jo + c;
                            if (j < n && gap_look(hb, c, gap_look16(k == 0 ? ec.tab_p.cdf[31] : ec.tab_q.cdf[31]) << 16))
                                queue[count++] = (uint16_t)((tid << 7) | (k << 5) | j);
                        }
                    }
            }
            for (int i = count - 1; i >= 0; --i) {
                const int item = queue[i], owner = item >> 7, k = (item >> 5) & 3, j = item & 31;
                const GapTable& tab = k == 0 ? ec.tab_p : ec.tab_q;
                uint32_t x, z;
                sample_site_word_gap(ec.seed, ec.first_word + (uint64_t)(wbase + owner), base + 32u * k + (uint32_t)j, tab,
                                     tab.cdf[31], x, z);
                uint32_t* mine = acc + owner;
                ec_fold_draw(px, pz, k, j, x, z, [mine](int row, uint32_t v) { mine[row * T] ^= v; });
            }
            for (int tid = 0; tid < live; ++tid) {
                uint32_t* mine = acc + tid;
                ec_apply_round(px, pz, sx[tid], lx[tid], sz[tid], lz[tid], lut_x, lut_z, wbase + tid, [mine](int row) {
                    const uint32_t v = mine[row * T];
                    mine[row * T] = 0u;
                    return v;
                });
            }
        }
        for (int tid = 0; tid < live; ++tid) {
            const int64_t w = wbase + tid;
            const uint32_t valid = (w == ec.words - 1) ? ec.tail_mask : 0xFFFFFFFFu;
            const WordOut ox = finish_side<true>(px, sx[tid], lx[tid], lut_x, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
            const WordOut oz = finish_side<true>(pz, sz[tid], lz[tid], lut_z, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
            tally[1] += popc32(ox.flip & valid); tally[2] += popc32(oz.flip & valid);
            tally[3] += popc32((ox.flip | oz.flip) & valid);
            tally[4] += popc32(ox.miss & valid); tally[5] += popc32(oz.miss & valid);
        }
    }
}

template <int NB, int MB>
void run_ec_generic(const GenericSide* x, const GenericSide* z, const EcParams& ec, uint64_t* tally) {
    GenericPolicy<NB, MB> px{x}, pz{z};
    run_ec(px, pz, ec, x, z, tally);
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int emu_sizeof_ec(void) { return (int)sizeof(EcParams); }

// small_common.cuh::launch_named / launch_generic on the gap path: W as GapqShape chooses it (Steane 4, QRM-15 2,
// Golay-23 1; generic 2 for the 5- and 8-row buckets, 1 for 16).  io->words must be a multiple of 4.
__attribute__((visibility("default")))
int emu_mc_gapq(const GenericSide* x, const GenericSide* z, const DecodeIO* io, int named_id, uint64_t* tally) {
    if (named_id == 0) { run_gapq<StaticPolicy<named::Steane_X>, StaticPolicy<named::Steane_Z>, 4>({}, {}, *io, x, z, tally); return 0; }
    if (named_id == 1) { run_gapq<StaticPolicy<named::Qrm15_X>, StaticPolicy<named::Qrm15_Z>, 2>({}, {}, *io, x, z, tally); return 0; }
    if (named_id == 2) { run_gapq<StaticPolicy<named::Golay23_X>, StaticPolicy<named::Golay23_Z>, 1>({}, {}, *io, x, z, tally); return 0; }
    const int m = x->m > z->m ? x->m : z->m;
    const int mb = m <= kSlicedM ? kSlicedM : (m <= 8 ? 8 : 16);
    if (x->n <= 16) {
        if (mb == kSlicedM) run_gapq_generic<16, kSlicedM, 2>(x, z, *io, tally);
        else if (mb == 8) run_gapq_generic<16, 8, 2>(x, z, *io, tally);
        else run_gapq_generic<16, 16, 1>(x, z, *io, tally);
    } else {
        if (mb == kSlicedM) run_gapq_generic<32, kSlicedM, 2>(x, z, *io, tally);
        else if (mb == 8) run_gapq_generic<32, 8, 2>(x, z, *io, tally);
        else run_gapq_generic<32, 16, 1>(x, z, *io, tally);
    }
    return 0;
}

// ec_kernels.cu::k_ec_named_q (static descriptors, both rates below 1/64)
__attribute__((visibility("default")))
int emu_ecq(const GenericSide* x, const GenericSide* z, const EcParams* ec, int named_id, uint64_t* tally) {
#define EMU_ECQ_CASE(ID, DX, DZ) \
    if (named_id == ID) { run_ecq(StaticPolicy<named::DX>{}, StaticPolicy<named::DZ>{}, *ec, x, z, tally); return 0; }
    QCSS_FOR_EACH_NAMED(EMU_ECQ_CASE)
#undef EMU_ECQ_CASE
    return -1;
}

// ec_kernels.cu::launch_ec_rounds, one word after the other on the host
__attribute__((visibility("default")))
int emu_ec(const GenericSide* x, const GenericSide* z, const EcParams* ec, int named_id, uint64_t* tally) {
#define EMU_EC_CASE(ID, DX, DZ) \
    if (named_id == ID) { run_ec(StaticPolicy<named::DX>{}, StaticPolicy<named::DZ>{}, *ec, x, z, tally); return 0; }
    QCSS_FOR_EACH_NAMED(EMU_EC_CASE)
#undef EMU_EC_CASE
    const int m = x->m > z->m ? x->m : z->m;
    const int mb = m <= kSlicedM ? kSlicedM : (m <= 8 ? 8 : 16);
    if (x->n <= 16) {
        if (mb == kSlicedM) run_ec_generic<16, kSlicedM>(x, z, *ec, tally);
        else if (mb == 8) run_ec_generic<16, 8>(x, z, *ec, tally);
        else run_ec_generic<16, 16>(x, z, *ec, tally);
    } else {
        if (mb == kSlicedM) run_ec_generic<32, kSlicedM>(x, z, *ec, tally);
        else if (mb == 8) run_ec_generic<32, 8>(x, z, *ec, tally);
        else run_ec_generic<32, 16>(x, z, *ec, tally);
    }
    return 0;
}

__attribute__((visibility("default"))) int emu_sizeof_side(void) { return (int)sizeof(GenericSide); }
__attribute__((visibility("default"))) int emu_sizeof_io(void) { return (int)sizeof(DecodeIO); }
// 1: run the FAST (tally-only, whole-unit) instantiation; caller guarantees words % VEC == 0,
// both sides present and no output planes.
__attribute__((visibility("default"))) void emu_set_fast(int on) { g_fast = on != 0; }

// named_id < 0: generic kernels with the launch_small bucket rule; else the static descriptor.
__attribute__((visibility("default")))
int emu_decode(const GenericSide* x, const GenericSide* z, const DecodeIO* io, int named_id, int sample,
               uint64_t* tally) {
#define EMU_CASE(ID, DX, DZ) \
    if (named_id == ID) { run_named<named::DX, named::DZ>(x, z, *io, sample, tally); return 0; }
    QCSS_FOR_EACH_NAMED(EMU_CASE)
#undef EMU_CASE
    const int m = x->m > z->m ? x->m : z->m;
    const int mb = m <= kSlicedM ? kSlicedM : (m <= 8 ? 8 : 16);
    if (x->n <= 16) {
        if (mb == kSlicedM) run_generic<16, kSlicedM, 4>(x, z, *io, sample, tally);
        else if (mb == 8) run_generic<16, 8, 4>(x, z, *io, sample, tally);
        else run_generic<16, 16, 4>(x, z, *io, sample, tally);
    } else {
        if (mb == kSlicedM) run_generic<32, kSlicedM, 2>(x, z, *io, sample, tally);
        else if (mb == 8) run_generic<32, 8, 2>(x, z, *io, sample, tally);
        else run_generic<32, 16, 2>(x, z, *io, sample, tally);
    }
    return 0;
}

__attribute__((visibility("default")))
void emu_transpose32(uint32_t* w) {
    uint32_t a[32];
    memcpy(a, w, sizeof(a));
    transpose_blocks<32>(a);
    memcpy(w, a, sizeof(a));
}

__attribute__((visibility("default")))
void emu_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
    Philox px;
    px.k0 = key[0];
    px.k1 = key[1];
    uint32_t o[4];
    px.block(ctr[0], ctr[1], ctr[2], ctr[3], o);
    memcpy(out, o, sizeof(o));
}

__attribute__((visibility("default")))
int emu_named_side(int id, int which_x, int* n, int* m, uint32_t* rows, uint32_t* l) {
    if (id < 0 || id >= named::kNumNamed) return -1;
    const named::SideInfo& s = which_x ? named::kNamed[id].x : named::kNamed[id].z;
    *n = s.n; *m = s.m; *l = s.l;
    for (int t = 0; t < 16; ++t) rows[t] = s.rows[t];
    return 0;
}

// The compacting host->device path of qcss_decode_xz: csrc/host_compact.cpp (the product's code, linked into this
// library) compacts one chunk the way api.cu::decode_xz_compacted lays it out -- `threads` workers with contiguous task
// ranges and their own value regions -- and the loop below rebuilds the chunk with the index arithmetic of
// format_kernels.cu::k_zs_expand (lane l: bitmap word l and the exclusive prefix of the population counts; round r:
// words 32 r .. 32 r + 31, value rank = prefix + bits below).  0 = ok, 1 = a region overflowed.
__attribute__((visibility("default")))
int emu_zs_roundtrip(const uint64_t* ex, const uint64_t* ez, int64_t e_stride, int n, int64_t w0, int64_t cw, int threads,
                     uint64_t* out_x, uint64_t* out_z) {
    const int rows = 2 * n, bpr = (int)((cw + kZsBlockWords - 1) / kZsBlockWords), tasks = rows * bpr;
    const size_t region_cap = ((size_t)rows * cw / 2) / threads + kZsBlockWords + 8;
    std::vector<uint64_t> bm((size_t)tasks * 32, ~0ull), vals((size_t)threads * region_cap, 0x5555555555555555ull);
    std::vector<uint32_t> off((size_t)tasks, 0xFFFFFFFFu);
    for (int t = 0; t < threads; ++t) {
        const size_t used = zs_compact_range(ex, ez, e_stride, n, w0, cw, bpr, (int)((int64_t)tasks * t / threads),
                                             (int)((int64_t)tasks * (t + 1) / threads), bm.data(), off.data(), vals.data(),
                                             (size_t)t * region_cap, region_cap);
        if (used == SIZE_MAX) return 1;
    }
    for (int task = 0; task < tasks; ++task) {
        const int row = task / bpr, blk = task % bpr;
        uint64_t* dst = (row < n ? out_x + (int64_t)row * cw : out_z + (int64_t)(row - n) * cw) + (int64_t)blk * kZsBlockWords;
        const int64_t left = cw - (int64_t)blk * kZsBlockWords;
        int excl[32], acc = 0;
        for (int l = 0; l < 32; ++l) { excl[l] = acc; acc += __builtin_popcountll(bm[(size_t)task * 32 + l]); }
        const uint64_t* v = vals.data() + off[task];
        for (int r = 0; r < 64; ++r)
            for (int lane = 0; lane < 32; ++lane) {
                const uint64_t word = bm[(size_t)task * 32 + (r >> 1)];
                const int bit = ((r & 1) << 5) + lane;
                uint64_t value = 0;
                if ((word >> bit) & 1ull) value = v[excl[r >> 1] + __builtin_popcountll(word & ((1ull << bit) - 1ull))];
                if (32 * r + lane < left) dst[32 * r + lane] = value;
            }
    }
    return 0;
}

}  // extern "C"
