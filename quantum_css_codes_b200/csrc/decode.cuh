// Small-code (n <= 32) syndrome + lookup-decode + logical-check pipeline for one "unit" of
// VEC 32-shot words.  Shared by the sm_100a kernels and the test-only host emulation.
#pragma once
#include "core.cuh"

namespace qcss {

struct TagTrue { static constexpr bool value = true; };
struct TagFalse { static constexpr bool value = false; };

struct Counters {
    uint32_t fail_x, fail_z, fail_any, miss_x, miss_z;
};

// Device-side view of qcss_decode_io in 32-bit words (all strides in uint32 units).
struct DecodeIO {
    const uint32_t* ex;
    const uint32_t* ez;
    int64_t e_stride;
    uint32_t* synd_x;
    uint32_t* synd_z;
    int64_t s_stride;
    uint32_t* corr_x;
    uint32_t* corr_z;
    int64_t c_stride;
    uint32_t* flip_x;
    uint32_t* flip_z;
    uint32_t* miss_x;
    uint32_t* miss_z;
    unsigned long long* tally;   // [6]
    int64_t words;               // ceil(shots / 32)
    uint32_t tail_mask;          // valid bits of word words-1
    int32_t sides;               // bit0: X side (which = 2), bit1: Z side (which = 1)
    // fused sampler
    uint32_t* ex_out;            // sampled planes written here when non-null
    uint32_t* ez_out;
    uint64_t seed;
    uint64_t first_word;         // global index of word 0 (first_shot / 32)
    uint32_t thr;                // floor(p * 2^32)
    uint32_t use_gap;            // p < 1/64: gap sampler (core.cuh) with the table below, else bit-serial
    GapTable gap;
};

// ---- loads ------------------------------------------------------------------------------------
template <int VEC>
QCSS_HD void load_words(const uint32_t* p, uint32_t (&out)[VEC]) {
#if defined(__CUDA_ARCH__)
    if constexpr (VEC == 4) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(out[0]), "=r"(out[1]), "=r"(out[2]), "=r"(out[3]) : "l"(p));
    } else if constexpr (VEC == 2) {
        asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                     : "=r"(out[0]), "=r"(out[1]) : "l"(p));
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) out[v] = __ldg(p + v);
    }
#else
    for (int v = 0; v < VEC; ++v) out[v] = p[v];
#endif
}

template <int VEC>
QCSS_HD void store_words(uint32_t* p, const uint32_t (&in)[VEC]) {
#if defined(__CUDA_ARCH__)
    if constexpr (VEC == 4) {
        *reinterpret_cast<uint4*>(p) = make_uint4(in[0], in[1], in[2], in[3]);
    } else if constexpr (VEC == 2) {
        *reinterpret_cast<uint2*>(p) = make_uint2(in[0], in[1]);
    } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) p[v] = in[v];
    }
#else
    for (int v = 0; v < VEC; ++v) p[v] = in[v];
#endif
}

// ---- lookup tables of one side as the kernels see them ------------------------------------------
// fm / e32 point to shared memory on the device (staged by run_small), corr to global memory.
struct SideLut {
    const uint8_t* fm;        // [2^m] bytes: bit0 = L.corr parity, bit1 = miss
    const uint32_t* corr;     // [2^m] correction masks
    const uint8_t* e32;       // [2^m] 32-bit entries addressed by byte offset (tally kernels)
    QCSS_HD uint32_t read_fm(uint32_t k) const { return (uint32_t)fm[k]; }
    QCSS_HD uint32_t read_corr(uint32_t k) const {
#if defined(__CUDA_ARCH__)
        return __ldg(corr + k);
#else
        return corr[k];
#endif
    }
    QCSS_HD uint32_t read_e32(uint32_t byte_off) const {
        return *reinterpret_cast<const uint32_t*>(e32 + byte_off);
    }
};

// ---- side policies ----------------------------------------------------------------------------
// Generic: H, L and truth tables are runtime data (kernel parameters -> constant bank operands).
//   MB == kSlicedM : fully bit-sliced decode (m <= 5)
//   MB == 8 or 16  : transpose + shared-memory table
template <int NB_, int MB_>
struct GenericPolicy {
    static constexpr int NB = NB_, MB = MB_;
    static constexpr bool kSliced = (MB_ == kSlicedM);
    static constexpr int kTallyM = kSliced ? 0 : (MB_ < kMaxE32M ? MB_ : kMaxE32M);   // lut_tally_word width
    const GenericSide* p;
    QCSS_HD bool use_e32() const { return p->lut_e32 != nullptr; }
    QCSS_HD int n() const { return p->n; }
    QCSS_HD int m() const { return p->m; }
    QCSS_HD bool has_table() const { return p->mode != kModeNone; }
    QCSS_HD bool has_miss() const { return p->has_miss != 0; }
    QCSS_HD uint32_t tt_flip() const { return p->tt_flip; }
    QCSS_HD uint32_t tt_miss() const { return p->tt_miss; }
    QCSS_HD uint32_t tt_corr(int j) const { return p->tt_corr[j]; }
    QCSS_HD void add(int j, uint32_t e, uint32_t (&s)[MB], uint32_t& le) const {
#pragma unroll
        for (int t = 0; t < MB; ++t) s[t] ^= e & p->mask[t][j];
        le ^= e & p->lexp[j];
    }
    // runtime qubit index (two-phase gap sampler): does key-bit row t / the logical row contain qubit j
    QCSS_HD bool rowbit(int t, int j) const { return (p->mask[t][j] & 1u) != 0u; }
    QCSS_HD bool lbit(int j) const { return (p->lexp[j] & 1u) != 0u; }
};

// Static: everything about H, L (and, for sliced sides, the truth tables) is a compile-time
// constant supplied by a descriptor D (named_codes.inc).  XORs against zero entries vanish.
template <class D>
struct StaticPolicy {
    static constexpr int NB = D::N, MB = D::MB;
    static constexpr bool kSliced = D::kSliced;
    static constexpr int kTallyM = (!D::kSliced && D::M <= kMaxE32M) ? D::M : 0;
    QCSS_HD bool use_e32() const { return kTallyM > 0; }
    QCSS_HD int n() const { return D::N; }
    QCSS_HD int m() const { return D::M; }
    QCSS_HD bool has_table() const { return true; }
    QCSS_HD bool has_miss() const { return D::kHasMiss; }
    QCSS_HD uint32_t tt_flip() const { return D::kTtFlip; }
    QCSS_HD uint32_t tt_miss() const { return D::kTtMiss; }
    QCSS_HD uint32_t tt_corr(int j) const { return D::tt_corr(j); }
    QCSS_HD void add(int j, uint32_t e, uint32_t (&s)[MB], uint32_t& le) const {
#pragma unroll
        for (int t = 0; t < MB; ++t)
            if ((D::row(t) >> j) & 1u) s[t] ^= e;
        if ((D::kL >> j) & 1u) le ^= e;
    }
    QCSS_HD bool rowbit(int t, int j) const { return ((D::row(t) >> j) & 1u) != 0u; }
    QCSS_HD bool lbit(int j) const { return ((D::kL >> j) & 1u) != 0u; }
};

// ---- finish one word of one side ---------------------------------------------------------------
// FAST = tally-only instantiation: no optional output planes, every bit of the word valid.
template <bool FAST, class P>
QCSS_HD WordOut finish_side(const P& pol, uint32_t (&s)[P::MB], uint32_t le, const SideLut& lut,
                            uint32_t* synd, int64_t s_stride, uint32_t* corr,
                            int64_t c_stride, uint32_t* flip_p, uint32_t* miss_p, int64_t w,
                            uint32_t valid) {
    constexpr int MB = P::MB;
    const int m = pol.m(), n = pol.n();
    const bool in_range = valid != 0u;
    if constexpr (!FAST) {
        if (synd != nullptr && in_range) {
#pragma unroll
            for (int t = 0; t < MB; ++t)
                if (t < m) synd[(int64_t)(m - 1 - t) * s_stride + w] = s[t] & valid;
        }
    }
    WordOut o;
    o.flip = 0u;
    o.miss = 0u;
    if constexpr (!FAST) {
        if (!pol.has_table()) return o;
    }
    uint32_t fc = 0u;
    if constexpr (P::kSliced) {
        fc = eval_truth_table<MB>(s, pol.tt_flip());
        if (pol.has_miss()) o.miss = eval_truth_table<MB>(s, pol.tt_miss());
        if constexpr (!FAST) {
            if (corr != nullptr && in_range) {
#pragma unroll
                for (int j = 0; j < P::NB; ++j)
                    if (j < n)
                        corr[(int64_t)j * c_stride + w] = eval_truth_table<MB>(s, pol.tt_corr(j)) & valid;
            }
        }
    } else {
        bool done = false;
        if constexpr (FAST && P::kTallyM > 0) {
            if (pol.use_e32()) {
                lut_tally_word<P::kTallyM>(reinterpret_cast<const uint32_t(&)[P::kTallyM]>(s), pol.has_miss(),
                                           [&lut](uint32_t off) { return lut.read_e32(off); }, fc, o.miss);
                done = true;
            }
        }
        if (!done)
            lut_flip_miss<MB>(s, pol.has_miss() ? 1 : 0, [&lut](uint32_t k) { return lut.read_fm(k); }, fc,
                              o.miss);
        if constexpr (!FAST) {
            if (corr != nullptr && in_range) {
                uint32_t planes[32];
                lut_corrections<MB>(s, [&lut](uint32_t k) { return lut.read_corr(k); }, planes);
#pragma unroll
                for (int j = 0; j < P::NB; ++j)
                    if (j < n) corr[(int64_t)j * c_stride + w] = planes[j] & valid;
            }
        }
    }
    o.flip = le ^ fc;
    if constexpr (!FAST) {
        if (flip_p != nullptr && in_range) flip_p[w] = o.flip & valid;
        if (miss_p != nullptr && in_range) miss_p[w] = o.miss & valid;
    }
    return o;
}

// ---- one unit: VEC consecutive words of both sides -------------------------------------------
// FAST units are the hot path of the Monte-Carlo tallies: both Pauli types present, all VEC words
// fully inside the batch, nothing written but the counters.  Everything else (single-side calls,
// output planes, the ragged tail of a batch) goes through the FAST = false instantiation.
//
// Loaded batches are processed one Pauli type after the other (X planes -> X flips, then Z), which
// halves the live syndrome registers; the fused sampler draws X and Z of a qubit from the same
// random words, so there both sides advance together.
template <class P, int VEC>
QCSS_HD void zero_side(uint32_t (&s)[VEC][P::MB], uint32_t (&le)[VEC]) {
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
#pragma unroll
        for (int t = 0; t < P::MB; ++t) s[v][t] = 0u;
        le[v] = 0u;
    }
}

template <class P, int VEC>
QCSS_HD void load_side(const P& pol, const uint32_t* base, int64_t stride, uint32_t (&s)[VEC][P::MB],
                       uint32_t (&le)[VEC]) {
    const int n = pol.n();
#pragma unroll
    for (int j = 0; j < P::NB; ++j) {
        if (j < n) {
            uint32_t e[VEC];
            load_words<VEC>(base + (int64_t)j * stride, e);
#pragma unroll
            for (int v = 0; v < VEC; ++v) pol.add(j, e[v], s[v], le[v]);
        }
    }
}

template <class PX, class PZ, int VEC, bool SAMPLE, bool FAST>
QCSS_HD void process_unit(const PX& px, const PZ& pz, const DecodeIO& io, int64_t unit,
                          const SideLut& lut_x, const SideLut& lut_z, Counters& c,
                          const GapTable* gap_tab = nullptr) {
    static_assert(PX::NB == PZ::NB, "sides share the qubit count");
    constexpr int NB = PX::NB;
    const int64_t w0 = unit * VEC;
    const bool do_x = FAST || SAMPLE || (io.sides & 1), do_z = FAST || SAMPLE || (io.sides & 2);
    uint32_t valid[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        valid[v] = 0xFFFFFFFFu;
        if constexpr (!FAST) {
            const int64_t w = w0 + v;
            valid[v] = (w < io.words) ? ((w == io.words - 1) ? io.tail_mask : 0xFFFFFFFFu) : 0u;
        }
    }
    WordOut ox[VEC], oz[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) ox[v].flip = ox[v].miss = oz[v].flip = oz[v].miss = 0u;

    if constexpr (SAMPLE) {
        uint32_t sx[VEC][PX::MB], sz[VEC][PZ::MB], lex[VEC], lez[VEC];
        zero_side<PX, VEC>(sx, lex);
        zero_side<PZ, VEC>(sz, lez);
        const int n = px.n();
        const GapTable& gtab = (gap_tab != nullptr) ? *gap_tab : io.gap;
        // the sampler choice is hoisted out of the (fully unrolled) site loop: one run executes one copy
        auto sample_all = [&](auto gap_tag) {
            constexpr bool GAP = decltype(gap_tag)::value;
            const uint32_t cdf31 = io.gap.cdf[31];
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                if (j < n) {
                    uint32_t xe[VEC], ze[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if constexpr (GAP)
                            sample_site_word_gap(io.seed, io.first_word + (uint64_t)(w0 + v), (uint32_t)j, gtab, cdf31,
                                                 xe[v], ze[v]);
                        else
                            sample_site_word(io.seed, io.first_word + (uint64_t)(w0 + v), (uint32_t)j, io.thr,
                                             xe[v], ze[v]);
                    }
                    if constexpr (!FAST) {
                        // padding bits (past the last shot) are written as zero, whatever the unit width
                        uint32_t xo[VEC], zo[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v) { xo[v] = xe[v] & valid[v]; zo[v] = ze[v] & valid[v]; }
                        if (io.ex_out != nullptr) store_words<VEC>(io.ex_out + (int64_t)j * io.e_stride + w0, xo);
                        if (io.ez_out != nullptr) store_words<VEC>(io.ez_out + (int64_t)j * io.e_stride + w0, zo);
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        px.add(j, xe[v], sx[v], lex[v]);
                        pz.add(j, ze[v], sz[v], lez[v]);
                    }
                }
            }
        };
        if (io.use_gap) sample_all(TagTrue{});
        else sample_all(TagFalse{});
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            ox[v] = finish_side<FAST>(px, sx[v], lex[v], lut_x, io.synd_x, io.s_stride, io.corr_x, io.c_stride,
                                      io.flip_x, io.miss_x, w0 + v, valid[v]);
            oz[v] = finish_side<FAST>(pz, sz[v], lez[v], lut_z, io.synd_z, io.s_stride, io.corr_z, io.c_stride,
                                      io.flip_z, io.miss_z, w0 + v, valid[v]);
        }
    } else {
        if (do_x) {
            uint32_t sx[VEC][PX::MB], lex[VEC];
            zero_side<PX, VEC>(sx, lex);
            load_side<PX, VEC>(px, io.ex + w0, io.e_stride, sx, lex);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                ox[v] = finish_side<FAST>(px, sx[v], lex[v], lut_x, io.synd_x, io.s_stride, io.corr_x,
                                          io.c_stride, io.flip_x, io.miss_x, w0 + v, valid[v]);
        }
        if (do_z) {
            uint32_t sz[VEC][PZ::MB], lez[VEC];
            zero_side<PZ, VEC>(sz, lez);
            load_side<PZ, VEC>(pz, io.ez + w0, io.e_stride, sz, lez);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                oz[v] = finish_side<FAST>(pz, sz[v], lez[v], lut_z, io.synd_z, io.s_stride, io.corr_z,
                                          io.c_stride, io.flip_z, io.miss_z, w0 + v, valid[v]);
        }
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        c.fail_x += popc32(ox[v].flip & valid[v]);
        c.fail_z += popc32(oz[v].flip & valid[v]);
        c.fail_any += popc32((ox[v].flip | oz[v].flip) & valid[v]);
        c.miss_x += popc32(ox[v].miss & valid[v]);
        c.miss_z += popc32(oz[v].miss & valid[v]);
    }
}

}  // namespace qcss
