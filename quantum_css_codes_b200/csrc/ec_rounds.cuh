// Pauli-frame Monte Carlo of repeated Steane error correction: the gadget the reference
// emits in CSSCode.error_correct (css_code.py:436-470) -- transversal CNOT data -> |+>_L ancilla, Z-basis
// measurement, quil_classical_correct with parity_check_c2 / _c2_syndromes on the X frame; then CNOT
// |0>_L ancilla -> data, X-basis measurement, quil_classical_correct with parity_check_c1 / _c1_syndromes
// on the Z frame -- tracked as Pauli errors instead of amplitudes, so the K1/K2/K3 device functions are its
// whole inner loop.  The reference runs this circuit on a QVM (test/test_fidelity.py); it has no Pauli
// model, so WHERE the noise enters is this file's to define; the gadget itself is pinned: oracle/ec_rounds.py states
// the same rounds in error space (bit-exact on identical Philox streams) and tests/test_ec_gadget.py checks that
// statement against the instruction stream the unmodified reference emits for error_correct:
//
//   per round r = 0 .. rounds-1, per shot
//     1. data block:   e ^= depolarising(p_data)                         Philox stream 3r
//     2. ancilla A (verified |+>_L, encode_plus css_code.py:345-366): a = depolarising(p_anc), stream 3r+1
//        CNOT data -> A copies X errors forward and Z errors back:       e_z ^= a_z
//        measured word = codeword ^ e_x ^ a_x; frame update as quil_classical_correct (css_code.py:649-685):
//            c = table2.get(key(H2 . (e_x ^ a_x ^ f_x)), 0);  f_x ^= c
//     3. ancilla B (verified |0>_L, encode_zero css_code.py:314-343): b = depolarising(p_anc), stream 3r+2
//        CNOT B -> data copies B's X errors onto the data (after step 2): e_x ^= b_x
//        measured word (X basis) = codeword ^ e_z ^ b_z;  c = table1.get(key(H1 . (e_z ^ b_z ^ f_z)), 0);  f_z ^= c
//   after the last round: ideal decode of the residual e ^ f with the same tables; a logical failure is
//   L . (e ^ f ^ c) = 1 exactly as in the single-shot Monte Carlo (decode.cuh); a miss is a residual
//   syndrome absent from the table.
//
// Everything above is linear in the errors except the table lookups, which only see syndromes.  So a thread
// keeps, per 32-shot word and Pauli type, only S = H . (e ^ f) (m bit-sliced words) and l = L . (e ^ f) (one
// word): errors are folded in as they are drawn (policy add()), a measurement looks up key(S ^ H . a), and a hit
// replaces S by H . a (the ancilla's own error is what the correction leaves behind) and flips l by L . c
// -- the flip bit the K2 tables already store.  No error or frame plane ever exists in memory.
//
// Philox counters: (word_lo, word_hi, site, block) with site = 32 * stream + qubit, key = seed; stream 0 of
// round 0 is the single-shot sampler's stream, so rounds = 1 with p_anc = 0 reproduces qcss_mc_run bit for bit.
#pragma once
#include "decode.cuh"

namespace qcss {

struct EcParams {
    unsigned long long* tally;   // [6]
    int64_t words;               // ceil(shots / 32)
    uint32_t tail_mask;          // valid bits of word words-1
    int32_t rounds;
    uint64_t seed;
    uint64_t first_word;
    uint32_t thr_p, thr_q;       // floor(p * 2^32) for the data and ancilla error rates
    uint32_t gap_p, gap_q;       // non-zero: gap sampler (rate < 1/64) with the table below
    GapTable tab_p, tab_q;
};

QCSS_HD void ec_draw(uint32_t use_gap, uint64_t seed, uint64_t g, uint32_t site, uint32_t thr, const GapTable& tab,
                     uint32_t& x, uint32_t& z) {
    if (use_gap) sample_site_word_gap(seed, g, site, tab, tab.cdf[31], x, z);
    else sample_site_word(seed, g, site, thr, x, z);
}

// lookup on the measured syndrome; a hit moves the frame: S <- H . a, l ^= L . c
template <class P>
QCSS_HD void ec_measure(const P& pol, uint32_t (&s)[P::MB], uint32_t& l, const uint32_t (&anc)[P::MB],
                        const SideLut& lut, int64_t w) {
    uint32_t meas[P::MB];
#pragma unroll
    for (int t = 0; t < P::MB; ++t) meas[t] = s[t] ^ anc[t];
    const WordOut o = finish_side<true>(pol, meas, 0u, lut, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
    l ^= o.flip;
#pragma unroll
    for (int t = 0; t < P::MB; ++t) s[t] = (s[t] & o.miss) | (anc[t] & ~o.miss);
}

template <class PX, class PZ>
QCSS_HD void process_ec_word(const PX& px, const PZ& pz, const EcParams& ec, const GapTable& tab_p,
                             const GapTable& tab_q, int64_t w, const SideLut& lut_x, const SideLut& lut_z,
                             Counters& c) {
    static_assert(PX::NB == PZ::NB, "sides share the qubit count");
    constexpr int NB = PX::NB;
    const int n = px.n();
    const uint64_t g = ec.first_word + (uint64_t)w;
    const uint32_t valid = (w == ec.words - 1) ? ec.tail_mask : 0xFFFFFFFFu;
    uint32_t sx[PX::MB], sz[PZ::MB], lx = 0u, lz = 0u;
#pragma unroll
    for (int t = 0; t < PX::MB; ++t) sx[t] = 0u;
#pragma unroll
    for (int t = 0; t < PZ::MB; ++t) sz[t] = 0u;

#pragma unroll 1
    for (int r = 0; r < ec.rounds; ++r) {
        uint32_t ax[PX::MB], bx[PX::MB], bz[PZ::MB], bxl = 0u, unused = 0u;
#pragma unroll
        for (int t = 0; t < PX::MB; ++t) ax[t] = bx[t] = 0u;
#pragma unroll
        for (int t = 0; t < PZ::MB; ++t) bz[t] = 0u;
        const uint32_t base = (uint32_t)(3 * r) << 5;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            if (j < n) {
                uint32_t x, z;
                ec_draw(ec.gap_p, ec.seed, g, base + (uint32_t)j, ec.thr_p, tab_p, x, z);          // data
                px.add(j, x, sx, lx);
                pz.add(j, z, sz, lz);
                ec_draw(ec.gap_q, ec.seed, g, base + 32u + (uint32_t)j, ec.thr_q, tab_q, x, z);    // ancilla A
                px.add(j, x, ax, unused);
                pz.add(j, z, sz, lz);                                                              // Z back-action
                ec_draw(ec.gap_q, ec.seed, g, base + 64u + (uint32_t)j, ec.thr_q, tab_q, x, z);    // ancilla B
                px.add(j, x, bx, bxl);                                                             // X back-action, applied below
                pz.add(j, z, bz, unused);
            }
        }
        ec_measure(px, sx, lx, ax, lut_x, w);
#pragma unroll
        for (int t = 0; t < PX::MB; ++t) sx[t] ^= bx[t];
        lx ^= bxl;
        ec_measure(pz, sz, lz, bz, lut_z, w);
    }
    const WordOut ox = finish_side<true>(px, sx, lx, lut_x, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
    const WordOut oz = finish_side<true>(pz, sz, lz, lut_z, nullptr, 0, nullptr, 0, nullptr, nullptr, w, 0xFFFFFFFFu);
    c.fail_x += popc32(ox.flip & valid);
    c.fail_z += popc32(oz.flip & valid);
    c.fail_any += popc32((ox.flip | oz.flip) & valid);
    c.miss_x += popc32(ox.miss & valid);
    c.miss_z += popc32(oz.miss & valid);
}

// ---- pieces of the CTA-wide two-phase form (ec_kernels.cu::k_ec_named_q), shared with the host replay ----------
// Per round the finished draws are folded into DELTA rows of the owning thread; where a draw goes is the model above:
//   data (stream 0)      x -> S_x, l_x          z -> S_z, l_z
//   ancilla A (stream 1) x -> H.a_x             z -> S_z, l_z   (back-action through CNOT data -> A)
//   ancilla B (stream 2) x -> H.b_x, L.b_x      z -> H.b_z      (b_x reaches the data after the X measurement)
template <int MBX, int MBZ>
struct EcDeltaRows {
    static constexpr int kSX = 0, kLX = MBX, kSZ = MBX + 1, kLZ = MBX + 1 + MBZ, kAX = kLZ + 1, kBX = kAX + MBX,
                         kBXL = kBX + MBX, kBZ = kBXL + 1, kRows = kBZ + MBZ;
};

// xor_row(row, value): XOR `value` into delta row `row` of the draw's owner
template <class PX, class PZ, class XorRow>
QCSS_HD void ec_fold_draw(const PX& px, const PZ& pz, int stream, int j, uint32_t x, uint32_t z, XorRow xor_row) {
    using R = EcDeltaRows<PX::MB, PZ::MB>;
    const int rx = stream == 0 ? R::kSX : (stream == 1 ? R::kAX : R::kBX);
    const int rz = stream == 2 ? R::kBZ : R::kSZ;
    if (x != 0u) {
#pragma unroll
        for (int t = 0; t < PX::MB; ++t)
            if (px.rowbit(t, j)) xor_row(rx + t, x);
        if (stream != 1 && px.lbit(j)) xor_row(stream == 0 ? R::kLX : R::kBXL, x);
    }
    if (z != 0u) {
#pragma unroll
        for (int t = 0; t < PZ::MB; ++t)
            if (pz.rowbit(t, j)) xor_row(rz + t, z);
        if (stream != 2 && pz.lbit(j)) xor_row(R::kLZ, z);
    }
}

// take(row): read-and-clear delta row `row` of this thread.  Applies one round to the register state.
template <class PX, class PZ, class Take>
QCSS_HD void ec_apply_round(const PX& px, const PZ& pz, uint32_t (&sx)[PX::MB], uint32_t& lx, uint32_t (&sz)[PZ::MB],
                            uint32_t& lz, const SideLut& lut_x, const SideLut& lut_z, int64_t w, Take take) {
    using R = EcDeltaRows<PX::MB, PZ::MB>;
    uint32_t ax[PX::MB], bx[PX::MB], bz[PZ::MB];
#pragma unroll
    for (int t = 0; t < PX::MB; ++t) { sx[t] ^= take(R::kSX + t); ax[t] = take(R::kAX + t); bx[t] = take(R::kBX + t); }
#pragma unroll
    for (int t = 0; t < PZ::MB; ++t) { sz[t] ^= take(R::kSZ + t); bz[t] = take(R::kBZ + t); }
    lx ^= take(R::kLX);
    lz ^= take(R::kLZ);
    const uint32_t bxl = take(R::kBXL);
    ec_measure(px, sx, lx, ax, lut_x, w);
#pragma unroll
    for (int t = 0; t < PX::MB; ++t) sx[t] ^= bx[t];
    lx ^= bxl;
    ec_measure(pz, sz, lz, bz, lut_z, w);
}

}  // namespace qcss
