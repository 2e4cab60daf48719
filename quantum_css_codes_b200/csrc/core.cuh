// Per-thread bit logic of the hot path, shared by the sm_100a kernels (kernels.cu) and by the
// test-only host emulation (tests/hostemu).  Everything here is branch-uniform integer work on
// 32-shot words: one bit of a word is one Monte-Carlo shot ("bit-sliced" / plane-major layout).
//
// Reference semantics implemented (jimpo/quantum-css-codes):
//   syndrome      s = H.e mod 2                         css_code.py:728
//   table key     big-endian vec_to_int(s)              bin_matrix.py:36-43
//   lookup decode errors ^= table.get(key, 0)           css_code.py:649-685 (miss => unchanged)
//   logical check L.(e ^ c) mod 2                       css_code.py:124-161, 641-646
#pragma once
#if defined(__CUDACC_RTC__)
// NVRTC (in-process specialisation, spec_nvrtc.cu): no system headers on the deployment box
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long size_t;
#else
#include <stdint.h>
#endif

#if defined(__CUDACC__)
#define QCSS_HD __host__ __device__ __forceinline__
#else
#define QCSS_HD inline
#endif

namespace qcss {

constexpr int kMaxN = 32;        // register-resident ("small code") kernels: n <= 32 qubits
constexpr int kMaxM = 16;        // lookup decode: m <= 16 syndrome bits (table in shared memory)
constexpr int kSlicedM = 5;      // fully bit-sliced decode when m <= 5
constexpr int kMaxE32M = 13;     // tally kernels use 32-bit table entries (4 * 2^m bytes of smem)
constexpr int kSparseMax = 3;    // <= this many non-zero syndromes per word: walk them one by one

enum Mode : int32_t { kModeNone = 0, kModeSliced = 1, kModeLut = 2 };

// One Pauli side of a small code, generic (runtime H).  Internal row order is *key-bit order*:
// row t of `mask` produces bit t of the big-endian table key, i.e. reference row m-1-t.
struct GenericSide {
    int32_t n, m, mode, has_miss;
    uint32_t tt_flip, tt_miss;            // sliced mode: bit k = value at key k
    uint32_t lexp[kMaxN];                 // 0 / ~0 : logical operator row L[j]
    uint32_t mask[kMaxM][kMaxN];          // 0 / ~0 : H[m-1-t][j]
    uint32_t tt_corr[kMaxN];              // sliced mode: correction truth table of qubit j
    const uint8_t* lut_fm;                // lut mode: [2^m] bytes, bit0 = L.corr parity, bit1 = miss
    const uint32_t* lut_corr;             // lut mode: [2^m] correction bit masks (bit j = qubit j)
    const uint32_t* lut_e32;              // lut mode, m <= kMaxE32M: [2^m] words, low half all-ones when
                                          // L.corr is odd, high half all-ones on a miss (tally kernels)
};

// ---------------------------------------------------------------------------------------------
// Bit-matrix transpose inside registers.  w[i] holds plane i for 32 shots; afterwards, for every
// block b of W shots, bits [b*W, b*W+W) of w[j] hold the W plane-bits of shot b*W + j
// (bit i of that field = plane i).  Classic butterfly network, W in {8, 16, 32}.
template <int W>
QCSS_HD void transpose_blocks(uint32_t (&w)[W]) {
#pragma unroll
    for (int d = W / 2; d >= 1; d >>= 1) {
        // mask with d ones then d zeros, repeating
        uint32_t mk = 0;
#pragma unroll
        for (int b = 0; b < 32; ++b) if (((b / d) & 1) == 0) mk |= (1u << b);
#pragma unroll
        for (int j = 0; j < W; ++j) {
            if ((j & d) == 0) {
                uint32_t a = w[j], b = w[j + d];
                uint32_t t = ((a >> d) ^ b) & mk;
                w[j + d] = b ^ t;
                w[j] = a ^ (t << d);
            }
        }
    }
}

// Arbitrary boolean function of M bit-sliced variables given its truth table (bit k = value when
// the variables spell k, v[0] = LSB).  Mux tree, 2^M - 1 three-input logic ops.
template <int M>
QCSS_HD uint32_t eval_truth_table(const uint32_t (&v)[M], uint32_t tt) {
    uint32_t node[1 << (M > 0 ? M - 1 : 0)];
    if (M == 0) return 0u - (tt & 1u);
#pragma unroll
    for (int k = 0; k < (1 << (M - 1)); ++k) {
        uint32_t l0 = 0u - ((tt >> (2 * k)) & 1u), l1 = 0u - ((tt >> (2 * k + 1)) & 1u);
        node[k] = (v[0] & l1) | (~v[0] & l0);
    }
#pragma unroll
    for (int lvl = 1; lvl < M; ++lvl) {
#pragma unroll
        for (int k = 0; k < (1 << (M - 1 - lvl)); ++k)
            node[k] = (v[lvl] & node[2 * k + 1]) | (~v[lvl] & node[2 * k]);
    }
    return node[0];
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11), the counter-based generator of the fused sampler.
struct Philox {
    uint32_t k0, k1;
    QCSS_HD static void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
        lo = a * b;
        hi = __umulhi(a, b);
#else
        uint64_t p = (uint64_t)a * (uint64_t)b;
        lo = (uint32_t)p;
        hi = (uint32_t)(p >> 32);
#endif
    }
    QCSS_HD void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t (&out)[4]) const {
        uint32_t ka = k0, kb = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ ka, n2 = hi0 ^ c3 ^ kb;
            c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};

// Depolarising draw for 32 shots of qubit j (one word): each shot independently suffers a Pauli
// error with probability thr / 2^32, uniformly X, Y or Z.  Returns x = X|Y and z = Z|Y planes.
// Random words come from Philox blocks with key = seed, counter = (g_lo, g_hi, j, block index),
// g = global index of the 32-shot word, so the draw is independent of launch shape and GPU count.
//   stage 1 ("r < thr" for 32 lanes at once): block q = 0..7 supplies the random words for
//     threshold bits 31-4q .. 28-4q (one word per bit, MSB first); lanes whose random bit differs
//     from the threshold bit are decided.  Stops after the first block that leaves no lane
//     undecided (2-3 blocks typically); lanes equal to thr in all 32 bits count as "no error".
//   stage 2 (only if some lane has an error): each further block gives two attempts
//     (w0,w1) and (w2,w3) = (x bits, z bits); (0,0) is rejected and retried, so X, Z, Y each
//     have probability exactly 1/3.
QCSS_HD void sample_site_word(uint64_t seed, uint64_t g, uint32_t j, uint32_t thr,
                              uint32_t& x, uint32_t& z) {
    Philox px;
    px.k0 = (uint32_t)seed;
    px.k1 = (uint32_t)(seed >> 32);
    const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
    uint32_t blk = 0u;
    uint32_t buf[4];
    uint32_t und = 0xFFFFFFFFu, err = 0u;
#pragma unroll 1
    for (int q = 0; q < 8 && und != 0u; ++q) {
        px.block(g_lo, g_hi, j, blk++, buf);
        const uint32_t tq = thr >> (28 - 4 * q);      // low 4 bits: threshold bits of this block
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t tb = 0u - ((tq >> (3 - k)) & 1u);   // all-ones when the bit is set
            const uint32_t r = buf[k];
            err |= und & ~r & tb;
            und &= r ^ ~tb;
        }
    }
    x = 0u;
    z = 0u;
    uint32_t need = err;
#pragma unroll 1
    while (need != 0u) {
        px.block(g_lo, g_hi, j, blk++, buf);
        uint32_t ok = need & (buf[0] | buf[1]);
        x |= ok & buf[0];
        z |= ok & buf[1];
        need &= ~ok;
        ok = need & (buf[2] | buf[3]);
        x |= ok & buf[2];
        z |= ok & buf[3];
        need &= ~ok;
    }
}

// Gap sampler (p < 1/64): instead of deciding 32 lanes bit by bit, draw the number of error-free lanes
// before the next error by inverse CDF.  cdf[k] = floor((1 - (1-p)^(k+1)) * 2^32) (host table, computed
// by repeated multiplication in double precision so that C and numpy agree bit for bit); a 32-bit
// uniform u gives d = #{k : cdf[k] <= u} clean lanes, d = 32 meaning "no further error in this word".
// A draw is a pair (u, tw): tw provides 16 two-bit attempts at the Pauli type ((x, z) != (0, 0), scanned from the
// low end); if all 16 are (0, 0) the whole draw is discarded and redrawn, which keeps gap and type exactly
// independent.  Where the words come from (key = seed throughout):
//   first look at site j  16 bits: half j & 1 (0 = low) of word (j & 7) >> 1 of the block with counter
//                         (g_lo, g_hi, j >> 3, 0) -- EIGHT sites share that block, because at p = 1e-3 97 % of the
//                         site-words are settled by one compare and the sampler's cost is its Philox blocks.  The 16
//                         bits are the HIGH half h of the site's first uniform u0: h > cdf[31] >> 16 already says
//                         u0 >= cdf[31], i.e. no error among the 32 lanes;
//   everything after it   the site's own blocks, counter (g_lo, g_hi, j, q), q = 1, 2, ...: block 1 gives the first
//                         draw's tw = w0, the LOW half of u0 = low 16 bits of w1 (u0 = h << 16 | w1 & 0xffff, tested
//                         exactly against cdf[31]) and the second draw (w2, w3); blocks q >= 2 give two draws each,
//                         (w0, w1) then (w2, w3).
// Expected blocks per site-word: 1/8 + 32 p (bit-serial: 2.2).
// Which sampler a rate takes: thr = floor(p * 2^32) below this bound -> gap sampler.  Measured on B200 with the queue
// kernels (Steane, 2^30 shots): the gap form costs 0.147 + 152 p picoseconds per shot, the bit-serial form a flat 3.4 --
// equal at p = 0.021; 1/64 is the last power of two on the right side (p = 0.0078: 7.3e11 against 3.1e11 shots/s).
constexpr uint32_t kGapThreshold = 1u << 26;

struct GapTable {
    uint32_t cdf[32];
    uint32_t inv;            // floor((2^32 - 1) / max(cdf[0], 1)): first guess d ~ u / cdf[0]
};

QCSS_HD uint32_t gap_count(const GapTable& t, uint32_t u) {
#if defined(__CUDA_ARCH__)
    uint32_t d = __umulhi(u, t.inv);
#else
    uint32_t d = (uint32_t)(((uint64_t)u * t.inv) >> 32);
#endif
    if (d > 32u) d = 32u;
    while (d < 32u && u >= t.cdf[d]) ++d;
    while (d > 0u && u < t.cdf[d - 1u]) --d;
    return d;
}

QCSS_HD uint32_t ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__ffs((int)v) - 1u;
#else
    return (uint32_t)__builtin_ctz(v);
#endif
}

// one draw (u, tw) at lane position pos; returns false when the word is finished
QCSS_HD bool gap_draw(const GapTable& t, uint32_t u, uint32_t tw, uint32_t& pos, uint32_t& x, uint32_t& z) {
    const uint32_t d = gap_count(t, u);
    if (pos + d >= 32u) return false;
    const uint32_t valid = (tw | (tw >> 1)) & 0x55555555u;
    if (valid != 0u) {
        const uint32_t b = ctz32(valid), lane = pos + d;
        x |= ((tw >> b) & 1u) << lane;
        z |= ((tw >> (b + 1u)) & 1u) << lane;
        pos = lane + 1u;
    }
    return pos < 32u;
}

// Rare continuation (a second error in the same word, 5e-4 per word at p = 1e-3): out of line, so the
// 4 x n unrolled call sites stay small.  Resumes with the block's second draw (w2, w3) at lane `pos`.
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
void sample_gap_rest(uint32_t k0, uint32_t k1, uint32_t g_lo, uint32_t g_hi, uint32_t j, const GapTable& t,
                     uint32_t pos, uint32_t w2, uint32_t w3, uint32_t& x, uint32_t& z) {
    Philox px;
    px.k0 = k0;
    px.k1 = k1;
    if (!gap_draw(t, w2, w3, pos, x, z)) return;
    uint32_t buf[4];
    for (uint32_t blk = 2u;; ++blk) {
        px.block(g_lo, g_hi, j, blk, buf);
        if (!gap_draw(t, buf[0], buf[1], pos, x, z)) return;
        if (!gap_draw(t, buf[2], buf[3], pos, x, z)) return;
    }
}

// First-look halves of the eight sites 8 jo .. 8 jo + 7 of word g: one block.
QCSS_HD void gap_first8(const Philox& px, uint32_t g_lo, uint32_t g_hi, uint32_t jo, uint32_t (&h)[4]) {
    px.block(g_lo, g_hi, jo, 0u, h);
}

// the 16 first-look bits of site 8 jo + c (c < 8, a compile-time constant at the unrolled call sites)
QCSS_HD uint32_t gap_half(const uint32_t (&h)[4], int c) {
    return (c & 1) ? (h[c >> 1] >> 16) : (h[c >> 1] & 0xFFFFu);
}

// h < gap_look16(cdf31)  <=>  the site may hold an error (its first uniform can still be below cdf31)
QCSS_HD uint32_t gap_look16(uint32_t cdf31) { return (cdf31 >> 16) + 1u; }

// the same test on the block words without extracting the half: look_hi = gap_look16(cdf31) << 16 (p < 1/64 keeps
// cdf31 below 0.4 * 2^32, so the shift cannot overflow)
QCSS_HD bool gap_look(const uint32_t (&h)[4], int c, uint32_t look_hi) {
    return (c & 1) ? (h[c >> 1] < look_hi) : ((h[c >> 1] << 16) < look_hi);
}

// The gap sampler of site j after its first look h (< gap_look16: the caller has looked) -- split off so that a caller
// can compute the shared first blocks of many sites back to back (independent 10-round chains) before looking at any.
QCSS_HD void gap_finish(const Philox& px, uint32_t g_lo, uint32_t g_hi, uint32_t j, const GapTable& t, uint32_t h,
                        uint32_t& x, uint32_t& z) {
    x = 0u;
    z = 0u;
    uint32_t buf[4];
    px.block(g_lo, g_hi, j, 1u, buf);
    const uint32_t u0 = (h << 16) | (buf[1] & 0xFFFFu);
    if (u0 >= t.cdf[31]) return;                       // no error among the 32 lanes after all (2^-16 of the looks)
    uint32_t pos = 0u;
    gap_draw(t, u0, buf[0], pos, x, z);                // first error: inline (3 % of the words)
    if (pos >= 32u || buf[2] >= t.cdf[31u - pos]) return;              // usually the only one
    sample_gap_rest(px.k0, px.k1, g_lo, g_hi, j, t, pos, buf[2], buf[3], x, z);
}

QCSS_HD void sample_site_word_gap(uint64_t seed, uint64_t g, uint32_t j, const GapTable& t, uint32_t cdf31,
                                  uint32_t& x, uint32_t& z) {
    Philox px;
    px.k0 = (uint32_t)seed;
    px.k1 = (uint32_t)(seed >> 32);
    const uint32_t g_lo = (uint32_t)g, g_hi = (uint32_t)(g >> 32);
    uint32_t hb[4];
    gap_first8(px, g_lo, g_hi, j >> 3, hb);
    const uint32_t wsel = (j & 4u) ? ((j & 2u) ? hb[3] : hb[2]) : ((j & 2u) ? hb[1] : hb[0]);
    const uint32_t h = (j & 1u) ? (wsel >> 16) : (wsel & 0xFFFFu);
    x = 0u;
    z = 0u;
    if (h >= gap_look16(cdf31)) return;                // no error among the 32 lanes
    gap_finish(px, g_lo, g_hi, j, t, h, x, z);
}

QCSS_HD uint32_t popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__popc(v);
#else
    return (uint32_t)__builtin_popcount(v);
#endif
}

// ---------------------------------------------------------------------------------------------
// Side accumulators.  add(j, e) folds one error plane word into the syndrome words; after all n
// planes, finish() yields the bit-sliced decode outputs.
struct WordOut {
    uint32_t flip;      // L.(e ^ c) per shot
    uint32_t miss;      // syndrome not in table
};

// Lookup over transposed keys.  LutRead(k) returns the byte at key k (shared memory on device).
template <int MB, class LutRead>
QCSS_HD void lut_flip_miss(uint32_t (&s)[MB], int has_miss, LutRead rd, uint32_t& fc, uint32_t& miss) {
    static_assert(MB == 8 || MB == 16, "lut key width");
    transpose_blocks<MB>(s);
    fc = 0u; miss = 0u;
    constexpr int kBlocks = 32 / MB;
    constexpr uint32_t kKeyMask = (1u << MB) - 1u;
#pragma unroll
    for (int j = 0; j < MB; ++j) {
#pragma unroll
        for (int b = 0; b < kBlocks; ++b) {
            uint32_t key = (s[j] >> (b * MB)) & kKeyMask;
            uint32_t ent = rd(key);
            fc |= (ent & 1u) << (b * MB + j);
            if (has_miss) miss |= ((ent >> 1) & 1u) << (b * MB + j);
        }
    }
}

// ---- tally-only lookup (FAST kernels), m <= kMaxE32M ------------------------------------------------
QCSS_HD uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7))) & 0xFF) << (8 * i);
    return r;
#endif
}

QCSS_HD int ffs32(uint32_t v) {       // index of the lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// True when `pred` holds for any lane of the warp (host emulation: for this thread).
QCSS_HD bool warp_any(bool pred) {
#if defined(__CUDA_ARCH__)
    return __any_sync(__activemask(), pred) != 0;
#else
    return pred;
#endif
}

// 16x16 transpose of both halves of 16 words; the distance-8 stage is two byte permutes per pair.
QCSS_HD void transpose16(uint32_t (&w)[16]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t a = w[j], b = w[j + 8];
        w[j] = prmt(a, b, 0x6240u);
        w[j + 8] = prmt(a, b, 0x7351u);
    }
#pragma unroll
    for (int d = 4; d >= 1; d >>= 1) {
        const uint32_t mk = d == 4 ? 0x0F0F0F0Fu : (d == 2 ? 0x33333333u : 0x55555555u);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            if ((j & d) == 0) {
                const uint32_t a = w[j], b = w[j + d];
                const uint32_t t = ((a >> d) ^ b) & mk;
                w[j + d] = b ^ t;
                w[j] = a ^ (t << d);
            }
        }
    }
}

// Flip / miss words of one 32-shot word from the key planes s[0..M) (key-bit order), M <= 13.
// rd32(byte_offset) reads the 32-bit entry at key = byte_offset / 4.  Two warp-uniform paths:
//   sparse: at most kSparseMax shots of every lane have a non-zero syndrome (the usual case at
//           physical error rates ~1e-3): gather each such shot's key bit by bit, one lookup each;
//   dense : planes are placed two bit positions up so the 16x16 transpose yields key*4 directly,
//           then 32 lookups whose all-ones entries are masked into place with one LOP3 each.
template <int M, class Rd32>
QCSS_HD void lut_tally_word(const uint32_t (&s)[M], bool has_miss, Rd32 rd32, uint32_t& fc, uint32_t& miss) {
    static_assert(M <= kMaxE32M, "32-bit entry table needs m <= 13");
    uint32_t nz = 0u;
#pragma unroll
    for (int t = 0; t < M; ++t) nz |= s[t];
    const bool dense = warp_any(popc32(nz) > (uint32_t)kSparseMax);
    if (!dense) {
        const uint32_t e0 = rd32(0u);                   // every zero-syndrome shot shares key 0
        fc = (e0 & 1u) ? ~nz : 0u;
        miss = (e0 >> 16) ? ~nz : 0u;
        while (nz != 0u) {
            const int k = ffs32(nz);
            nz &= nz - 1u;
            uint32_t key4 = 0u;
#pragma unroll
            for (int t = M - 1; t >= 0; --t) key4 = (key4 << 1) | ((s[t] >> k) & 1u);
            const uint32_t e = rd32(key4 << 2);
            fc |= (e & 1u) << k;
            miss |= (e >> 31) << k;
        }
        if (!has_miss) miss = 0u;
        return;
    }
    uint32_t w[16];
#pragma unroll
    for (int t = 0; t < 16; ++t) w[t] = (t >= 2 && t - 2 < M) ? s[t - 2] : 0u;
    transpose16(w);
    uint32_t acc_lo = 0u, acc_hi = 0u;      // shots 0..15 / 16..31: low half flips, high half misses
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const uint32_t e_lo = rd32(w[j] & 0xFFFFu);
        const uint32_t e_hi = rd32(w[j] >> 16);
        acc_lo |= e_lo & (0x00010001u << j);
        acc_hi |= e_hi & (0x00010001u << j);
    }
    fc = prmt(acc_lo, acc_hi, 0x5410u);
    miss = has_miss ? prmt(acc_lo, acc_hi, 0x7632u) : 0u;
}

// Same walk, but fetching the n-bit correction of every shot and re-slicing it into planes.
template <int MB, class CorrRead>
QCSS_HD void lut_corrections(const uint32_t (&keys_t)[MB], CorrRead rd, uint32_t (&planes)[32]) {
    constexpr int kBlocks = 32 / MB;
    constexpr uint32_t kKeyMask = (1u << MB) - 1u;
    // planes[] first holds one correction mask per shot, then is transposed 32x32 in place:
    // afterwards planes[q] bit s = bit q of shot s's mask.
#pragma unroll
    for (int j = 0; j < MB; ++j)
#pragma unroll
        for (int b = 0; b < kBlocks; ++b)
            planes[b * MB + j] = rd((keys_t[j] >> (b * MB)) & kKeyMask);
    transpose_blocks<32>(planes);
}

}  // namespace qcss
