"""
Oracle (test infrastructure only): the depolarising sampler of the fused Monte-Carlo kernel,
restated in numpy.  The reference has NO sampler -- parity unpinned by the
reference; the CUDA sampler is pinned bit-for-bit against this restatement, this restatement
against the Random123 known-answer vectors for Philox4x32-10, and the resulting rates against
the exact enumerator rates (montecarlo.exact_rate).

Specification (what quantum_css_codes_b200/csrc/core.cuh::sample_site_word implements):
  * word g covers shots 32g .. 32g+31 (bit b of the word is shot 32g+b); qubit j.
  * random words: Philox4x32-10 blocks, key = (seed_lo, seed_hi), counter = (g_lo, g_hi, j, q),
    q = 0, 1, 2, ... ; a block yields words w0..w3.
  * stage 1: for q = 0..7, for k = 0..3: threshold bit b = 31 - (4q + k) of thr = floor(p*2^32);
    with r = w_k:  bit set:  err |= und & ~r, und &= r ;  bit clear: und &= ~r.
    Stop after the first block with und == 0.  (Per lane: error iff its 32-bit uniform < thr.)
  * stage 2, while some error lane is untyped: next block gives attempts (w0, w1) then (w2, w3)
    as (x bits, z bits); a lane accepts the first attempt where x|z = 1.
  * outputs: x plane word (X or Y), z plane word (Z or Y).
"""

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (..., 2) uint32 -> (..., 4) uint32."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK32
        k1 = (k1 + np.uint64(W1)) & MASK32
    return np.stack(c, axis=-1).astype(np.uint32)


def threshold(p):
    return int(min(max(int(np.floor(p * 4294967296.0)), 0), 0xFFFFFFFF))


def _blocks(seed, g, j, q):
    g = np.asarray(g, dtype=np.uint64)
    ctr = np.stack([g & MASK32, g >> np.uint64(32),
                    np.full(g.shape, j, dtype=np.uint64), np.asarray(q, dtype=np.uint64)], axis=-1)
    key = np.empty(g.shape + (2,), dtype=np.uint32)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr.astype(np.uint32), key)


def sample_words(seed, first_word, n_words, n, p):
    """Sampled planes as uint32 words: (ex, ez), each (n, n_words)."""
    thr = threshold(p)
    g = np.arange(first_word, first_word + n_words, dtype=np.uint64)
    ex = np.zeros((n, n_words), dtype=np.uint32)
    ez = np.zeros((n, n_words), dtype=np.uint32)
    full = np.uint32(0xFFFFFFFF)
    for j in range(n):
        und = np.full(n_words, full, dtype=np.uint32)
        err = np.zeros(n_words, dtype=np.uint32)
        blk = np.zeros(n_words, dtype=np.uint64)
        for q in range(8):
            active = und != 0
            if not active.any():
                break
            w = _blocks(seed, g[active], j, blk[active])
            u, e = und[active], err[active]
            for k in range(4):
                b = 31 - (4 * q + k)
                r = w[:, k]
                if (thr >> b) & 1:
                    e |= u & ~r
                    u &= r
                else:
                    u &= ~r
            und[active], err[active] = u, e
            blk[active] += np.uint64(1)
        need = err.copy()
        x = np.zeros(n_words, dtype=np.uint32)
        z = np.zeros(n_words, dtype=np.uint32)
        while True:
            active = need != 0
            if not active.any():
                break
            w = _blocks(seed, g[active], j, blk[active])
            nd, xa, za = need[active], x[active], z[active]
            for a in (0, 2):
                ok = nd & (w[:, a] | w[:, a + 1])
                xa |= ok & w[:, a]
                za |= ok & w[:, a + 1]
                nd &= ~ok
            need[active], x[active], z[active] = nd, xa, za
            blk[active] += np.uint64(1)
        ex[j], ez[j] = x, z
    return ex, ez


def sample_bits(seed, first_shot, shots, n, p):
    """(ex, ez) as (shots, n) uint8; first_shot must be a multiple of 32."""
    assert first_shot % 32 == 0
    n_words = (shots + 31) // 32
    ex, ez = sample_words(seed, first_shot // 32, n_words, n, p)
    def bits(planes):
        b = np.unpackbits(planes.view(np.uint8).reshape(n, -1), axis=1, bitorder="little")
        return np.ascontiguousarray(b[:, :shots].T)
    return bits(ex), bits(ez)
