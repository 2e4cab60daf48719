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

Gap sampler (core.cuh::sample_site_word_gap), used instead when thr < 2^26 (p < 1/64):
  * table cdf[k] = floor((1 - (1-p)^(k+1)) * 2^32), k = 0..31, (1-p)^(k+1) by repeated multiplication
    in IEEE double (gap_table); a uniform word u gives d = #{k : cdf[k] <= u} clean lanes before the
    next error (d = 32: none left in this word).
  * a draw is a pair (u, tw).  With pos = next lane to decide: pos + d >= 32 ends the word; otherwise
    the error sits at lane pos + d, its type is the first two-bit field of tw (from the low end) that is
    not 00, read as (x, z); if all 16 fields are 00 the draw is discarded (pos unchanged).
  * the first look at site j is 16 bits: half j & 1 (0 = low) of word (j & 7) >> 1 of the block with counter
    (g_lo, g_hi, j >> 3, 0) -- eight sites share that block.  They are the HIGH half h of the site's first
    uniform u0; h > cdf[31] >> 16 means no error among the 32 lanes.  Everything after it comes from the
    site's own blocks, counter (g_lo, g_hi, j, q), q = 1, 2, ...: block 1 gives the first draw's tw = w0, the low
    half of u0 (u0 = h << 16 | w1 & 0xffff) and the second draw (w2, w3); blocks q >= 2 give two draws each,
    (w0, w1) then (w2, w3).
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


def gap_table(p):
    """(cdf[32] uint32, inv) exactly as api.cu::gap_table_from_p computes them."""
    q = 1.0 - float(p)
    acc = 1.0
    cdf = []
    for _ in range(32):
        acc = acc * q
        v = np.floor((1.0 - acc) * 4294967296.0)
        cdf.append(int(min(max(v, 0.0), 4294967295.0)))
    inv = (4294967295 // cdf[0]) if cdf[0] else 0xFFFFFFFF
    return np.array(cdf, dtype=np.uint64), inv


def uses_gap_sampler(p):
    return threshold(p) < (1 << 26)


def _blocks(seed, g, j, q):
    g = np.asarray(g, dtype=np.uint64)
    ctr = np.stack([g & MASK32, g >> np.uint64(32),
                    np.full(g.shape, j, dtype=np.uint64), np.asarray(q, dtype=np.uint64)], axis=-1)
    key = np.empty(g.shape + (2,), dtype=np.uint32)
    key[..., 0] = seed & 0xFFFFFFFF
    key[..., 1] = (seed >> 32) & 0xFFFFFFFF
    return philox4x32_10(ctr.astype(np.uint32), key)


def _sample_words_gap(seed, first_word, n_words, n, p, site0=0):
    cdf, _ = gap_table(p)
    g = np.arange(first_word, first_word + n_words, dtype=np.uint64)
    ex = np.zeros((n, n_words), dtype=np.uint32)
    ez = np.zeros((n, n_words), dtype=np.uint32)
    look16 = (int(cdf[31]) >> 16) + 1
    shared = {}
    for row in range(n):
        j = site0 + row
        if j >> 3 not in shared:
            shared[j >> 3] = _blocks(seed, g, j >> 3, np.zeros(n_words, dtype=np.uint64))
        word = shared[j >> 3][:, (j & 7) >> 1].astype(np.uint64)
        first = (word >> np.uint64(16)) if (j & 1) else (word & np.uint64(0xFFFF))   # high half of the first uniform
        hit = np.flatnonzero(first < look16)                               # words that may hold an error
        for idx in hit:
            buf = _blocks(seed, g[idx:idx + 1], j, np.array([1], dtype=np.uint64))[0]
            u0 = (int(first[idx]) << 16) | (int(buf[1]) & 0xFFFF)
            draws = [(u0, int(buf[0])), (int(buf[2]), int(buf[3]))]
            pos, blk, x, z = 0, 2, 0, 0
            done = False
            while not done:
                for u, tw in draws:
                    d = int(np.count_nonzero(cdf <= u))
                    if pos + d >= 32:
                        done = True
                        break
                    valid = (tw | (tw >> 1)) & 0x55555555
                    if valid == 0:
                        continue
                    b = (valid & -valid).bit_length() - 1
                    lane = pos + d
                    x |= ((tw >> b) & 1) << lane
                    z |= ((tw >> (b + 1)) & 1) << lane
                    pos = lane + 1
                    if pos >= 32:
                        done = True
                        break
                if not done:
                    buf = _blocks(seed, g[idx:idx + 1], j, np.array([blk], dtype=np.uint64))[0]
                    draws = [(int(buf[0]), int(buf[1])), (int(buf[2]), int(buf[3]))]
                    blk += 1
            ex[row, idx], ez[row, idx] = x, z
    return ex, ez


def sample_words(seed, first_word, n_words, n, p, force_bit_serial=False, site0=0):
    """Sampled planes as uint32 words: (ex, ez), each (n, n_words).  Row r is Philox site ``site0 + r``
    (the third counter word); the single-shot sampler uses sites 0..n-1, the error-correction rounds
    (oracle/ec_rounds.py) site 32 * stream + qubit."""
    if uses_gap_sampler(p) and not force_bit_serial:
        return _sample_words_gap(seed, first_word, n_words, n, p, site0)
    thr = threshold(p)
    g = np.arange(first_word, first_word + n_words, dtype=np.uint64)
    ex = np.zeros((n, n_words), dtype=np.uint32)
    ez = np.zeros((n, n_words), dtype=np.uint32)
    full = np.uint32(0xFFFFFFFF)
    for row in range(n):
        j = site0 + row
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
        ex[row], ez[row] = x, z
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
