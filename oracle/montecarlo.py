"""
Oracle (test infrastructure only): the composed Monte-Carlo path.

The reference has no Monte-Carlo driver; this composes its primitives shot by shot
(``decode_literal``) and in a batched numpy form that gives identical bits
(``decode_batch``).  Parity unpinned by the reference itself: pinned here by exhaustive
failure-weight enumerators (``failure_enumerator``).

Conventions shared with the CUDA path
-------------------------------------
* X errors are decoded with parity_check_c2 / _c2_syndromes / Lz ("which" = 2), Z errors with
  parity_check_c1 / _c1_syndromes / Lx ("which" = 1)  (css_code.py:461-470, 641-646).
* Table key = vec_to_int(H.e mod 2), big-endian (bin_matrix.py:36-43).
* A syndrome absent from the table is a *miss*: no correction applied (css_code.py:652-656).
* Depolarising draw from uniforms r: X if r < p/3, Z if p/3 <= r < 2p/3,
  Y if 2p/3 <= r < p;  e_x = X|Y, e_z = Z|Y.
"""

import numpy as np

from . import css as ocss


def syndromes_batch(h, errs):
    """(B, n) 0/1 -> (B, m) 0/1; batched form of np.mod(np.matmul(H, e), 2) (css_code.py:728)."""
    return (errs.astype(np.int64) @ np.asarray(h, dtype=np.int64).T) % 2


def keys_batch(synd):
    """Big-endian integer keys (bin_matrix.py:36-43), batched.  Valid for m <= 62."""
    m = synd.shape[1]
    weights = (1 << np.arange(m - 1, -1, -1, dtype=np.int64))
    return synd.astype(np.int64) @ weights


def dense_table(table, m, n):
    """Flatten {key -> correction} into (present[2^m] bool, corr[2^m, n] uint8)."""
    present = np.zeros(1 << m, dtype=bool)
    corr = np.zeros((1 << m, n), dtype=np.uint8)
    for key, vec in table.items():
        present[int(key)] = True
        corr[int(key)] = np.asarray(vec, dtype=np.uint8)
    return present, corr


def decode_batch(h, table, logical, errs):
    """Batched A.3 for one Pauli type.  Returns dict(synd, keys, corr, miss, flip)."""
    h = np.asarray(h)
    m, n = h.shape
    synd = syndromes_batch(h, errs)
    keys = keys_batch(synd)
    present, corr_tab = dense_table(table, m, n)
    miss = ~present[keys]
    corr = corr_tab[keys]
    resid = (errs.astype(np.uint8) ^ corr)
    # k = 1 in the reference; with several logical rows (allow_multi_logical) a shot fails when any of them flips
    flip = ((resid.astype(np.int64) @ np.asarray(logical, dtype=np.int64).T) % 2).any(axis=1)
    return dict(synd=synd.astype(np.uint8), keys=keys, corr=corr,
                miss=miss.astype(np.uint8), flip=flip.astype(np.uint8))


def decode_literal(h, table, logical, errs):
    """Per-shot loop through ocss.decode_one -- the reference-literal composition."""
    out = dict(synd=[], keys=[], corr=[], miss=[], flip=[])
    for e in errs:
        s, key, c, miss, flip = ocss.decode_one(h, table, logical, np.asarray(e, dtype='int'))
        out['synd'].append(s); out['keys'].append(key); out['corr'].append(c)
        out['miss'].append(miss); out['flip'].append(flip)
    return {k: np.array(v) for k, v in out.items()}


def tally_xz(code, ex, ez):
    """Tallies of A.3 over a batch: shots, fail_x, fail_z, fail_any, miss_x, miss_z."""
    hx, tx, lz = ocss.pauli_side(code, 2)
    hz, tz, lx = ocss.pauli_side(code, 1)
    dx = decode_batch(hx, tx, lz, ex)
    dz = decode_batch(hz, tz, lx, ez)
    return dict(shots=int(ex.shape[0]),
                fail_x=int(dx['flip'].sum()), fail_z=int(dz['flip'].sum()),
                fail_any=int((dx['flip'] | dz['flip']).sum()),
                miss_x=int(dx['miss'].sum()), miss_z=int(dz['miss'].sum()))


def depolarizing_from_uniform(r, p):
    """SURVEY 8d draw: uniforms r (B, n) -> (e_x, e_z) uint8."""
    is_x = r < p / 3
    is_z = (r >= p / 3) & (r < 2 * p / 3)
    is_y = (r >= 2 * p / 3) & (r < p)
    return (is_x | is_y).astype(np.uint8), (is_z | is_y).astype(np.uint8)


def sample_depolarizing(rng, shots, n, p):
    return depolarizing_from_uniform(rng.random((shots, n)), p)


def all_patterns(n, lo=0, hi=None):
    """Rows lo..hi-1 of the (2^n, n) matrix of all bit patterns; bit j of the row index is
    qubit j (little-endian in the index, which only fixes the enumeration order)."""
    hi = (1 << n) if hi is None else hi
    idx = np.arange(lo, hi, dtype=np.int64)
    return ((idx[:, None] >> np.arange(n, dtype=np.int64)[None, :]) & 1).astype(np.uint8)


def failure_enumerator(h, table, logical, chunk=1 << 20):
    """Counts, by Hamming weight 0..n, of the error patterns (one Pauli type) that end in a
    logical flip and of those that miss the table -- SURVEY Appendix A.4."""
    n = np.asarray(h).shape[1]
    flips = np.zeros(n + 1, dtype=np.int64)
    misses = np.zeros(n + 1, dtype=np.int64)
    for lo in range(0, 1 << n, chunk):
        pats = all_patterns(n, lo, min(1 << n, lo + chunk))
        d = decode_batch(h, table, logical, pats)
        wt = pats.sum(axis=1)
        flips += np.bincount(wt[d['flip'] == 1], minlength=n + 1)
        misses += np.bincount(wt[d['miss'] == 1], minlength=n + 1)
    return flips, misses


def exact_rate(enumerator, q):
    """P(event) when each bit is set independently with probability q."""
    n = len(enumerator) - 1
    w = np.arange(n + 1)
    return float(np.sum(np.asarray(enumerator, dtype=np.float64) * q ** w * (1 - q) ** (n - w)))
