"""
Oracle (test infrastructure only): GF(2) toolkit, restating reference ``bin_matrix.py``.

Every function names the reference lines it follows.  ``rref_literal`` keeps the reference's
exact arithmetic (integer row adds, reduction mod 2 only at the end, dtype preserved);
``rref_packed`` is a fast bit-packed variant used for big cases once it has been checked
against the literal one.  ``rank`` / ``null_space`` / ``solve`` are NOT in the reference
(bin_matrix.py has exactly four functions) -- parity unpinned; semantics documented here.
"""

import itertools

import numpy as np


def rref_literal(mat):
    """Canonical reduced row echelon form over GF(2).  Follows bin_matrix.py:8-34.

    Walks the columns left to right with a pivot-row counter.  The first row at or below the
    counter with an odd entry supplies the pivot; if the counter row itself is even there the
    supplier row is *added* to it (bin_matrix.py:23-24 -- there is no swap), then every other
    row with an odd entry in the column gets the counter row added (bin_matrix.py:27-29).
    Entries are only reduced mod 2 on return (bin_matrix.py:34), so intermediate values grow
    in the input dtype; only parity matters.
    """
    work = np.array(mat, copy=True)
    n_rows, n_cols = work.shape
    lead = 0
    for col in range(n_cols):
        odd_rows = np.flatnonzero(work[lead:, col] % 2 == 1) if lead < n_rows else ()
        if len(odd_rows) == 0:
            continue
        supplier = lead + int(odd_rows[0])
        if work[lead, col] % 2 == 0:
            work[lead, :] += work[supplier, :]
        for other in range(n_rows):
            if other != lead and work[other, col] % 2 == 1:
                work[other, :] += work[lead, :]
        lead += 1
    return np.mod(work, 2)


def pack_rows(mat):
    """(m, n) 0/1 array -> (m, ceil(n/64)) uint64, bit j of word w <-> column 64*w + j."""
    mat = np.asarray(mat)
    m, n = mat.shape
    words = (n + 63) // 64
    padded = np.zeros((m, words * 64), dtype=np.uint8)
    padded[:, :n] = mat & 1
    return np.packbits(padded, axis=1, bitorder='little').view(np.uint64).reshape(m, words)


def unpack_rows(packed, n):
    """Inverse of pack_rows -> (m, n) int64 0/1."""
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    bits = np.unpackbits(packed.view(np.uint8), axis=1, bitorder='little')
    return bits[:, :n].astype(np.int64)


def rref_packed(packed, n):
    """Bit-packed Gauss-Jordan; returns (rref_packed, pivot_columns).

    Same canonical RREF as ``rref_literal`` (RREF is unique, so pivot order is free); rows are
    uint64 words and the elimination is one vectorised XOR per pivot.
    """
    a = np.array(packed, dtype=np.uint64, copy=True)
    m = a.shape[0]
    lead = 0
    pivots = []
    one = np.uint64(1)
    for col in range(n):
        if lead == m:
            break
        w, b = divmod(col, 64)
        colbits = (a[:, w] >> np.uint64(b)) & one
        cand = np.flatnonzero(colbits[lead:])
        if len(cand) == 0:
            continue
        src = lead + int(cand[0])
        if src != lead:
            a[[lead, src]] = a[[src, lead]]
            colbits[[lead, src]] = colbits[[src, lead]]
        hit = colbits.astype(bool)
        hit[lead] = False
        a[hit] ^= a[lead]
        pivots.append(col)
        lead += 1
    return a, np.array(pivots, dtype=np.int32)


def rref_fast(mat):
    """RREF of a 0/1 integer matrix through the packed path; dtype preserved like the reference."""
    mat = np.asarray(mat)
    out, _ = rref_packed(pack_rows(np.mod(mat, 2).astype(np.uint8)), mat.shape[1])
    return unpack_rows(out, mat.shape[1]).astype(mat.dtype)


def vec_to_int(vec):
    """Big-endian bit vector -> integer; vec[0] is the MSB.  Follows bin_matrix.py:36-43."""
    acc = 0
    for bit in np.asarray(vec).ravel():
        acc = (acc << 1) + bit
    return acc


def int_to_vec(value, n):
    """Integer -> big-endian bit vector of length n (dtype 'int'); ValueError('n is too small')
    when bits are left over.  Follows bin_matrix.py:45-55."""
    out = np.zeros(n, dtype='int')
    for pos in range(n - 1, -1, -1):
        out[pos] = value & 1
        value >>= 1
    if value != 0:
        raise ValueError("n is too small")
    return out


def weight_w_vectors(n, w):
    """All length-n 0/1 vectors of Hamming weight w, in the reference's order: lexicographic in
    the sorted support (bin_matrix.py:57-72 recursion sets positions start..n-1 in turn, which
    is exactly itertools.combinations order).  Yields fresh dtype='int' arrays."""
    for support in itertools.combinations(range(n), w):
        v = np.zeros(n, dtype='int')
        v[list(support)] = 1
        yield v


# --- not in the reference (parity unpinned): derived from the canonical RREF -------------

def rank(mat):
    """Number of non-zero rows of the RREF."""
    r = rref_fast(np.asarray(mat))
    return int(np.count_nonzero(r.any(axis=1)))


def pivot_columns(mat):
    mat = np.asarray(mat)
    _, piv = rref_packed(pack_rows(np.mod(mat, 2).astype(np.uint8)), mat.shape[1])
    return piv


def null_space(mat):
    """Basis of {x : mat @ x = 0 mod 2}, shape (n - rank, n).

    Convention (documented, the builder's own): one basis vector per free column f, in
    increasing f; vector has x[f] = 1, x[pivot_i] = RREF[i, f], zero elsewhere.
    """
    mat = np.asarray(mat)
    m, n = mat.shape
    r, piv = rref_packed(pack_rows(np.mod(mat, 2).astype(np.uint8)), n)
    rr = unpack_rows(r, n)
    free = [c for c in range(n) if c not in set(piv.tolist())]
    basis = np.zeros((len(free), n), dtype=np.int64)
    for k, f in enumerate(free):
        basis[k, f] = 1
        for i, p in enumerate(piv):
            basis[k, p] = rr[i, f]
    return basis


def solve(mat, rhs):
    """One solution x of mat @ x = rhs (mod 2) with all free variables 0, or None if
    inconsistent.  Convention: RREF of the augmented matrix [mat | rhs]; x[pivot_i] = rhs'_i."""
    mat = np.asarray(mat)
    rhs = np.asarray(rhs).reshape(-1, 1)
    m, n = mat.shape
    aug = np.concatenate([np.mod(mat, 2), np.mod(rhs, 2)], axis=1).astype(np.uint8)
    r, piv = rref_packed(pack_rows(aug), n + 1)
    if len(piv) and piv[-1] == n:
        return None
    rr = unpack_rows(r, n + 1)
    x = np.zeros(n, dtype=np.int64)
    for i, p in enumerate(piv):
        x[p] = rr[i, n]
    return x
