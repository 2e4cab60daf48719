"""
Drop-in for the reference's ``bin_matrix.py`` (same four functions, same argument meaning and
error behaviour), plus the batched GF(2) toolkit the reference lacks.

``reduced_row_echelon_form`` runs on the GPU (kernel K4, ``qcss_gf2_rref``); there is no host
fallback -- without the CUDA library it raises ``NativeLibraryError``.  The bit <-> integer
helpers and the weight-w enumerator are host-side index bookkeeping, as in the reference.
"""

import itertools

import numpy as np

from . import _native


def reduced_row_echelon_form(mat):
    """New copy of a binary matrix in reduced row echelon form over GF(2).

    Reference: bin_matrix.py:8-34.  Same contract: the input is not modified, the result has
    the input's shape and dtype with entries 0/1, zero rows last (the canonical RREF, so the
    pivot order used on the device cannot change the bits).  Entries are taken mod 2 on the way
    in, as the reference only ever tests parity (bin_matrix.py:18,23,28).  Boolean arrays are
    rejected: the reference silently computes garbage for them (``+=`` on bool is OR).
    """
    mat = np.asarray(mat)
    if mat.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    if mat.dtype == np.bool_:
        raise TypeError("boolean matrices are not supported; use an integer dtype")
    m, n = mat.shape
    if m == 0 or n == 0:
        return np.mod(np.copy(mat), 2)
    out, _, _ = _native.gf2_rref_bits(np.mod(mat, 2).astype(np.uint8)[None, :, :])
    return out[0].astype(mat.dtype)


def vec_to_int(vec):
    """Big-endian bit vector -> integer, ``vec[0]`` is the most significant bit.

    Reference: bin_matrix.py:36-43 (the result keeps numpy's integer type when the elements are
    numpy integers, and wraps silently at 64 bits exactly like the reference)."""
    vec = np.asarray(vec)
    value = 0
    for idx in range(vec.size):
        value = (value << 1) + vec.flat[idx]
    return value


def int_to_vec(int_repr, n):
    """Integer -> length-n big-endian bit vector (dtype 'int'); raises
    ``ValueError("n is too small")`` when the value needs more than n bits.
    Reference: bin_matrix.py:45-55."""
    out = np.zeros(n, dtype='int')
    rest = int_repr
    for pos in range(n - 1, -1, -1):
        out[pos] = rest & 1
        rest = rest >> 1
    if rest != 0:
        raise ValueError("n is too small")
    return out


def weight_w_vectors(n, w):
    """Generate every length-n binary vector of Hamming weight w as a fresh dtype='int' array,
    in the reference's order (supports in lexicographic order).  Reference: bin_matrix.py:57-72."""
    for support in itertools.combinations(range(n), w):
        vec = np.zeros(n, dtype='int')
        if support:
            vec[np.fromiter(support, dtype=np.intp, count=w)] = 1
        yield vec


# ---- batched toolkit (not in the reference; semantics documented in DESIGN.md) ---------------

def rref_batched(mats):
    """(batch, m, n) 0/1 integer array -> (rref, rank, pivots): rref same shape/dtype,
    rank (batch,) int32, pivots (batch, min(m,n)) int32 padded with -1."""
    mats = np.asarray(mats)
    if mats.ndim != 3:
        raise ValueError("expected a (batch, m, n) array")
    out, rank, piv = _native.gf2_rref_bits(np.mod(mats, 2).astype(np.uint8))
    return out.astype(mats.dtype), rank, piv


def rref_packed_batched(packed, n):
    """(batch, m, ceil(n/64)) uint64 packed rows (bit j of word w = column 64w+j) ->
    (rref_packed, rank, pivots).  The zero-copy form used by the C5 benchmark."""
    return _native.gf2_rref_packed(packed, n)


def rank(mat):
    mat = np.asarray(mat)
    _, rk, _ = _native.gf2_rref_bits(np.mod(mat, 2).astype(np.uint8)[None, :, :])
    return int(rk[0])


def null_space(mat):
    """Basis of {x : mat.x = 0 mod 2}, shape (n - rank, n): one vector per free column f in
    increasing order, with x[f] = 1 and x[pivot_i] = RREF[i, f] (built on the device,
    ``qcss_gf2_nullspace``)."""
    mat = np.asarray(mat)
    m, n = mat.shape
    basis, rk = null_space_batched(mat[None, :, :])
    return basis[0][: n - int(rk[0])]


def null_space_batched(mats):
    """(batch, m, n) 0/1 integer array -> (basis (batch, n, n) int64, rank (batch,)); matrix b's basis is
    ``basis[b, : n - rank[b]]``, the remaining rows are zero."""
    mats = np.asarray(mats)
    if mats.ndim != 3:
        raise ValueError("expected a (batch, m, n) array")
    n = mats.shape[2]
    packed = _native.pack_bits(np.mod(mats, 2).astype(np.uint8))
    basis, rk = _native.gf2_nullspace_packed(packed, n)
    return _native.unpack_bits(basis, n).astype(np.int64), rk


def null_space_packed_batched(packed, n, max_basis_rows=None):
    """Packed form of ``null_space_batched`` (the C5 benchmark shape): (batch, m, ceil(n/64)) uint64 ->
    (basis (batch, rows, ceil(n/64)) uint64, rank); ``rows`` defaults to n."""
    return _native.gf2_nullspace_packed(packed, n, max_basis_rows)


def solve(mat, rhs):
    """One solution of mat.x = rhs (mod 2) with every free variable 0, or None when the
    system is inconsistent (RREF of the augmented matrix on the device, ``qcss_gf2_solve``)."""
    mat = np.asarray(mat)
    m, n = mat.shape
    x, ok = solve_batched(mat[None, :, :], np.asarray(rhs).reshape(1, m))
    return x[0] if ok[0] else None


def solve_batched(mats, rhs):
    """(batch, m, n) matrices and (batch, m) right-hand sides -> (x (batch, n) int64, ok (batch,) bool)."""
    mats = np.asarray(mats)
    rhs = np.asarray(rhs)
    if mats.ndim != 3 or rhs.shape != mats.shape[:2]:
        raise ValueError("expected (batch, m, n) matrices and (batch, m) right-hand sides")
    n = mats.shape[2]
    packed = _native.pack_bits(np.mod(mats, 2).astype(np.uint8))
    x, ok = _native.gf2_solve_packed(packed, _native.pack_bits(np.mod(rhs, 2).astype(np.uint8)), n)
    return _native.unpack_bits(x, n).astype(np.int64), ok.astype(bool)
