"""
Oracle (test infrastructure only): numeric core of reference ``css_code.py``.

Restates, with the reference lines each step follows:
  CSSCode.__init__ numerics          css_code.py:32-75
  z_operator_matrix / x_operator_matrix   css_code.py:124-136 / 149-161
  decode semantics of quil_classical_correct   css_code.py:649-685
  syndrome_table                     css_code.py:715-735
  swap_columns                       css_code.py:783-785
  normalize_parity_check             css_code.py:809-836
  codes_equal / is_doubly_even       css_code.py:838-850
Exceptions are plain ValueError / OracleInvalidCode with the reference's messages so the
parity tests can compare error behaviour as well.
"""

from dataclasses import dataclass, field

import numpy as np

from . import gf2


class OracleInvalidCode(Exception):
    """Stands in for errors.InvalidCodeError (errors.py:5)."""


def swap_columns(mat, pair):
    """In-place exchange of two columns (css_code.py:783-785)."""
    a, b = pair
    col_a = mat[:, a].copy()
    mat[:, a] = mat[:, b]
    mat[:, b] = col_a


def normalize_parity_check(h, offset):
    """Standard form with an identity block starting at column ``offset`` (css_code.py:809-836).

    For row i the target column is i+offset.  If some row j >= i is odd there, row i is made odd
    by *adding* that row when needed (:816-821).  Otherwise the first odd entry of row i at or
    right of the target column names a qubit swap, recorded and applied (:823-829); no such entry
    -> "rows are not independent" (:825-826).  Then the target column is cleared from every
    other row by adding row i (:832-834).  ``h`` is mutated in place with un-reduced integers;
    the return value is (h mod 2, swaps) (:836).
    """
    rows, cols = h.shape
    if cols < offset + rows:
        raise ValueError("not enough columns")
    swaps = []
    for i in range(rows):
        target = i + offset
        below = [j for j in range(i, rows) if h[j, target] % 2 == 1]
        if below:
            if h[i, target] % 2 == 0:
                h[i, :] += h[below[0], :]
        else:
            right = [j for j in range(target, cols) if h[i, j] % 2 == 1]
            if not right:
                raise OracleInvalidCode("rows are not independent")
            swaps.append((target, right[0]))
            swap_columns(h, swaps[-1])
        for j in range(rows):
            if j != i and h[j, target] % 2 == 1:
                h[j, :] += h[i, :]
    return np.mod(h, 2), swaps


def syndrome_table(parity_check):
    """Unique-decoding radius t and {key -> min-weight error} (css_code.py:715-735).

    Layers of weight w = 0, 1, ... are enumerated in ``weight_w_vectors`` order; the first
    syndrome that repeats (against earlier layers or within the layer) ends the search with
    (w - 1, table-of-complete-layers) (:730-731).  Keys are ``vec_to_int`` of H.e mod 2 (:728-729).
    """
    n = parity_check.shape[1]
    table = {}
    for w in range(n + 1):
        layer = {}
        for e in gf2.weight_w_vectors(n, w):
            key = gf2.vec_to_int(np.mod(np.matmul(parity_check, e), 2))
            if key in table or key in layer:
                return w - 1, table
            layer[key] = e
        table = {**table, **layer}
    return n, table


def codes_equal(h_a, h_b):
    """Same row space <=> same RREF (css_code.py:838-844)."""
    if h_a.shape != h_b.shape:
        return False
    return bool(np.array_equal(gf2.rref_literal(h_a), gf2.rref_literal(h_b)))


def is_doubly_even(mat):
    """Every row weight divisible by 4 (css_code.py:846-850)."""
    return not np.any(np.mod(np.sum(mat, axis=1), 4))


@dataclass
class OracleCSS:
    n: int
    k: int
    t: int
    r_1: int
    r_2: int
    parity_check_c1: np.ndarray
    parity_check_c2: np.ndarray
    c1_syndromes: dict
    c2_syndromes: dict
    transversal_gates: list
    lz: np.ndarray = field(default=None)
    lx: np.ndarray = field(default=None)


def z_operator_matrix(h1, n, r_1, r_2, k):
    """[A2^T 0 I] (css_code.py:124-136)."""
    out = np.zeros((k, n), dtype='int')
    out[:, 0:r_1] = np.transpose(h1[:, (r_1 + r_2):n])
    out[:, (r_1 + r_2):n] = np.identity(k)
    return out


def x_operator_matrix(h2, n, r_1, r_2, k):
    """[0 E^T I] (css_code.py:149-161)."""
    out = np.zeros((k, n), dtype='int')
    out[:, r_1:(r_1 + r_2)] = np.transpose(h2[:, (r_1 + r_2):n])
    out[:, (r_1 + r_2):n] = np.identity(k)
    return out


def build_css(parity_check_c1, parity_check_c2, allow_k_not_1=False):
    """Everything CSSCode.__init__ computes (css_code.py:32-75), in the reference's order of
    evaluation so the same exception fires first."""
    r_1, n_1 = parity_check_c1.shape
    r_2, n_2 = parity_check_c2.shape
    if n_1 != n_2:
        raise ValueError("C_1 and C_2 must have the same code word length")
    h_1 = np.mod(np.array(parity_check_c1, dtype='int'), 2)
    h_2 = np.mod(np.array(parity_check_c2, dtype='int'), 2)
    if not np.array_equal(h_1, parity_check_c1):
        raise ValueError("C_1 parity check matrix must be binary")
    if not np.array_equal(h_2, parity_check_c2):
        raise ValueError("C_2 parity check matrix must be binary")
    if np.any(np.mod(np.matmul(h_1, np.transpose(h_2)), 2)):
        raise ValueError("C_2 dual code must be a subspace of C_1")

    h_1, swaps = normalize_parity_check(h_1, offset=0)
    for pair in swaps:
        swap_columns(h_2, pair)
    h_2, swaps = normalize_parity_check(h_2, offset=r_1)
    for pair in swaps:
        swap_columns(h_1, pair)

    n = n_1
    k = n_1 - r_1 - r_2
    t_1, c1 = syndrome_table(h_1)
    t_2, c2 = syndrome_table(h_2)
    gates = ['I', 'CNOT']
    if codes_equal(h_1, h_2):
        gates += ['H', 'CZ']
        if is_doubly_even(h_1):
            gates.append('S')
    if k != 1 and not allow_k_not_1:
        raise OracleInvalidCode("currently only supports CSS codes for a single logical qubit")
    code = OracleCSS(n=n, k=k, t=min(t_1, t_2), r_1=r_1, r_2=r_2,
                     parity_check_c1=h_1, parity_check_c2=h_2,
                     c1_syndromes=c1, c2_syndromes=c2, transversal_gates=gates)
    if k >= 0:
        code.lz = z_operator_matrix(h_1, n, r_1, r_2, k)
        code.lx = x_operator_matrix(h_2, n, r_1, r_2, k)
    return code


def pauli_side(code, which):
    """(H, table, L) for one Pauli type, reference naming (css_code.py:461-470, 641-646):
    which=2 -> X errors: parity_check_c2, _c2_syndromes, Lz;
    which=1 -> Z errors: parity_check_c1, _c1_syndromes, Lx."""
    if which == 2:
        return code.parity_check_c2, code.c2_syndromes, code.lz
    if which == 1:
        return code.parity_check_c1, code.c1_syndromes, code.lx
    raise ValueError("which must be 1 or 2")


def decode_one(h, table, logical, e):
    """One shot, literally as the reference does it:
    s = H.e mod 2 (css_code.py:728); key = vec_to_int(s) (bin_matrix.py:36-43);
    correction = table.get(key), a miss leaves the frame unchanged (css_code.py:652-656,677-682);
    flip = L.r mod 2 on the residual (css_code.py:641-646)."""
    s = np.mod(np.matmul(h, e), 2)
    key = gf2.vec_to_int(s)
    c = table.get(key)
    miss = c is None
    r = e if miss else (e + c) % 2
    flip = int(np.mod(np.matmul(logical, r), 2)[0])
    corr = np.zeros_like(e) if miss else c
    return s, key, corr, int(miss), flip
