"""
Parity-check constructions for the named benchmark codes.

None of these exist in the reference beyond the Hamming(7,4) matrix in
test/test_css_code.py:13-17; they are input generators for the tests and ``bench.py``.
All return plain numpy 0/1 ``int`` matrices in the form ``CSSCode(parity_check_c1,
parity_check_c2)`` expects.
"""

import numpy as np


def hamming_7_4():
    """Hamming(7,4) parity check: column v (1..7) is the binary expansion of v, MSB in row 0
    (test/test_css_code.py:13-17)."""
    cols = np.arange(1, 8)
    return np.array([(cols >> b) & 1 for b in (2, 1, 0)], dtype='int')


def steane():
    """Steane [[7,1,3]]: CSSCode(h, h) with h = Hamming(7,4)."""
    h = hamming_7_4()
    return h, h.copy()


def qrm15():
    """Quantum Reed-Muller [[15,1,3]].  P = 4x15, column v (1..15) is the binary expansion of v
    with bit b in row b.  HX = P (-> parity_check_c1), HZ = P plus all pairwise row products
    (10x15 -> parity_check_c2)."""
    cols = np.arange(1, 16)
    p = np.array([(cols >> b) & 1 for b in range(4)], dtype='int')
    pairs = [p[i] * p[j] for i in range(4) for j in range(i + 1, 4)]
    return p, np.vstack([p] + pairs).astype('int')


def golay23():
    """Golay [[23,1,7]] from the cyclic [23,12,7] code.  h(x) = (x^23+1)/g(x) with
    g(x) = 1+x^2+x^4+x^5+x^6+x^10+x^11; rows of H are the 11 cyclic shifts of reversed h(x)."""
    h_coeffs = [1, 0, 1, 0, 0, 1, 0, 0, 1, 1, 1, 1, 1]      # low -> high degree
    first = np.zeros(23, dtype='int')
    first[:13] = h_coeffs[::-1]
    h = np.array([np.roll(first, i) for i in range(11)], dtype='int')
    return h, h.copy()


def gallager_ldpc(seed=1600, blocks=8, row_weight=4, layers=3):
    """(layers*blocks) x (blocks*row_weight) Gallager parity check: each layer is a column
    permutation of kron(I_blocks, ones(1,row_weight)); permutations from default_rng(seed)."""
    rng = np.random.default_rng(seed)
    n = blocks * row_weight
    base = np.kron(np.eye(blocks, dtype='int'), np.ones((1, row_weight), dtype='int'))
    return np.vstack([base[:, rng.permutation(n)] for _ in range(layers)]).astype('int')


def hypergraph_product(h):
    """HX = [H (x) I_n | I_m (x) H^T], HZ = [I_n (x) H | H^T (x) I_m] for a classical m x n H."""
    m, n = h.shape
    hx = np.hstack([np.kron(h, np.eye(n, dtype='int')), np.kron(np.eye(m, dtype='int'), h.T)])
    hz = np.hstack([np.kron(np.eye(n, dtype='int'), h), np.kron(h.T, np.eye(m, dtype='int'))])
    return hx.astype('int'), hz.astype('int')


def hgp1600(seed=1600):
    """Hypergraph-product code, n = 32*32 + 24*24 = 1600, HX and HZ 768x1600, row weight 7.
    Syndrome-only: the reference constructor rejects it (k = 104, dependent rows).
    X errors are checked by HZ, Z errors by HX."""
    return hypergraph_product(gallager_ldpc(seed))


def random_matrices_c5(count, m=1024, n=2048, seed=5, offset=0):
    """``count`` uniform random m x n GF(2) matrices as packed little-endian bytes viewed as
    uint64 words: shape (count, m, n//64).  SURVEY 8d config C5 (default_rng(5) bytes).  The
    generator is advanced so that matrices [offset, offset+count) equal that slice of the full
    4096-matrix draw."""
    rng = np.random.default_rng(seed)
    per = m * (n // 8)
    if offset:
        # integers(uint8) consumes the stream bytewise in draw order; skipping by redrawing is
        # exact and cheap enough for the sizes used in tests.
        left = offset * per
        while left > 0:
            step = min(left, 1 << 26)
            rng.integers(0, 256, size=step, dtype=np.uint8)
            left -= step
    raw = rng.integers(0, 256, size=(count, m, n // 8), dtype=np.uint8)
    return raw.view(np.uint64).reshape(count, m, n // 64)


def shor9():
    """Shor [[9,1,3]] as a CSS code: H1 (X-type generators, 2 rows of weight 6), H2 (Z-type pairs, 6 rows)."""
    hx = np.zeros((2, 9), dtype=np.int64)
    hx[0, 0:6] = 1
    hx[1, 3:9] = 1
    hz = np.zeros((6, 9), dtype=np.int64)
    for b in range(3):
        hz[2 * b, [3 * b, 3 * b + 1]] = 1
        hz[2 * b + 1, [3 * b + 1, 3 * b + 2]] = 1
    return hx, hz


def rotated_surface(d):
    """Rotated surface code [[d^2, 1, d]] (d odd): (d^2 - 1)/2 X-type and Z-type checks, weight 4 in the bulk
    and weight 2 on the boundary.  Returns (HX, HZ) for ``CSSCode(HX, HZ)``."""
    if d % 2 == 0 or d < 3:
        raise ValueError("d must be odd and >= 3")
    hx, hz = [], []
    for i in range(-1, d):
        for j in range(-1, d):
            cells = [(a, b) for a in (i, i + 1) for b in (j, j + 1) if 0 <= a < d and 0 <= b < d]
            x_type = (i + j) % 2 == 0
            if len(cells) == 2:
                horizontal_edge = i in (-1, d - 1)              # top / bottom boundary
                if horizontal_edge != x_type:
                    continue
            elif len(cells) != 4:
                continue
            row = np.zeros(d * d, dtype=np.int64)
            for a, b in cells:
                row[a * d + b] = 1
            (hx if x_type else hz).append(row)
    return np.array(hx), np.array(hz)
