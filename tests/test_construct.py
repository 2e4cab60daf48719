"""CSS construction numerics on the device -- the standard form with the reference's
column-swap pivot rule (css_code.normalize_parity_check, css_code.py:809-836) and the two-matrix sequence
of CSSCode.__init__ (css_code.py:47-61) -- against outputs of the UNMODIFIED reference
(tests/golden/normalize_golden.npz, made by oracle/gen_normalize_golden.py) and against the oracle on
larger random matrices.  The CPU part pins the oracle and the host implementation to the same vectors."""

import os

import numpy as np
import pytest

import css_code
from css_code import CSSCode, InvalidCodeError
from oracle import css as ocss
from quantum_css_codes_b200 import _native, codes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def norm_golden():
    with np.load(os.path.join(HERE, "golden", "normalize_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def golden_cases(g):
    for i in range(int(g["norm_count"])):
        yield (g[f"n{i}_in"].astype('int'), int(g[f"n{i}_offset"]), g[f"n{i}_out"].astype('int'),
               [tuple(int(v) for v in pair) for pair in g[f"n{i}_swaps"]])


def random_full_rank(rng, r, n, offset, dead=3):
    """Random r x n matrix of rank r with dead / repeated pivot columns so that qubit swaps fire."""
    while True:
        a = rng.integers(0, 2, size=(r, n), dtype=np.int64)
        for _ in range(dead):
            a[:, offset + rng.integers(0, r)] = 0
        try:
            ocss.normalize_parity_check(a.copy(), offset)
        except ocss.OracleInvalidCode:
            continue
        return a


# ---- CPU: oracle and host implementation against the reference's outputs ------------------------------

def test_oracle_and_host_normalize_golden(norm_golden):
    for a, offset, want, swaps in golden_cases(norm_golden):
        got, got_swaps = ocss.normalize_parity_check(a.copy(), offset)
        assert np.array_equal(got, want) and [tuple(s) for s in got_swaps] == swaps
        got, got_swaps = css_code.normalize_parity_check(a.copy(), offset)
        assert np.array_equal(got, want) and [tuple(s) for s in got_swaps] == swaps


def test_oracle_and_host_pairs_golden(norm_golden):
    for i in range(int(norm_golden["pair_count"])):
        p1, p2 = norm_golden[f"p{i}_in1"].astype('int'), norm_golden[f"p{i}_in2"].astype('int')
        ref = ocss.build_css(p1.copy(), p2.copy())
        assert np.array_equal(ref.parity_check_c1, norm_golden[f"p{i}_h1"])
        assert np.array_equal(ref.parity_check_c2, norm_golden[f"p{i}_h2"])
        code = CSSCode(p1.copy(), p2.copy())
        assert np.array_equal(code.parity_check_c1, norm_golden[f"p{i}_h1"])
        assert np.array_equal(code.parity_check_c2, norm_golden[f"p{i}_h2"])


def test_standard_form_argument():
    h = np.array(codes.hamming_7_4())
    with pytest.raises(ValueError, match="standard_form must be"):
        CSSCode(h, h, standard_form="tpu")


# ---- GPU: the kernels ------------------------------------------------------------------------------

@pytest.mark.gpu
def test_normalize_gpu_golden(norm_golden):
    for a, offset, want, swaps in golden_cases(norm_golden):
        work = a.copy()
        got, got_swaps = css_code.normalize_parity_check_gpu(work, offset)
        assert np.array_equal(got, want)
        assert got_swaps == swaps
        assert np.array_equal(work, want)                    # in place, reduced


@pytest.mark.gpu
@pytest.mark.parametrize("r,n,offset", [(1, 1, 0), (2, 130, 64), (24, 60, 0), (24, 60, 36), (100, 300, 17),
                                         (257, 640, 300), (768, 1600, 0), (768, 1600, 768)])
def test_normalize_gpu_random_vs_oracle(r, n, offset):
    """(768, 1600) rows fit shared memory (150 KB); see the next test for the global-memory path."""
    rng = np.random.default_rng(r * 1000 + n + offset)
    if n > offset + r:
        a = random_full_rank(rng, r, n, offset)
    else:                                                    # no room for swaps: any invertible trailing block
        a = rng.integers(0, 2, size=(r, n), dtype=np.int64)
        a[:, offset:] = np.triu(rng.integers(0, 2, size=(r, r)), 1) + np.eye(r, dtype=np.int64)
    want, want_swaps = ocss.normalize_parity_check(a.copy(), offset)
    got, got_swaps = css_code.normalize_parity_check_gpu(a.copy(), offset)
    assert np.array_equal(got, want)
    assert got_swaps == [tuple(int(v) for v in s) for s in want_swaps]


@pytest.mark.gpu
def test_normalize_gpu_global_memory_path():
    """1024 x 2048 (256 KB of rows) does not fit shared memory: the CTA works in place in global memory."""
    rng = np.random.default_rng(1024)
    a = random_full_rank(rng, 1024, 2048, 512, dead=5)
    want, want_swaps = ocss.normalize_parity_check(a.copy(), 512)
    got, got_swaps = css_code.normalize_parity_check_gpu(a.copy(), 512)
    assert np.array_equal(got, want)
    assert got_swaps == [tuple(int(v) for v in s) for s in want_swaps] and len(got_swaps) >= 5


@pytest.mark.gpu
def test_normalize_gpu_batched_mixed_status():
    """One launch, one CTA per matrix; dependent matrices are flagged, the others unaffected."""
    rng = np.random.default_rng(77)
    r, n, offset = 20, 70, 5
    mats, want = [], []
    for i in range(64):
        if i % 5 == 4:
            a = rng.integers(0, 2, size=(r, n), dtype=np.int64)
            a[r - 1] = a[0] ^ a[1]                           # dependent rows
            want.append(None)
        else:
            a = random_full_rank(rng, r, n, offset)
            want.append(ocss.normalize_parity_check(a.copy(), offset))
        mats.append(a)
    packed = _native.pack_bits(np.array(mats, dtype=np.uint8))
    out, swaps, status = _native.gf2_normalize_packed(packed, n, offset)
    bits = _native.unpack_bits(out, n)
    for i, w in enumerate(want):
        if w is None:
            assert status[i] == 2
        else:
            assert status[i] == 0
            assert np.array_equal(bits[i], w[0])
            assert swaps[i] == [tuple(int(v) for v in s) for s in w[1]]


@pytest.mark.gpu
def test_normalize_gpu_errors():
    with pytest.raises(ValueError, match="not enough columns"):
        css_code.normalize_parity_check_gpu(np.ones((3, 2), dtype='int'), 0)
    with pytest.raises(InvalidCodeError, match="rows are not independent"):
        css_code.normalize_parity_check_gpu(np.array([[1, 1, 0], [1, 1, 0]]), 0)
    with pytest.raises(ValueError, match="not enough columns"):
        _native.gf2_normalize_packed(np.zeros((1, 3, 1), dtype=np.uint64), 4, 2)


@pytest.mark.gpu
def test_standard_form_gpu_pairs_golden(norm_golden):
    for i in range(int(norm_golden["pair_count"])):
        p1, p2 = norm_golden[f"p{i}_in1"].astype('int'), norm_golden[f"p{i}_in2"].astype('int')
        code = CSSCode(p1.copy(), p2.copy(), standard_form="gpu")
        assert np.array_equal(code.parity_check_c1, norm_golden[f"p{i}_h1"])
        assert np.array_equal(code.parity_check_c2, norm_golden[f"p{i}_h2"])
        assert code.parity_check_c1.dtype == np.dtype('int')


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23", "shor9"])
def test_standard_form_gpu_equals_host(name):
    h1, h2 = [np.array(h) for h in getattr(codes, name)()]
    host = CSSCode(h1.copy(), h2.copy())
    dev = CSSCode(h1.copy(), h2.copy(), standard_form="gpu", table_builder="gpu")
    assert np.array_equal(host.parity_check_c1, dev.parity_check_c1)
    assert np.array_equal(host.parity_check_c2, dev.parity_check_c2)
    assert (host.n, host.k, host.t) == (dev.n, dev.k, dev.t)
    assert list(host._c2_syndromes) == list(dev._c2_syndromes)
    assert np.array_equal(host.z_operator_matrix(), dev.z_operator_matrix())
    errs = np.eye(host.n, dtype=np.uint8)
    assert np.array_equal(dev.decode(errs, 2)["correction"], errs)


@pytest.mark.gpu
def test_standard_form_gpu_surface_codes():
    for d in (3, 5):
        hx, hz = [np.array(h) for h in codes.rotated_surface(d)]
        host = CSSCode(hx.copy(), hz.copy())
        dev = CSSCode(hx.copy(), hz.copy(), standard_form="gpu")
        assert np.array_equal(host.parity_check_c1, dev.parity_check_c1)
        assert np.array_equal(host.parity_check_c2, dev.parity_check_c2)


@pytest.mark.gpu
def test_standard_form_gpu_errors_in_reference_order():
    """Same exceptions, and the same one when several apply (css_code.py:47-61 order)."""
    h = np.array(codes.hamming_7_4())
    bad = h.copy(); bad[0, 0] = 1
    with pytest.raises(ValueError, match="dual code must be a subspace"):
        CSSCode(h, bad, standard_form="gpu")
    dep = np.vstack([h, h[0] ^ h[1]])                        # CSS holds, rows of C_1 dependent
    with pytest.raises(InvalidCodeError, match="rows are not independent"):
        CSSCode(dep, h, standard_form="gpu")
    with pytest.raises(InvalidCodeError, match="rows are not independent"):
        CSSCode(h, dep, standard_form="gpu")
    # n < r1 + r2: "not enough columns" from the second normalisation (css_code.py:811-812 via :58)
    a = np.array([[1, 1, 0, 0], [0, 0, 1, 1]]); b = np.array([[1, 1, 1, 1], [1, 1, 0, 0], [0, 0, 1, 1]])
    for kwargs in ({}, {"standard_form": "gpu"}):
        with pytest.raises((ValueError, InvalidCodeError)) as host_exc:
            CSSCode(a.copy(), b.copy(), **kwargs)
        if not kwargs:
            expected = (type(host_exc.value), str(host_exc.value))
        else:
            assert (type(host_exc.value), str(host_exc.value)) == expected
    # k != 1 is still rejected after construction (css_code.py:74-75)
    hx, hz = [np.array(m) for m in codes.rotated_surface(3)]
    with pytest.raises(InvalidCodeError, match="single logical qubit"):
        CSSCode(hx[:-1], hz, standard_form="gpu")


@pytest.mark.gpu
def test_css_condition_large():
    """CSS condition kernel on the hypergraph-product pair (768 x 1600 each): holds; one flipped bit breaks it."""
    hx, hz = codes.hgp1600()
    hx, hz = np.array(hx, dtype=np.uint8), np.array(hz, dtype=np.uint8)
    status, _, _, _ = _native.css_standard_form_bits(hx, hz)
    assert status != _native.FORM_NOT_CSS                    # dependent rows (status 2 / 3) are expected here
    assert status in (_native.FORM_DEPENDENT_ROWS_C1, _native.FORM_DEPENDENT_ROWS_C2)
    hz2 = hz.copy(); hz2[700, 1599] ^= 1
    if np.any((hx.astype(np.int64) @ hz2.T.astype(np.int64)) % 2):
        status, _, _, _ = _native.css_standard_form_bits(hx, hz2)
        assert status == _native.FORM_NOT_CSS
