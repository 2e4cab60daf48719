"""The oracle against (i) the reference's own known-answer tests and (ii) outputs of the
unmodified reference captured by oracle/gen_golden.py.  CPU only."""

import numpy as np
import pytest

from oracle import gf2, css as ocss, montecarlo as omc
from quantum_css_codes_b200 import codes


# ---- reference KATs: test/test_bin_matrix.py:8-31 ---------------------------------------

def test_rref_kat():
    mat = np.array([[1, 0, 1, 1, 0, 1, 0], [0, 1, 1, 0, 0, 1, 1], [1, 0, 1, 0, 1, 0, 1]], dtype='int')
    want = np.array([[1, 0, 1, 0, 1, 0, 1], [0, 1, 1, 0, 0, 1, 1], [0, 0, 0, 1, 1, 1, 1]], dtype='int')
    assert np.array_equal(gf2.rref_literal(mat), want)
    assert np.array_equal(gf2.rref_fast(mat), want)


def test_vec_int_kat():
    assert gf2.vec_to_int(np.array([0, 1, 0, 1, 1])) == 11
    assert np.array_equal(gf2.int_to_vec(11, 5), np.array([0, 1, 0, 1, 1]))
    with pytest.raises(ValueError, match="n is too small"):
        gf2.int_to_vec(11, 3)


# ---- captured reference outputs -----------------------------------------------------------

def test_rref_golden(golden):
    for i in range(int(golden["rref_count"])):
        a, want = golden[f"rref_in_{i}"], golden[f"rref_out_{i}"]
        assert np.array_equal(gf2.rref_literal(a), want)
        assert np.array_equal(gf2.rref_fast(a), want)
    for tag in ("wide", "u8"):
        got = gf2.rref_literal(golden[f"rref_in_{tag}"])
        assert got.dtype == golden[f"rref_out_{tag}"].dtype
        assert np.array_equal(got, golden[f"rref_out_{tag}"])


def test_vec_int_golden(golden):
    for v, k, back in zip(golden["v2i_in"], golden["v2i_out"], golden["i2v_out"]):
        assert gf2.vec_to_int(v) == k
        assert np.array_equal(gf2.int_to_vec(int(k), 40), back)


def test_weight_w_vectors_golden(golden):
    assert np.array_equal(np.array(list(gf2.weight_w_vectors(6, 3))), golden["wwv_6_3"])
    assert np.array_equal(np.array(list(gf2.weight_w_vectors(5, 0))), golden["wwv_5_0"])
    assert np.array_equal(np.array(list(gf2.weight_w_vectors(4, 4))), golden["wwv_4_4"])


@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23"])
def test_build_css_golden(golden, name):
    h1, h2 = getattr(codes, name)()
    assert np.array_equal(h1, golden[f"{name}_in1"]) and np.array_equal(h2, golden[f"{name}_in2"])
    code = ocss.build_css(np.array(h1), np.array(h2))
    assert [code.n, code.k, code.t, code.r_1, code.r_2] == golden[f"{name}_nkt"].tolist()
    assert np.array_equal(code.parity_check_c1, golden[f"{name}_h1"])
    assert np.array_equal(code.parity_check_c2, golden[f"{name}_h2"])
    assert np.array_equal(code.lz, golden[f"{name}_lz"])
    assert np.array_equal(code.lx, golden[f"{name}_lx"])
    for tab, tag in ((code.c1_syndromes, "c1"), (code.c2_syndromes, "c2")):
        keys = np.array([int(k) for k in tab.keys()])
        vals = np.array(list(tab.values()))
        assert np.array_equal(keys, golden[f"{name}_{tag}_keys"])       # insertion order too
        assert np.array_equal(vals, golden[f"{name}_{tag}_vals"])
    assert sorted(code.transversal_gates) == golden[f"{name}_gates"].tolist()


def test_steane_pinned_by_reference_tests():
    """test/test_css_code.py:20-53,108-118: identity blocks, Lz = Z1 Z2 Z6, Lx = X3 X4 X6."""
    code = ocss.build_css(*[np.array(h) for h in codes.steane()])
    assert np.array_equal(code.parity_check_c1[:, 0:3], np.identity(3))
    assert np.array_equal(code.parity_check_c2[:, 3:6], np.identity(3))
    assert code.parity_check_c1.tolist() == [[1, 0, 0, 1, 1, 1, 0], [0, 1, 0, 1, 0, 1, 1], [0, 0, 1, 0, 1, 1, 1]]
    assert code.parity_check_c2.tolist() == [[1, 0, 1, 1, 0, 0, 1], [1, 1, 0, 0, 1, 0, 1], [1, 1, 1, 0, 0, 1, 0]]
    assert code.lz.tolist() == [[0, 1, 1, 0, 0, 0, 1]]
    assert code.lx.tolist() == [[0, 0, 0, 1, 1, 0, 1]]
    t, table = ocss.syndrome_table(code.parity_check_c1)
    assert t == 1 and len(table) == 8
    for s, e in table.items():
        assert s == gf2.vec_to_int(np.mod(np.matmul(code.parity_check_c1, e), 2))
    assert list(table.keys()) == [0, 4, 2, 1, 6, 5, 7, 3]
    assert list(code.c2_syndromes.keys()) == [0, 7, 3, 5, 4, 2, 1, 6]


@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23"])
@pytest.mark.parametrize("which", [1, 2])
def test_decode_golden(golden, name, which):
    code = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    h, table, lop = ocss.pauli_side(code, which)
    pre = f"{name}_w{which}"
    errs = golden[pre + "_errs"]
    lit = omc.decode_literal(h, table, lop, errs[:64])
    bat = omc.decode_batch(h, table, lop, errs)
    for key in ("synd", "keys", "corr", "miss", "flip"):
        assert np.array_equal(bat[key], golden[pre + "_" + key]), key
        assert np.array_equal(lit[key], golden[pre + "_" + key][:64]), key


def test_module_functions_golden(golden):
    h = np.array(codes.hamming_7_4())
    out, swaps = ocss.normalize_parity_check(h, 0)
    assert np.array_equal(out, golden["norm_steane_out"])
    assert np.array_equal(h, golden["norm_steane_mutated"])               # in-place mutation
    assert np.array_equal(np.array(swaps).reshape(-1, 2), golden["norm_steane_swaps"])
    assert bool(golden["doubly_even_true"]) is True and bool(golden["doubly_even_false"]) is False
    assert ocss.codes_equal(golden["ce_a"], golden["ce_b"]) == bool(golden["ce_equal"]) is True
    assert ocss.codes_equal(golden["ce_a"], golden["ce_c"]) == bool(golden["ce_unequal"])
    assert int(golden["golay_table_t"]) == 3


def test_is_doubly_even_kat():
    """test/test_css_code.py:120-143."""
    base = [[0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 1, 1, 0], [1, 1, 1, 0, 0, 0, 0, 1], [1, 1, 1, 1, 1, 1, 1, 1]]
    assert ocss.is_doubly_even(np.array(base))
    bad = [r[:] for r in base]; bad[2][0] = 0
    assert not ocss.is_doubly_even(np.array(bad))
    bad = [r[:] for r in base]; bad[1][0] = 1
    assert not ocss.is_doubly_even(np.array(bad))


def test_constructor_errors():
    h = np.array(codes.hamming_7_4())
    with pytest.raises(ValueError, match="same code word length"):
        ocss.build_css(h, h[:, :6])
    with pytest.raises(ValueError, match="C_1 parity check matrix must be binary"):
        ocss.build_css(h * 2, h)
    with pytest.raises(ValueError, match="C_2 parity check matrix must be binary"):
        ocss.build_css(h, h + 2)
    bad = h.copy(); bad[0, 0] = 1
    with pytest.raises(ValueError, match="dual code must be a subspace"):
        ocss.build_css(h, bad)
    with pytest.raises(ValueError, match="not enough columns"):
        ocss.normalize_parity_check(np.ones((3, 2), dtype='int'), 0)
    with pytest.raises(ocss.OracleInvalidCode, match="rows are not independent"):
        ocss.normalize_parity_check(np.array([[1, 1, 0], [1, 1, 0]]), 0)


def test_hgp_golden(golden):
    hx, hz = codes.hgp1600()
    assert hx.shape == hz.shape == (768, 1600)
    assert np.all(hx.sum(axis=1) == 7) and np.all(hz.sum(axis=1) == 7)
    assert not np.any((hx @ hz.T) % 2)
    assert "rows are not independent" in str(golden["hgp_rejected"]) or "single logical" in str(golden["hgp_rejected"])
    errs = np.unpackbits(golden["hgp_errs"], axis=1, bitorder="little")[:, :1600]
    for h, key in ((hz, "hgp_synd_hz"), (hx, "hgp_synd_hx")):
        want = np.unpackbits(golden[key], axis=1, bitorder="little")[:, :768]
        assert np.array_equal(omc.syndromes_batch(h, errs), want)


# ---- exact enumerators: pins the Monte-Carlo composition ------------------------

A4 = {
    ("steane", 2): [0, 0, 21, 7, 28, 0, 7, 1],
    ("steane", 1): [0, 0, 21, 7, 28, 0, 7, 1],
    ("qrm15", 1): [0, 0, 105, 35, 1260, 168, 4725, 435, 6000, 280, 2835, 105, 420, 0, 15, 1],
    ("qrm15", 2): [0, 0, 0, 0, 965, 1211, 3625, 2055, 4380, 1380, 1792, 400, 455, 105, 15, 1],
}
A4_MISS_QRM_X = [0, 0, 0, 0, 840, 1848, 1960, 2520, 2520, 1960, 1848, 840, 0, 0, 0, 0]


@pytest.mark.parametrize("name,which", list(A4.keys()))
def test_failure_enumerators(name, which):
    code = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    h, table, lop = ocss.pauli_side(code, which)
    flips, misses = omc.failure_enumerator(h, table, lop)
    assert flips.tolist() == A4[(name, which)]
    if (name, which) == ("qrm15", 2):
        assert misses.tolist() == A4_MISS_QRM_X
    else:
        assert misses.sum() == 0


def test_exact_rates_steane():
    q = 2e-3 / 3
    assert omc.exact_rate(A4[("steane", 2)], q) == pytest.approx(9.304338e-06, rel=1e-6)


def test_rank_nullspace_solve_properties():
    rng = np.random.default_rng(3)
    for m, n in [(5, 9), (12, 12), (20, 7), (40, 100)]:
        a = rng.integers(0, 2, size=(m, n), dtype=np.int64)
        if m > 3:
            a[3] = (a[0] + a[1]) % 2
        r = gf2.rank(a)
        ns = gf2.null_space(a)
        assert ns.shape == (n - r, n)
        assert not np.any((a @ ns.T) % 2)
        assert gf2.rank(ns) == n - r if n - r else True
        x0 = rng.integers(0, 2, size=n, dtype=np.int64)
        b = (a @ x0) % 2
        x = gf2.solve(a, b)
        assert x is not None and np.array_equal((a @ x) % 2, b)


def test_oracle_rref_full_size_c5_vs_reference_goldens():
    """The oracle's packed Gauss-Jordan against the UNMODIFIED reference at BASELINE config 5's full size: SHA-256 and
    rank of bin_matrix.reduced_row_echelon_form on sixteen 1024 x 2048 matrices (oracle/gen_c5_golden.py)."""
    import hashlib
    import os
    from oracle.gen_c5_golden import COUNT, variant
    from quantum_css_codes_b200 import codes
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c5_rref_golden.npz")
    with np.load(path, allow_pickle=False) as z:
        sha, rank_want, full0 = z["sha256"], z["rank"], z["rref_0"]
    packed = codes.random_matrices_c5(COUNT)
    for i in range(COUNT):
        mat = gf2.pack_rows(variant(i, gf2.unpack_rows(packed[i], 2048)).astype(np.uint8))
        out, pivots = gf2.rref_packed(mat, 2048)
        if i == 0:
            assert np.array_equal(out, full0)
        assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == sha[i], i
        assert len(pivots) == rank_want[i]
