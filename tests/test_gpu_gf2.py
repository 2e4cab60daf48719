"""GPU parity for K4 (batched GF(2) Gauss-Jordan) through the C ABI, against the oracle and the
reference's captured outputs."""

import os

import numpy as np
import pytest

import bin_matrix
import css_code
from oracle import gf2 as ogf2
from quantum_css_codes_b200 import codes, _native

pytestmark = pytest.mark.gpu


def test_reference_kat():
    """test/test_bin_matrix.py:8-20."""
    mat = np.array([[1, 0, 1, 1, 0, 1, 0], [0, 1, 1, 0, 0, 1, 1], [1, 0, 1, 0, 1, 0, 1]], dtype='int')
    want = np.array([[1, 0, 1, 0, 1, 0, 1], [0, 1, 1, 0, 0, 1, 1], [0, 0, 0, 1, 1, 1, 1]], dtype='int')
    got = bin_matrix.reduced_row_echelon_form(mat)
    assert np.array_equal(got, want) and got.dtype == mat.dtype
    assert np.array_equal(mat[0], [1, 0, 1, 1, 0, 1, 0])            # input untouched


def test_rref_golden(golden):
    for i in range(int(golden["rref_count"])):
        got = bin_matrix.reduced_row_echelon_form(golden[f"rref_in_{i}"])
        assert np.array_equal(got, golden[f"rref_out_{i}"]), i
    for tag in ("wide", "u8"):
        got = bin_matrix.reduced_row_echelon_form(golden[f"rref_in_{tag}"])
        assert got.dtype == golden[f"rref_out_{tag}"].dtype
        assert np.array_equal(got, golden[f"rref_out_{tag}"])


@pytest.mark.parametrize("m,n", [(1, 1), (1, 200), (200, 1), (64, 64), (65, 129), (300, 100), (100, 300),
                                 (512, 1024), (777, 1500)])
def test_rref_random_shapes(m, n):
    rng = np.random.default_rng(m * 1000 + n)
    batch = 3
    mats = rng.integers(0, 2, size=(batch, m, n), dtype=np.int64)
    if m > 4:
        mats[1, m - 1] = mats[1, 0] ^ mats[1, 2]                     # dependent row
        mats[2, :, n // 3] = 0                                       # zero column
        mats[2, m // 2] = 0                                          # zero row
    out, rank, piv = bin_matrix.rref_batched(mats)
    for b in range(batch):
        packed, pv = ogf2.rref_packed(ogf2.pack_rows(mats[b].astype(np.uint8)), n)
        assert np.array_equal(out[b], ogf2.unpack_rows(packed, n)), b
        assert rank[b] == len(pv)
        assert np.array_equal(piv[b, : rank[b]], pv) and np.all(piv[b, rank[b]:] == -1)


@pytest.mark.parametrize("m,n", [(1024, 2048), (1000, 3000), (768, 1600), (1024, 100), (33, 4000)])
def test_rref_structured(m, n):
    """Sparse, identity-first and low-rank matrices: strips with few or no pivots, many slabs."""
    rng = np.random.default_rng(m + n)
    sparse = (rng.random((m, n)) < 0.004).astype(np.uint8)
    ident = np.zeros((m, n), dtype=np.uint8)
    ident[np.arange(min(m, n)), np.arange(min(m, n))] = 1
    ident[:, min(m, n) - 1:] ^= (rng.random((m, n - min(m, n) + 1)) < 0.5).astype(np.uint8)
    basis = rng.integers(0, 2, size=(7, n), dtype=np.uint8)
    lowrank = ((rng.integers(0, 2, size=(m, 7), dtype=np.int64) @ basis.astype(np.int64)) % 2).astype(np.uint8)
    mats = np.stack([sparse, ident, lowrank, np.zeros((m, n), dtype=np.uint8)])
    out, rank, piv = bin_matrix.rref_batched(mats)
    for b in range(len(mats)):
        want, pv = ogf2.rref_packed(ogf2.pack_rows(mats[b]), n)
        assert np.array_equal(out[b], ogf2.unpack_rows(want, n)), b
        assert rank[b] == len(pv) and np.array_equal(piv[b, : rank[b]], pv)
    assert rank[2] <= 7 and rank[3] == 0


@pytest.mark.parametrize("knob", [1, 2, 3, 4])
def test_rref_every_kernel_generation(knob):
    """The dispatcher picks a kernel by shape; option "gf2_kernel" forces each implementation (1 the general
    column-by-column kernel, 2 gf2_m4r, 3 gf2_m4r2, 4 gf2_m4r4) over small, ragged, rank-deficient and full-size shapes."""
    with _native.option("gf2_kernel", knob):
        _rref_every_shape(knob)


def _rref_every_shape(knob):
    rng = np.random.default_rng(11 + knob)
    for m, n in [(1, 1), (7, 70), (33, 31), (64, 200), (130, 1100), (300, 100), (640, 640), (1024, 1100)]:
        mats = rng.integers(0, 2, size=(3, m, n), dtype=np.int64)
        if m > 4:
            mats[1, m - 1] = mats[1, 0] ^ mats[1, 2]
            mats[2, :, n // 2] = 0
            mats[2] *= (rng.random((m, n)) < 0.05)                   # sparse: strips with few pivots
        out, rank, piv = bin_matrix.rref_batched(mats)
        for b in range(3):
            packed, pv = ogf2.rref_packed(ogf2.pack_rows(mats[b].astype(np.uint8)), n)
            assert np.array_equal(out[b], ogf2.unpack_rows(packed, n)), (knob, m, n, b)
            assert rank[b] == len(pv) and np.array_equal(piv[b, : rank[b]], pv)


def test_rref_more_than_1024_rows_uses_general_kernel():
    rng = np.random.default_rng(77)
    mat = rng.integers(0, 2, size=(1100, 300), dtype=np.int64)
    out, rank, piv = bin_matrix.rref_batched(mat[None])
    want, pv = ogf2.rref_packed(ogf2.pack_rows(mat.astype(np.uint8)), 300)
    assert np.array_equal(out[0], ogf2.unpack_rows(want, 300)) and rank[0] == len(pv)


def test_rref_c5_full_size():
    """BASELINE config 5 shape: random 1024 x 2048 matrices, plus rank-deficient variants."""
    packed = codes.random_matrices_c5(4).copy()
    packed[1, 1000] = packed[1, 3] ^ packed[1, 5]
    packed[2, :, 0] &= ~np.uint64(0xFFFF)                            # 16 zero columns
    out, rank, piv = bin_matrix.rref_packed_batched(packed, 2048)
    for b in range(4):
        want, pv = ogf2.rref_packed(packed[b], 2048)
        assert np.array_equal(out[b], want), b
        assert rank[b] == len(pv) and np.array_equal(piv[b, : rank[b]], pv)
    assert rank[0] == 1024 and rank[1] == 1023


def test_rref_c5_sixteen_full_size_matrices_vs_reference_goldens():
    """>= 16 full-size 1024 x 2048 matrices bit-exact against bin_matrix.reduced_row_echelon_form ITSELF --
    tests/golden/c5_rref_golden.npz holds SHA-256 + rank of the UNMODIFIED reference's output for sixteen
    default_rng(5) matrices (full rank, dependent rows, zero columns, zero leading columns; oracle/gen_c5_golden.py)
    and two outputs in full.  Every kernel that takes this shape is run."""
    import hashlib
    from oracle.gen_c5_golden import COUNT, variant
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c5_rref_golden.npz")
    with np.load(path, allow_pickle=False) as z:
        sha, rank_want, full = z["sha256"], z["rank"], [z["rref_0"], z["rref_1"]]
    packed = codes.random_matrices_c5(COUNT)
    mats = np.stack([ogf2.pack_rows(variant(i, ogf2.unpack_rows(packed[i], 2048)).astype(np.uint8)) for i in range(COUNT)])
    for knob in (0, 2, 3, 4):
        with _native.option("gf2_kernel", knob):
            out, rank, piv = bin_matrix.rref_packed_batched(mats, 2048)
        assert np.array_equal(out[0], full[0]) and np.array_equal(out[1], full[1]), knob
        assert [hashlib.sha256(out[i].tobytes()).hexdigest() for i in range(COUNT)] == sha.tolist(), knob
        assert rank.tolist() == rank_want.tolist(), knob
        for i in range(COUNT):                                     # pivots: first set column of each non-zero RREF row
            rows = ogf2.unpack_rows(out[i][: rank[i]], 2048)
            assert np.array_equal(piv[i, : rank[i]], rows.argmax(axis=1)), (knob, i)


def test_rank_nullspace_solve():
    rng = np.random.default_rng(8)
    for m, n in [(5, 9), (12, 12), (20, 7), (40, 100), (130, 260)]:
        a = rng.integers(0, 2, size=(m, n), dtype=np.int64)
        if m > 3:
            a[3] = (a[0] + a[1]) % 2
        r = bin_matrix.rank(a)
        assert r == ogf2.rank(a)
        ns = bin_matrix.null_space(a)
        assert np.array_equal(ns, ogf2.null_space(a))
        assert not np.any((a @ ns.T) % 2)
        x0 = rng.integers(0, 2, size=n, dtype=np.int64)
        b = (a @ x0) % 2
        x = bin_matrix.solve(a, b)
        assert np.array_equal(x, ogf2.solve(a, b)) and np.array_equal((a @ x) % 2, b)
    assert bin_matrix.solve(np.array([[1, 1], [1, 1]]), np.array([0, 1])) is None


@pytest.mark.parametrize("m,n", [(1, 1), (3, 70), (70, 3), (64, 128), (100, 333), (333, 100), (640, 1300)])
def test_nullspace_and_solve_batched(m, n):
    """Device null space / solve (qcss_gf2_nullspace, qcss_gf2_solve) against the oracle's RREF-derived
    restatement: same basis vectors in the same order, H.N^T = 0, dim = n - rank, A.x = b."""
    rng = np.random.default_rng(31 * m + n)
    batch = 4
    mats = rng.integers(0, 2, size=(batch, m, n), dtype=np.int64)
    mats[1] *= (rng.random((m, n)) < 0.08)                           # sparse: scattered pivots
    if m > 3:
        mats[2, m - 1] = mats[2, 0] ^ mats[2, 1]
    mats[3, :, : n // 2] = 0                                         # leading zero columns
    basis, rank = bin_matrix.null_space_batched(mats)
    for b in range(batch):
        want = ogf2.null_space(mats[b])
        assert rank[b] == ogf2.rank(mats[b])
        assert np.array_equal(basis[b, : n - rank[b]], want), (m, n, b)
        assert not basis[b, n - rank[b]:].any()
        assert not np.any((mats[b] @ basis[b].T) % 2)
    x0 = rng.integers(0, 2, size=(batch, n), dtype=np.int64)
    rhs = np.einsum('bij,bj->bi', mats, x0) % 2
    rhs_bad = rhs.copy()
    x, ok = bin_matrix.solve_batched(mats, rhs)
    for b in range(batch):
        assert ok[b] and np.array_equal((mats[b] @ x[b]) % 2, rhs[b])
        assert np.array_equal(x[b], ogf2.solve(mats[b], rhs[b]))
    if m > 3:                                                        # row m-1 = row 0 + row 1: flip one side only
        rhs_bad[2, m - 1] ^= 1
        x, ok = bin_matrix.solve_batched(mats, rhs_bad)
        assert not ok[2] and not x[2].any() and ok[0]
        assert ogf2.solve(mats[2], rhs_bad[2]) is None


def test_nullspace_capacity_error_and_c5_shape():
    packed = codes.random_matrices_c5(2).copy()
    packed[1, 7] = packed[1, 3] ^ packed[1, 4]                       # rank 1023 -> 1025 basis vectors
    with pytest.raises(ValueError):
        bin_matrix.null_space_packed_batched(packed, 2048, max_basis_rows=1024)
    basis, rank = bin_matrix.null_space_packed_batched(packed, 2048, max_basis_rows=1025)
    assert rank.tolist() == [1024, 1023]
    for b in range(2):
        bits = ogf2.unpack_rows(basis[b], 2048).astype(np.int64)
        mat = ogf2.unpack_rows(packed[b], 2048).astype(np.int64)
        assert not np.any((mat @ bits.T) % 2)
        assert np.array_equal(bits[: 2048 - rank[b]], ogf2.null_space(mat))


def test_codes_equal_and_transversal_gates(golden):
    """css_code.py:182-201, 838-844 through the GPU RREF."""
    assert css_code.codes_equal(golden["ce_a"], golden["ce_b"]) is True
    assert css_code.codes_equal(golden["ce_a"], golden["ce_c"]) is False
    assert css_code.codes_equal(golden["ce_a"], golden["ce_a"][:, :5]) is False
    for name in ("steane", "qrm15", "golay23"):
        code = css_code.CSSCode(*[np.array(h) for h in getattr(codes, name)()])
        assert sorted(code._transversal_gates) == golden[f"{name}_gates"].tolist()
    steane = css_code.CSSCode(*[np.array(h) for h in codes.steane()])
    for gate in ('I', 'CNOT', 'H', 'CZ', 'S'):                       # test/test_css_code.py:24-26 ('S' is
        assert steane.is_transversal(gate)                           # what css_code.py:199 registers)
    assert not steane.is_transversal('T')


def test_bool_rejected_and_empty():
    with pytest.raises(TypeError):
        bin_matrix.reduced_row_echelon_form(np.zeros((2, 2), dtype=bool))
    assert bin_matrix.reduced_row_echelon_form(np.zeros((0, 5), dtype=int)).shape == (0, 5)
