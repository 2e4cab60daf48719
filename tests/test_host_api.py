"""CPU tests of the host-side drop-in surface (no GPU): constructor numerics against the
reference's captured outputs, error behaviour, bit conventions, and that libqcss.so loads and
exports every symbol include/qcss.h declares."""

import os
import re

import numpy as np
import pytest

import bin_matrix
import css_code
import errors
from css_code import CSSCode
from quantum_css_codes_b200 import codes, planes, _native

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vec_int_reference_kats():
    """test/test_bin_matrix.py:22-31."""
    assert bin_matrix.vec_to_int(np.array([0, 1, 0, 1, 1])) == 11
    assert np.array_equal(bin_matrix.int_to_vec(11, 5), np.array([0, 1, 0, 1, 1]))
    with pytest.raises(ValueError, match="n is too small"):
        bin_matrix.int_to_vec(11, 3)
    assert isinstance(bin_matrix.vec_to_int(np.array([1, 0], dtype=np.int64)), np.int64)


def test_vec_int_golden(golden):
    for v, k, back in zip(golden["v2i_in"], golden["v2i_out"], golden["i2v_out"]):
        assert bin_matrix.vec_to_int(v) == k
        assert np.array_equal(bin_matrix.int_to_vec(int(k), 40), back)
    assert np.array_equal(np.array(list(bin_matrix.weight_w_vectors(6, 3))), golden["wwv_6_3"])
    assert np.array_equal(np.array(list(bin_matrix.weight_w_vectors(5, 0))), golden["wwv_5_0"])
    assert np.array_equal(np.array(list(bin_matrix.weight_w_vectors(4, 4))), golden["wwv_4_4"])


@pytest.mark.parametrize("name", ["steane", "qrm15", "golay23"])
def test_constructor_matches_reference(golden, name):
    h1, h2 = getattr(codes, name)()
    code = CSSCode(np.array(h1), np.array(h2))
    assert [code.n, code.k, code.t, code.r_1, code.r_2] == golden[f"{name}_nkt"].tolist()
    assert np.array_equal(code.parity_check_c1, golden[f"{name}_h1"])
    assert np.array_equal(code.parity_check_c2, golden[f"{name}_h2"])
    assert np.array_equal(code.z_operator_matrix(), golden[f"{name}_lz"])
    assert np.array_equal(code.x_operator_matrix(), golden[f"{name}_lx"])
    for tab, tag in ((code._c1_syndromes, "c1"), (code._c2_syndromes, "c2")):
        assert np.array_equal(np.array([int(k) for k in tab]), golden[f"{name}_{tag}_keys"])
        assert np.array_equal(np.array(list(tab.values())), golden[f"{name}_{tag}_vals"])
        assert all(isinstance(k, np.int64) for k in tab)
        assert all(v.dtype == np.dtype('int') for v in tab.values())


def test_steane_like_reference_tests():
    """test/test_css_code.py:20-22, 28-30, 108-118 (the PauliTerm tests need pyquil)."""
    steane = CSSCode(*[np.array(h) for h in codes.steane()])
    for mat, off in ((steane.parity_check_c1, 0), (steane.parity_check_c2, 3)):
        assert np.array_equal(mat[:, off:off + 3], np.identity(3))
    t, table = css_code.syndrome_table(steane.parity_check_c1)
    assert t == 1 and len(table) == 8
    for s, e in table.items():
        assert s == bin_matrix.vec_to_int(np.mod(np.matmul(steane.parity_check_c1, e), 2))
    assert steane.z_operator_matrix().tolist() == [[0, 1, 1, 0, 0, 0, 1]]      # Z1 Z2 Z6
    assert steane.x_operator_matrix().tolist() == [[0, 0, 0, 1, 1, 0, 1]]      # X3 X4 X6


def test_is_doubly_even_reference_kats():
    base = [[0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 1, 1, 0], [1, 1, 1, 0, 0, 0, 0, 1], [1, 1, 1, 1, 1, 1, 1, 1]]
    assert css_code.is_doubly_even(np.array(base))
    bad = [r[:] for r in base]; bad[2][0] = 0
    assert not css_code.is_doubly_even(np.array(bad))
    bad = [r[:] for r in base]; bad[1][0] = 1
    assert not css_code.is_doubly_even(np.array(bad))


def test_module_functions(golden):
    h = np.array(codes.hamming_7_4())
    out, swaps = css_code.normalize_parity_check(h, 0)
    assert np.array_equal(out, golden["norm_steane_out"])
    assert np.array_equal(h, golden["norm_steane_mutated"])
    assert np.array_equal(np.array(swaps).reshape(-1, 2), golden["norm_steane_swaps"])
    m = np.arange(12).reshape(3, 4)
    css_code.swap_columns(m, (0, 2))
    assert m[:, 0].tolist() == [2, 6, 10] and m[:, 2].tolist() == [0, 4, 8]


def test_constructor_errors_in_reference_order():
    h = np.array(codes.hamming_7_4())
    with pytest.raises(ValueError, match="C_1 and C_2 must have the same code word length"):
        CSSCode(h, h[:, :6])
    with pytest.raises(ValueError, match="C_1 parity check matrix must be binary"):
        CSSCode(h * 2, h)
    with pytest.raises(ValueError, match="C_2 parity check matrix must be binary"):
        CSSCode(h, h + 2)
    bad = h.copy(); bad[0, 0] = 1
    with pytest.raises(ValueError, match="C_2 dual code must be a subspace of C_1"):
        CSSCode(h, bad)
    with pytest.raises(ValueError, match="not enough columns"):
        css_code.normalize_parity_check(np.ones((3, 2), dtype='int'), 0)
    with pytest.raises(errors.InvalidCodeError, match="rows are not independent"):
        css_code.normalize_parity_check(np.array([[1, 1, 0], [1, 1, 0]]), 0)
    # k != 1 is only rejected after all the tables are built (css_code.py:69-75)
    four = np.array([[1, 1, 1, 1]])                                   # the [[4,2,2]] code: k = 2
    with pytest.raises(errors.InvalidCodeError, match="single logical qubit"):
        CSSCode(four, four)
    hx, hz = codes.hgp1600()
    with pytest.raises(errors.InvalidCodeError):
        CSSCode(hx, hz)


def test_plane_packing_roundtrip():
    rng = np.random.default_rng(0)
    for shots, n in ((1, 3), (64, 7), (65, 23), (1000, 40)):
        v = rng.integers(0, 2, size=(shots, n), dtype=np.uint8)
        p = planes.pack_planes(v)
        assert p.shape == (n, planes.stride_words(shots)) and p.shape[1] % 8 == 0
        assert np.array_equal(planes.unpack_planes(p, shots), v)
        assert (int(p[0, 0]) >> 0) & 1 == v[0, 0]                      # shot 0 is bit 0 of word 0
    assert planes.pack_planes(np.array([[0], [1], [1]]))[0, 0] == 6


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(REPO, "include", "qcss.h")).read()
    declared = set(re.findall(r"QCSS_API\s+(?:const\s+char\*|int)\s+(qcss_\w+)\s*\(", header))
    assert len(declared) >= 20
    assert declared == set(_native.PROTOTYPES), declared ^ set(_native.PROTOTYPES)
    lib = _native.load()                                               # builds nothing: must exist
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.qcss_version() >= 100


def test_library_reads_no_environment_knobs(monkeypatch):
    """Kernel selection goes through qcss_set_option only: the shipped library contains none of the round-1
    environment-variable names (nor any other QCSS_* string), so an inherited variable cannot change which
    kernel runs or what it computes; the superseded kernels are not in the binary either."""
    blob = open(_native.LIB_PATH, "rb").read()
    for name in (b"QCSS_RING_DBG", b"QCSS_TMA_DBG", b"QCSS_DENSE_DBG", b"QCSS_TILES", b"QCSS_RING", b"QCSS_GAPQ",
                 b"QCSS_DENSE", b"QCSS_SAMPLER_BITS", b"QCSS_DISABLE_NAMED", b"QCSS_GF2_", b"QCSS_TILED_"):
        assert name not in blob, name
    for kernel in (b"k_syndrome_tma", b"k_syndrome_wide", b"k_syndrome_l1", b"k_gf2_fast"):
        assert kernel not in blob, kernel
    # setting the old variables changes nothing the library can observe: options keep their defaults
    for name in ("QCSS_GAPQ", "QCSS_DENSE", "QCSS_RING_DBG", "QCSS_GF2_SIMPLE"):
        monkeypatch.setenv(name, "1")
    lib = _native.load()
    import ctypes
    for key, default in (("gapq", 1), ("dense", -1), ("named", 1), ("gf2_kernel", 0), ("host_compact", 1), ("host_threads", 0)):
        val = ctypes.c_int(99)
        assert lib.qcss_get_option(key.encode(), ctypes.byref(val)) == 0 and val.value == default
    with _native.option("gf2_kernel", 2):
        assert lib.qcss_get_option(b"gf2_kernel", ctypes.byref(val)) == 0 and val.value == 2
    assert lib.qcss_get_option(b"gf2_kernel", ctypes.byref(val)) == 0 and val.value == 0
    with pytest.raises(ValueError):
        _native.set_option("gapq", 7)
    with pytest.raises(ValueError):
        _native.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        _native.set_option("host_threads", 1000)
    with _native.option("host_compact", 0), _native.option("host_threads", 12):
        assert lib.qcss_get_option(b"host_threads", ctypes.byref(val)) == 0 and val.value == 12
    assert lib.qcss_get_option(b"host_compact", ctypes.byref(val)) == 0 and val.value == 1


def test_no_cpu_fallback_without_gpu():
    """Without a device every numeric entry point fails loudly instead of computing on the host."""
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("GPU present")
    steane = CSSCode(*[np.array(h) for h in codes.steane()])
    with pytest.raises(_native.NativeLibraryError):
        steane.syndromes(np.zeros((2, 7), dtype=np.uint8), 1)
    with pytest.raises(_native.NativeLibraryError):
        bin_matrix.reduced_row_echelon_form(np.eye(3, dtype=int))
    # every entry point added since: decode, Monte Carlo, EC rounds, device construction, tiles, null space
    import css_code
    from quantum_css_codes_b200 import SyndromeCode
    h = np.array(codes.hamming_7_4())
    big = SyndromeCode(np.eye(40, 60, dtype=int), np.eye(40, 60, dtype=int))
    calls = [
        lambda: steane.decode(np.zeros((2, 7), dtype=np.uint8), 2),
        lambda: steane.decode_xz(np.zeros((2, 7), dtype=np.uint8), np.zeros((2, 7), dtype=np.uint8)),
        lambda: steane.monte_carlo(1e-3, 1000),
        lambda: steane.error_correct_monte_carlo(1e-3, 1e-3, 2, 1000),
        lambda: steane.syndrome_histogram(np.zeros((2, 7), dtype=np.uint8), 1),
        lambda: css_code.normalize_parity_check_gpu(h.copy(), 0),
        lambda: CSSCode(h, h, standard_form="gpu"),
        lambda: css_code.syndrome_table_gpu(h),
        lambda: big.syndromes_tiled(np.zeros((5, 60), dtype=np.uint8), 1),
        lambda: big.syndromes_tiled(np.zeros((0, 60), dtype=np.uint8), 1),
        lambda: big.sample_syndromes(1e-3, 100),
        lambda: bin_matrix.null_space(np.eye(3, 5, dtype=int)),
        lambda: bin_matrix.solve(np.eye(3, dtype=int), np.ones(3, dtype=int)),
    ]
    for call in calls:
        with pytest.raises(_native.NativeLibraryError):
            call()


def test_product_never_imports_oracle():
    pkg = os.path.join(REPO, "quantum_css_codes_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
    for f in ("bin_matrix.py", "css_code.py", "errors.py"):
        assert "oracle" not in open(os.path.join(REPO, f)).read()


def test_tile_major_pack_helpers_round_trip():
    """planes.pack_tiles / unpack_tiles / planes_to_tiles (host-side layout helpers, no GPU): round trip, zero
    padding of the last tile, agreement with the plane-major packing."""
    from quantum_css_codes_b200 import planes
    rng = np.random.default_rng(12)
    for shots, n in ((0, 5), (1, 3), (1023, 9), (1024, 9), (2500, 9), (4096, 40)):
        v = (rng.random((shots, n)) < 0.3).astype(np.uint8)
        tiles = planes.pack_tiles(v)
        assert tiles.shape == ((shots + 1023) // 1024, n, 16) and tiles.dtype == np.uint64
        assert np.array_equal(planes.unpack_tiles(tiles, shots), v)
        assert np.array_equal(planes.planes_to_tiles(planes.pack_planes(v), shots), tiles)
        if shots % 1024:
            bits = np.unpackbits(tiles[-1].view(np.uint8), axis=1, bitorder="little")
            assert not bits[:, shots % 1024:].any()
    one = np.zeros((1030, 2), dtype=np.uint8)
    one[1029, 1] = 1                                           # shot 1029 = tile 1, bit 5 of word 0 of plane 1
    assert planes.pack_tiles(one)[1, 1, 0] == np.uint64(1 << 5)


class _ReferenceShapedCode:
    """Duck-typed stand-in for a CSSCode built by the unmodified reference: exactly the attributes the
    reference object has on this path (css_code.py:63-72, 124-161), filled from the oracle restatement."""

    def __init__(self, name):
        from oracle import css as ocss
        ref = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
        self._n, self._k, self._t = ref.n, ref.k, ref.t
        self.r_1, self.r_2 = ref.r_1, ref.r_2
        self.parity_check_c1, self.parity_check_c2 = ref.parity_check_c1, ref.parity_check_c2
        self._c1_syndromes, self._c2_syndromes = ref.c1_syndromes, ref.c2_syndromes
        self._transversal_gates = frozenset(ref.transversal_gates)
        self._lx, self._lz = ref.lx, ref.lz
        self.gate_emitter_marker = "stays with the reference object"

    n = property(lambda self: self._n)
    k = property(lambda self: self._k)
    t = property(lambda self: self._t)

    def x_operator_matrix(self):
        return self._lx

    def z_operator_matrix(self):
        return self._lz


def test_attach_binds_device_path_to_a_reference_shaped_object():
    """``attach(obj)`` wraps an existing code object (a reference CSSCode, or anything
    with its attributes) without recomputing tables; missing attributes are reported by name; other
    attribute reads fall through to the wrapped object."""
    from quantum_css_codes_b200 import AttachedCode, attach
    obj = _ReferenceShapedCode("steane")
    bound = attach(obj)
    assert isinstance(bound, AttachedCode)
    assert bound.n == 7 and bound.t == 1 and bound.gate_emitter_marker.startswith("stays")
    assert bound.parity_check_c1 is obj.parity_check_c1 and bound._side(2)[1] is obj._c2_syndromes
    with pytest.raises(TypeError, match="_c1_syndromes"):
        class Partial:
            n = 7
            parity_check_c1 = parity_check_c2 = np.eye(3, 7, dtype=int)
        attach(Partial())
    from conftest import has_cuda
    if not has_cuda():
        with pytest.raises(_native.NativeLibraryError):        # reaches qcss_code_create: loud, no fallback
            bound.monte_carlo(1e-3, 1000)


def test_transversal_gates_is_a_frozenset_like_the_reference():
    """css_code.py:72, 201: ``_transversal_gates`` is a frozenset built by ``_determine_transversal_gates``."""
    steane = CSSCode(*[np.array(h) for h in codes.steane()])
    assert hasattr(steane, "_determine_transversal_gates")
    from conftest import has_cuda
    if has_cuda():
        assert steane._transversal_gates == frozenset(['I', 'CNOT', 'H', 'CZ', 'S'])
        assert isinstance(steane._transversal_gates, frozenset)


def test_attach_accepts_the_unmodified_reference_object():
    """Where the reference checkout exists (the build container), attach() takes a CSSCode built by the UNMODIFIED
    reference (imported next to an inert pyquil stub, as oracle/gen_golden.py does): its attributes are exactly what
    DeviceCode.from_csscode uploads, and equal our own constructor's."""
    if not os.path.exists("/root/reference/css_code.py"):
        pytest.skip("reference checkout not present (GPU box)")
    import subprocess, sys, textwrap
    script = textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r)
        from oracle import gen_golden
        ref_bm, ref_css = gen_golden.load_reference()              # the reference's modules, unmodified
        sys.path.insert(0, %r)
        from quantum_css_codes_b200 import attach, codes, _native
        from quantum_css_codes_b200.css_code import CSSCode as Ours
        for name in ("steane", "golay23"):
            h1, h2 = [np.array(h) for h in getattr(codes, name)()]
            theirs = ref_css.CSSCode(h1.copy(), h2.copy())
            assert type(theirs).__module__ == "css_code" and ref_css.__file__.startswith("/root/reference")
            bound, ours = attach(theirs), Ours(h1.copy(), h2.copy())
            assert np.array_equal(bound.parity_check_c1, ours.parity_check_c1)
            assert np.array_equal(bound.x_operator_matrix(), ours.x_operator_matrix())
            assert list(bound._c2_syndromes) == list(ours._c2_syndromes)
            assert bound.is_transversal("H") == (name in ("steane", "golay23"))       # falls through to the reference object
            try:
                bound.device                                       # uploads through qcss_code_create
                uploaded = True
            except _native.NativeLibraryError:
                uploaded = False                                   # no GPU in the build container: loud, not silent
            print(name, "uploaded" if uploaded else "no-device")
        """ % (REPO, REPO))
    res = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, cwd=REPO)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "steane" in res.stdout and "golay23" in res.stdout


def test_allow_multi_logical_constructor_matches_oracle_and_default_still_raises():
    """k > 1 objects only on request; the reference's exception stays the default
    (css_code.py:74-75).  Operator matrices get k rows (css_code.py:124-161 already slice k)."""
    from oracle import css as ocss
    h = np.array(codes.hamming_7_4())
    with pytest.raises(errors.InvalidCodeError, match="single logical qubit"):
        CSSCode(h, h[:2])
    code = CSSCode(h, h[:2], allow_multi_logical=True)
    ref = ocss.build_css(h, h[:2], allow_k_not_1=True)
    assert (code.n, code.k, code.r_1, code.r_2) == (7, 2, 3, 2) == (ref.n, ref.k, ref.r_1, ref.r_2)
    assert np.array_equal(code.x_operator_matrix(), ref.lx) and code.x_operator_matrix().shape == (2, 7)
    assert np.array_equal(code.z_operator_matrix(), ref.lz)
    assert np.array_equal(code.parity_check_c1, ref.parity_check_c1) and list(code._c2_syndromes) == list(ref.c2_syndromes)
    # logical Z rows commute with the X-type stabilisers (rows of H1), logical X rows with the Z-type ones (H2)
    assert not np.any(code.z_operator_matrix() @ code.parity_check_c1.T % 2)
    assert not np.any(code.x_operator_matrix() @ code.parity_check_c2.T % 2)
