"""
Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in the build
container.  Test infrastructure only; run once here, the fixtures are committed because
/root/reference does not exist on the GPU box.

    python oracle/gen_golden.py            # writes tests/golden/reference_golden.npz

``css_code.py`` imports pyquil at module top (css_code.py:7-11, qecc.py:6-8,
quil_classical.py:5-7) and pyquil is not installed, so an inert stub package that only
provides the imported names is put first on sys.path (in a temp dir, never committed as
product code).  None of the numeric functions exercised below touch pyquil.
"""

import os
import sys
import tempfile
import textwrap

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("QCSS_REFERENCE", "/root/reference")


def _install_pyquil_stub(root):
    pkg = os.path.join(root, "pyquil")
    os.makedirs(pkg)
    files = {
        "__init__.py": "class Program:\n    pass\ndef get_qc(*a, **k):\n    raise RuntimeError('stub')\n",
        "gates.py": "def __getattr__(name):\n    raise AttributeError(name)\n",
        "paulis.py": "class PauliTerm:\n    pass\ndef ID():\n    raise RuntimeError('stub')\nsX = sY = sZ = ID\n",
        "quil.py": "class Program:\n    pass\n",
        "quilatom.py": "class MemoryReference:\n    pass\nclass Qubit:\n    pass\nclass QubitPlaceholder:\n    pass\n",
        "quilbase.py": "class Gate:\n    pass\n",
    }
    for name, body in files.items():
        with open(os.path.join(pkg, name), "w") as fh:
            fh.write(textwrap.dedent(body))


def load_reference():
    stub_root = tempfile.mkdtemp(prefix="pyquil_stub_")
    _install_pyquil_stub(stub_root)
    sys.path.insert(0, stub_root)
    sys.path.insert(0, REFERENCE)
    import warnings
    warnings.simplefilter("ignore")
    import bin_matrix as ref_bm           # noqa: E402  (the reference's, from /root/reference)
    import css_code as ref_css            # noqa: E402
    assert os.path.dirname(ref_bm.__file__) == REFERENCE, ref_bm.__file__
    return ref_bm, ref_css


def table_arrays(table, n):
    keys = np.array([int(k) for k in table.keys()], dtype=np.int64)      # insertion order
    vals = np.array([np.asarray(v, dtype=np.int64) for v in table.values()],
                    dtype=np.int64).reshape(len(keys), n)
    return keys, vals


def main():
    sys.path.insert(0, REPO)
    ref_bm, ref_css = load_reference()
    # our own input generators (not reference code)
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "qcss_codes", os.path.join(REPO, "quantum_css_codes_b200", "codes.py"))
    codes = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(codes)

    out = {}

    # ---- bin_matrix: RREF on seeded random matrices of assorted shapes -------------------
    rng = np.random.default_rng(20261018)
    shapes = [(3, 7), (1, 1), (5, 5), (8, 3), (3, 8), (16, 40), (33, 70), (64, 64), (65, 130),
              (20, 100), (100, 20), (128, 256)]
    for idx, (m, n) in enumerate(shapes):
        mat = rng.integers(0, 2, size=(m, n), dtype=np.int64)
        if idx % 3 == 2 and m > 2:                      # force rank deficiency / zero columns
            mat[m - 1] = (mat[0] + mat[1]) % 2
            mat[:, n // 2] = 0
        out[f"rref_in_{idx}"] = mat
        out[f"rref_out_{idx}"] = ref_bm.reduced_row_echelon_form(mat)
    out["rref_count"] = np.array(len(shapes))
    # un-reduced integer entries and uint8 dtype (reference only tests parity; dtype preserved)
    mat = rng.integers(0, 7, size=(6, 11), dtype=np.int64)
    out["rref_in_wide"] = mat
    out["rref_out_wide"] = ref_bm.reduced_row_echelon_form(mat)
    mat8 = rng.integers(0, 2, size=(9, 17), dtype=np.uint8)
    out["rref_in_u8"] = mat8
    out["rref_out_u8"] = ref_bm.reduced_row_echelon_form(mat8)

    # ---- vec<->int, weight_w_vectors -----------------------------------------------------
    vecs = rng.integers(0, 2, size=(32, 40), dtype=np.int64)
    out["v2i_in"] = vecs
    out["v2i_out"] = np.array([int(ref_bm.vec_to_int(v)) for v in vecs], dtype=np.int64)
    out["i2v_out"] = np.array([ref_bm.int_to_vec(int(k), 40) for k in out["v2i_out"]])
    out["wwv_6_3"] = np.array(list(ref_bm.weight_w_vectors(6, 3)))
    out["wwv_5_0"] = np.array(list(ref_bm.weight_w_vectors(5, 0)))
    out["wwv_4_4"] = np.array(list(ref_bm.weight_w_vectors(4, 4)))

    # ---- CSSCode for the named codes -----------------------------------------------------
    named = {"steane": codes.steane(), "qrm15": codes.qrm15(), "golay23": codes.golay23()}
    for name, (h1, h2) in named.items():
        code = ref_css.CSSCode(np.array(h1), np.array(h2))
        n = code.n
        out[f"{name}_in1"], out[f"{name}_in2"] = np.array(h1), np.array(h2)
        out[f"{name}_nkt"] = np.array([code.n, code.k, code.t, code.r_1, code.r_2])
        out[f"{name}_h1"], out[f"{name}_h2"] = code.parity_check_c1, code.parity_check_c2
        out[f"{name}_lz"], out[f"{name}_lx"] = code.z_operator_matrix(), code.x_operator_matrix()
        out[f"{name}_c1_keys"], out[f"{name}_c1_vals"] = table_arrays(code._c1_syndromes, n)
        out[f"{name}_c2_keys"], out[f"{name}_c2_vals"] = table_arrays(code._c2_syndromes, n)
        out[f"{name}_gates"] = np.array(sorted(code._transversal_gates))

        # per-shot reference-literal decode of a seeded batch (both Pauli types)
        brng = np.random.default_rng(len(name) * 7919)
        for which, h, tab, lop in ((2, code.parity_check_c2, code._c2_syndromes, code.z_operator_matrix()),
                                   (1, code.parity_check_c1, code._c1_syndromes, code.x_operator_matrix())):
            m = h.shape[0]
            errs = (brng.random((512, n)) < 0.12).astype(np.int64)
            errs[0] = 0
            errs[1] = 1
            synd = np.zeros((512, m), dtype=np.int64)
            keys = np.zeros(512, dtype=np.int64)
            corr = np.zeros((512, n), dtype=np.int64)
            miss = np.zeros(512, dtype=np.int64)
            flip = np.zeros(512, dtype=np.int64)
            for i, e in enumerate(errs):
                s = np.mod(np.matmul(h, e), 2)                    # css_code.py:728
                key = ref_bm.vec_to_int(s)                        # bin_matrix.py:36
                c = tab.get(key)
                r = e if c is None else (e + c) % 2               # css_code.py:677-682
                synd[i], keys[i] = s, int(key)
                miss[i] = c is None
                if c is not None:
                    corr[i] = c
                flip[i] = np.mod(np.matmul(lop, r), 2)[0]         # css_code.py:641-646
            pre = f"{name}_w{which}"
            out[pre + "_errs"], out[pre + "_synd"], out[pre + "_keys"] = errs, synd, keys
            out[pre + "_corr"], out[pre + "_miss"], out[pre + "_flip"] = corr, miss, flip

    # ---- module functions ----------------------------------------------------------------
    h = np.array(codes.hamming_7_4())
    norm_in = h.copy()
    norm_out, swaps = ref_css.normalize_parity_check(norm_in, 0)
    out["norm_steane_out"], out["norm_steane_mutated"] = norm_out, norm_in
    out["norm_steane_swaps"] = np.array(swaps, dtype=np.int64).reshape(-1, 2)
    t, tab = ref_css.syndrome_table(code.parity_check_c1)      # golay c1 (last code in loop)
    out["golay_table_t"] = np.array(t)
    de = np.array([[0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 1, 0, 1, 1, 0],
                   [1, 1, 1, 0, 0, 0, 0, 1], [1, 1, 1, 1, 1, 1, 1, 1]])
    out["doubly_even_true"] = np.array(ref_css.is_doubly_even(de))
    de2 = de.copy(); de2[2, 0] = 0
    out["doubly_even_false"] = np.array(ref_css.is_doubly_even(de2))
    a = rng.integers(0, 2, size=(5, 12), dtype=np.int64)
    mix = rng.integers(0, 2, size=(5, 5), dtype=np.int64)
    while round(abs(np.linalg.det(mix))) % 2 == 0:
        mix = rng.integers(0, 2, size=(5, 5), dtype=np.int64)
    out["ce_a"], out["ce_b"] = a, (mix @ a) % 2
    out["ce_equal"] = np.array(ref_css.codes_equal(a, (mix @ a) % 2))
    b = a.copy(); b[0, 0] ^= 1
    out["ce_c"] = b
    out["ce_unequal"] = np.array(ref_css.codes_equal(a, b))

    # ---- HGP-1600: the reference constructor must reject it; syndromes via the raw idiom --
    hx, hz = codes.hgp1600()
    try:
        ref_css.CSSCode(np.array(hx), np.array(hz))
        rejected = "accepted"
    except Exception as exc:                                    # InvalidCodeError
        rejected = type(exc).__name__ + ": " + str(exc)
    out["hgp_rejected"] = np.array(rejected)
    errs = (rng.random((64, 1600)) < 0.01).astype(np.int64)
    out["hgp_errs"] = np.packbits(errs.astype(np.uint8), axis=1, bitorder="little")
    out["hgp_synd_hz"] = np.packbits(
        np.array([np.mod(np.matmul(hz, e), 2) for e in errs], dtype=np.uint8), axis=1, bitorder="little")
    out["hgp_synd_hx"] = np.packbits(
        np.array([np.mod(np.matmul(hx, e), 2) for e in errs], dtype=np.uint8), axis=1, bitorder="little")

    dst = os.path.join(REPO, "tests", "golden", "reference_golden.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
