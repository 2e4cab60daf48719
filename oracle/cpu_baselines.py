"""
CPU baselines of SURVEY 8(d), measured with the UNMODIFIED reference where it is present
(/root/reference, this build container) and with the oracle's literal restatement otherwise:

  (1) reference-literal Monte-Carlo loop: per shot  np.mod(np.matmul(H, e), 2) -> vec_to_int ->
      table.get -> L.r  (css_code.py:728, bin_matrix.py:36-43, css_code.py:649-685, 641-646), 1 core;
  (2) batched-numpy restatement of the same arithmetic (oracle.montecarlo.tally_xz), 1 core;
  (4) bin_matrix.reduced_row_echelon_form (bin_matrix.py:8-34) on C5 matrices (1024 x 2048), 1 core.

Test infrastructure only (never imported by the product).  Writes one JSON document.

    python oracle/cpu_baselines.py [--shots 200000] [--matrices 2] > profiles/r01_cpu_baselines.json
"""

import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from oracle import css as ocss, gf2 as ogf2, montecarlo as omc       # noqa: E402
from quantum_css_codes_b200 import codes                            # noqa: E402

REFERENCE = os.environ.get("QCSS_REFERENCE", "/root/reference")


def reference_bin_matrix():
    """The reference's own bin_matrix module (numpy only), or None on machines without the checkout."""
    path = os.path.join(REFERENCE, "bin_matrix.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_bin_matrix", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def literal_loop(code, ex, ez, vec_to_int):
    """SURVEY A.3, one Python iteration per shot and Pauli type, reference primitives only."""
    fails = 0
    for which, errs in ((2, ex), (1, ez)):
        h, table, lop = ocss.pauli_side(code, which)
        for e in errs:
            s = np.mod(np.matmul(h, e), 2)
            c = table.get(vec_to_int(s))
            r = e if c is None else (e + c) % 2
            fails += int(np.mod(np.matmul(lop, r), 2)[0])
    return fails


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shots", type=int, default=200000)
    ap.add_argument("--matrices", type=int, default=2)
    args = ap.parse_args()
    ref = reference_bin_matrix()
    out = {"host_cpus": os.cpu_count(), "reference_present": ref is not None,
           "implementation": "reference bin_matrix.py" if ref is not None else "oracle literal restatement"}
    vec_to_int = ref.vec_to_int if ref is not None else ogf2.vec_to_int
    rref = ref.reduced_row_echelon_form if ref is not None else ogf2.rref_literal

    for name in ("steane", "golay23"):
        code = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
        rng = np.random.default_rng(7)
        ex, ez = omc.sample_depolarizing(rng, args.shots, code.n, 1e-3)
        ex, ez = ex.astype(np.int64), ez.astype(np.int64)
        t0 = time.perf_counter()
        fails = literal_loop(code, ex, ez, vec_to_int)
        t_lit = time.perf_counter() - t0
        t0 = time.perf_counter()
        tally = omc.tally_xz(code, ex, ez)
        t_bat = time.perf_counter() - t0
        assert fails == tally["fail_x"] + tally["fail_z"]
        out[f"mc_{name}"] = {"shots": args.shots, "p": 1e-3, "cores": 1,
                             "reference_literal_shots_per_s": args.shots / t_lit,
                             "batched_numpy_shots_per_s": args.shots / t_bat}

    mats = codes.random_matrices_c5(args.matrices)
    times = []
    for b in range(args.matrices):
        mat = ogf2.unpack_rows(mats[b], 2048).astype(np.int64)
        t0 = time.perf_counter()
        got = rref(mat)
        times.append(time.perf_counter() - t0)
        want, _ = ogf2.rref_packed(mats[b], 2048)
        assert np.array_equal(got, ogf2.unpack_rows(want, 2048))
    out["rref_c5"] = {"shape": [1024, 2048], "matrices": args.matrices, "cores": 1,
                      "seconds_per_matrix": float(np.mean(times)),
                      "extrapolated_4096_matrices_s": float(np.mean(times)) * 4096}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
