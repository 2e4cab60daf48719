"""
Golden vectors for the standard form (css_code.normalize_parity_check, css_code.py:809-836, and the
two-matrix sequence of CSSCode.__init__, css_code.py:55-61) from the UNMODIFIED reference.  Inputs are
chosen so that the qubit-swap branch (:822-829) fires: columns of full-rank matrices are permuted and
zero columns inserted.  Test infrastructure only; run here once (needs /root/reference):

    python oracle/gen_normalize_golden.py        # writes tests/golden/normalize_golden.npz
"""

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_golden import REPO, load_reference                      # noqa: E402


def full_rank_with_swaps(rng, r, n, offset):
    """Random r x n matrix of rank r whose leading block is singular (so swaps are needed)."""
    while True:
        a = rng.integers(0, 2, size=(r, n), dtype=np.int64)
        a[:, offset + rng.integers(0, r)] = 0                     # a dead pivot column
        if r > 2:
            a[:, offset + rng.integers(0, r)] = a[:, offset]      # and a repeated one
        if np.linalg.matrix_rank(a[:, offset:].astype(float)) < r:   # cheap necessary screen, exact check below
            continue
        return a


def main():
    ref_bm, ref_css = load_reference()
    sys.path.insert(0, REPO)
    from quantum_css_codes_b200 import codes

    rng = np.random.default_rng(809836)
    out = {}
    count = 0
    for r, n, offset in [(3, 7, 0), (3, 7, 2), (4, 15, 0), (10, 15, 4), (11, 23, 0), (11, 23, 11), (20, 70, 5),
                         (40, 130, 64), (64, 64, 0), (33, 200, 100)]:
        for _ in range(40):
            a = full_rank_with_swaps(rng, r, n, offset) if n > offset + r else rng.integers(0, 2, size=(r, n))
            work = a.copy()
            try:
                res, swaps = ref_css.normalize_parity_check(work, offset)
            except Exception:                                     # dependent rows: keep one per shape below
                continue
            out[f"n{count}_in"] = a.astype(np.uint8)
            out[f"n{count}_offset"] = np.int64(offset)
            out[f"n{count}_out"] = res.astype(np.uint8)
            out[f"n{count}_swaps"] = np.array(swaps, dtype=np.int64).reshape(-1, 2)
            count += 1
            break
        else:
            raise SystemExit(f"no independent sample for {(r, n, offset)}")
    out["norm_count"] = np.int64(count)

    # two-matrix standard form of CSSCode.__init__ on column-permuted named codes
    pairs = 0
    for name in ("steane", "qrm15", "golay23", "shor9"):
        h1, h2 = [np.array(h) for h in getattr(codes, name)()]
        for trial in range(3):
            perm = rng.permutation(h1.shape[1])
            p1, p2 = h1[:, perm], h2[:, perm]
            code = ref_css.CSSCode(p1.copy(), p2.copy())
            out[f"p{pairs}_in1"], out[f"p{pairs}_in2"] = p1.astype(np.uint8), p2.astype(np.uint8)
            out[f"p{pairs}_h1"] = code.parity_check_c1.astype(np.uint8)
            out[f"p{pairs}_h2"] = code.parity_check_c2.astype(np.uint8)
            pairs += 1
    out["pair_count"] = np.int64(pairs)

    dst = os.path.join(REPO, "tests", "golden", "normalize_golden.npz")
    np.savez_compressed(dst, **out)
    nsw = [len(out[f"n{i}_swaps"]) for i in range(count)]
    print("wrote", dst, "matrices", count, "swaps per matrix", nsw, "pairs", pairs)


if __name__ == "__main__":
    main()
