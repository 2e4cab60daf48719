"""
Golden vectors for BASELINE config 5 at FULL size, made by the UNMODIFIED reference
(/root/reference/bin_matrix.py:8-34 ``reduced_row_echelon_form``) in the build container:
16 matrices of 1024 x 2048 from SURVEY 8d's ``default_rng(5)`` draw, a quarter of them made rank deficient,
a quarter with zero columns, a quarter with zero leading columns.  The fixture holds the SHA-256 of each
packed RREF (bit j of little-endian uint64 word w = column 64 w + j), its rank, and the full packed RREF of
matrix 0 and matrix 1 for diffing.  Test infrastructure only.

    python oracle/gen_c5_golden.py         # writes tests/golden/c5_rref_golden.npz  (~1 minute on 8 cores)
"""

import hashlib
import importlib.util
import multiprocessing as mp
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
REFERENCE = os.environ.get("QCSS_REFERENCE", "/root/reference")
COUNT = 16


def variant(index, bits):
    """The same edits tests/test_gpu_gf2.py applies before calling the CUDA kernel."""
    bits = bits.copy()
    kind = index % 4
    if kind == 1:                                   # dependent rows: rank 1022
        bits[1000] = bits[3] ^ bits[7]
        bits[555] = bits[4]
    elif kind == 2:                                 # zero columns inside the pivot range
        bits[:, [5, 700, 1023]] = 0
    elif kind == 3:                                 # zero leading columns and a duplicated column
        bits[:, :3] = 0
        bits[:, 10] = bits[:, 9]
    return bits


def work(index):
    from quantum_css_codes_b200 import codes
    from oracle import gf2 as ogf2
    spec = importlib.util.spec_from_file_location("reference_bin_matrix", os.path.join(REFERENCE, "bin_matrix.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    packed = codes.random_matrices_c5(1, offset=index)[0]
    bits = variant(index, ogf2.unpack_rows(packed, 2048))
    out = ref.reduced_row_echelon_form(bits.astype(np.int64))
    assert out.shape == (1024, 2048) and set(np.unique(out)) <= {0, 1}
    out_packed = ogf2.pack_rows(out.astype(np.uint8))
    rank = int(np.count_nonzero(out.any(axis=1)))
    return hashlib.sha256(out_packed.tobytes()).hexdigest(), rank, out_packed


def main():
    with mp.get_context("fork").Pool(min(COUNT, os.cpu_count() or 1)) as pool:
        res = pool.map(work, range(COUNT))
    np.savez_compressed(os.path.join(REPO, "tests", "golden", "c5_rref_golden.npz"),
                        sha256=np.array([r[0] for r in res]), rank=np.array([r[1] for r in res], dtype=np.int32),
                        rref_0=res[0][2], rref_1=res[1][2])
    print([r[1] for r in res])


if __name__ == "__main__":
    main()
