"""Top-level drop-in module: the reference's tests do ``import bin_matrix``
(test/test_bin_matrix.py:4).  Everything lives in quantum_css_codes_b200.bin_matrix."""
from quantum_css_codes_b200.bin_matrix import *          # noqa: F401,F403
from quantum_css_codes_b200.bin_matrix import (          # noqa: F401
    reduced_row_echelon_form, vec_to_int, int_to_vec, weight_w_vectors,
    rref_batched, rref_packed_batched, rank, null_space, solve,
    null_space_batched, null_space_packed_batched, solve_batched)
