"""Top-level drop-in module: the reference's tests do ``from css_code import CSSCode``
(test/test_css_code.py:5-7).  Everything lives in quantum_css_codes_b200.css_code."""
from quantum_css_codes_b200.css_code import (            # noqa: F401
    CSSCode, SyndromeCode, syndrome_table, syndrome_table_gpu, swap_columns, normalize_parity_check,
    normalize_parity_check_gpu,
    codes_equal, is_doubly_even, quil_classical_correct, quil_classical_detect, InvalidCodeError, UnsupportedGateError)
