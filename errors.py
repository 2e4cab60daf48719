"""Top-level drop-in module mirroring the reference's errors.py:5,8."""
from quantum_css_codes_b200.errors import (              # noqa: F401
    InvalidCodeError, UnsupportedGateError, NativeLibraryError)
