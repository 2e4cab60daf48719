"""
quantum_css_codes_b200 -- B200-native (sm_100a) hot path of jimpo/quantum-css-codes:
Monte-Carlo syndrome extraction and lookup decoding for CSS codes plus the GF(2) toolkit,
behind the reference's ``bin_matrix`` / ``css_code`` Python surface.

Importing the package does not load the CUDA library; the first numeric call does, and fails
loudly (``NativeLibraryError``) when ``libqcss.so`` has not been built.  No CPU fallback.
"""

from . import bin_matrix, codes, css_code, errors, planes            # noqa: F401
from .css_code import AttachedCode, CSSCode, SyndromeCode, attach     # noqa: F401
from .errors import InvalidCodeError, NativeLibraryError, UnsupportedGateError  # noqa: F401

__all__ = ["bin_matrix", "css_code", "codes", "errors", "planes", "CSSCode", "SyndromeCode", "AttachedCode", "attach",
           "InvalidCodeError", "UnsupportedGateError", "NativeLibraryError"]
