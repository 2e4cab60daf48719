"""Exception types of the drop-in surface (reference errors.py:5,8)."""


class InvalidCodeError(Exception):
    pass


class UnsupportedGateError(Exception):
    pass


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing, failed to load, or a call into it failed.  There is no CPU
    fallback: every numeric entry point raises this instead of computing on the host."""
