"""
ctypes binding of libqcss.so (C ABI in include/qcss.h).

The library is built in-tree (``python -m quantum_css_codes_b200.build`` or
``__graft_entry__.build()``) as ``quantum_css_codes_b200/libqcss.so``.  There is no CPU
fallback: if the library is missing or a CUDA call fails, ``NativeLibraryError`` is raised.
"""

import contextlib
import ctypes
import os
import threading

import numpy as np

from . import planes as _planes
from .errors import NativeLibraryError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqcss.so")

QCSS_ERR_INVALID = -1
QCSS_ERR_CUDA = -2
QCSS_ERR_UNSUPPORTED = -3
QCSS_ERR_NOMEM = -4

_c_u64p = ctypes.POINTER(ctypes.c_uint64)
_c_u8p = ctypes.POINTER(ctypes.c_uint8)
_c_i64p = ctypes.POINTER(ctypes.c_int64)
_c_i32p = ctypes.POINTER(ctypes.c_int32)


class Tally(ctypes.Structure):
    _fields_ = [(name, ctypes.c_uint64)
                for name in ("shots", "fail_x", "fail_z", "fail_any", "miss_x", "miss_z")]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


TALLY_FIELDS = tuple(name for name, _ in Tally._fields_)


class DecodeIO(ctypes.Structure):
    _fields_ = [("ex", ctypes.c_void_p), ("ez", ctypes.c_void_p), ("e_stride", ctypes.c_int64),
                ("synd_x", ctypes.c_void_p), ("synd_z", ctypes.c_void_p), ("s_stride", ctypes.c_int64),
                ("corr_x", ctypes.c_void_p), ("corr_z", ctypes.c_void_p), ("c_stride", ctypes.c_int64),
                ("flip_x", ctypes.c_void_p), ("flip_z", ctypes.c_void_p),
                ("miss_x", ctypes.c_void_p), ("miss_z", ctypes.c_void_p),
                ("tally", ctypes.c_void_p)]


# name -> (restype, argtypes); every symbol include/qcss.h declares
PROTOTYPES = {
    "qcss_version": (ctypes.c_int, []),
    "qcss_last_error": (ctypes.c_char_p, []),
    "qcss_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "qcss_set_device": (ctypes.c_int, [ctypes.c_int]),
    "qcss_host_alloc": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]),
    "qcss_host_free": (ctypes.c_int, [ctypes.c_void_p]),
    "qcss_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    "qcss_get_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]),
    "qcss_code_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_u8p, ctypes.c_int, _c_u8p,
                                        _c_u8p, _c_u8p,
                                        ctypes.c_int64, _c_i64p, _c_u8p,
                                        ctypes.c_int64, _c_i64p, _c_u8p,
                                        ctypes.POINTER(ctypes.c_void_p)]),
    "qcss_code_create_multi": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_u8p, ctypes.c_int, _c_u8p, ctypes.c_int,
                                              _c_u8p, _c_u8p,
                                              ctypes.c_int64, _c_i64p, _c_u8p,
                                              ctypes.c_int64, _c_i64p, _c_u8p,
                                              ctypes.POINTER(ctypes.c_void_p)]),
    "qcss_code_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "qcss_code_kernel_name": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int]),
    "qcss_code_last_transfer": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)]),
    "qcss_code_spec_source": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int64,
                                             ctypes.POINTER(ctypes.c_int64)]),
    "qcss_code_load_specialized": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]),
    "qcss_code_specialize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]),
    "qcss_syndrome": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                     ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]),
    "qcss_syndrome_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                         ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]),
    "qcss_syndrome_tiles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                           ctypes.c_void_p]),
    "qcss_syndrome_tiles_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                               ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_sample_syndrome_tiles": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64,
                                                  ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p]),
    "qcss_sample_syndrome_tiles_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64,
                                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                      ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_syndrome_hist": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int64, ctypes.c_void_p]),
    "qcss_syndrome_hist_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_decode": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                   ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.POINTER(Tally)]),
    "qcss_decode_xz": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                      ctypes.c_int64, ctypes.POINTER(Tally)]),
    "qcss_decode_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(DecodeIO), ctypes.c_int64,
                                       ctypes.c_void_p]),
    "qcss_syndrome_shots": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                           ctypes.c_void_p]),
    "qcss_decode_shots": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(Tally)]),
    "qcss_decode_xz_shots": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int64,
                                            ctypes.POINTER(Tally)]),
    "qcss_pack_shots_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_int64, ctypes.c_void_p]),
    "qcss_unpack_planes_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                              ctypes.c_void_p]),
    "qcss_decode_xz_sparse": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                             ctypes.POINTER(Tally)]),
    "qcss_decode_xz_sparse_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_events_from_planes_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                   ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                                   ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_mc_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64,
                                   ctypes.c_int64, ctypes.POINTER(Tally)]),
    "qcss_mc_run_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64,
                                       ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_ec_run": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int64,
                                   ctypes.c_uint64, ctypes.c_int64, ctypes.POINTER(Tally)]),
    "qcss_ec_run_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_int64,
                                       ctypes.c_uint64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_mc_sample": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64,
                                      ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]),
    "qcss_mc_sample_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_uint64,
                                          ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_void_p]),
    "qcss_gf2_rref": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_void_p, _c_i32p, _c_i32p]),
    "qcss_gf2_rref_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_gf2_normalize": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, _c_i32p, _c_i32p, _c_i32p]),
    "qcss_gf2_normalize_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_css_standard_form": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, _c_i32p, _c_i32p, _c_i32p]),
    "qcss_table_build": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _c_u8p, ctypes.c_int64,
                                        ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int),
                                        ctypes.POINTER(ctypes.c_int64)]),
    "qcss_table_read": (ctypes.c_int, [ctypes.c_void_p, _c_i64p, ctypes.c_void_p]),
    "qcss_table_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "qcss_gf2_nullspace": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, _c_i32p]),
    "qcss_gf2_nullspace_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "qcss_gf2_solve": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_void_p, _c_i32p]),
    "qcss_gf2_solve_dev": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libqcss.so once; raise NativeLibraryError (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryError(
                f"{LIB_PATH} not found: build it with `python -m quantum_css_codes_b200.build` "
                "(needs nvcc).  There is no CPU fallback.")
        try:
            lib = ctypes.CDLL(LIB_PATH)
        except OSError as exc:
            raise NativeLibraryError(f"cannot load {LIB_PATH}: {exc}") from exc
        for name, (restype, argtypes) in PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return _lib


def check(rc):
    if rc == 0:
        return
    msg = load().qcss_last_error().decode("utf-8", "replace")
    if rc == QCSS_ERR_INVALID:
        raise ValueError(msg)
    if rc == QCSS_ERR_NOMEM:
        raise MemoryError(msg)
    raise NativeLibraryError(f"libqcss error {rc}: {msg}")


def set_option(name, value):
    """qcss_set_option: choose between bit-identical kernel implementations (include/qcss.h lists them).
    Returns the previous value.  The library reads nothing from the environment."""
    lib = load()
    old = ctypes.c_int()
    check(lib.qcss_get_option(name.encode(), ctypes.byref(old)))
    check(lib.qcss_set_option(name.encode(), int(value)))
    return old.value


@contextlib.contextmanager
def option(name, value):
    """``with option("gapq", 0): ...`` -- set a selection option for a block and restore it."""
    old = set_option(name, value)
    try:
        yield
    finally:
        set_option(name, old)


def _ptr(arr):
    return ctypes.c_void_p(arr.ctypes.data) if arr is not None else ctypes.c_void_p(0)


def _table_arrays(table, n):
    if table is None:
        return 0, None, None
    keys = np.array([int(k) for k in table.keys()], dtype=np.int64)
    corr = np.zeros((len(keys), n), dtype=np.uint8)
    for row, vec in enumerate(table.values()):
        corr[row] = np.asarray(vec, dtype=np.int64) & 1
    return len(keys), keys, corr


def _u8(mat):
    return None if mat is None else np.ascontiguousarray(np.asarray(mat) & 1, dtype=np.uint8)


class DeviceCode:
    """Owns a qcss_code handle; numpy-in / numpy-out wrappers over the host-buffer entry points
    and thin pass-throughs for the device-pointer ones."""

    def __init__(self, n, h1, h2, lx, lz, table1, table2):
        lib = load()
        h1, h2 = _u8(h1), _u8(h2)
        self.n, self.m1, self.m2 = int(n), h1.shape[0], h2.shape[0]
        lx, lz = _u8(lx), _u8(lz)                       # (n,) for one logical qubit, (k, n) for several
        self.k = 1 if lx is None or lx.ndim == 1 else lx.shape[0]
        n1, keys1, corr1 = _table_arrays(table1, n)
        n2, keys2, corr2 = _table_arrays(table2, n)
        handle = ctypes.c_void_p()
        as_u8 = lambda a: a.ctypes.data_as(_c_u8p) if a is not None else None
        as_i64 = lambda a: a.ctypes.data_as(_c_i64p) if a is not None else None
        if self.k == 1:
            lx1 = None if lx is None else np.ascontiguousarray(lx.reshape(-1))
            lz1 = None if lz is None else np.ascontiguousarray(lz.reshape(-1))
            check(lib.qcss_code_create(self.n, self.m1, as_u8(h1), self.m2, as_u8(h2), as_u8(lx1), as_u8(lz1),
                                       n1, as_i64(keys1), as_u8(corr1), n2, as_i64(keys2), as_u8(corr2),
                                       ctypes.byref(handle)))
        else:
            check(lib.qcss_code_create_multi(self.n, self.m1, as_u8(h1), self.m2, as_u8(h2), self.k, as_u8(lx), as_u8(lz),
                                             n1, as_i64(keys1), as_u8(corr1), n2, as_i64(keys2), as_u8(corr2),
                                             ctypes.byref(handle)))
        self._lib = lib
        self.handle = handle

    @classmethod
    def from_csscode(cls, code):
        """Device object for ANY code object carrying the reference's attributes (css_code.py:63-72, 124-161):
        ``n``, ``parity_check_c1/2`` (normalised), ``_c1/_c2_syndromes`` (dict key -> correction) and the
        logical rows ``x_operator_matrix()`` / ``z_operator_matrix()`` -- e.g. a ``CSSCode`` built by the
        unmodified reference.  Nothing is recomputed: the tables and matrices are uploaded as they stand."""
        lx = np.asarray(code.x_operator_matrix())
        lz = np.asarray(code.z_operator_matrix())
        if lx.shape[0] == 1 and lz.shape[0] == 1:
            lx, lz = lx[0], lz[0]
        return cls(int(code.n), np.asarray(code.parity_check_c1), np.asarray(code.parity_check_c2),
                   lx, lz, code._c1_syndromes, code._c2_syndromes)

    def __del__(self):
        handle, self.handle = getattr(self, "handle", None), None
        if handle:
            try:
                self._lib.qcss_code_destroy(handle)
            except Exception:
                pass

    def m(self, which):
        if which not in (1, 2):
            raise ValueError("which must be 1 (Z errors, C_1) or 2 (X errors, C_2)")
        return self.m1 if which == 1 else self.m2

    def kernel_name(self):
        buf = ctypes.create_string_buffer(256)
        check(self._lib.qcss_code_kernel_name(self.handle, buf, 256))
        return buf.value.decode()

    # ---- host-buffer entry points ---------------------------------------------------------
    def _check_planes(self, e_planes, shots):
        e_planes = np.ascontiguousarray(e_planes, dtype=np.uint64)
        if e_planes.ndim != 2 or e_planes.shape[0] != self.n:
            raise ValueError(f"expected ({self.n}, stride) uint64 planes")
        if e_planes.shape[1] % 2 or e_planes.shape[1] * 64 < shots:
            raise ValueError("plane stride must be even and cover all shots")
        return e_planes

    def syndrome_planes(self, e_planes, shots, which):
        e_planes = self._check_planes(e_planes, shots)
        stride = e_planes.shape[1]
        out = np.zeros((self.m(which), stride), dtype=np.uint64)
        check(self._lib.qcss_syndrome(self.handle, which, _ptr(e_planes), stride, shots, _ptr(out), stride))
        return out

    def syndrome_tiles(self, e_tiles, shots, which):
        """Tile-major batches: (tiles, n, 16) uint64 -> (tiles, m, 16) uint64 (qcss_syndrome_tiles)."""
        e_tiles = np.ascontiguousarray(e_tiles, dtype=np.uint64)
        tiles = (shots + 1023) // 1024
        if e_tiles.shape != (tiles, self.n, 16):
            raise ValueError("expected (ceil(shots / 1024), n, 16) uint64 tiles")
        out = np.zeros((tiles, self.m(which), 16), dtype=np.uint64)
        if tiles == 0:
            load()                                   # an empty batch still needs the library (no silent CPU path)
            return out
        check(self._lib.qcss_syndrome_tiles(self.handle, which, _ptr(e_tiles), shots, _ptr(out)))
        return out

    def syndrome_tiles_dev(self, which, e_ptr, shots, s_ptr, stream=0):
        check(self._lib.qcss_syndrome_tiles_dev(self.handle, which, ctypes.c_void_p(e_ptr), shots,
                                                ctypes.c_void_p(s_ptr), ctypes.c_void_p(stream)))

    def sample_syndrome_tiles(self, p, shots, seed=0, first_shot=0, errors=False):
        """Fused sampler + sparse syndromes (qcss_sample_syndrome_tiles): returns (sx, sz) tile arrays, plus
        (ex, ez) when ``errors`` is set."""
        tiles = (shots + 1023) // 1024
        sx = np.zeros((tiles, self.m2, 16), dtype=np.uint64)
        sz = np.zeros((tiles, self.m1, 16), dtype=np.uint64)
        ex = np.zeros((tiles, self.n, 16), dtype=np.uint64) if errors else None
        ez = np.zeros((tiles, self.n, 16), dtype=np.uint64) if errors else None
        if tiles == 0:
            if first_shot % 1024:
                raise ValueError("first_shot must be a multiple of 1024 (one tile)")
            return (sx, sz, ex, ez) if errors else (sx, sz)
        check(self._lib.qcss_sample_syndrome_tiles(self.handle, float(p), int(shots), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                   int(first_shot), _ptr(sx), _ptr(sz), _ptr(ex), _ptr(ez)))
        return (sx, sz, ex, ez) if errors else (sx, sz)

    def sample_syndrome_tiles_dev(self, p, shots, seed, first_shot, sx_ptr, sz_ptr, ex_ptr=0, ez_ptr=0, stream=0):
        check(self._lib.qcss_sample_syndrome_tiles_dev(self.handle, float(p), int(shots), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                       int(first_shot), ctypes.c_void_p(sx_ptr), ctypes.c_void_p(sz_ptr),
                                                       ctypes.c_void_p(ex_ptr), ctypes.c_void_p(ez_ptr),
                                                       ctypes.c_void_p(stream)))

    def syndrome_hist_planes(self, e_planes, shots, which):
        """uint64[2^m] counts of the big-endian syndrome keys of a batch of error planes."""
        e_planes = self._check_planes(e_planes, shots)
        hist = np.zeros(1 << self.m(which), dtype=np.uint64)
        check(self._lib.qcss_syndrome_hist(self.handle, which, _ptr(e_planes), e_planes.shape[1], shots, _ptr(hist)))
        return hist

    def syndrome_hist_dev(self, which, e_ptr, e_stride, shots, hist_ptr, stream=0):
        check(self._lib.qcss_syndrome_hist_dev(self.handle, which, ctypes.c_void_p(e_ptr), e_stride, shots,
                                               ctypes.c_void_p(hist_ptr), ctypes.c_void_p(stream)))

    def decode_planes(self, e_planes, shots, which):
        e_planes = self._check_planes(e_planes, shots)
        self.m(which)
        stride = e_planes.shape[1]
        corr = np.zeros((self.n, stride), dtype=np.uint64)
        flip = np.zeros(stride, dtype=np.uint64)
        miss = np.zeros(stride, dtype=np.uint64)
        tally = Tally()
        check(self._lib.qcss_decode(self.handle, which, _ptr(e_planes), stride, shots,
                                    _ptr(corr), _ptr(flip), _ptr(miss), ctypes.byref(tally)))
        return corr, flip, miss, tally.as_dict()

    def decode_xz_planes(self, ex_planes, ez_planes, shots):
        ex_planes = self._check_planes(ex_planes, shots)
        ez_planes = self._check_planes(ez_planes, shots)
        if ex_planes.shape != ez_planes.shape:
            raise ValueError("x and z planes must have the same stride")
        tally = Tally()
        check(self._lib.qcss_decode_xz(self.handle, _ptr(ex_planes), _ptr(ez_planes),
                                       ex_planes.shape[1], shots, ctypes.byref(tally)))
        return tally.as_dict()

    def last_transfer(self):
        """(bytes sent host -> device, size of the compacting host team) of the last ``qcss_decode_xz`` call."""
        nbytes, threads = ctypes.c_int64(0), ctypes.c_int(0)
        check(self._lib.qcss_code_last_transfer(self.handle, ctypes.byref(nbytes), ctypes.byref(threads)))
        return int(nbytes.value), int(threads.value)

    def decode_xz_host_ptr(self, ex_ptr, ez_ptr, stride, shots):
        """Same call with raw (e.g. pinned) host pointers -- the e2e benchmark path."""
        tally = Tally()
        check(self._lib.qcss_decode_xz(self.handle, ctypes.c_void_p(ex_ptr), ctypes.c_void_p(ez_ptr),
                                       stride, shots, ctypes.byref(tally)))
        return tally.as_dict()

    # ---- the reference's own layout: (shots, n) arrays, transposed on the device ----------------------
    def _shot_rows(self, errors):
        """C-contiguous (shots, n) uint8 or int64 view of ``errors`` (bool is reinterpreted, other dtypes are
        reduced mod 2 into uint8); returns (array, elem_bytes)."""
        errors = np.asarray(errors)
        if errors.ndim != 2 or errors.shape[1] != self.n:
            raise ValueError(f"expected a (shots, {self.n}) array")
        if errors.dtype == np.bool_:
            errors = errors.view(np.uint8)
        elif errors.dtype not in (np.uint8, np.int64):
            errors = np.mod(errors, 2).astype(np.uint8)
        return np.ascontiguousarray(errors), errors.dtype.itemsize

    def syndrome_shots(self, errors, which):
        """(shots, n) errors -> (shots, m) uint8 syndromes (qcss_syndrome_shots)."""
        rows, eb = self._shot_rows(errors)
        out = np.zeros((rows.shape[0], self.m(which)), dtype=np.uint8)
        check(self._lib.qcss_syndrome_shots(self.handle, which, _ptr(rows), eb, rows.shape[0], _ptr(out)))
        return out

    def decode_shots(self, errors, which, corrections=True):
        """(shots, n) errors -> (correction (shots, n) uint8 or None, flip (shots,), miss (shots,), tally)."""
        rows, eb = self._shot_rows(errors)
        self.m(which)
        shots = rows.shape[0]
        corr = np.zeros((shots, self.n), dtype=np.uint8) if corrections else None
        flip = np.zeros(shots, dtype=np.uint8)
        miss = np.zeros(shots, dtype=np.uint8)
        tally = Tally()
        check(self._lib.qcss_decode_shots(self.handle, which, _ptr(rows), eb, shots, _ptr(corr), _ptr(flip), _ptr(miss),
                                          ctypes.byref(tally)))
        return corr, flip, miss, tally.as_dict()

    def decode_xz_shots(self, x_errors, z_errors):
        """Tallies of a shared (shots, n) batch of X and Z errors (qcss_decode_xz_shots)."""
        rx, ebx = self._shot_rows(x_errors)
        rz, ebz = self._shot_rows(z_errors)
        if rx.shape != rz.shape:
            raise ValueError("x and z errors must have the same shape")
        if ebx != ebz:
            rx, rz, ebx = (rx & 1).astype(np.uint8), (rz & 1).astype(np.uint8), 1
        tally = Tally()
        check(self._lib.qcss_decode_xz_shots(self.handle, _ptr(rx), _ptr(rz), ebx, rx.shape[0], ctypes.byref(tally)))
        return tally.as_dict()

    def decode_xz_shots_host_ptr(self, x_ptr, z_ptr, elem_bytes, shots):
        tally = Tally()
        check(self._lib.qcss_decode_xz_shots(self.handle, ctypes.c_void_p(x_ptr), ctypes.c_void_p(z_ptr), elem_bytes,
                                             int(shots), ctypes.byref(tally)))
        return tally.as_dict()

    # ---- sparse batches: events sorted by shot -----------------------------------------------------
    def decode_xz_sparse(self, events, shots):
        """Tallies of a batch given as uint64 events ``shot << 18 | qubit << 2 | pauli`` (qcss_decode_xz_sparse)."""
        events = np.ascontiguousarray(events, dtype=np.uint64)
        tally = Tally()
        check(self._lib.qcss_decode_xz_sparse(self.handle, _ptr(events), events.size, int(shots), ctypes.byref(tally)))
        return tally.as_dict()

    def decode_xz_sparse_host_ptr(self, events_ptr, n_events, shots):
        tally = Tally()
        check(self._lib.qcss_decode_xz_sparse(self.handle, ctypes.c_void_p(events_ptr), int(n_events), int(shots),
                                              ctypes.byref(tally)))
        return tally.as_dict()

    def decode_xz_sparse_dev(self, events_ptr, n_events, shots, tally_ptr, status_ptr=0, stream=0):
        check(self._lib.qcss_decode_xz_sparse_dev(self.handle, ctypes.c_void_p(events_ptr), int(n_events), int(shots),
                                                  ctypes.c_void_p(tally_ptr), ctypes.c_void_p(status_ptr),
                                                  ctypes.c_void_p(stream)))

    def events_from_planes_dev(self, ex_ptr, ez_ptr, e_stride, shots, first_shot, events_ptr, capacity, count_ptr, stream=0):
        check(self._lib.qcss_events_from_planes_dev(self.handle, ctypes.c_void_p(ex_ptr), ctypes.c_void_p(ez_ptr), e_stride,
                                                    int(shots), int(first_shot), ctypes.c_void_p(events_ptr), int(capacity),
                                                    ctypes.c_void_p(count_ptr), ctypes.c_void_p(stream)))

    def mc_run(self, p, shots, seed=0, first_shot=0):
        tally = Tally()
        check(self._lib.qcss_mc_run(self.handle, float(p), int(shots), int(seed), int(first_shot),
                                    ctypes.byref(tally)))
        return tally.as_dict()

    def ec_run(self, p_data, p_ancilla, rounds, shots, seed=0, first_shot=0):
        tally = Tally()
        check(self._lib.qcss_ec_run(self.handle, float(p_data), float(p_ancilla), int(rounds), int(shots),
                                   int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_shot), ctypes.byref(tally)))
        return tally.as_dict()

    def ec_run_dev(self, p_data, p_ancilla, rounds, shots, seed, first_shot, tally_ptr, stream=0):
        check(self._lib.qcss_ec_run_dev(self.handle, float(p_data), float(p_ancilla), int(rounds), int(shots),
                                       int(seed) & 0xFFFFFFFFFFFFFFFF, int(first_shot),
                                       ctypes.c_void_p(tally_ptr), ctypes.c_void_p(stream)))

    def mc_sample(self, p, shots, seed=0, first_shot=0):
        stride = _planes.stride_words(shots)
        ex = np.zeros((self.n, stride), dtype=np.uint64)
        ez = np.zeros((self.n, stride), dtype=np.uint64)
        check(self._lib.qcss_mc_sample(self.handle, float(p), int(shots), int(seed), int(first_shot),
                                       _ptr(ex), _ptr(ez), stride))
        return ex, ez

    # ---- device-pointer entry points (pointers are ints, e.g. torch.Tensor.data_ptr()) -----
    def decode_dev(self, shots, stream=0, **ptrs):
        io = DecodeIO()
        for name, value in ptrs.items():
            setattr(io, name, value)
        check(self._lib.qcss_decode_dev(self.handle, ctypes.byref(io), int(shots), ctypes.c_void_p(stream)))

    def syndrome_dev(self, which, e_ptr, e_stride, shots, s_ptr, s_stride, stream=0):
        check(self._lib.qcss_syndrome_dev(self.handle, which, ctypes.c_void_p(e_ptr), e_stride, int(shots),
                                          ctypes.c_void_p(s_ptr), s_stride, ctypes.c_void_p(stream)))

    def mc_run_dev(self, p, shots, seed, first_shot, tally_ptr, stream=0):
        check(self._lib.qcss_mc_run_dev(self.handle, float(p), int(shots), int(seed), int(first_shot),
                                        ctypes.c_void_p(tally_ptr), ctypes.c_void_p(stream)))

    def mc_sample_dev(self, p, shots, seed, first_shot, ex_ptr, ez_ptr, e_stride, stream=0):
        check(self._lib.qcss_mc_sample_dev(self.handle, float(p), int(shots), int(seed), int(first_shot),
                                           ctypes.c_void_p(ex_ptr), ctypes.c_void_p(ez_ptr), e_stride,
                                           ctypes.c_void_p(stream)))


# ---- GF(2) toolkit (K4) ------------------------------------------------------------------------

def gf2_rref_packed(packed, n):
    """(batch, m, words) uint64 packed rows -> (rref, rank (batch,), pivots (batch, min(m,n)))."""
    lib = load()
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    if packed.ndim != 3:
        raise ValueError("expected (batch, m, words) packed matrices")
    batch, m, words = packed.shape
    if words != (n + 63) // 64:
        raise ValueError("words must be ceil(n / 64)")
    out = np.empty_like(packed)
    rank = np.zeros(batch, dtype=np.int32)
    npiv = min(m, n)
    piv = np.full((batch, max(npiv, 1)), -1, dtype=np.int32)
    check(lib.qcss_gf2_rref(_ptr(packed), batch, m, n, _ptr(out),
                            rank.ctypes.data_as(_c_i32p), piv.ctypes.data_as(_c_i32p)))
    return out, rank, piv[:, :npiv]


def gf2_rref_bits(mats_u8):
    """(batch, m, n) uint8 0/1 -> (rref uint8, rank, pivots) through the packed entry point."""
    batch, m, n = mats_u8.shape
    words = (n + 63) // 64
    padded = np.zeros((batch, m, words * 64), dtype=np.uint8)
    padded[:, :, :n] = mats_u8
    packed = np.packbits(padded, axis=2, bitorder="little").view(np.uint64).reshape(batch, m, words)
    out, rank, piv = gf2_rref_packed(packed, n)
    bits = np.unpackbits(out.view(np.uint8).reshape(batch, m, words * 8), axis=2, bitorder="little")
    return np.ascontiguousarray(bits[:, :, :n]), rank, piv


def pack_bits(bits_u8):
    """(..., n) uint8 0/1 -> (..., ceil(n/64)) uint64, bit j of word w = element 64w + j."""
    bits_u8 = np.asarray(bits_u8, dtype=np.uint8)
    n = bits_u8.shape[-1]
    words = (n + 63) // 64
    padded = np.zeros(bits_u8.shape[:-1] + (words * 64,), dtype=np.uint8)
    padded[..., :n] = bits_u8
    return np.packbits(padded, axis=-1, bitorder="little").view(np.uint64).reshape(bits_u8.shape[:-1] + (words,))


def unpack_bits(packed, n):
    """Inverse of pack_bits."""
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    bits = np.unpackbits(packed.view(np.uint8).reshape(packed.shape[:-1] + (packed.shape[-1] * 8,)), axis=-1,
                         bitorder="little")
    return np.ascontiguousarray(bits[..., :n])


def gf2_nullspace_packed(packed, n, max_basis_rows=None):
    """(batch, m, words) packed matrices -> (basis (batch, rows, words), rank (batch,)): matrix b has
    n - rank[b] basis vectors (free columns in increasing order), zero rows after them."""
    lib = load()
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    batch, m, words = packed.shape
    if words != (n + 63) // 64:
        raise ValueError("words must be ceil(n / 64)")
    rows = n if max_basis_rows is None else int(max_basis_rows)
    basis = np.zeros((batch, rows, words), dtype=np.uint64)
    rank = np.zeros(batch, dtype=np.int32)
    check(lib.qcss_gf2_nullspace(_ptr(packed), batch, m, n, rows, _ptr(basis), rank.ctypes.data_as(_c_i32p)))
    return basis, rank


def gf2_solve_packed(packed, rhs_packed, n):
    """(batch, m, words) matrices and (batch, ceil(m/64)) right-hand sides -> (x (batch, words), ok (batch,))."""
    lib = load()
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    rhs_packed = np.ascontiguousarray(rhs_packed, dtype=np.uint64)
    batch, m, words = packed.shape
    if words != (n + 63) // 64 or rhs_packed.shape != (batch, (m + 63) // 64):
        raise ValueError("bad packed shapes")
    x = np.zeros((batch, words), dtype=np.uint64)
    ok = np.zeros(batch, dtype=np.int32)
    check(lib.qcss_gf2_solve(_ptr(packed), _ptr(rhs_packed), batch, m, n, _ptr(x), ok.ctypes.data_as(_c_i32p)))
    return x, ok


FORM_NOT_CSS, FORM_DEPENDENT_ROWS_C1, FORM_DEPENDENT_ROWS_C2, FORM_FEW_COLUMNS_C1, FORM_FEW_COLUMNS_C2 = 1, 2, 3, 4, 5


def gf2_normalize_packed(packed, n, offset):
    """(batch, m, words) packed matrices -> (normalised, swaps list per matrix, status (batch,)) through
    qcss_gf2_normalize: the standard form of css_code.normalize_parity_check with its column-swap rule."""
    lib = load()
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    if packed.ndim != 3:
        raise ValueError("expected (batch, m, words) packed matrices")
    batch, m, words = packed.shape
    if words != (n + 63) // 64:
        raise ValueError("words must be ceil(n / 64)")
    out = np.empty_like(packed)
    swaps = np.full((batch, max(n, 1), 2), -1, dtype=np.int32)
    counts = np.zeros(batch, dtype=np.int32)
    status = np.zeros(batch, dtype=np.int32)
    check(lib.qcss_gf2_normalize(_ptr(packed), batch, m, n, int(offset), _ptr(out), swaps.ctypes.data_as(_c_i32p),
                                 counts.ctypes.data_as(_c_i32p), status.ctypes.data_as(_c_i32p)))
    lists = [[(int(a), int(b)) for a, b in swaps[i, :counts[i]]] for i in range(batch)]
    return out, lists, status


def css_standard_form_bits(h1_u8, h2_u8):
    """qcss_css_standard_form on 0/1 byte matrices -> (status, H1', H2', swaps)."""
    lib = load()
    h1 = np.ascontiguousarray(h1_u8, dtype=np.uint8)
    h2 = np.ascontiguousarray(h2_u8, dtype=np.uint8)
    (r1, n), (r2, _) = h1.shape, h2.shape
    p1, p2 = pack_bits(h1), pack_bits(h2)
    o1, o2 = np.empty_like(p1), np.empty_like(p2)
    swaps = np.full((max(n, 1), 2), -1, dtype=np.int32)
    count = ctypes.c_int32()
    status = ctypes.c_int32()
    check(lib.qcss_css_standard_form(_ptr(p1), r1, _ptr(p2), r2, n, _ptr(o1), _ptr(o2),
                                     swaps.ctypes.data_as(_c_i32p), ctypes.byref(count), ctypes.byref(status)))
    pairs = [(int(a), int(b)) for a, b in swaps[:count.value]]
    return status.value, unpack_bits(o1, n), unpack_bits(o2, n), pairs


def syndrome_table_arrays(parity_check_u8, max_entries=1 << 28):
    """Device weight-layer search (qcss_table_build): returns (t, keys int64[count], supports uint64[count])
    in the reference's insertion order."""
    lib = load()
    h = np.ascontiguousarray(parity_check_u8, dtype=np.uint8)
    m, n = h.shape
    handle = ctypes.c_void_p()
    t = ctypes.c_int()
    count = ctypes.c_int64()
    check(lib.qcss_table_build(n, m, h.ctypes.data_as(_c_u8p), int(max_entries), ctypes.byref(handle),
                               ctypes.byref(t), ctypes.byref(count)))
    try:
        keys = np.zeros(max(count.value, 1), dtype=np.int64)
        supports = np.zeros(max(count.value, 1), dtype=np.uint64)
        check(lib.qcss_table_read(handle, keys.ctypes.data_as(_c_i64p), _ptr(supports)))
    finally:
        lib.qcss_table_destroy(handle)
    return t.value, keys[: count.value], supports[: count.value]


def host_alloc(nbytes):
    """Pinned host buffer as a uint64 numpy array (freed by host_free(arr))."""
    lib = load()
    ptr = ctypes.c_void_p()
    check(lib.qcss_host_alloc(ctypes.byref(ptr), nbytes))
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=np.uint64)
    arr.flags.writeable = True
    return arr, ptr.value


def host_free(ptr_value):
    check(load().qcss_host_free(ctypes.c_void_p(ptr_value)))
