"""Test helper: builds and drives tests/hostemu (host emulation of the kernels' per-thread code).
Test infrastructure only."""

import ctypes
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "hostemu.cu")
LIB = os.path.join(HERE, "hostemu", "libqcss_hostemu.so")
CSRC = os.path.join(os.path.dirname(HERE), "quantum_css_codes_b200", "csrc")

MAX_N, MAX_M = 32, 16


class GenericSide(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int32), ("m", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("has_miss", ctypes.c_int32), ("tt_flip", ctypes.c_uint32), ("tt_miss", ctypes.c_uint32),
                ("lexp", ctypes.c_uint32 * MAX_N), ("mask", (ctypes.c_uint32 * MAX_N) * MAX_M),
                ("tt_corr", ctypes.c_uint32 * MAX_N),
                ("lut_fm", ctypes.c_void_p), ("lut_corr", ctypes.c_void_p), ("lut_e32", ctypes.c_void_p)]


class DecodeIO(ctypes.Structure):
    _fields_ = [("ex", ctypes.c_void_p), ("ez", ctypes.c_void_p), ("e_stride", ctypes.c_int64),
                ("synd_x", ctypes.c_void_p), ("synd_z", ctypes.c_void_p), ("s_stride", ctypes.c_int64),
                ("corr_x", ctypes.c_void_p), ("corr_z", ctypes.c_void_p), ("c_stride", ctypes.c_int64),
                ("flip_x", ctypes.c_void_p), ("flip_z", ctypes.c_void_p),
                ("miss_x", ctypes.c_void_p), ("miss_z", ctypes.c_void_p), ("tally", ctypes.c_void_p),
                ("words", ctypes.c_int64), ("tail_mask", ctypes.c_uint32), ("sides", ctypes.c_int32),
                ("ex_out", ctypes.c_void_p), ("ez_out", ctypes.c_void_p),
                ("seed", ctypes.c_uint64), ("first_word", ctypes.c_uint64), ("thr", ctypes.c_uint32),
                ("use_gap", ctypes.c_uint32), ("gap_cdf", ctypes.c_uint32 * 32), ("gap_inv", ctypes.c_uint32)]


class GapTable(ctypes.Structure):
    _fields_ = [("cdf", ctypes.c_uint32 * 32), ("inv", ctypes.c_uint32)]


class EcParams(ctypes.Structure):
    _fields_ = [("tally", ctypes.c_void_p), ("words", ctypes.c_int64), ("tail_mask", ctypes.c_uint32),
                ("rounds", ctypes.c_int32), ("seed", ctypes.c_uint64), ("first_word", ctypes.c_uint64),
                ("thr_p", ctypes.c_uint32), ("thr_q", ctypes.c_uint32), ("gap_p", ctypes.c_uint32),
                ("gap_q", ctypes.c_uint32), ("tab_p", GapTable), ("tab_q", GapTable)]


def _needs_build():
    if not os.path.exists(LIB):
        return True
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("core.cuh", "decode.cuh", "ec_rounds.cuh", "named_codes.inc",
                                                     "host_compact.h", "host_compact.cpp")]
    return os.path.getmtime(LIB) < max(os.path.getmtime(p) for p in deps)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if _needs_build():
            nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
            cmd = [nvcc, "-O1", "-std=c++17", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
                   "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, SRC,
                   os.path.join(CSRC, "host_compact.cpp")]
            subprocess.run(cmd, check=True, capture_output=True)
        _lib = ctypes.CDLL(LIB)
        assert _lib.emu_sizeof_side() == ctypes.sizeof(GenericSide)
        assert _lib.emu_sizeof_io() == ctypes.sizeof(DecodeIO)
    return _lib


class Side:
    """Python mirror of api.cu::build_side (flattening of H, L and the syndrome table)."""

    def __init__(self, h, lrow, table):
        h = np.asarray(h) & 1
        m, n = h.shape
        s = GenericSide()
        s.n, s.m = n, m
        lmask = 0
        for t in range(m):
            for j in range(n):
                if h[m - 1 - t, j]:
                    s.mask[t][j] = 0xFFFFFFFF
        if lrow is not None:
            for j in range(n):
                if int(lrow[j]) & 1:
                    s.lexp[j] = 0xFFFFFFFF
                    lmask |= 1 << j
        self.fm = np.full(1 << m, 2, dtype=np.uint8)
        self.co = np.zeros(1 << m, dtype=np.uint32)
        s.mode = 0
        if table is not None:
            for key, vec in table.items():
                cm = int(sum(int(b) << j for j, b in enumerate(vec)))
                self.co[int(key)] = cm
                self.fm[int(key)] = bin(cm & lmask).count("1") & 1
            s.has_miss = int(bool((self.fm & 2).any()))
            s.mode = 1 if m <= 5 else 2
            if m <= 5:
                for k in range(1 << m):
                    s.tt_flip |= int(self.fm[k] & 1) << k
                    s.tt_miss |= int((self.fm[k] >> 1) & 1) << k
                    for j in range(n):
                        s.tt_corr[j] |= ((int(self.co[k]) >> j) & 1) << k
            s.lut_fm = self.fm.ctypes.data
            s.lut_corr = self.co.ctypes.data
            if 5 < m <= 13:
                self.e32 = (np.where(self.fm & 1, 0x0000FFFF, 0) | np.where(self.fm & 2, 0xFFFF0000, 0)).astype(np.uint32)
                s.lut_e32 = self.e32.ctypes.data
        self.c = s
        self.n, self.m = n, m


def decode(side_x, side_z, ex_planes=None, ez_planes=None, shots=0, named_id=-1, want=(),
           sample=None, fast=False):
    """Run the emulated kernel.  ex/ez planes: (n, stride) uint64.  want: subset of
    {'synd', 'corr', 'flip', 'miss'}.  sample: dict(seed, first_shot, p_thr) for the fused sampler.
    Returns dict with tally and requested planes."""
    L = lib()
    n = side_x.n
    io = DecodeIO()
    out = {}
    keep = []
    if sample is None:
        stride = (ex_planes if ex_planes is not None else ez_planes).shape[1]
    else:
        stride = max(2, ((shots + 127) // 128) * 2)
    io.e_stride = io.s_stride = io.c_stride = stride * 2
    io.words = (shots + 31) // 32
    io.tail_mask = (1 << (shots % 32)) - 1 if shots % 32 else 0xFFFFFFFF
    if sample is None:
        if ex_planes is not None:
            ex_planes = np.ascontiguousarray(ex_planes, dtype=np.uint64); keep.append(ex_planes)
            io.ex = ex_planes.ctypes.data
            io.sides |= 1
        if ez_planes is not None:
            ez_planes = np.ascontiguousarray(ez_planes, dtype=np.uint64); keep.append(ez_planes)
            io.ez = ez_planes.ctypes.data
            io.sides |= 2
    else:
        io.sides = 3
        io.seed, io.first_word, io.thr = sample["seed"], sample["first_shot"] // 32, sample["thr"]
        if "p" in sample:                                   # gap sampler for p < 1/64, as api.cu chooses
            from oracle import philox as _ophilox
            cdf, inv = _ophilox.gap_table(sample["p"])
            io.use_gap = 1 if _ophilox.uses_gap_sampler(sample["p"]) else 0
            for k in range(32):
                io.gap_cdf[k] = int(cdf[k])
            io.gap_inv = inv
        out["ex"] = np.zeros((n, stride), dtype=np.uint64)
        out["ez"] = np.zeros((n, stride), dtype=np.uint64)
        io.ex_out, io.ez_out = out["ex"].ctypes.data, out["ez"].ctypes.data
    do_x, do_z = bool(io.sides & 1), bool(io.sides & 2)
    for tag, side, active in (("x", side_x, do_x), ("z", side_z, do_z)):
        if not active:
            continue
        if "synd" in want:
            out["synd_" + tag] = np.zeros((side.m, stride), dtype=np.uint64)
            setattr(io, "synd_" + tag, out["synd_" + tag].ctypes.data)
        if "corr" in want:
            out["corr_" + tag] = np.zeros((n, stride), dtype=np.uint64)
            setattr(io, "corr_" + tag, out["corr_" + tag].ctypes.data)
        for name in ("flip", "miss"):
            if name in want:
                out[f"{name}_{tag}"] = np.zeros(stride, dtype=np.uint64)
                setattr(io, f"{name}_{tag}", out[f"{name}_{tag}"].ctypes.data)
    tally = np.zeros(6, dtype=np.uint64)
    if fast:
        assert shots % 128 == 0 and not want and io.sides == 3
        io.ex_out = io.ez_out = None
    L.emu_set_fast(int(fast))
    rc = L.emu_decode(ctypes.byref(side_x.c), ctypes.byref(side_z.c), ctypes.byref(io), named_id,
                      int(sample is not None), tally.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    assert rc == 0
    out["tally"] = dict(shots=shots, fail_x=int(tally[1]), fail_z=int(tally[2]), fail_any=int(tally[3]),
                        miss_x=int(tally[4]), miss_z=int(tally[5]))
    return out


def ec_run(side_x, side_z, p_data, p_ancilla, rounds, shots, seed=0, first_shot=0, named_id=-1, queue_form=False):
    """Host emulation of qcss_ec_run (api.cu::launch_ec + ec_kernels.cu): returns the tally dict.
    ``queue_form``: replay the CTA-wide two-phase kernel (k_ec_named_q; static descriptors, both rates < 1/64)."""
    from oracle import philox as _ophilox
    L = lib()
    assert L.emu_sizeof_ec() == ctypes.sizeof(EcParams)
    ec = EcParams()
    ec.words = (shots + 31) // 32
    ec.tail_mask = (1 << (shots % 32)) - 1 if shots % 32 else 0xFFFFFFFF
    ec.rounds, ec.seed, ec.first_word = rounds, seed, first_shot // 32
    for tag, p in (("p", p_data), ("q", p_ancilla)):
        setattr(ec, "thr_" + tag, _ophilox.threshold(p))
        setattr(ec, "gap_" + tag, 1 if _ophilox.uses_gap_sampler(p) else 0)
        cdf, inv = _ophilox.gap_table(p)
        tab = getattr(ec, "tab_" + tag)
        for k in range(32):
            tab.cdf[k] = int(cdf[k])
        tab.inv = inv
    tally = np.zeros(6, dtype=np.uint64)
    fn = L.emu_ecq if queue_form else L.emu_ec
    rc = fn(ctypes.byref(side_x.c), ctypes.byref(side_z.c), ctypes.byref(ec), named_id,
            tally.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    assert rc == 0
    return dict(shots=shots, fail_x=int(tally[1]), fail_z=int(tally[2]), fail_any=int(tally[3]),
                miss_x=int(tally[4]), miss_z=int(tally[5]))


def mc_gapq(side_x, side_z, p, shots, seed=0, first_shot=0, named_id=-1):
    """Host replay of the CTA-wide two-phase gap sampler (small_common.cuh::run_small_gapq); shots must be a
    multiple of 128.  Returns the tally dict."""
    from oracle import philox as _ophilox
    assert shots % 128 == 0 and _ophilox.uses_gap_sampler(p)
    L = lib()
    io = DecodeIO()
    io.words = shots // 32
    io.tail_mask = 0xFFFFFFFF
    io.sides = 3
    io.seed, io.first_word, io.thr, io.use_gap = seed, first_shot // 32, _ophilox.threshold(p), 1
    cdf, inv = _ophilox.gap_table(p)
    for k in range(32):
        io.gap_cdf[k] = int(cdf[k])
    io.gap_inv = inv
    tally = np.zeros(6, dtype=np.uint64)
    rc = L.emu_mc_gapq(ctypes.byref(side_x.c), ctypes.byref(side_z.c), ctypes.byref(io), named_id,
                       tally.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
    assert rc == 0
    return dict(shots=shots, fail_x=int(tally[1]), fail_z=int(tally[2]), fail_any=int(tally[3]),
                miss_x=int(tally[4]), miss_z=int(tally[5]))


def zs_roundtrip(ex, ez, w0, cw, threads):
    """Host side of the compacting qcss_decode_xz path (csrc/host_compact.cpp: the product's own code) followed by a
    host restatement of k_zs_expand's index arithmetic: returns the rebuilt (n, cw) chunks of both plane sets, or None
    when a worker's share of non-zero words exceeds its region (the product then sends the chunk uncompacted)."""
    ex = np.ascontiguousarray(ex, dtype=np.uint64)
    ez = np.ascontiguousarray(ez, dtype=np.uint64)
    n, stride = ex.shape
    assert ez.shape == ex.shape and w0 + cw <= stride
    out_x = np.full((n, cw), 0xDEADBEEF, dtype=np.uint64)
    out_z = np.full((n, cw), 0xDEADBEEF, dtype=np.uint64)
    fn = lib().emu_zs_roundtrip
    fn.restype = ctypes.c_int
    rc = fn(ctypes.c_void_p(ex.ctypes.data), ctypes.c_void_p(ez.ctypes.data), ctypes.c_int64(stride), n,
            ctypes.c_int64(w0), ctypes.c_int64(cw), threads, ctypes.c_void_p(out_x.ctypes.data),
            ctypes.c_void_p(out_z.ctypes.data))
    if rc == 1:
        return None
    assert rc == 0
    return out_x, out_z
