"""
Event-timed probes of every kernel family on one GPU; one JSON line per probe.

    python tools/perf_probe.py [--quick] > gpurun_out/probe.jsonl

Not the benchmark of record (bench.py is); used to fill DESIGN.md / profiles with per-config
throughput and roofline fractions.
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from quantum_css_codes_b200 import CSSCode, SyndromeCode, codes, _native   # noqa: E402

HBM_PEAK = 6549.1
try:
    HBM_PEAK = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    return float(np.median(times)), float(np.min(times))


def emit(**kw):
    print(json.dumps(kw), flush=True)


def probe_decode(name, shots, p=1e-3):
    code = CSSCode(*[np.array(h) for h in getattr(codes, name)()])
    dev = code.device
    n = code.n
    stride = ((shots + 127) // 128) * 2
    stream = torch.cuda.current_stream().cuda_stream
    ex = torch.empty((n, stride), dtype=torch.int64, device="cuda")
    ez = torch.empty((n, stride), dtype=torch.int64, device="cuda")
    tally = torch.zeros(6, dtype=torch.int64, device="cuda")
    dev.mc_sample_dev(p, shots, 1, 0, ex.data_ptr(), ez.data_ptr(), stride, stream)
    med, best = timed(lambda: dev.decode_dev(shots, stream, ex=ex.data_ptr(), ez=ez.data_ptr(),
                                             e_stride=stride, tally=tally.data_ptr()))
    gbs = 2 * n / 8 * shots / (med / 1e3) / 1e9
    emit(probe="decode_resident", code=name, kernel=dev.kernel_name(), shots=shots, p=p, ms=med, ms_best=best,
         shots_per_s=shots / (med / 1e3), gbs=gbs, hbm_frac=gbs / HBM_PEAK)
    # fused sampler
    tally.zero_()
    med, best = timed(lambda: dev.mc_run_dev(p, shots, 7, 0, tally.data_ptr(), stream))
    emit(probe="mc_fused", code=name, kernel=dev.kernel_name(), shots=shots, p=p, ms=med, ms_best=best,
         shots_per_s=shots / (med / 1e3))
    del ex, ez
    torch.cuda.empty_cache()


def probe_specialized(shots, p=1e-3):
    """A code the library has no built-in descriptor for (Shor [[9,1,3]]): generic kernels vs kernels
    compiled for the code (CSSCode.specialize)."""
    hx, hz = codes.shor9()
    for jit in (False, True):
        code = CSSCode(hx, hz)
        if jit:
            code.specialize()
        dev = code.device
        n = code.n
        stride = ((shots + 127) // 128) * 2
        stream = torch.cuda.current_stream().cuda_stream
        ex = torch.empty((n, stride), dtype=torch.int64, device="cuda")
        ez = torch.empty((n, stride), dtype=torch.int64, device="cuda")
        tally = torch.zeros(6, dtype=torch.int64, device="cuda")
        dev.mc_sample_dev(p, shots, 1, 0, ex.data_ptr(), ez.data_ptr(), stride, stream)
        med, best = timed(lambda: dev.decode_dev(shots, stream, ex=ex.data_ptr(), ez=ez.data_ptr(), e_stride=stride,
                                                 tally=tally.data_ptr()))
        gbs = 2 * n / 8 * shots / (med / 1e3) / 1e9
        emit(probe="decode_resident", code="shor9", kernel=dev.kernel_name(), shots=shots, p=p, ms=med, ms_best=best,
             shots_per_s=shots / (med / 1e3), gbs=gbs, hbm_frac=gbs / HBM_PEAK)
        del ex, ez
        torch.cuda.empty_cache()


def probe_hgp(shots):
    hx, hz = codes.hgp1600()
    code = SyndromeCode(hx, hz)
    dev = code.device
    n, m = 1600, 768
    stride = ((shots + 127) // 128) * 2
    stream = torch.cuda.current_stream().cuda_stream
    e = torch.randint(-2**62, 2**62, (n, stride), dtype=torch.int64, device="cuda")
    s = torch.empty((m, stride), dtype=torch.int64, device="cuda")
    for which in (1, 2):
        med, best = timed(lambda: dev.syndrome_dev(which, e.data_ptr(), stride, shots, s.data_ptr(), stride, stream))
        gbs = (n + m) / 8 * shots / (med / 1e3) / 1e9
        emit(probe="hgp_syndrome", which=which, kernel=dev.kernel_name(), shots=shots, ms=med, ms_best=best,
             shots_per_s_one_type=shots / (med / 1e3), gbs=gbs, hbm_frac=gbs / HBM_PEAK)
    del e, s
    torch.cuda.empty_cache()


def probe_dense(shots, m=1024, n=2048):
    """C4-dense: H = a random dense 1024 x 2048 matrix; tcgen05 int8 MMA vs the bit-sliced kernel."""
    rng = np.random.default_rng(5)
    h = rng.integers(0, 2, size=(m, n))
    stride = ((shots + 127) // 128) * 2
    stream = torch.cuda.current_stream().cuda_stream
    e = torch.randint(-2**62, 2**62, (n, stride), dtype=torch.int64, device="cuda")
    s = torch.empty((m, stride), dtype=torch.int64, device="cuda")
    ref = None
    for force in ("1", "0"):
        _native.set_option("dense", int(force))
        dev = SyndromeCode(h, h[:8]).device
        n_shots = shots if force == "1" else min(shots, 1 << 18)
        med, best = timed(lambda: dev.syndrome_dev(1, e.data_ptr(), stride, n_shots, s.data_ptr(), stride, stream),
                          warmup=1, iters=3)
        chk = int(s[:, : n_shots // 64].sum().item())
        if ref is None:
            ref = s[:, : (1 << 18) // 64].clone()
            same = None
        else:
            same = bool(torch.equal(ref[:, : n_shots // 64], s[:, : n_shots // 64]))
        emit(probe="dense_syndrome", kernel=dev.kernel_name(), m=m, n=n, shots=n_shots, ms=med, ms_best=best,
             shots_per_s=n_shots / (med / 1e3), int_ops_per_s=2.0 * m * n * n_shots / (med / 1e3),
             matches_other_path=same, checksum=chk)
    _native.set_option("dense", -1)


def probe_gf2(batch, m=1024, n=2048):
    lib = _native.load()
    words = n // 64
    mats = torch.randint(-2**62, 2**62, (batch, m, words), dtype=torch.int64, device="cuda")
    out = torch.empty_like(mats)
    rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
    piv = torch.zeros((batch, min(m, n)), dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def run():
        _native.check(lib.qcss_gf2_rref_dev(mats.data_ptr(), batch, m, n, out.data_ptr(), rank.data_ptr(),
                                            piv.data_ptr(), stream))
    med, best = timed(run, warmup=1, iters=3)
    word_ops = 2.52e7 * batch * (m / 1024) ** 2 * (n / 2048)
    emit(probe="gf2_rref", batch=batch, m=m, n=n, ms=med, ms_best=best, matrices_per_s=batch / (med / 1e3),
         xor_word_ops_per_s=word_ops / (med / 1e3), full_rank=int((rank == min(m, n)).sum().item()))


def probe_gf2_nullspace(batch, m=1024, n=2048):
    lib = _native.load()
    words = n // 64
    rows = n - m + 8
    mats = torch.randint(-2**62, 2**62, (batch, m, words), dtype=torch.int64, device="cuda")
    basis = torch.empty((batch, rows, words), dtype=torch.int64, device="cuda")
    rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
    ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def run():
        _native.check(lib.qcss_gf2_nullspace_dev(mats.data_ptr(), batch, m, n, rows, basis.data_ptr(), rank.data_ptr(),
                                                 ovf.data_ptr(), stream))
    med, best = timed(run, warmup=1, iters=3)
    emit(probe="gf2_nullspace", batch=batch, m=m, n=n, ms=med, ms_best=best, matrices_per_s=batch / (med / 1e3),
         includes="RREF + rank + basis (n - m + 8 rows per matrix)", overflow=int(ovf.item()),
         full_rank=int((rank == min(m, n)).sum().item()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    emit(probe="env", gpu=torch.cuda.get_device_name(0), hbm_peak_gbs=HBM_PEAK,
         disable_named=bool(os.environ.get("PROBE_DISABLE_NAMED")))
    big = 1 << 28 if args.quick else 1_000_000_000
    only = set(args.only.split(",")) if args.only else None
    if only is None or "decode" in only:
        for name in ("steane", "qrm15", "golay23"):
            probe_decode(name, big)
        for name in ("qrm15", "golay23"):
            probe_decode(name, big // 4, p=0.05)
    if only is None or "jit" in only:
        probe_specialized(big)
    if only is None or "hgp" in only:
        probe_hgp(1 << 22 if args.quick else 100_000_000)
    if only is None or "dense" in only:
        probe_dense(1 << 18 if args.quick else 1 << 21)
    if only is None or "gf2" in only:
        probe_gf2(64 if args.quick else 4096)
        probe_gf2(4096, 256, 512)
        probe_gf2(2048, 768, 1600)
        probe_gf2_nullspace(64 if args.quick else 4096)


if __name__ == "__main__":
    main()
