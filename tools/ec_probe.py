"""Throughput of the repeated-EC Pauli-frame kernel (qcss_ec_run_dev) and of the standard-form kernels.
Run on the GPU box: python tools/ec_probe.py > gpurun_out/probe_ec.jsonl"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import css_code                                                    # noqa: E402
from quantum_css_codes_b200 import _native, codes                 # noqa: E402


def time_ec(name, p, q, rounds, shots, reps=5):
    code = css_code.CSSCode(*[np.array(h) for h in getattr(codes, name)()])
    tally = torch.zeros(6, dtype=torch.int64, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    dev = code.device
    for _ in range(2):
        dev.ec_run_dev(p, q, rounds, shots, 1, 0, tally.data_ptr(), stream)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    start.record()
    for i in range(reps):
        dev.ec_run_dev(p, q, rounds, shots, 2 + i, 0, tally.data_ptr(), stream)
    stop.record()
    torch.cuda.synchronize()
    ms = start.elapsed_time(stop) / reps
    print(json.dumps(dict(probe="ec_rounds", code=name, p=p, q=q, rounds=rounds, shots=shots, ms=ms,
                          shot_rounds_per_s=shots * rounds / (ms * 1e-3), kernel=dev.kernel_name())), flush=True)


def time_normalize(r, n, offset, batch):
    rng = np.random.default_rng(r + n)
    mats = rng.integers(0, 2, size=(batch, r, n), dtype=np.uint8)
    mats[:, :, offset:offset + r] |= np.eye(r, dtype=np.uint8)     # mostly independent rows; status is reported
    packed = _native.pack_bits(mats)
    _native.gf2_normalize_packed(packed[:1], n, offset)
    t0 = time.perf_counter()
    out, swaps, status = _native.gf2_normalize_packed(packed, n, offset)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(probe="gf2_normalize", r=r, n=n, offset=offset, batch=batch, wall_ms=dt * 1e3,
                          ok=int((status == 0).sum()), swaps=int(sum(len(s) for s in swaps)))), flush=True)


if __name__ == "__main__":
    for name in ("steane", "qrm15", "golay23"):
        for p, q in ((1e-3, 1e-3), (0.05, 0.05)):
            time_ec(name, p, q, 10, 10**8 if p < 0.01 else 10**7)
    time_ec("steane", 1e-3, 1e-3, 100, 10**8)
    time_normalize(768, 1600, 0, 1)
    time_normalize(768, 1600, 0, 148)
    time_normalize(1024, 2048, 512, 1)
    time_normalize(1024, 2048, 512, 16)
    time_normalize(24, 60, 0, 4096)
