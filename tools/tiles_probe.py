"""HGP-1600 syndromes: plane-major ring vs tile-major ring (qcss_syndrome_tiles_dev).
python tools/tiles_probe.py [shots] > gpurun_out/probe_tiles.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import SyndromeCode, codes            # noqa: E402

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
hx, hz = codes.hgp1600()
dev = SyndromeCode(hx, hz).device
n, m = 1600, 768
tiles = (shots + 1023) // 1024
gen = torch.Generator(device="cuda").manual_seed(1)
e = torch.randint(-2**62, 2**62, (tiles, n, 16), dtype=torch.int64, device="cuda", generator=gen)
s = torch.empty((tiles, m, 16), dtype=torch.int64, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
bytes_per_shot = (n + m) / 8


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for which in (1, 2):
    ms = timed(lambda: dev.syndrome_tiles_dev(which, e.data_ptr(), shots, s.data_ptr(), stream))
    print(json.dumps(dict(probe="hgp_tiles", which=which, shots=shots, ms=ms, shots_per_s=shots / (ms * 1e-3),
                          gbs=bytes_per_shot * shots / (ms * 1e-3) / 1e9)), flush=True)
# plane-major on the same bits: e viewed as planes needs a transpose; time it on fresh random planes instead
stride = tiles * 16
ep = e.view(-1)[: n * stride].view(n, stride)
sp = s.view(-1)[: m * stride].view(m, stride)
for which in (1, 2):
    ms = timed(lambda: dev.syndrome_dev(which, ep.data_ptr(), stride, shots, sp.data_ptr(), stride, stream))
    print(json.dumps(dict(probe="hgp_planes", which=which, shots=shots, ms=ms, shots_per_s=shots / (ms * 1e-3),
                          gbs=bytes_per_shot * shots / (ms * 1e-3) / 1e9)), flush=True)

# fused sampler + syndromes of both Pauli types (no HBM input)
sx = torch.empty((tiles, m, 16), dtype=torch.int64, device="cuda")
for p in (1e-3, 0.05):
    ms = timed(lambda: dev.sample_syndrome_tiles_dev(p, shots, 7, 0, sx.data_ptr(), s.data_ptr(), 0, 0, stream), reps=3)
    print(json.dumps(dict(probe="hgp_fused_sampler", p=p, shots=shots, ms=ms, shots_per_s=shots / (ms * 1e-3),
                          site_words_per_s=shots / 32 * n / (ms * 1e-3))), flush=True)
