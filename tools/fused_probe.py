"""Fused sampler + syndromes on HGP-1600 only (for ncu).  python tools/fused_probe.py [shots] [p]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import SyndromeCode, codes
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
p = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
hx, hz = codes.hgp1600()
dev = SyndromeCode(hx, hz).device
tiles = (shots + 1023) // 1024
sx = torch.empty((tiles, 768, 16), dtype=torch.int64, device="cuda")
sz = torch.empty_like(sx)
st = torch.cuda.current_stream().cuda_stream
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dev.sample_syndrome_tiles_dev(p, shots, 7, 0, sx.data_ptr(), sz.data_ptr(), 0, 0, st)
torch.cuda.synchronize()
a.record()
for i in range(3):
    dev.sample_syndrome_tiles_dev(p, shots, 8 + i, 0, sx.data_ptr(), sz.data_ptr(), 0, 0, st)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(json.dumps(dict(probe="hgp_fused_sampler", p=p, shots=shots, ms=ms, shots_per_s=shots / ms * 1e3)))
