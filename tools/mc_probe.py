"""A few launches of the fused Philox sampler + decode kernel (for ncu captures).
    python tools/mc_probe.py [code] [shots] [p] [library]"""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from quantum_css_codes_b200 import CSSCode, codes, _native
if len(sys.argv) > 4: _native.LIB_PATH = os.path.abspath(sys.argv[4])     # A/B: another build of the library
name = sys.argv[1] if len(sys.argv) > 1 else "steane"
shots = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1 << 30
p = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
code = CSSCode(*[np.array(h) for h in getattr(codes, name)()])
dev = code.device
tally = torch.zeros(6, dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    dev.mc_run_dev(p, shots, 7, 0, tally.data_ptr(), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
dev.mc_run_dev(p, shots, 7, 0, tally.data_ptr(), st)
b.record(); torch.cuda.synchronize()
print(dev.kernel_name(), "ms", a.elapsed_time(b), "shots/s", shots / a.elapsed_time(b) * 1e3, tally.tolist())
