"""A few launches of the HGP-1600 syndrome kernel (for ncu captures and quick timing).
    python tools/hgp_probe.py [shots]"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from quantum_css_codes_b200 import SyndromeCode, codes
shots = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 24
hx, hz = codes.hgp1600()
dev = SyndromeCode(hx, hz).device
stride = ((shots + 127) // 128) * 2
e = torch.randint(-2**62, 2**62, (1600, stride), dtype=torch.int64, device="cuda")
s = torch.empty((768, stride), dtype=torch.int64, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    dev.syndrome_dev(2, e.data_ptr(), stride, shots, s.data_ptr(), stride, st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    dev.syndrome_dev(2, e.data_ptr(), stride, shots, s.data_ptr(), stride, st)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
print(dev.kernel_name(), "ms", ms, "GB/s", 296 * shots / ms / 1e6, "checksum", int(s.sum()))
