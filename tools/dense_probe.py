"""Time the dense tcgen05 syndrome kernel .
    python tools/dense_probe.py [shots]"""
import os, sys
import numpy as np
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from quantum_css_codes_b200 import SyndromeCode
shots = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 21
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(5)
h = rng.integers(0, 2, size=(1024, n), dtype=np.uint8)
dev = SyndromeCode(h, h).device
stride = ((shots + 127) // 128) * 2
e = torch.randint(-2**62, 2**62, (n, stride), dtype=torch.int64, device="cuda")
s = torch.empty((1024, stride), dtype=torch.int64, device="cuda")
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
print(dev.kernel_name(), "ms", round(ms, 3), "n", n, "POPS", round(2 * 1024 * n * shots / ms / 1e12, 3),
      "checksum", int(s.sum()))
