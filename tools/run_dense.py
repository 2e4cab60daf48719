"""Small K1' (tcgen05 dense syndrome) run for profiling: python tools/run_dense.py [shots]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import SyndromeCode, _native
_native.set_option("dense", 1)
shots = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 19
m, n = 1024, 2048
h = np.random.default_rng(5).integers(0, 2, size=(m, n))
dev = SyndromeCode(h, h[:8]).device
stride = ((shots + 127) // 128) * 2
e = torch.randint(-2**62, 2**62, (n, stride), dtype=torch.int64, device="cuda")
s = torch.empty((m, stride), dtype=torch.int64, device="cuda")
for _ in range(3):
    dev.syndrome_dev(1, e.data_ptr(), stride, shots, s.data_ptr(), stride, 0)
torch.cuda.synchronize()
print(dev.kernel_name(), int(s.sum().item()))
