"""Compare the CTA-pair dense kernel (option dense_pair = 1) with the single-SM kernel: equality and time.
   python tools/dense_pair_probe.py [shots] [n] [m]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from quantum_css_codes_b200 import SyndromeCode, _native
shots = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 21
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
m = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
rng = np.random.default_rng(5)
h = rng.integers(0, 2, size=(m, n), dtype=np.uint8)
_native.set_option("dense", 1)
dev = SyndromeCode(h, h).device
stride = ((shots + 127) // 128) * 2
e = torch.randint(-2**31, 2**31, (n, stride * 2), dtype=torch.int32, device="cuda").view(torch.int64)
st = torch.cuda.current_stream().cuda_stream
res = {}
for pair in (0, 1):
    _native.set_option("dense_pair", pair)
    s = torch.zeros((m, stride), dtype=torch.int64, device="cuda")
    for _ in range(2):
        dev.syndrome_dev(2, e.data_ptr(), stride, shots, s.data_ptr(), stride, st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        dev.syndrome_dev(2, e.data_ptr(), stride, shots, s.data_ptr(), stride, st)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    res[pair] = s
    print("pair" if pair else "single", "ms", round(ms, 3), "POPS", round(2 * m * n * shots / ms / 1e12, 3), flush=True)
print("equal", bool(torch.equal(res[0], res[1])))
