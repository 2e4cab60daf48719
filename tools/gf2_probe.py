"""One launch of the batched GF(2) RREF kernel (for ncu captures).
    python tools/gf2_probe.py [batch] [m] [n]"""
import os, sys
import torch
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from quantum_css_codes_b200 import _native
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 148
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
lib = _native.load()
torch.manual_seed(5)
mats = torch.randint(-2**62, 2**62, (batch, m, n // 64), dtype=torch.int64, device="cuda")
out = torch.empty_like(mats)
rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
for _ in range(2):
    _native.check(lib.qcss_gf2_rref_dev(mats.data_ptr(), batch, m, n, out.data_ptr(), rank.data_ptr(), 0, 0))
torch.cuda.synchronize()
print("rank min", int(rank.min()), "checksum", int(out.sum()))
