"""Null-space probe: RREF + rank + basis of 4096 x (1024 x 2048) (qcss_gf2_nullspace_dev), few iterations,
suitable for an ncu launch list.  python tools/ns_probe.py [batch]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import _native                       # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m, n, words = 1024, 2048, 32
rows = n - m + 8
lib = _native.load()
mats = torch.randint(-2**62, 2**62, (batch, m, words), dtype=torch.int64, device="cuda")
basis = torch.empty((batch, rows, words), dtype=torch.int64, device="cuda")
rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream


def run():
    _native.check(lib.qcss_gf2_nullspace_dev(mats.data_ptr(), batch, m, n, rows, basis.data_ptr(), rank.data_ptr(),
                                             ovf.data_ptr(), stream))


run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    run()
b.record()
torch.cuda.synchronize()
print(json.dumps(dict(probe="gf2_nullspace", batch=batch, ms=a.elapsed_time(b) / 3, overflow=int(ovf.item()))))
