"""K4 timing over widths at 1024 rows: separates discovery from replay cost per block.
   python tools/gf2_shapes.py [batch] [gf2_kernel option] [library]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import _native
if len(sys.argv) > 3: _native.LIB_PATH = os.path.abspath(sys.argv[3])     # A/B: another build of the library
lib = _native.load()
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
if len(sys.argv) > 2: _native.check(lib.qcss_set_option(b"gf2_kernel", int(sys.argv[2])))
for m, n in ((1024, 1024), (1024, 1280), (1024, 1536), (1024, 2048), (1024, 2560), (1024, 3072), (1024, 4096), (512, 1024), (640, 1280), (768, 1600), (768, 2048), (896, 1792)):
    mats = torch.randint(-2**31, 2**31, (batch, m, n // 32), dtype=torch.int32, device="cuda").view(torch.int64)
    out = torch.empty_like(mats)
    rank = torch.zeros(batch, dtype=torch.int32, device="cuda")
    piv = torch.zeros((batch, m), dtype=torch.int32, device="cuda")
    run = lambda: _native.check(lib.qcss_gf2_rref_dev(mats.data_ptr(), batch, m, n, out.data_ptr(), rank.data_ptr(), piv.data_ptr(), 0))
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(json.dumps(dict(knob=int(sys.argv[2]) if len(sys.argv) > 2 else 0, m=m, n=n, batch=batch, ms=best, us_per_matrix_per_sm_slot=best * 1e3 * 296 / batch, full_rank=int((rank == m).sum()))))
