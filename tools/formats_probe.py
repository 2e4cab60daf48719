"""One call each of the host-format paths on a 2^28-shot Steane batch (for ncu and quick timing)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import CSSCode, codes, _native
code = CSSCode(*[np.array(h) for h in codes.steane()])
dev, n, shots = code.device, code.n, 1 << 28
lib = _native.load()
stride = ((shots + 127) // 128) * 2
ex = torch.empty((n, stride), dtype=torch.int64, device="cuda"); ez = torch.empty_like(ex)
dev.mc_sample_dev(1e-3, shots, 7, 0, ex.data_ptr(), ez.data_ptr(), stride, 0)
cap = int(shots * 2 * n * 1e-3 * 1.3) + (1 << 20)
events = torch.empty(cap, dtype=torch.int64, device="cuda"); count = torch.zeros(1, dtype=torch.int64, device="cuda")
dev.events_from_planes_dev(ex.data_ptr(), ez.data_ptr(), stride, shots, 0, events.data_ptr(), cap, count.data_ptr(), 0)
torch.cuda.synchronize(); k = int(count.item())
hev = events[:k].cpu().numpy().view(np.uint64)
rows_x = torch.empty((shots, n), dtype=torch.uint8, device="cuda"); rows_z = torch.empty_like(rows_x)
_native.check(lib.qcss_unpack_planes_dev(ex.data_ptr(), stride, n, shots, rows_x.data_ptr(), 0))
_native.check(lib.qcss_unpack_planes_dev(ez.data_ptr(), stride, n, shots, rows_z.data_ptr(), 0))
hx, hz = rows_x.cpu().numpy(), rows_z.cpu().numpy()
out = {}
for name, fn in (("sparse", lambda: code.decode_xz_sparse(hev, shots)), ("shot_major", lambda: code.decode_xz(hx, hz))):
    fn(); t0 = time.perf_counter(); res = fn(); dt = time.perf_counter() - t0
    out[name] = dict(shots_per_s=shots / dt, ms=dt * 1e3, tally=res)
print(json.dumps(out))
