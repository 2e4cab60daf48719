"""Debug helper: small HGP syndrome call (run under compute-sanitizer on the GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quantum_css_codes_b200 import SyndromeCode, codes
from oracle import montecarlo as omc
hx, hz = codes.hgp1600()
code = SyndromeCode(hx, hz)
rng = np.random.default_rng(1)
for shots in (64, 513, 5000):
    errs = (rng.random((shots, 1600)) < 0.3).astype(np.uint8)
    got = code.syndromes(errs, 2)
    print(shots, "match", bool(np.array_equal(got, omc.syndromes_batch(hz, errs))), flush=True)
