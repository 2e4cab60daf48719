"""Build quantum_css_codes_b200/libqcss_experiments.so: the product sources with -DQCSS_EXPERIMENTS, i.e. plus the
superseded kernel generations (tools/experiments/tiled_variants.inc, gf2_fast.cu) and their QCSS_* environment
knobs.  Used only to reproduce the round-1 measurements quoted in DESIGN.md; the product library has neither.

    python tools/experiments/build_experiments.py
    then, in a probe script: _native.LIB_PATH = ".../libqcss_experiments.so" before _native.load(), and e.g.
    QCSS_TILED_WIDE=1 python tools/hgp_probe.py
"""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
from quantum_css_codes_b200 import build as qbuild   # noqa: E402

out = os.path.join(qbuild.HERE, "libqcss_experiments.so")
srcs = [os.path.join(qbuild.CSRC, s) for s in qbuild.SOURCES] + [os.path.join(REPO, "tools", "experiments", "gf2_fast.cu")]
cmd = [qbuild.nvcc(), *qbuild.ARCH, *qbuild.FLAGS, "-DQCSS_EXPERIMENTS", "-I", qbuild.CSRC, "-shared", "-o", out, *srcs]
print(" ".join(cmd))
subprocess.run(cmd, check=True)
print(out)
