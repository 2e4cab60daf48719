"""Per-source-line instruction and stall-sample shares from an .ncu-rep captured with
--import-source on (kernels built with -lineinfo).
    python tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
from collections import defaultdict
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
inst = defaultdict(int); samp = defaultdict(int); text = {}
fname = ''
for r in rows:
    if r and r[0] == 'File Path':
        fname = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No':
        hdr = r
        iI = hdr.index('Instructions Executed'); iW = hdr.index('# Samples')
        continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        key = (fname, int(r[0]))
    except ValueError:
        continue
    text[key] = r[1]
    try:
        inst[key] += int(r[iI]); samp[key] += int(r[iW])
    except (ValueError, IndexError):
        pass
ti = sum(inst.values()) or 1; ts = sum(samp.values()) or 1
print(f"total warp-inst {ti}  samples {ts}")
for key in sorted(inst):
    pi, ps = 100 * inst[key] / ti, 100 * samp[key] / ts
    if pi >= minpct or ps >= minpct:
        print(f"{key[0]}:{key[1]:4d} {pi:5.1f}% inst {ps:5.1f}% stall  | {text[key].strip()[:90]}")
