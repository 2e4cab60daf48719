#!/bin/bash
# One GPU-box round trip: smoke, GPU tests, bench at N = 1 (and the reference arm), launch list.
# usage (from the build container): gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r02_x'
tag=${1:-r02}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
tools/i8_mma_peak > gpurun_out/${tag}_i8_mma_peak.json 2>&1; cat gpurun_out/${tag}_i8_mma_peak.json
python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${tag}_bench_n1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${tag}_bench_n1.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["frac"], d["e2e"])
for k, v in (d.get("other_configs") or {}).items():
    print(k, {a: b for a, b in v.items() if a in ("ms", "value", "error")}, (v.get("roofline") or {}).get("frac"))
print(d.get("strong_scaling")); print(d.get("cpu_reference"))
PY
