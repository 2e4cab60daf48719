#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, launch list + full ncu capture of the headline kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests.log
tail -3 gpurun_out/gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -2 gpurun_out/bench.err
CMD="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_small_named -s 3 -c 1 -o gpurun_out/prof_bench $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
timeout 300 python tools/perf_probe.py --only gf2 > gpurun_out/probe_gf2.jsonl 2>&1; cat gpurun_out/probe_gf2.jsonl | cut -c1-300
