#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full captures of the kernels changed in round 2.
# usage: gpurun --timeout 1800 -- 'bash tools/gpu_profiles.sh r02'
tag=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_bench_for_ncu.json 2> gpurun_out/${tag}_bench_for_ncu.err; echo "plain bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_syndrome_mma -c 1 -o gpurun_out/${tag}_dense_v3 \
    python tools/run_dense.py 2097152 > gpurun_out/${tag}_ncu_dense.log 2>&1; echo "dense rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_decode_events|k_pack_shots|k_small_named" -c 6 -o gpurun_out/${tag}_formats \
    python tools/formats_probe.py > gpurun_out/${tag}_ncu_formats.log 2>&1; echo "formats rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gf2_m4r2 -c 1 -o gpurun_out/${tag}_gf2_m4r2_lazy \
    python tools/run_gf2.py 296 > gpurun_out/${tag}_ncu_gf2.log 2>&1; echo "gf2 rc=$?"
