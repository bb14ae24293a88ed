#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
tail -2 gpurun_out/r2_bench_under_ncu.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_orient_group -c 1 -f -o gpurun_out/r2_orient_group \
    python tools/run_desc.py 256 > gpurun_out/r2_ncu_orient_group.log 2>&1
tail -1 gpurun_out/r2_ncu_orient_group.log
