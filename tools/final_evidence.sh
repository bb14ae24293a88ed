#!/usr/bin/env bash
# One GPU visit that produces everything cited in profiles/ for the round.
set -x
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.txt 2>&1
python bench.py > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final/bench_ref.json 2>&1
python tools/time_blur.py 512 0 > gpurun_out/final/blur_timing.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/final/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --blur-reps 1 > gpurun_out/final/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_descriptor2 -s 1 -c 1 -o gpurun_out/final/desc \
    python tools/run_desc.py 192 > /dev/null 2>&1
ls -la gpurun_out/final
