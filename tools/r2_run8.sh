#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -x 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 2 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms')}), d['config']['keypoints'], d['e2e']['ms_per_step'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 1 > gpurun_out/r2_bench_under_ncu.log 2>&1
