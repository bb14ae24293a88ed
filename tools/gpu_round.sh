#!/usr/bin/env bash
# One GPU visit: parity tests, bench line, launch list.
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --blur-reps 1 > $out/bench_under_ncu.log 2>&1
tail -n 3 $out/pytest_gpu.txt
python - <<PY
import json
d=json.loads(open("$out/bench_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"],d["e2e"]["ms_per_step"],d["stages_ms"],d["config"]["candidates"],d["config"]["keypoints"])
PY
