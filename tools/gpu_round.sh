#!/usr/bin/env bash
# One GPU visit: parity tests, smoke, full bench line, launch list.
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
timeout 600 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --blur-reps 1 > $out/bench_under_ncu.log 2>&1
cat $out/pytest_gpu.txt $out/smoke.txt
python - <<PY
import json
d=json.loads(open("$out/bench_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"],d["e2e"]["ms_per_step"],d["stages_ms"],d["config"]["candidates"],d["config"]["keypoints"],d["gpu_launches"])
PY
