#!/usr/bin/env bash
# One GPU visit: parity tests, bench line, descriptor kernel metrics.
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_gpu.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_n1.json 2> $out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_descriptor2 -s 1 -c 1 -o $out/desc \
    python tools/run_desc.py 192 > $out/ncu_desc.log 2>&1
tail -n 4 $out/pytest_gpu.txt
python - <<PY
import json
d=json.loads(open("$out/bench_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"],d["e2e"]["ms_per_step"],d["stages_ms"])
PY
