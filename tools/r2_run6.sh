#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2_run6_pytest.txt
for o in "" "--opt orient_v1=1"; do
  echo "== bench $o"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 2 $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms')}), d['config']['keypoints'], d['e2e']['ms_per_step'])
"
done 2>&1 | tee gpurun_out/r2_run6_bench.txt
