#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_resample.py -m gpu -q -x 2>&1 | tail -3
for o in "" "--opt orient_g=16" "--opt orient_g=32" "--opt desc_pre=1" "--opt desc_pre=1 --opt desc_occ=3"; do
  echo "== bench $o"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 1 $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms')}), d['config']['keypoints'], d['e2e']['ms_per_step'])
"
done
