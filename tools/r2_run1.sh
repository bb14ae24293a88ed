#!/bin/bash
# round-2 GPU session 1: parity of the cell-owner descriptor kernel, then A/B timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2_run1_pytest.txt
for o in "" "--opt desc_v2=1" "--opt orient_stage=1"; do
  echo "== bench $o"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --blur-reps 2 $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms', 'config')}))
print('e2e', d['e2e'])
"
done 2>&1 | tee gpurun_out/r2_run1_bench.txt
