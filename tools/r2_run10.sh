#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --blur-reps 2 2>&1 | tail -1 > gpurun_out/r2_run10_bench.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2_run10_bench.json').read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms', 'e2e', 'dense')}))
PY
S3D_COPY_THREADS=8 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('8 threads e2e', d['e2e']['ms_per_step'], d['e2e']['pinned']['ms_per_step'])"
