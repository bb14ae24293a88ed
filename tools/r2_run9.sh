#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "orientation or golden" 2>&1 | tail -3
for o in "" "--opt orient_scalar=1"; do
  echo "== bench $o"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 2 $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms')}), d['config']['keypoints'], d['e2e']['ms_per_step'])
"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_descriptor3 -c 1 -f -o gpurun_out/r2_desc3c \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 1 > gpurun_out/r2_ncu_desc3c.log 2>&1
tail -c 300 gpurun_out/r2_ncu_desc3c.log
