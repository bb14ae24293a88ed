#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py tests/test_gpu_slab.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_run3_pytest.txt
for o in "" "--opt desc_occ=3" "--opt desc_v2=1"; do
  echo "== bench $o"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --blur-reps 2 $o 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({k: d.get(k) for k in ('ms_per_step', 'stages_ms')}), d['config']['keypoints'])
"
done 2>&1 | tee gpurun_out/r2_run3_bench.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_descriptor3 -s 1 -c 1 -f -o gpurun_out/r2_desc3b \
    python tools/run_desc.py 256 > gpurun_out/r2_ncu_desc3b.log 2>&1
tail -2 gpurun_out/r2_ncu_desc3b.log
