#!/usr/bin/env bash
# One GPU visit: orientation batch A/B (S3D_ORIENT_BATCH) -- parity + bench stage split.
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
S3D_ORIENT_BATCH=8 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5 > $out/pytest_gpu_b8.txt
for b in 4 8; do
S3D_ORIENT_BATCH=$b timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_b$b.json 2> $out/bench_b$b.err
done
tail -n 3 $out/pytest_gpu_b8.txt
python - <<PY
import json
for b in (4,8):
    d=json.loads(open("$out/bench_b%d.json"%b).read().strip().splitlines()[-1])
    print(b, d["ms_per_step"],d["e2e"]["ms_per_step"],d["stages_ms"], d["config"]["keypoints"])
PY
