#!/usr/bin/env bash
# One GPU visit: parity tests, bench line, descriptor-kernel variants (S3D_DESC_OCC) under ncu.
# usage: tools/gpu_round.sh <tag>
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_gpu.txt
S3D_DESC_OCC=3 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > $out/pytest_gpu_occ3.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_occ4.json 2> $out/bench_occ4.err
S3D_DESC_OCC=3 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_occ3.json 2> $out/bench_occ3.err
for occ in 4 3; do
S3D_DESC_OCC=$occ timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:k_descriptor2 -s 1 -c 1 python tools/run_desc.py 192 > $out/ncu_desc_occ$occ.txt 2>&1
done
tail -3 $out/pytest_gpu.txt $out/pytest_gpu_occ3.txt
python - <<PY
import json
for occ in (4,3):
    try:
        d=json.loads(open("$out/bench_occ%d.json"%occ).read().strip().splitlines()[-1])
        print("occ",occ,d["ms_per_step"],d["e2e"]["ms_per_step"],d["stages_ms"])
    except Exception as ex: print("occ",occ,"failed",ex)
PY
grep -E "gpu__time|inst_executed|issue_active|lsu_wave" $out/ncu_desc_occ4.txt $out/ncu_desc_occ3.txt
