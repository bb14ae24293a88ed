#!/usr/bin/env bash
# One GPU visit: descriptor vertex-order rotation A/B (S3D_DESC_NOROT) -- parity, bench, ncu counters.
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5 > $out/pytest_gpu.txt
for r in 1 0; do
S3D_DESC_NOROT=$r timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_norot$r.json 2> $out/bench_norot$r.err
S3D_DESC_NOROT=$r timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum \
    --clock-control none -k regex:k_descriptor2 -s 1 -c 1 python tools/run_desc.py 192 > $out/ncu_norot$r.txt 2>&1
done
tail -n 3 $out/pytest_gpu.txt
python - <<PY
import json
for b in (1,0):
    d=json.loads(open("$out/bench_norot%d.json"%b).read().strip().splitlines()[-1])
    print("norot",b, d["ms_per_step"],d["e2e"]["ms_per_step"],d["stages_ms"])
PY
grep -E "gpu__time|inst_executed|issue_active|lsu_wave|bank_conf" $out/ncu_norot1.txt $out/ncu_norot0.txt
