#!/usr/bin/env bash
# One GPU visit that produces everything cited under profiles/r02_* (round 2).
# usage (from the repo root, on a B200 box): bash tools/final_evidence_r2.sh
out=gpurun_out/r02_final
mkdir -p $out
set -x
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dense --blur-reps 1 > $out/bench_under_ncu.log 2>&1
python tools/ncu_blur_traffic.py 512 $out/r02_ncu_blur_traffic.json > $out/blur_traffic.log 2>&1
for w in 0 5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_blur_tma -s 2 -c 1 -o $out/blur_tma_f$w -f \
      python tools/run_blur.py 512 $w 3 > $out/ncu_tma_f$w.log 2>&1
done
cat $out/pytest_gpu.txt $out/smoke.txt
tail -c 1500 $out/bench_n1.json
