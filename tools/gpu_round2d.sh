#!/usr/bin/env bash
# Final GPU visit of the round: smoke + the N = 1 bench record (with the CPU baseline leg).
out=gpurun_out/r02d
mkdir -p $out
set -x
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err
cat $out/smoke.txt; tail -c 3000 $out/bench_n1.json; tail -5 $out/bench_n1.err
