#!/usr/bin/env bash
# One GPU visit that produces everything cited in profiles/ for the round.
# usage: tools/final_evidence.sh <tag>
tag=${1:-final}
out=gpurun_out/$tag
set -x
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
python bench.py > $out/bench_n1.json 2> $out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref.json 2>&1
python tools/time_blur.py 512 0 > $out/blur_timing.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1700 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 2 --no-cpu-baseline --blur-reps 1 > $out/bench_under_ncu.log 2>&1
S3D_TRACE=1 timeout 300 python tools/dense_time.py 256 > $out/dense_256.txt 2>&1
timeout 300 python tools/e2e_breakdown.py 512 > $out/e2e_breakdown.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_conv_dyadic|k_extrema_mark" -c 37 --csv --page raw \
    --log-file $out/ncu_dyadic_extrema.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --blur-reps 1 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_descriptor2 -s 1 -c 1 -o $out/desc \
    python tools/run_desc.py 192 > $out/ncu_desc.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $out/sanitizer_memcheck.txt 2>&1
echo "memcheck rc=$?" >> $out/sanitizer_memcheck.txt
ls -la $out
cat $out/pytest_gpu.txt $out/smoke.txt $out/bench_n1.json $out/dense_256.txt $out/e2e_breakdown.txt
tail -5 $out/sanitizer_memcheck.txt
