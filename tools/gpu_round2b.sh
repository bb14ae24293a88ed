#!/usr/bin/env bash
# One short GPU visit (round 2, third session): the GPU suite on the refactored host copy team,
# the A/B of the chunk pipeline (host_pipe.h), the suite's large-volume tests with the pipeline on.
out=gpurun_out/r02b
mkdir -p $out
set -x
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $out/pytest_gpu_default.txt
timeout 240 python tools/copy_pipe_ab.py 512 256 5 > $out/copy_pipe_ab.txt 2>&1
best=$(python tools/pick_pipe.py $out/copy_pipe_ab.txt)
echo "best: $best" > $out/best.txt
pipe_env=${best:-S3D_COPY_PIPE=1}
env $pipe_env timeout 300 python -m pytest tests/test_gpu_large.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -6 > $out/pytest_gpu_pipe.txt
for t in 5 12; do
  S3D_COPY_THREADS=$t timeout 120 python tools/copy_pipe_ab.py 512 256 4 "copy_pipe=0,copy_pipe=1+pipe_chunk_kb=4096+pipe_slots=8,copy_pipe=1+pipe_chunk_kb=2048+pipe_slots=16" > $out/copy_pipe_ab_t$t.txt 2>&1
done
nproc > $out/host.txt; lscpu | head -25 >> $out/host.txt
cat $out/pytest_gpu_default.txt $out/best.txt $out/pytest_gpu_pipe.txt
cat $out/copy_pipe_ab.txt
