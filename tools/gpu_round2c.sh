#!/usr/bin/env bash
# Short GPU visit: two-stream descriptor chunks (A/B + tests), the copy pipeline as the default.
out=gpurun_out/r02c
mkdir -p $out
set -x
timeout 200 python tools/copy_pipe_ab.py 512 64 6 "desc_streams=1,desc_streams=2,desc_streams=2+desc_chunk=2048,desc_streams=2+desc_chunk=8192,desc_streams=1+desc_chunk=8192,desc_streams=1,desc_streams=2" > $out/desc_streams_ab.txt 2>&1
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $out/pytest_gpu.txt
cat $out/desc_streams_ab.txt $out/pytest_gpu.txt
