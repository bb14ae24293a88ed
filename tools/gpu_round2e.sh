#!/usr/bin/env bash
# Last short GPU visit: the copy pipeline in the geometry it takes when several ranks share the
# host ($LOCAL_WORLD_SIZE > 1: 8 x 1 MB ring, the polling caller replaces a worker), on one GPU.
out=gpurun_out/r02e
mkdir -p $out
set -x
LOCAL_WORLD_SIZE=8 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pageable_copy or chunks_on_two or golden" 2>&1 | tail -4 > $out/pytest_shared_host.txt
for w in 8 4; do
  LOCAL_WORLD_SIZE=$w timeout 100 python tools/copy_pipe_ab.py 512 256 4 "copy_pipe=0,copy_pipe=1,copy_pipe=1+pipe_chunk_kb=2048+pipe_slots=16" > $out/copy_pipe_ab_lws$w.txt 2>&1
done
cat $out/pytest_shared_host.txt $out/copy_pipe_ab_lws8.txt $out/copy_pipe_ab_lws4.txt
