#!/usr/bin/env bash
# Very short GPU visit: the poller-free copy pipeline (copy_pipe=2) -- bytes, then timing with
# the thread counts a rank gets on a shared host.
out=gpurun_out/r02f
mkdir -p $out
LOCAL_WORLD_SIZE=8 timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pageable_copy" 2>&1 | tail -3 > $out/pytest_copy.txt
LOCAL_WORLD_SIZE=8 timeout 40 python tools/copy_pipe_ab.py 512 256 3 "copy_pipe=0,copy_pipe=2" > $out/ab_lws8.txt 2>&1
LOCAL_WORLD_SIZE=4 timeout 40 python tools/copy_pipe_ab.py 512 256 3 "copy_pipe=0,copy_pipe=2" > $out/ab_lws4.txt 2>&1
timeout 40 python tools/copy_pipe_ab.py 512 256 3 "copy_pipe=1,copy_pipe=2" > $out/ab_lws1.txt 2>&1
cat $out/pytest_copy.txt $out/ab_lws8.txt $out/ab_lws4.txt $out/ab_lws1.txt
