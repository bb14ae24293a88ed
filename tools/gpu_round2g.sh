#!/usr/bin/env bash
# The automatic copy path of a rank on a shared host (copy_pipe=-1 with $LOCAL_WORLD_SIZE > 1 -> the
# poller-free pipeline, the caller in worker 0's place), on one GPU.
out=gpurun_out/r02g
mkdir -p $out
LOCAL_WORLD_SIZE=4 timeout 50 python tools/copy_pipe_ab.py 512 256 3 "copy_pipe=0,copy_pipe=-1" > $out/ab_auto_lws4.txt 2>&1
cat $out/ab_auto_lws4.txt
