#!/usr/bin/env bash
# Two-GPU visit (gpurun --gpus 2): NCCL slab check, independent-volume bench at N=2, slab bench.
tag=${1:-n2}
out=gpurun_out/$tag
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 $TR tools/slab_nccl_check.py > $out/slab_nccl_check.txt 2>&1
timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
SLAB_STEPS=2 timeout 300 $TR tools/slab_bench.py > $out/slab_bench_n2.txt 2>&1
tail -n 2 $out/slab_nccl_check.txt
tail -n 1 $out/bench_n2.json | cut -c1-900
tail -n 4 $out/slab_bench_n2.txt
