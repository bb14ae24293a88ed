#!/usr/bin/env bash
# Two-GPU visit (gpurun --gpus 2): independent-volume bench at N=2, NCCL slab check, slab bench.
tag=${1:-n2}
out=gpurun_out/$tag
mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $out/bench_n2.json 2> $out/bench_n2.err
timeout 300 $TR tools/slab_nccl_check.py > $out/slab_nccl_check.txt 2>&1
timeout 600 $TR tools/slab_bench.py > $out/slab_bench_n2.txt 2>&1
tail -n 1 $out/bench_n2.json | cut -c1-700
tail -n 3 $out/slab_nccl_check.txt
tail -n 6 $out/slab_bench_n2.txt
