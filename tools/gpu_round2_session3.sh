#!/usr/bin/env bash
# The short GPU visits of round 2's last session (180 GPU-minutes per round; 18 were left), one
# function per visit; results under profiles/r02_copy_pipe_ab.txt, r02_desc_streams_ab.txt,
# r02_bench_n1.json, r02_pytest_gpu.txt.   usage: bash tools/gpu_round2_session3.sh <visit>
out=gpurun_out/r02_s3
mkdir -p $out
AB="python tools/copy_pipe_ab.py"
case "$1" in
copy_pipe)      # GPU suite on the refactored host team + A/B of the chunk pipeline
    timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $out/pytest_gpu_default.txt
    timeout 240 $AB 512 256 5 > $out/copy_pipe_ab.txt 2>&1
    S3D_COPY_PIPE=1 timeout 300 python -m pytest tests/test_gpu_large.py tests/test_gpu_cli.py -m gpu -q -x 2>&1 | tail -6 > $out/pytest_gpu_pipe.txt
    for t in 5 12; do
        S3D_COPY_THREADS=$t timeout 120 $AB 512 256 4 "copy_pipe=0,copy_pipe=1+pipe_chunk_kb=4096+pipe_slots=8,copy_pipe=1+pipe_chunk_kb=2048+pipe_slots=16" > $out/copy_pipe_ab_t$t.txt 2>&1
    done
    nproc > $out/host.txt; lscpu | head -25 >> $out/host.txt ;;
desc_streams)   # two-stream descriptor chunks: A/B, then the whole GPU suite
    timeout 200 $AB 512 64 6 "desc_streams=1,desc_streams=2,desc_streams=2+desc_chunk=2048,desc_streams=2+desc_chunk=8192,desc_streams=1+desc_chunk=8192,desc_streams=1,desc_streams=2" > $out/desc_streams_ab.txt 2>&1
    timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $out/pytest_gpu.txt ;;
bench)          # smoke + the N = 1 bench record (with the CPU baseline leg)
    python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1
    timeout 400 python bench.py --steps 20 --warmup 5 > $out/bench_n1.json 2> $out/bench_n1.err ;;
shared_host)    # the copy paths with the few threads a rank gets when ranks share the host
    LOCAL_WORLD_SIZE=8 timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pageable_copy or chunks_on_two or golden" 2>&1 | tail -4 > $out/pytest_shared_host.txt
    for w in 8 4; do
        LOCAL_WORLD_SIZE=$w timeout 100 $AB 512 256 4 "copy_pipe=0,copy_pipe=1,copy_pipe=2,copy_pipe=-1" > $out/copy_pipe_ab_lws$w.txt 2>&1
    done
    timeout 40 $AB 512 256 3 "copy_pipe=1,copy_pipe=2" > $out/copy_pipe_ab_lws1.txt 2>&1 ;;
*) echo "usage: $0 copy_pipe|desc_streams|bench|shared_host"; exit 2 ;;
esac
cat $out/*.txt 2>/dev/null | tail -60
