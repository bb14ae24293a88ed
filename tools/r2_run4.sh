#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/r2_run4_pytest.txt
( time python bench.py --steps 10 --warmup 3 ) > gpurun_out/r2_run4_bench.json 2> gpurun_out/r2_run4_bench.err
tail -c 6000 gpurun_out/r2_run4_bench.json; tail -5 gpurun_out/r2_run4_bench.err
python tools/ncu_blur_traffic.py 512 gpurun_out/r02_ncu_blur_traffic.json 2>&1 | tail -8
