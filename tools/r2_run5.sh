#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "materialis or copy_sift3d or kernels_agree" 2>&1 | tail -3
for w in 0 2; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_blur_fused -s 2 -c 1 -f -o gpurun_out/r2_blur_f$w \
    python tools/run_blur.py 512 $w 3 > gpurun_out/r2_ncu_blur_f$w.log 2>&1
tail -1 gpurun_out/r2_ncu_blur_f$w.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_orient -c 1 -f -o gpurun_out/r2_orient \
    python tools/run_desc.py 256 > gpurun_out/r2_ncu_orient.log 2>&1
tail -1 gpurun_out/r2_ncu_orient.log
