#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_run12_pytest.txt
python tools/time_blur.py 512 0 2>&1 | tail -8
python - <<'PY'
# octave-0 blur on the example volume's shape (181 x 217 x 181): fused (unaligned) vs per-axis kernels
import ctypes as C, sys
sys.path.insert(0, '.')
import torch
from sift3d_b200.engine_api import Engine
from bench import gauss_taps, pyramid_filters
e = Engine(0); stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); e.set_stream(C.c_void_p(stream.cuda_stream))
src = torch.rand((181, 217, 181), device='cuda'); dst = torch.empty_like(src)
for mode in (0, 1):
    tot = 0
    for sg in pyramid_filters():
        taps = gauss_taps(sg)
        for _ in range(2): e.blur_device(src.data_ptr(), dst.data_ptr(), 181, 217, 181, taps, mode=mode)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(stream)
        for _ in range(5): e.blur_device(src.data_ptr(), dst.data_ptr(), 181, 217, 181, taps, mode=mode)
        b.record(stream); torch.cuda.synchronize(); tot += a.elapsed_time(b) / 5
    print('181x217x181 six blurs, mode', mode, '(0 = auto/fused, 1 = per-axis):', round(tot, 3), 'ms')
PY
