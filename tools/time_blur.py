"""Time the blur at n^3 for the six pyramid filters (CUDA events on the engine's stream)."""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from sift3d_b200.engine_api import Engine
from bench import gauss_taps, pyramid_filters, hbm_peak
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
modes = [int(m) for m in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
reps = 5
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
src = torch.rand((n, n, n), device="cuda")
dst = torch.empty_like(src)
e = Engine(0)
e.set_stream(C.c_void_p(stream.cuda_stream))
peak, _ = hbm_peak()
for mode in modes:
    tot = 0.0
    for sg in pyramid_filters():
        taps = gauss_taps(sg)
        for _ in range(2):
            e.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps, mode=mode)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(reps):
            e.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps, mode=mode)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        tot += ms
        gbs = 8.0 * n ** 3 / ms / 1e6
        print(f"mode {mode} w={len(taps):2d}  {ms:8.4f} ms  {gbs:8.1f} GB/s  frac {gbs/peak:.3f}")
    gbs = 6 * 8.0 * n ** 3 / tot / 1e6
    print(f"mode {mode} total {tot:.3f} ms  aggregate {gbs:.1f} GB/s  frac {gbs/peak:.3f}")
