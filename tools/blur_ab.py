"""A/B of the fused blur kernels on the GPU: parity first (k_blur_tma vs k_blur_fused, bit for bit,
on shapes that hit interior / edge / overlapping tiles and both tile heights), then CUDA-event
timings of the six pyramid filters at n^3 for every variant.  Usage: blur_ab.py [n] [variants]
variants: comma list of name=value option sets separated by '+', e.g. "blur_v1=1,blur_v1=0,blur_v1=0+blur_rpt4_hw=0"."""
import ctypes as C
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from sift3d_b200.engine_api import Engine
from bench import gauss_taps, pyramid_filters, hbm_peak

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
variants = sys.argv[2].split(",") if len(sys.argv) > 2 else ["blur_v1=1", "blur_v1=0"]
e = Engine(0)
e.L.s3d_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]


def setopts(spec):
    for kv in spec.split("+"):
        k, v = kv.split("=")
        assert e.L.s3d_set_option(e.h, k.encode(), int(v)) == 0, kv


# ---- parity: new vs old, bit for bit ---------------------------------------------------------
rng = np.random.default_rng(5)
bad = 0
import os
for shape in ([] if os.environ.get("SKIP_PARITY") else [(40, 96, 84), (37, 100, 132), (70, 130, 200), (24, 64, 76), (19, 49, 88), (50, 67, 256), (33, 80, 80), (40, 200, 260)]):
    vol = rng.random(shape, dtype=np.float32)
    for sg in pyramid_filters() + [2.2, 0.3]:
        taps = gauss_taps(sg)
        setopts("blur_v1=1")
        want = e.blur(vol, taps)
        for spec in ("blur_v1=0+blur_rpt4_hw=4", "blur_v1=0+blur_rpt4_hw=0"):
            setopts(spec)
            got = e.blur(vol, taps)
            if not np.array_equal(got.view(np.uint32), want.view(np.uint32)):
                d = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
                print("MISMATCH", shape, len(taps), spec, len(d), "first", d[:3].tolist(), flush=True)
                bad += 1
print("parity:", "OK" if bad == 0 else f"{bad} mismatching cases", flush=True)
setopts("blur_v1=0+blur_rpt4_hw=4")

# ---- timing -------------------------------------------------------------------------------------
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
src = torch.rand((n, n, n), device="cuda")
dst = torch.empty_like(src)
e.set_stream(C.c_void_p(stream.cuda_stream))
peak, _ = hbm_peak()
reps = 5
for spec in variants:
    setopts(spec)
    tot = 0.0
    line = []
    for sg in pyramid_filters():
        taps = gauss_taps(sg)
        for _ in range(2):
            e.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(stream)
        for _ in range(reps):
            e.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps)
        b.record(stream)
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        tot += ms
        line.append(f"w{len(taps)}={ms:.3f}")
    gbs = 6 * 8.0 * n ** 3 / tot / 1e6
    print(f"{spec:40s} {' '.join(line)}  total {tot:.3f} ms  frac {gbs/peak:.3f}", flush=True)
