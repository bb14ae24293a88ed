"""Wall-clock split of the end-to-end step (C API, host buffers) vs the device-resident stages."""
import ctypes as C
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume_torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dev = torch.device("cuda", 0)
vol_dev = blob_volume_torch((n, n, n), 1234, dev)
pinned = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
pinned.copy_(vol_dev)
host = pinned.numpy()
pageable = host.copy()
lib = capi.load_b200()
cu = C.CDLL(str(capi.CUDA_LIB))
lib.lib.sift3d_b200_engine.restype = C.c_void_p
lib.lib.sift3d_b200_engine.argtypes = [C.POINTER(capi.SIFT3D)]
for f, at in (("s3d_image_from_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
              ("s3d_build_pyramid", [C.c_void_p]),
              ("s3d_detect_extrema", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
              ("s3d_assign_orientations", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
              ("s3d_extract_descriptors_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
              ("s3d_engine_sync", [C.c_void_p])):
    getattr(cu, f).argtypes = at
cu.s3d_device_keypoints.argtypes = [C.c_void_p]
cu.s3d_device_keypoints.restype = C.c_void_p
s = capi.Sift3D(lib)
for src, name in ((host, "pinned"), (pageable, "pageable")):
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kp = s.detect_keypoints(src, copy=False)
        t1 = time.perf_counter()
        d = s.extract_descriptors(copy=False)
        t2 = time.perf_counter()
    print(f"e2e {name}: detect {1e3*(t1-t0):.1f} ms  extract {1e3*(t2-t1):.1f} ms  total {1e3*(t2-t0):.1f} ms  kp {len(kp)}")
eng = lib.lib.sift3d_b200_engine(C.byref(s.s))
desc = torch.empty(len(kp) * 3104 + 4096, dtype=torch.uint8, device=dev)
for it in range(3):
    ts = [time.perf_counter()]
    nc, nk = C.c_int(0), C.c_int(0)
    cu.s3d_image_from_device(eng, vol_dev.data_ptr(), n, n, n); cu.s3d_engine_sync(eng); ts.append(time.perf_counter())
    cu.s3d_build_pyramid(eng); cu.s3d_engine_sync(eng); ts.append(time.perf_counter())
    cu.s3d_detect_extrema(eng, s.s.peak_thresh, C.byref(nc)); cu.s3d_engine_sync(eng); ts.append(time.perf_counter())
    cu.s3d_assign_orientations(eng, s.s.corner_thresh, C.byref(nk)); cu.s3d_engine_sync(eng); ts.append(time.perf_counter())
    cu.s3d_extract_descriptors_device(eng, cu.s3d_device_keypoints(eng), nk.value, desc.data_ptr()); cu.s3d_engine_sync(eng); ts.append(time.perf_counter())
names = ["copy", "pyramid+dog", "extrema", "orient", "descriptors"]
print("device stages (ms): " + "  ".join(f"{nm} {1e3*(b-a):.2f}" for nm, a, b in zip(names, ts, ts[1:])),
      f" total {1e3*(ts[-1]-ts[0]):.1f}  cand {nc.value} kp {nk.value}")
