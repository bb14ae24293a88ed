"""A/B of the two host-side paths between PAGEABLE host memory and the device (engine.cu):
the chunk-at-a-time parallel memcpy (copy_pipe=0) against the whole-chunk worker pipeline
(copy_pipe=1, host_pipe.h) for several chunk sizes / ring depths.  Through the C API, like a stock
caller: SIFT3D_detect_keypoints + SIFT3D_extract_descriptors on a 512^3 malloc'ed volume, and
SIFT3D_extract_dense_descriptors on 256^3.  Results must be bit-identical between the variants.
Usage: copy_pipe_ab.py [n] [dense_n] [reps] [variants: comma list of name=value sets joined by +]"""
import ctypes as C
import hashlib
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume_torch
from bench import bind_cuda, engine_of

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
dn = int(sys.argv[2]) if len(sys.argv) > 2 else 256
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
vol = blob_volume_torch((n, n, n), 1234, dev).cpu().numpy().copy()   # pageable
dvol = blob_volume_torch((dn, dn, dn), 1234, dev).cpu().numpy().copy()
lib = capi.load_b200()
cu = bind_cuda(capi)
variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["copy_pipe=0", "copy_pipe=1+pipe_chunk_kb=4096+pipe_slots=8", "copy_pipe=1+pipe_chunk_kb=2048+pipe_slots=16",
            "copy_pipe=1+pipe_chunk_kb=1024+pipe_slots=16", "copy_pipe=1+pipe_chunk_kb=8192+pipe_slots=4",
            "copy_pipe=1+pipe_chunk_kb=1024+pipe_slots=48", "copy_pipe=1+pipe_chunk_kb=16384+pipe_slots=3"]


def setopts(s, spec):
    eng = engine_of(lib, s, capi)
    for kv in spec.split("+"):
        k, v = kv.split("=")
        assert cu.s3d_set_option(eng, k.encode(), int(v)) == 0, kv


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()[:16]


ref = None
print(f"sparse {n}^3, pageable volume in, keypoints + descriptors out; {reps} reps, min / median ms")
with capi.Sift3D(lib) as s:
    s.detect_keypoints(vol[:64, :64, :64].copy(), copy=False)  # creates the engine
    for spec in variants:
        setopts(s, spec)
        td, te = [], []
        for it in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            kp = s.detect_keypoints(vol, copy=False)
            t1 = time.perf_counter()
            d = s.extract_descriptors(copy=False)
            t2 = time.perf_counter()
            if it:
                td.append(1e3 * (t1 - t0)), te.append(1e3 * (t2 - t1))
        tot = np.array(td) + np.array(te)
        sig = (len(kp), digest(kp["xd"]), digest(d["hists"]))
        if ref is None:
            ref = sig
        print(f"{spec:50s} detect {min(td):6.2f} / {np.median(td):6.2f}  extract {min(te):6.2f} / {np.median(te):6.2f}"
              f"  total {tot.min():6.2f} / {np.median(tot):6.2f}  {'same result' if sig == ref else 'RESULT DIFFERS ' + str(sig)}",
              flush=True)

print(f"dense {dn}^3, pageable in and out (805 MB at 256^3): e2e ms min, device-side upload / kernels / download")
dref = None
with capi.Sift3D(lib) as s:
    im = capi.make_image(dvol)
    res = capi.empty_image()
    assert s.L.SIFT3D_extract_dense_descriptors(C.byref(s.s), C.byref(im), C.byref(res)) == 0
    for spec in variants:
        setopts(s, spec)
        ts = []
        for it in range(4):
            t0 = time.perf_counter()
            assert s.L.SIFT3D_extract_dense_descriptors(C.byref(s.s), C.byref(im), C.byref(res)) == 0
            ts.append(1e3 * (time.perf_counter() - t0))
        ms3 = (C.c_double * 3)()
        cu.s3d_dense_last_timing(engine_of(lib, s, capi), ms3)
        arr = np.ctypeslib.as_array(res.data, shape=(res.nx * res.ny * res.nz * res.nc,))
        sig = digest(arr)
        if dref is None:
            dref = sig
        print(f"{spec:50s} e2e {min(ts[1:]):6.2f}  upload {ms3[0]:5.2f} kernels {ms3[1]:5.2f} download {ms3[2]:6.2f}"
              f"  {'same result' if sig == dref else 'RESULT DIFFERS'}", flush=True)
    lib._libc.free(C.cast(res.data, C.c_void_p))
