"""Time SIFT3D_extract_dense_descriptors with dense_rotate = 1 at n^3 through the C API."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
vol = blob_volume(n, seed=1234)
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    s.s.dense_rotate = 1
    for rep in range(2):
        t0 = time.perf_counter()
        d = s.extract_dense_descriptors(vol)
        dt = time.perf_counter() - t0
        print(f"dense_rotate {n}^3 rep{rep}: {dt*1e3:.1f} ms  ({vol.size/dt/1e6:.2f} Mvox/s, host buffers)")
    print("shape", d.shape, "finite", bool(np.isfinite(d).all()))
