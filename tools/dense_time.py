"""Time SIFT3D_extract_dense_descriptors at n^3 through the C API (config 3) and check it on a crop."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
vol = blob_volume(n, seed=1234)
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    for rep in range(3):
        t0 = time.perf_counter()
        d = s.extract_dense_descriptors(vol)
        dt = time.perf_counter() - t0
        print(f"dense {n}^3 rep{rep}: {dt*1e3:.1f} ms  ({vol.size/dt/1e6:.1f} Mvox/s, host buffers, 48 B/voxel out)")
    print("shape", d.shape, "finite", bool(np.isfinite(d).all()))
