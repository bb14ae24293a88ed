"""Time SIFT3D_extract_dense_descriptors at n^3 through the C API (BASELINE.json configs[2]):
the C call alone (host buffers in and out, no Python-side copy of the 48 B/voxel result)."""
import ctypes as C
import sys
import time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume_torch
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
rotate = int(sys.argv[2]) if len(sys.argv) > 2 else 0
vol = blob_volume_torch((n, n, n), 1234, torch.device("cuda", 0)).cpu().numpy()
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    s.s.dense_rotate = rotate
    im = capi.make_image(vol)
    out = capi.empty_image()
    for rep in range(4):
        t0 = time.perf_counter()
        rc = s.L.SIFT3D_extract_dense_descriptors(C.byref(s.s), C.byref(im), C.byref(out))
        dt = time.perf_counter() - t0
        assert rc == 0
        print(f"dense {n}^3 rotate={rotate} rep{rep}: {dt*1e3:.1f} ms  ({vol.size/dt/1e6:.1f} Mvox/s, "
              f"host buffers, {out.nc * 4} B/voxel out)")
    a = np.ctypeslib.as_array(out.data, shape=(out.nx * out.ny * out.nz * out.nc,))
    print("finite", bool(np.isfinite(a[::97]).all()), "nonzero frac", float((a[::97] != 0).mean()))
