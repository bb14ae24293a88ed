"""Small run touching every kernel (compute-sanitizer target)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume
vol = blob_volume((40, 48, 72), seed=2)   # fused blur eligible (nx>=64, ny>=32), odd tile overlap
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    kp = s.detect_keypoints(vol)
    d = s.extract_descriptors()
    dd = s.extract_dense_descriptors(np.ascontiguousarray(vol[:20, :24, :32]))
    im = capi.make_image(vol)
    assert lib.lib.SIFT3D_extract_raw_descriptors(C.byref(s.s), C.byref(im), C.byref(s.kp), C.byref(s.desc)) == 0
    conf = C.POINTER(C.c_double)()
    assert lib.lib.SIFT3D_assign_orientations(C.byref(s.s), C.byref(im), C.byref(s.kp), C.byref(conf)) == 0
    # row length not a multiple of 4: octave 0 takes the register-blocked per-axis kernels
    vol2 = blob_volume((30, 33, 70), seed=5)
    kp2 = s.detect_keypoints(vol2)
    d2 = s.extract_descriptors() if len(kp2) else None
print("ok", len(kp), d.shape, dd.shape, len(kp2))
