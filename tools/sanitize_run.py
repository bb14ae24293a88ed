"""Small run touching every kernel (compute-sanitizer target)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume
vol = blob_volume((56, 64, 96), seed=2)   # fused blur, per-octave DoG, 4-voxel extrema, cell-owner descriptors
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    kp = s.detect_keypoints(vol)
    d = s.extract_descriptors()
    # the round-1 kernels that remain as fallbacks / A-B references
    import ctypes as C2
    cu = C2.CDLL(str(capi.CUDA_LIB))
    cu.s3d_set_option.argtypes = [C2.c_void_p, C2.c_char_p, C2.c_int]
    lib.lib.sift3d_b200_engine.restype = C2.c_void_p
    lib.lib.sift3d_b200_engine.argtypes = [C2.POINTER(capi.SIFT3D)]
    eng = lib.lib.sift3d_b200_engine(C2.byref(s.s))
    for opt in (b"desc_v2", b"orient_v1"):
        cu.s3d_set_option(eng, opt, 1)
        s.detect_keypoints(vol)
        s.extract_descriptors()
        cu.s3d_set_option(eng, opt, 0)
    kp = s.detect_keypoints(vol)
    f = lib.lib.sift3d_b200_materialize_pyramids
    f.argtypes = [C2.POINTER(capi.SIFT3D)]
    assert f(C2.byref(s.s)) == 0
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
