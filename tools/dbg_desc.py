import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent / 'oracle'))
from oracle_api import Oracle
from sift3d_b200.volumes import noise_volume
vol = noise_volume((40, 48, 56), 1)
lib = capi.load_b200(); orc = Oracle()
okp = orc.detect(vol); od, _ = orc.describe(okp)
cu = C.CDLL(str(capi.CUDA_LIB)); cu.s3d_set_option.argtypes=[C.c_void_p, C.c_char_p, C.c_int]
lib.lib.sift3d_b200_engine.restype = C.c_void_p
for fast in (1, 0):
    with capi.Sift3D(lib) as s:
        kp = s.detect_keypoints(vol)
        eng = lib.lib.sift3d_b200_engine(C.byref(s.s))
        cu.s3d_set_option(eng, b"icos_fast", fast)
        d = s.extract_descriptors()["hists"]
        rel = np.linalg.norm(d - od, axis=1) / np.linalg.norm(od, axis=1)
        w = np.argsort(rel)[::-1][:4]
        print("icos_fast", fast, "max", rel.max(), "n>1e-6:", int((rel > 1e-6).sum()))
        for i in w:
            diff = d[i] - od[i]
            top = np.argsort(np.abs(diff))[::-1][:6]
            print("  kp", i, "o,s", kp["o"][i], kp["s"][i], "xyz", kp["xd"][i], kp["yd"][i], kp["zd"][i], "rel", rel[i])
            print("    bins", [(int(b), float(diff[b]), float(od[i][b])) for b in top])
