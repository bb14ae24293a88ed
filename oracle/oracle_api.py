"""ctypes access to oracle/_build/liboracle.so (the CPU restatement) and to the compiled
reference in oracle/_ref -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module (`sys.path` gets the `oracle/` directory there).  Nothing under
sift3d_b200/ imports it; the product path is sift3d_b200.capi -> libsift3D.so ->
libsift3d_cuda.so and fails loudly without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
ORACLE_LIB = REPO / "oracle" / "_build" / "liboracle.so"
REF_DIR = REPO / "oracle" / "_ref"
REF_LIB = REF_DIR / "libsift3D_ref.so"
REF_IMUTIL = REF_DIR / "libimutil_ref.so"


def load_reference():
    """The UNMODIFIED reference compiled by oracle/build_ref.sh, behind the same ctypes
    wrapper class the product library uses (byte-identical struct layouts)."""
    import sys
    sys.path.insert(0, str(REPO))
    from sift3d_b200 import capi
    return capi.Sift3DLib(REF_LIB, "reference")


class OrcParams(C.Structure):
    _fields_ = [("peak_thresh", C.c_double), ("corner_thresh", C.c_double),
                ("sigma_n", C.c_double), ("sigma0", C.c_double), ("num_kp_levels", C.c_int)]


class OrcLevel(C.Structure):
    _fields_ = [("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int), ("ux", C.c_double),
                ("uy", C.c_double), ("uz", C.c_double), ("s", C.c_double),
                ("data", C.POINTER(C.c_float))]


ORC_KP_DTYPE = np.dtype({"names": ["xd", "yd", "zd", "sd", "o", "s", "R"],
                         "formats": ["<f8", "<f8", "<f8", "<f8", "<i4", "<i4", ("<f4", (3, 3))],
                         "offsets": [0, 8, 16, 24, 32, 36, 40], "itemsize": 80})


class Oracle:
    def __init__(self, peak_thresh=0.1, corner_thresh=0.4, sigma_n=1.15, sigma0=1.6,
                 num_kp_levels=3):
        if not ORACLE_LIB.exists():
            raise FileNotFoundError(f"{ORACLE_LIB} missing: run `make -C oracle`")
        L = self.L = C.CDLL(str(ORACLE_LIB))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcParams)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(OrcLevel)]
        for f in ("orc_num_octaves", "orc_num_candidates", "orc_num_keypoints"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_candidates", "orc_keypoints"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.c_void_p
        L.orc_describe.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_dense.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p]
        L.orc_dense_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, C.c_void_p]
        L.orc_gauss_taps.argtypes = [C.c_double, C.c_void_p, C.c_int]
        L.orc_blur.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.orc_blur.restype = None
        L.orc_eig3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_mesh.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_mesh.restype = None
        L.orc_resample_affine.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                          C.c_int]
        L.orc_nn_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float,
                                   C.c_void_p]
        p = OrcParams(peak_thresh, corner_thresh, sigma_n, sigma0, num_kp_levels)
        self.ctx = L.orc_create(C.byref(p))

    def __del__(self):
        if getattr(self, "ctx", None):
            self.L.orc_destroy(self.ctx)
            self.ctx = None

    def detect(self, vol, units=(1.0, 1.0, 1.0)):
        vol = np.ascontiguousarray(vol, np.float32)
        nz, ny, nx = vol.shape
        u = np.asarray(units, np.float64)
        rc = self.L.orc_detect(self.ctx, vol.ctypes.data, nx, ny, nz, u.ctypes.data)
        if rc:
            raise RuntimeError("orc_detect failed")
        return self.keypoints()

    def _kps(self, ptr, n):
        if n == 0:
            return np.zeros(0, ORC_KP_DTYPE)
        buf = (C.c_char * (n * 80)).from_address(ptr)
        return np.frombuffer(buf, ORC_KP_DTYPE, n).copy()

    def keypoints(self):
        return self._kps(self.L.orc_keypoints(self.ctx), self.L.orc_num_keypoints(self.ctx))

    def candidates(self):
        return self._kps(self.L.orc_candidates(self.ctx), self.L.orc_num_candidates(self.ctx))

    def num_octaves(self):
        return self.L.orc_num_octaves(self.ctx)

    def level(self, which, o, s):
        lv = OrcLevel()
        if self.L.orc_get_level(self.ctx, 0 if which == "gpyr" else 1, o, s, C.byref(lv)):
            raise IndexError((which, o, s))
        n = lv.nx * lv.ny * lv.nz
        return np.ctypeslib.as_array(lv.data, shape=(n,)).reshape(lv.nz, lv.ny, lv.nx).copy(), lv

    def describe(self, kps):
        kps = np.ascontiguousarray(kps)
        n = len(kps)
        desc = np.zeros((n, 768), np.float32)
        coords = np.zeros((n, 4), np.float64)
        rc = self.L.orc_describe(self.ctx, kps.ctypes.data, n, desc.ctypes.data,
                                 coords.ctypes.data)
        if rc:
            raise RuntimeError("orc_describe failed")
        return desc, coords

    def dense(self, vol, units=(1.0, 1.0, 1.0), rotate=False):
        vol = np.ascontiguousarray(vol, np.float32)
        nz, ny, nx = vol.shape
        u = np.asarray(units, np.float64)
        out = np.zeros((nz, ny, nx, 12), np.float32)
        self.L.orc_dense_ex(self.ctx, vol.ctypes.data, nx, ny, nz, u.ctypes.data,
                            1 if rotate else 0, out.ctypes.data)
        return out

    def gauss_taps(self, sigma):
        t = np.zeros(256, np.float32)
        w = self.L.orc_gauss_taps(sigma, t.ctypes.data, 256)
        return t[:w].copy()

    def blur(self, vol, taps, units=(1.0, 1.0, 1.0), unit=1.0, nc=1):
        vol = np.ascontiguousarray(vol, np.float32)
        nz, ny, nx = vol.shape[:3]
        u = np.asarray(units, np.float64)
        taps = np.ascontiguousarray(taps, np.float32)
        out = np.empty_like(vol)
        self.L.orc_blur(vol.ctypes.data, out.ctypes.data, nx, ny, nz, nc, u.ctypes.data,
                        taps.ctypes.data, len(taps), unit)
        return out

    def mesh(self):
        v = np.zeros((20, 3, 3), np.float32)
        idx = np.zeros((20, 3), np.int32)
        self.L.orc_mesh(v.ctypes.data, idx.ctypes.data)
        return v, idx

    def resample_affine(self, vol, A, out_shape, interp=0):
        """im_inv_transform for an affine 3x4 `A`; vol [z][y][x] or [z][y][x][c]."""
        vol = np.ascontiguousarray(vol, np.float32)
        nc = vol.shape[3] if vol.ndim == 4 else 1
        nz, ny, nx = vol.shape[:3]
        A = np.ascontiguousarray(A, np.float64).reshape(12)
        dnz, dny, dnx = out_shape
        out = np.zeros((dnz, dny, dnx) + ((nc,) if vol.ndim == 4 else ()), np.float32)
        if self.L.orc_resample_affine(vol.ctypes.data, nx, ny, nz, nc, A.ctypes.data, interp,
                                      out.ctypes.data, dnx, dny, dnz):
            raise RuntimeError("orc_resample_affine failed")
        return out

    def nn_match(self, d1, d2, nn_thresh=0.8):
        d1 = np.ascontiguousarray(d1, np.float32)
        d2 = np.ascontiguousarray(d2, np.float32)
        out = np.full(len(d1), -2, np.int32)
        if self.L.orc_nn_match(d1.ctypes.data, len(d1), d2.ctypes.data, len(d2), nn_thresh,
                               out.ctypes.data):
            raise RuntimeError("orc_nn_match failed")
        return out
