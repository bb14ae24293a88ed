"""ctypes mirror of the SIFT3D C ABI (structs + the hot-path entry points).

This is the host-side mirror of the reference interface for the accelerated path:
the same structs (`Image`, `Keypoint`, `Keypoint_store`, `SIFT3D_Descriptor`,
`SIFT3D`, ... -- reference `imutil/imtypes.h:136-334`) and the same function
names and argument meaning (`sift3d/sift.h:19-108`).  Because the layouts are
byte-identical, ONE wrapper class drives either

  * the B200 library built from this repo (`sift3d_b200/lib/libsift3D.so`), or
  * any other build of the same ABI -- the tests hand it the unmodified reference
    (`oracle/oracle_api.py:load_reference`, test infrastructure outside this package),

which is what makes the parity tests read like "same calls, two libraries".

Nothing here computes: every number comes out of the loaded shared library.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
LIB_DIR = Path(__file__).resolve().parent / "lib"
B200_LIB = LIB_DIR / "libsift3D.so"
CUDA_LIB = LIB_DIR / "libsift3d_cuda.so"

NHIST_PER_DIM = 4
ICOS_NVERT = 12
DESC_NUM_TOTAL_HIST = NHIST_PER_DIM ** 3
DESC_NUMEL = DESC_NUM_TOTAL_HIST * ICOS_NVERT  # 768

SIFT3D_SUCCESS = 0
SIFT3D_FAILURE = -1


class Mat_rm(C.Structure):  # imtypes.h:136-149  (32 bytes)
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t), ("num_cols", C.c_int),
                ("num_rows", C.c_int), ("static_mem", C.c_int), ("type", C.c_int)]


class Image(C.Structure):  # imtypes.h:156-168  (104 bytes)
    _fields_ = [("data", C.POINTER(C.c_float)), ("cl_image", C.c_int), ("s", C.c_double),
                ("size", C.c_size_t), ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
                ("ux", C.c_double), ("uy", C.c_double), ("uz", C.c_double),
                ("xs", C.c_size_t), ("ys", C.c_size_t), ("zs", C.c_size_t),
                ("nc", C.c_int), ("cl_valid", C.c_int)]


class Sep_FIR_filter(C.Structure):  # imtypes.h:171-179
    _fields_ = [("cl_apply_unrolled", C.c_int), ("kernel", C.POINTER(C.c_float)),
                ("dim", C.c_int), ("width", C.c_int), ("symmetric", C.c_int)]


class Gauss_filter(C.Structure):  # imtypes.h:182-187
    _fields_ = [("sigma", C.c_double), ("f", Sep_FIR_filter)]


class GSS_filters(C.Structure):  # imtypes.h:190-197
    _fields_ = [("first_gauss", Gauss_filter), ("gauss_octave", C.POINTER(Gauss_filter)),
                ("num_filters", C.c_int), ("first_level", C.c_int)]


class Pyramid(C.Structure):  # imtypes.h:207-223  (48 bytes)
    _fields_ = [("levels", C.POINTER(Image)), ("sigma_n", C.c_double), ("sigma0", C.c_double),
                ("num_kp_levels", C.c_int), ("first_octave", C.c_int), ("num_octaves", C.c_int),
                ("first_level", C.c_int), ("num_levels", C.c_int)]


class Slab(C.Structure):  # imtypes.h:244-250
    _fields_ = [("buf", C.c_void_p), ("num", C.c_size_t), ("buf_size", C.c_size_t)]


class Keypoint(C.Structure):  # imtypes.h:253-261  (112 bytes)
    _fields_ = [("r_data", C.c_float * 9), ("R", Mat_rm), ("xd", C.c_double), ("yd", C.c_double),
                ("zd", C.c_double), ("sd", C.c_double), ("o", C.c_int), ("s", C.c_int)]


class Keypoint_store(C.Structure):  # imtypes.h:264-270  (48 bytes)
    _fields_ = [("buf", C.POINTER(Keypoint)), ("slab", Slab), ("nx", C.c_int), ("ny", C.c_int),
                ("nz", C.c_int)]


class SIFT3D_Descriptor(C.Structure):  # imtypes.h:291-296  (3104 bytes)
    _fields_ = [("hists", C.c_float * DESC_NUMEL), ("xd", C.c_double), ("yd", C.c_double),
                ("zd", C.c_double), ("sd", C.c_double)]


class SIFT3D_Descriptor_store(C.Structure):  # imtypes.h:299-305  (32 bytes)
    _fields_ = [("buf", C.POINTER(SIFT3D_Descriptor)), ("num", C.c_size_t), ("nx", C.c_int),
                ("ny", C.c_int), ("nz", C.c_int)]


class Mesh(C.Structure):  # imtypes.h:285-288
    _fields_ = [("tri", C.c_void_p), ("num", C.c_int)]


class SIFT3D(C.Structure):  # imtypes.h:309-334  (304 bytes)
    _fields_ = [("mesh", Mesh), ("gss", GSS_filters), ("kernels", C.c_int), ("gpyr", Pyramid),
                ("dog", Pyramid), ("im", Image), ("peak_thresh", C.c_double),
                ("corner_thresh", C.c_double), ("dense_rotate", C.c_int)]


ABI_SIZES = {Mat_rm: 32, Image: 104, Pyramid: 48, Keypoint: 112, Keypoint_store: 48,
             SIFT3D_Descriptor: 3104, SIFT3D_Descriptor_store: 32, SIFT3D: 304,
             GSS_filters: 56, Slab: 24}
for _t, _n in ABI_SIZES.items():
    assert C.sizeof(_t) == _n, (_t.__name__, C.sizeof(_t), _n)

KEYPOINT_DTYPE = np.dtype({
    "names": ["R", "xd", "yd", "zd", "sd", "o", "s"],
    "formats": [("<f4", (3, 3)), "<f8", "<f8", "<f8", "<f8", "<i4", "<i4"],
    "offsets": [0, 72, 80, 88, 96, 104, 108], "itemsize": 112})
DESCRIPTOR_DTYPE = np.dtype({
    "names": ["hists", "xd", "yd", "zd", "sd"],
    "formats": [("<f4", (DESC_NUMEL,)), "<f8", "<f8", "<f8", "<f8"],
    "offsets": [0, 3072, 3080, 3088, 3096], "itemsize": 3104})

# every function sift.h declares (sift3d/sift.h:19-108) -- the drop-in symbol set
SIFT_H_SYMBOLS = [
    "init_Keypoint_store", "init_Keypoint", "resize_Keypoint_store", "copy_Keypoint",
    "cleanup_Keypoint_store", "init_SIFT3D_Descriptor_store", "cleanup_SIFT3D_Descriptor_store",
    "set_peak_thresh_SIFT3D", "set_corner_thresh_SIFT3D", "set_num_kp_levels_SIFT3D",
    "set_sigma_n_SIFT3D", "set_sigma0_SIFT3D", "init_SIFT3D", "copy_SIFT3D", "cleanup_SIFT3D",
    "print_opts_SIFT3D", "parse_args_SIFT3D", "SIFT3D_assign_orientations",
    "SIFT3D_detect_keypoints", "SIFT3D_have_gpyr", "SIFT3D_extract_descriptors",
    "SIFT3D_extract_raw_descriptors", "SIFT3D_extract_dense_descriptors", "SIFT3D_nn_match",
    "Keypoint_store_to_Mat_rm", "SIFT3D_Descriptor_coords_to_Mat_rm",
    "SIFT3D_Descriptor_store_to_Mat_rm", "Mat_rm_to_SIFT3D_Descriptor_store",
    "SIFT3D_matches_to_Mat_rm", "draw_matches", "write_Keypoint_store",
    "write_SIFT3D_Descriptor_store",
]


def make_image(vol: np.ndarray, units=(1.0, 1.0, 1.0), nc: int = 1) -> Image:
    """Wrap a C-contiguous float32 array [z][y][x] (or [z][y][x][c]) as an `Image`
    (no copy; the caller keeps `vol` alive).  Strides are the reference's default
    (`im_default_stride`, imutil.c:1453): xs=nc, ys=nc*nx, zs=nc*nx*ny."""
    assert vol.dtype == np.float32 and vol.flags.c_contiguous
    if nc == 1:
        assert vol.ndim == 3
        nz, ny, nx = vol.shape
    else:
        assert vol.ndim == 4 and vol.shape[3] == nc
        nz, ny, nx, _ = vol.shape
    im = Image()
    im.data = vol.ctypes.data_as(C.POINTER(C.c_float))
    im.cl_image = 0
    im.s = -1.0
    im.size = nx * ny * nz * nc
    im.nx, im.ny, im.nz = nx, ny, nz
    im.ux, im.uy, im.uz = units
    im.xs, im.ys, im.zs = nc, nc * nx, nc * nx * ny
    im.nc = nc
    im.cl_valid = 0
    return im


def empty_image() -> Image:
    """Equivalent of `init_im` (imutil.c:3626-3648)."""
    im = Image()
    C.memset(C.byref(im), 0, C.sizeof(im))
    im.ux = im.uy = im.uz = 1.0
    im.s = -1.0
    return im


@dataclass
class DetectResult:
    keypoints: np.ndarray  # KEYPOINT_DTYPE
    nx: int
    ny: int
    nz: int


class Sift3DLib:
    """One loaded implementation of the SIFT3D C API (B200 build or reference)."""

    def __init__(self, path: os.PathLike | str = B200_LIB, name: str | None = None):
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(
                f"{path} is missing - run `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path = path
        self.name = name or path.stem
        self.lib = C.CDLL(str(path))
        L = self.lib
        P = C.POINTER
        L.init_SIFT3D.argtypes = [P(SIFT3D)]
        L.init_SIFT3D.restype = C.c_int
        L.cleanup_SIFT3D.argtypes = [P(SIFT3D)]
        L.cleanup_SIFT3D.restype = None
        L.copy_SIFT3D.argtypes = [P(SIFT3D), P(SIFT3D)]
        L.copy_SIFT3D.restype = C.c_int
        for fn in ("set_peak_thresh_SIFT3D", "set_corner_thresh_SIFT3D", "set_sigma_n_SIFT3D",
                   "set_sigma0_SIFT3D"):
            getattr(L, fn).argtypes = [P(SIFT3D), C.c_double]
            getattr(L, fn).restype = C.c_int
        L.set_num_kp_levels_SIFT3D.argtypes = [P(SIFT3D), C.c_uint]
        L.set_num_kp_levels_SIFT3D.restype = C.c_int
        L.init_Keypoint_store.argtypes = [P(Keypoint_store)]
        L.init_Keypoint_store.restype = None
        L.cleanup_Keypoint_store.argtypes = [P(Keypoint_store)]
        L.cleanup_Keypoint_store.restype = None
        L.resize_Keypoint_store.argtypes = [P(Keypoint_store), C.c_size_t]
        L.resize_Keypoint_store.restype = C.c_int
        L.init_SIFT3D_Descriptor_store.argtypes = [P(SIFT3D_Descriptor_store)]
        L.init_SIFT3D_Descriptor_store.restype = None
        L.cleanup_SIFT3D_Descriptor_store.argtypes = [P(SIFT3D_Descriptor_store)]
        L.cleanup_SIFT3D_Descriptor_store.restype = None
        L.SIFT3D_detect_keypoints.argtypes = [P(SIFT3D), P(Image), P(Keypoint_store)]
        L.SIFT3D_detect_keypoints.restype = C.c_int
        L.SIFT3D_extract_descriptors.argtypes = [P(SIFT3D), P(Keypoint_store),
                                                 P(SIFT3D_Descriptor_store)]
        L.SIFT3D_extract_descriptors.restype = C.c_int
        L.SIFT3D_extract_raw_descriptors.argtypes = [P(SIFT3D), P(Image), P(Keypoint_store),
                                                     P(SIFT3D_Descriptor_store)]
        L.SIFT3D_extract_raw_descriptors.restype = C.c_int
        L.SIFT3D_extract_dense_descriptors.argtypes = [P(SIFT3D), P(Image), P(Image)]
        L.SIFT3D_extract_dense_descriptors.restype = C.c_int
        L.SIFT3D_assign_orientations.argtypes = [P(SIFT3D), P(Image), P(Keypoint_store),
                                                 P(P(C.c_double))]
        L.SIFT3D_assign_orientations.restype = C.c_int
        L.SIFT3D_have_gpyr.argtypes = [P(SIFT3D)]
        L.SIFT3D_have_gpyr.restype = C.c_int
        L.SIFT3D_nn_match.argtypes = [P(SIFT3D_Descriptor_store), P(SIFT3D_Descriptor_store),
                                      C.c_float, P(P(C.c_int))]
        L.SIFT3D_nn_match.restype = C.c_int
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]

    def exports(self, sym: str) -> bool:
        try:
            getattr(self.lib, sym)
            return True
        except AttributeError:
            return False


class Sift3D:
    """A `SIFT3D` object (reference `init_SIFT3D` ... `cleanup_SIFT3D` lifecycle)."""

    def __init__(self, lib: Sift3DLib, peak_thresh=None, corner_thresh=None, num_kp_levels=None,
                 sigma_n=None, sigma0=None):
        self.lib = lib
        self.L = lib.lib
        self.s = SIFT3D()
        if self.L.init_SIFT3D(C.byref(self.s)) != 0:
            raise RuntimeError("init_SIFT3D failed")
        self._alive = True
        if sigma_n is not None and self.L.set_sigma_n_SIFT3D(C.byref(self.s), sigma_n):
            raise ValueError("sigma_n")
        if sigma0 is not None and self.L.set_sigma0_SIFT3D(C.byref(self.s), sigma0):
            raise ValueError("sigma0")
        if peak_thresh is not None and self.L.set_peak_thresh_SIFT3D(C.byref(self.s), peak_thresh):
            raise ValueError("peak_thresh")
        if corner_thresh is not None and \
                self.L.set_corner_thresh_SIFT3D(C.byref(self.s), corner_thresh):
            raise ValueError("corner_thresh")
        if num_kp_levels is not None and \
                self.L.set_num_kp_levels_SIFT3D(C.byref(self.s), num_kp_levels):
            raise ValueError("num_kp_levels")
        self.kp = Keypoint_store()
        self.L.init_Keypoint_store(C.byref(self.kp))
        self.desc = SIFT3D_Descriptor_store()
        self.L.init_SIFT3D_Descriptor_store(C.byref(self.desc))

    # -- the three hot API calls ------------------------------------------------
    def detect_keypoints(self, vol: np.ndarray, units=(1.0, 1.0, 1.0), copy: bool = True):
        """`SIFT3D_detect_keypoints` (sift.c:1609).  Returns a KEYPOINT_DTYPE array."""
        im = make_image(np.ascontiguousarray(vol, np.float32), units)
        rc = self.L.SIFT3D_detect_keypoints(C.byref(self.s), C.byref(im), C.byref(self.kp))
        if rc != 0:
            raise RuntimeError(f"SIFT3D_detect_keypoints returned {rc}")
        return self.keypoints(copy)

    def keypoints(self, copy: bool = True) -> np.ndarray:
        """Keypoint records; copy=False returns a view of the store (valid until the next call)."""
        n = self.kp.slab.num
        if n == 0:
            return np.zeros(0, KEYPOINT_DTYPE)
        buf = (C.c_char * (n * 112)).from_address(C.addressof(self.kp.buf.contents))
        a = np.frombuffer(buf, dtype=KEYPOINT_DTYPE, count=n)
        return a.copy() if copy else a

    def extract_descriptors(self, copy: bool = True) -> np.ndarray:
        """`SIFT3D_extract_descriptors` (sift.c:2025) on the stored keypoints."""
        rc = self.L.SIFT3D_extract_descriptors(C.byref(self.s), C.byref(self.kp),
                                               C.byref(self.desc))
        if rc != 0:
            raise RuntimeError(f"SIFT3D_extract_descriptors returned {rc}")
        return self.descriptors(copy)

    def descriptors(self, copy: bool = True) -> np.ndarray:
        n = self.desc.num
        buf = (C.c_char * (n * 3104)).from_address(C.addressof(self.desc.buf.contents))
        a = np.frombuffer(buf, dtype=DESCRIPTOR_DTYPE, count=n)
        return a.copy() if copy else a

    def extract_dense_descriptors(self, vol: np.ndarray, units=(1.0, 1.0, 1.0)) -> np.ndarray:
        """`SIFT3D_extract_dense_descriptors` (sift.c:2354).  Returns [z][y][x][12] float32."""
        vol = np.ascontiguousarray(vol, np.float32)
        im = make_image(vol, units)
        out = empty_image()
        rc = self.L.SIFT3D_extract_dense_descriptors(C.byref(self.s), C.byref(im), C.byref(out))
        if rc != 0:
            raise RuntimeError(f"SIFT3D_extract_dense_descriptors returned {rc}")
        n = out.nx * out.ny * out.nz * out.nc
        arr = np.ctypeslib.as_array(out.data, shape=(n,)).reshape(
            out.nz, out.ny, out.nx, out.nc).copy()
        self.lib._libc.free(C.cast(out.data, C.c_void_p))
        return arr

    # -- introspection used by the parity tests ----------------------------------
    def _pyr_level(self, pyr: Pyramid, o: int, s: int) -> Image:
        idx = (o - pyr.first_octave) * pyr.num_levels + (s - pyr.first_level)
        return pyr.levels[idx]

    def num_octaves(self) -> int:
        return self.s.gpyr.num_octaves

    def level_meta(self, which: str, o: int, s: int):
        im = self._pyr_level(self.s.gpyr if which == "gpyr" else self.s.dog, o, s)
        return dict(nx=im.nx, ny=im.ny, nz=im.nz, ux=im.ux, uy=im.uy, uz=im.uz, s=im.s)

    def level_data(self, which: str, o: int, s: int) -> np.ndarray:
        """Host copy of pyramid level (o, s).  For the B200 build this asks the library
        to materialise the level from HBM (`sift3d_b200_fetch_level`)."""
        pyr = self.s.gpyr if which == "gpyr" else self.s.dog
        im = self._pyr_level(pyr, o, s)
        n = im.nx * im.ny * im.nz
        if self.lib.exports("sift3d_b200_fetch_level"):
            out = np.empty((im.nz, im.ny, im.nx), np.float32)
            f = self.L.sift3d_b200_fetch_level
            f.argtypes = [C.POINTER(SIFT3D), C.c_int, C.c_int, C.c_int, C.c_void_p]
            f.restype = C.c_int
            rc = f(C.byref(self.s), 0 if which == "gpyr" else 1, o, s, out.ctypes.data)
            if rc != 0:
                raise RuntimeError("sift3d_b200_fetch_level failed")
            return out
        return np.ctypeslib.as_array(im.data, shape=(n,)).reshape(im.nz, im.ny, im.nx).copy()

    def close(self):
        if self._alive:
            self.L.cleanup_Keypoint_store(C.byref(self.kp))
            self.L.cleanup_SIFT3D_Descriptor_store(C.byref(self.desc))
            self.L.cleanup_SIFT3D(C.byref(self.s))
            self._alive = False

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def resample_affine(lib: Sift3DLib, vol: np.ndarray, A, out_shape=None, interp: int = 0,
                    units=(1.0, 1.0, 1.0)) -> np.ndarray:
    """`im_inv_transform` (imutil.c:2040) for an affine 3x4 matrix through
    `sift3d_b200_im_inv_transform_affine`; out_shape None = same dims as the input (resize)."""
    vol = np.ascontiguousarray(vol, np.float32)
    nc = vol.shape[3] if vol.ndim == 4 else 1
    src = make_image(vol, units, nc)
    dst = empty_image()
    A = np.ascontiguousarray(A, np.float64).reshape(12)
    f = lib.lib.sift3d_b200_im_inv_transform_affine
    f.argtypes = [C.c_void_p, C.POINTER(Image), C.c_int, C.c_int, C.POINTER(Image)]
    f.restype = C.c_int
    out = None
    if out_shape is not None:
        dnz, dny, dnx = out_shape
        out = np.zeros((dnz, dny, dnx) + ((nc,) if vol.ndim == 4 else ()), np.float32)
        dst = make_image(out, units, nc)        # `out` stays alive until we return it
    rc = f(A.ctypes.data, C.byref(src), interp, 1 if out_shape is None else 0, C.byref(dst))
    if rc != 0:
        raise RuntimeError(f"sift3d_b200_im_inv_transform_affine returned {rc}")
    if out is not None:
        return out
    n = dst.nx * dst.ny * dst.nz * dst.nc
    arr = np.ctypeslib.as_array(dst.data, shape=(n,)).reshape(
        (dst.nz, dst.ny, dst.nx) + ((dst.nc,) if vol.ndim == 4 else ())).copy()
    lib._libc.free(C.cast(dst.data, C.c_void_p))
    return arr


def im_resample(lib: Sift3DLib, vol: np.ndarray, units_in, units_out, interp: int = 0):
    """`im_resample` (imutil.c:2191) through `sift3d_b200_im_resample`."""
    vol = np.ascontiguousarray(vol, np.float32)
    src = make_image(vol, units_in)
    dst = empty_image()
    u = np.asarray(units_out, np.float64)
    f = lib.lib.sift3d_b200_im_resample
    f.argtypes = [C.POINTER(Image), C.c_void_p, C.c_int, C.POINTER(Image)]
    f.restype = C.c_int
    if f(C.byref(src), u.ctypes.data, interp, C.byref(dst)) != 0:
        raise RuntimeError("sift3d_b200_im_resample failed")
    n = dst.nx * dst.ny * dst.nz
    arr = np.ctypeslib.as_array(dst.data, shape=(n,)).reshape(dst.nz, dst.ny, dst.nx).copy()
    meta = (dst.ux, dst.uy, dst.uz)
    lib._libc.free(C.cast(dst.data, C.c_void_p))
    return arr, meta


def load_b200() -> Sift3DLib:
    return Sift3DLib(B200_LIB, "b200")
