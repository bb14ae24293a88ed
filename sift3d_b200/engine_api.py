"""ctypes access to the kernel-level entry points of libsift3d_cuda.so (tests / bench)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .capi import CUDA_LIB


class Engine:
    def __init__(self, device: int = -1):
        if not CUDA_LIB.exists():
            raise FileNotFoundError(f"{CUDA_LIB} missing: run __graft_entry__.build()")
        L = self.L = C.CDLL(str(CUDA_LIB))
        L.s3d_engine_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
        L.s3d_engine_destroy.argtypes = [C.c_void_p]
        L.s3d_engine_destroy.restype = None
        L.s3d_engine_error.argtypes = [C.c_void_p]
        L.s3d_engine_error.restype = C.c_char_p
        L.s3d_last_create_error.restype = C.c_char_p
        L.s3d_engine_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.s3d_engine_sync.argtypes = [C.c_void_p]
        L.s3d_engine_launch_count.argtypes = [C.c_void_p]
        L.s3d_engine_launch_count.restype = C.c_longlong
        L.s3d_set_blur_mode.argtypes = [C.c_void_p, C.c_int]
        L.s3d_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.s3d_dev_alloc.argtypes = [C.c_void_p, C.c_size_t]
        L.s3d_dev_alloc.restype = C.c_void_p
        L.s3d_dev_free.argtypes = [C.c_void_p, C.c_void_p]
        L.s3d_dev_free.restype = None
        L.s3d_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.s3d_memcpy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.s3d_host_roundtrip.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.s3d_blur_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double,
                                      C.c_void_p]
        h = C.c_void_p()
        if L.s3d_engine_create(C.byref(h), device) != 0:
            raise RuntimeError("s3d_engine_create failed: " +
                               L.s3d_last_create_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.s3d_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {self.L.s3d_engine_error(self.h).decode()}")

    def set_stream(self, stream_ptr):
        self._check(self.L.s3d_engine_set_stream(self.h, stream_ptr), "set_stream")

    def sync(self):
        self._check(self.L.s3d_engine_sync(self.h), "sync")

    def host_roundtrip(self, src: np.ndarray) -> np.ndarray:
        """src -> device -> a fresh host array through the library's pageable-buffer copy paths."""
        src = np.ascontiguousarray(src)
        dst = np.empty_like(src)
        self._check(self.L.s3d_host_roundtrip(self.h, src.ctypes.data, dst.ctypes.data, src.nbytes),
                    "host_roundtrip")
        return dst

    def set_option(self, name: str, value: int):
        self._check(self.L.s3d_set_option(self.h, name.encode(), int(value)), f"set_option {name}")

    def launches(self) -> int:
        return int(self.L.s3d_engine_launch_count(self.h))

    def blur_device(self, d_src, d_dst, nx, ny, nz, taps, units=(1.0, 1.0, 1.0), unit=1.0, nc=1,
                    mode=0):
        taps = np.ascontiguousarray(taps, np.float32)
        u = np.asarray(units, np.float64)
        self.L.s3d_set_blur_mode(self.h, mode)
        self._check(self.L.s3d_blur_device(self.h, d_src, d_dst, nx, ny, nz, nc,
                                           taps.ctypes.data, len(taps), unit, u.ctypes.data),
                    "s3d_blur_device")

    def blur(self, vol, taps, units=(1.0, 1.0, 1.0), unit=1.0, nc=1, mode=0):
        """Host array in, host array out (kernel-level parity tests)."""
        vol = np.ascontiguousarray(vol, np.float32)
        nz, ny, nx = vol.shape[:3]
        nb = vol.nbytes
        a = self.L.s3d_dev_alloc(self.h, nb)
        b = self.L.s3d_dev_alloc(self.h, nb)
        try:
            self._check(self.L.s3d_memcpy_h2d(self.h, a, vol.ctypes.data, nb), "h2d")
            self.blur_device(a, b, nx, ny, nz, taps, units, unit, nc, mode)
            out = np.empty_like(vol)
            self._check(self.L.s3d_memcpy_d2h(self.h, out.ctypes.data, b, nb), "d2h")
        finally:
            self.L.s3d_dev_free(self.h, a)
            self.L.s3d_dev_free(self.h, b)
        return out
