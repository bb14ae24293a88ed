"""Synthetic input volumes shared by tests and bench.

`blob_volume` is the generator the survey used for its CPU probes: a sum of
signed Gaussian blobs at several scales plus a little uniform noise, giving a
realistic ~0.15 % candidate rate.  Deterministic for a given (n, seed).
"""
from __future__ import annotations

import numpy as np


def blob_volume(shape, seed: int = 1234) -> np.ndarray:
    import scipy.ndimage as ndi
    if isinstance(shape, int):
        shape = (shape, shape, shape)
    nz, ny, nx = shape
    rng = np.random.Generator(np.random.PCG64(seed))
    vol = np.zeros(shape, np.float32)
    nvox = nz * ny * nx
    for sig in (2, 3, 4, 6, 8):
        k = max(1, int(nvox / (64 * sig ** 3)))
        imp = np.zeros(shape, np.float32)
        idx = np.stack([rng.integers(0, nz, k), rng.integers(0, ny, k), rng.integers(0, nx, k)], 1)
        amp = rng.uniform(-1, 1, k).astype(np.float32) * sig ** 3
        np.add.at(imp, tuple(idx.T), amp)
        vol += ndi.gaussian_filter(imp, sig, mode="reflect")
    vol += 0.01 * rng.random(shape, dtype=np.float32)
    vol -= vol.min()
    return np.ascontiguousarray(vol, np.float32)


def noise_volume(shape, seed: int = 1) -> np.ndarray:
    """White noise U[0,1): the keypoint-dense stress case."""
    if isinstance(shape, int):
        shape = (shape, shape, shape)
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.random(shape, dtype=np.float32)


def smooth_noise_volume(shape, seed: int = 0, sigma: float = 2.0) -> np.ndarray:
    import scipy.ndimage as ndi
    if isinstance(shape, int):
        shape = (shape, shape, shape)
    rng = np.random.Generator(np.random.PCG64(seed))
    v = ndi.gaussian_filter(rng.standard_normal(shape).astype(np.float32), sigma)
    v = v.astype(np.float32)
    v -= v.min()
    return np.ascontiguousarray(v)
