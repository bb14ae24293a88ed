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


def blob_volume_torch(shape, seed, device):
    """SURVEY.md Appendix C generator evaluated with torch on `device` (fast at 512^3): signed
    Gaussian blobs at sigma 2,3,4,6,8 + 0.01*U[0,1) noise, shifted to min 0.  Returns a float32
    [z][y][x] torch tensor.  (Same recipe as `blob_volume`, different random stream.)"""
    import torch
    import torch.nn.functional as F
    if isinstance(shape, int):
        shape = (shape, shape, shape)
    nz, ny, nx = shape
    nvox = nz * ny * nx
    g = torch.Generator(device="cpu").manual_seed(seed)
    vol = torch.zeros(shape, dtype=torch.float32, device=device)
    for sig in (2, 3, 4, 6, 8):
        k = max(1, int(nvox / (64 * sig ** 3)))
        idx = torch.stack([torch.randint(0, nz, (k,), generator=g),
                           torch.randint(0, ny, (k,), generator=g),
                           torch.randint(0, nx, (k,), generator=g)], 1)
        amp = (torch.rand(k, generator=g) * 2 - 1).float() * sig ** 3
        imp = torch.zeros(shape, dtype=torch.float32, device=device)
        imp.index_put_((idx[:, 0].to(device), idx[:, 1].to(device), idx[:, 2].to(device)),
                       amp.to(device), accumulate=True)
        r = int(4 * sig)
        x = torch.arange(-r, r + 1, dtype=torch.float32, device=device)
        w = torch.exp(-0.5 * (x / sig) ** 2)
        w = w / w.sum()
        t = imp[None, None]
        for ax in range(3):
            wshape = [1, 1, 1, 1, 1]
            wshape[2 + ax] = 2 * r + 1
            pad = [0, 0, 0, 0, 0, 0]
            pad[2 * (2 - ax)] = pad[2 * (2 - ax) + 1] = r
            t = F.conv3d(F.pad(t, pad, mode="reflect"), w.view(wshape))
        vol += t[0, 0]
        del imp, t
    noise = torch.rand(shape, generator=g, dtype=torch.float32)
    vol += 0.01 * noise.to(device)
    vol -= vol.min()
    return vol.contiguous()
