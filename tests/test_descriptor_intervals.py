"""Host-side check of the row intervals of the descriptor kernel (k_descriptor2, phase A).

The kernel visits, per (y, z) row of a keypoint's window, ONE x interval: a superset computed in
approximate arithmetic (sphere chord intersected with three slabs, widened) that is then trimmed
inwards with the exact f32 tests of the reference (sphere `sift.c:96-119`, descriptor cube
`sift.c:1870-1880`).  That is only correct if (1) the voxels passing the exact tests form one
interval and (2) the approximate interval contains it.  Both are properties of f32 arithmetic,
so they are checked here with numpy float32 (one rounding per operation, like the kernel's
`__fmul_rn` / `__fadd_rn` code) over random centres, scales, rotations and anisotropic units --
a missed voxel would change a descriptor by ~1e-5, below what the GPU parity tolerance can see.
"""
import numpy as np

f32 = np.float32


def _rotation(rng):
    q, _ = np.linalg.qr(rng.standard_normal((3, 3)))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q.astype(f32)


def _bounds(c, rad, u, n):
    lo = int(max(np.floor(f32(c) - f32(rad) / f32(u)), f32(1.0)))
    hi = int(min(np.ceil(f32(c) + f32(rad) / f32(u)), f32(n - 2)))
    return lo, hi


def _check_keypoint(rng, kp, sd, R, units, dims):
    nx, ny, nz = dims
    uxf, uyf, uzf = (f32(u) for u in units)
    iux = f32(1.0) / uxf
    sigma = f32(np.float64(sd) * 7.071067812)
    win_radius = f32(2.0 * np.float64(sigma))
    half = f32(np.float64(win_radius) / np.sqrt(2.0))
    hist_width = (f32(2.0) * half) / f32(4.0)
    bin_fctr = f32(1.0) / hist_width
    r2 = win_radius * win_radius
    kx, ky, kz = (f32(v) for v in kp)
    x0, x1 = _bounds(kx, win_radius, uxf, nx)
    y0, y1 = _bounds(ky, win_radius, uyf, ny)
    z0, z1 = _bounds(kz, win_radius, uzf, nz)
    if x1 < x0 or y1 < y0 or z1 < z0:
        return 0
    Rt = R.T.copy().reshape(-1)  # Rt[3*i + j] = R[j][i], as in the kernel
    yy, zz = np.meshgrid(np.arange(y0, y1 + 1), np.arange(z0, z1 + 1), indexing="ij")
    yy, zz = yy.reshape(-1), zz.reshape(-1)
    # ---- phase A, approximate interval (same operation order, f32)
    vy = (yy.astype(f32) - ky) * uyf
    vz = (zz.astype(f32) - kz) * uzf
    rem = r2 - (vy * vy + vz * vz)
    ok = rem >= f32(-1e-3) * r2
    hx = np.sqrt(np.maximum(rem, f32(0.0))) * iux
    lo, hi = -hx, hx.copy()
    empty = np.zeros(len(yy), bool)
    for a in range(3):
        sl = Rt[3 * a] * uxf * bin_fctr
        off = ((Rt[3 * a + 1] * vy + Rt[3 * a + 2] * vz) + half) * bin_fctr
        if abs(sl) < f32(1e-6):
            empty |= (off < f32(-1e-3)) | (off > f32(4.001))
        else:
            t0, t1 = (f32(0.0) - off) / sl, (f32(4.0) - off) / sl
            lo = np.maximum(lo, np.minimum(t0, t1))
            hi = np.minimum(hi, np.maximum(t0, t1))
    have = ok & ~empty & (lo <= hi + f32(1.0))
    xa = np.maximum(x0, np.floor(kx + lo).astype(np.int64) - 1)
    xb = np.minimum(x1, np.ceil(kx + hi).astype(np.int64) + 1)
    have &= xb >= xa
    # ---- exact tests for every voxel of the box (reference operation order)
    xs = np.arange(x0, x1 + 1)
    vx = ((xs.astype(f32) - kx) * uxf)[None, :]
    evy = ((yy.astype(f32) - ky) * uyf)[:, None]
    evz = ((zz.astype(f32) - kz) * uzf)[:, None]
    sq = (vx * vx + evy * evy) + evz * evz
    passed = ~(sq > r2)
    for a in range(3):
        vk = (Rt[3 * a] * vx + Rt[3 * a + 1] * evy) + Rt[3 * a + 2] * evz
        vb = (vk + half) * bin_fctr
        passed &= ~((vb < f32(0.0)) | (vb >= f32(4.0)))
    cnt = passed.sum(axis=1)
    first = np.where(cnt > 0, passed.argmax(axis=1), 0) + x0
    last = x1 - np.where(cnt > 0, passed[:, ::-1].argmax(axis=1), 0)
    # (1) one interval per row
    assert np.all((cnt == 0) | (last - first + 1 == cnt)), "passing voxels of a row are not contiguous"
    # (2) contained in the approximate, widened interval
    nonempty = cnt > 0
    assert np.all(have[nonempty]), "a row with passing voxels was declared empty"
    assert np.all(xa[nonempty] <= first[nonempty]) and np.all(last[nonempty] <= xb[nonempty]), \
        "the approximate interval misses passing voxels"
    return int(cnt.sum())


def test_row_intervals_are_exact_supersets():
    rng = np.random.default_rng(2024)
    total = 0
    unit_sets = [(1.0, 1.0, 1.0), (2.0, 2.0, 2.0), (1.0, 1.3, 0.8), (0.7, 1.0, 2.5), (4.0, 4.0, 8.0)]
    for case in range(240):
        units = unit_sets[case % len(unit_sets)]
        dims = (int(rng.integers(40, 160)), int(rng.integers(40, 160)), int(rng.integers(40, 160)))
        kp = [float(rng.integers(2, d - 2)) for d in dims]
        if case % 3 == 2:  # raw-image API: non-integer centres
            kp = [v + float(rng.random()) * 0.9 for v in kp]
        sd = float(rng.uniform(1.3, 4.2)) * min(units)
        R = _rotation(rng)
        if case % 7 == 0:  # axis-aligned frames: slabs that do not depend on x
            R = np.eye(3, dtype=f32)[rng.permutation(3)]
        total += _check_keypoint(rng, kp, sd, R, units, dims)
    assert total > 1000000
