"""GPU parity of the resampling row: im_inv_transform with an affine map and
im_resample, through the C ABI of libsift3D.so, against the oracle restatement (itself pinned
bit-exact to the unmodified reference in tests/test_oracle.py).  LINEAR (trilinear, f64) must
be bit-identical; LANCZOS2 evaluates sin() with a different libm: tolerance 1e-6 absolute on
values of order 1 (measured ~1e-7: the f32 store dominates)."""
import numpy as np
import pytest

from test_oracle import AFFINES

pytestmark = pytest.mark.gpu
LANCZOS_TOL = 1e-6


@pytest.mark.parametrize("name", list(AFFINES))
def test_linear_bit_exact(b200_lib, oracle_cls, name):
    from sift3d_b200 import capi
    rng = np.random.default_rng(3)
    vol = rng.random((14, 17, 19), dtype=np.float32)
    for out_shape in ((16, 15, 23), None):
        got = capi.resample_affine(b200_lib, vol, AFFINES[name], out_shape, 0)
        want = oracle_cls().resample_affine(vol, AFFINES[name], out_shape or vol.shape, 0)
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    v4 = rng.random((6, 7, 8, 3), dtype=np.float32)       # channel-interleaved
    got = capi.resample_affine(b200_lib, v4, AFFINES[name], (7, 6, 9), 0)
    want = oracle_cls().resample_affine(v4, AFFINES[name], (7, 6, 9), 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("name", list(AFFINES))
def test_lanczos_within_tolerance(b200_lib, oracle_cls, name):
    from sift3d_b200 import capi
    vol = np.random.default_rng(4).random((12, 13, 15), dtype=np.float32)
    got = capi.resample_affine(b200_lib, vol, AFFINES[name], (13, 12, 16), 1)
    want = oracle_cls().resample_affine(vol, AFFINES[name], (13, 12, 16), 1)
    assert np.array_equal(got == 0, want == 0)            # same out-of-bounds set
    assert np.abs(got - want).max() <= LANCZOS_TOL


def test_im_resample_units(b200_lib, oracle_cls):
    """im_resample (imutil.c:2191): dims = ceil(n * u_in / u_out), diagonal affine, new units."""
    from sift3d_b200 import capi
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume((20, 24, 28), seed=2)
    u_in, u_out = (0.8, 1.0, 2.5), (1.0, 1.0, 1.0)
    got, units = capi.im_resample(b200_lib, vol, u_in, u_out, 0)
    f = [u_in[i] / u_out[i] for i in range(3)]
    dims = [int(np.ceil(n * fi)) for n, fi in zip((28, 24, 20), f)]   # x, y, z
    assert got.shape == (dims[2], dims[1], dims[0]) and units == u_out
    A = np.zeros((3, 4))
    for i in range(3):
        A[i, i] = 1.0 / f[i]
    want = oracle_cls().resample_affine(vol, A, got.shape, 0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_linear_large_matches_property(b200_lib):
    """Full-size property (no CPU oracle needed): the identity map returns the input bit for
    bit, and a pure integer shift returns the shifted input with zeros outside."""
    from sift3d_b200 import capi
    vol = np.random.default_rng(7).random((96, 128, 160), dtype=np.float32)
    same = capi.resample_affine(b200_lib, vol, AFFINES["identity"], None, 0)
    assert np.array_equal(same, vol)
    shift = [1, 0, 0, 3, 0, 1, 0, -2, 0, 0, 1, 5]
    got = capi.resample_affine(b200_lib, vol, shift, None, 0)
    want = np.zeros_like(vol)
    want[:96 - 5, 2:, :160 - 3] = vol[5:, :128 - 2, 3:]
    assert np.array_equal(got, want)


def test_inv_transform_honours_dst_strides(b200_lib, oracle_cls):
    """im_inv_transform writes through SIFT3D_IM_GET_VOX (imutil.c:2062-2076): a caller-owned
    dst with padded rows keeps its strides and gets the voxels where they say."""
    import ctypes as C
    from sift3d_b200 import capi
    rng = np.random.default_rng(4)
    vol = rng.random((12, 14, 16), dtype=np.float32)
    A = np.array([[0.9, 0.1, 0.0, 0.5], [-0.1, 1.0, 0.05, 0.2], [0.0, 0.02, 1.1, -0.3]])
    want = oracle_cls().resample_affine(vol, A, (10, 11, 13))
    big = np.full((10, 11, 20), -7.0, np.float32)          # rows padded from 13 to 20 floats
    src = capi.make_image(vol)
    dst = capi.make_image(big)
    dst.nx = 13
    dst.size = 10 * 11 * 13
    Aflat = np.ascontiguousarray(A, np.float64).reshape(12)
    f = b200_lib.lib.sift3d_b200_im_inv_transform_affine
    f.argtypes = [C.c_void_p, C.POINTER(capi.Image), C.c_int, C.c_int, C.POINTER(capi.Image)]
    f.restype = C.c_int
    assert f(Aflat.ctypes.data, C.byref(src), 0, 0, C.byref(dst)) == 0
    assert (dst.xs, dst.ys, dst.zs) == (1, 20, 220)       # untouched
    assert np.array_equal(big[:, :, :13].view(np.uint32), want.view(np.uint32))
    assert (big[:, :, 13:] == -7.0).all()                 # the padding was not written
