"""CPU tests: the oracle restatement is pinned against (a) golden fixtures generated
from the unmodified reference and (b) the compiled reference itself (oracle/_ref)
when it is present.  No GPU needed."""
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN_CASES, load_golden


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(oracle_cls, name):
    g = load_golden(name)
    orc = oracle_cls(**g["kwargs"])
    kp = orc.detect(g["input"], tuple(g["units"]))
    assert orc.num_octaves() == int(g["noct"])
    # every pyramid level bit-identical to the reference's
    K = g["kwargs"]["num_kp_levels"]
    want = {}
    for row in g["level_sha256"]:
        w, o, lv, h = str(row).split(",")
        want[(w, int(o), int(lv))] = h
    for o in range(orc.num_octaves()):
        for lv in range(-1, K + 2):
            assert digest(orc.level("gpyr", o, lv)[0]) == want[("g", o, lv)], ("gpyr", o, lv)
        for lv in range(-1, K + 1):
            assert digest(orc.level("dog", o, lv)[0]) == want[("d", o, lv)], ("dog", o, lv)
    # keypoints: identical integer coordinates / octave / level, in the same order
    assert len(kp) == len(g["kp_xd"])
    for f in ("xd", "yd", "zd", "o", "s", "sd"):
        assert np.array_equal(kp[f], g["kp_" + f]), f
    assert np.abs(kp["R"] - g["kp_R"]).max() <= 1e-6  # Jacobi vs LAPACK dsyevd
    if len(kp):
        desc, coords = orc.describe(kp)
        rel = np.linalg.norm(desc - g["desc"], axis=1) / np.linalg.norm(g["desc"], axis=1)
        assert rel.max() <= 1e-6, rel.max()
        assert np.array_equal(coords, g["desc_coords"])
    if "dense" in g:
        sl = tuple(slice(int(a), int(b)) for a, b in g["dense_in_slices"])
        sub = np.ascontiguousarray(g["input"][sl])
        dd = orc.dense(sub, tuple(g["units"]))
        assert np.abs(dd - g["dense"]).max() <= 1e-6 * max(1.0, np.abs(g["dense"]).max())


def test_oracle_dense_rotate_matches_golden(oracle_cls):
    """dense_rotate = 1 (sift.c:2521-2588): per-voxel orientation + rotated histogram."""
    from conftest import GOLDEN
    z = np.load(GOLDEN / "dense_rotate.npz")
    orc = oracle_cls()
    for key in ("iso", "aniso"):
        vol, units, want = z[key + "_input"], tuple(z[key + "_units"]), z[key + "_dense"]
        got = orc.dense(vol, units, rotate=True)
        assert got.shape == want.shape
        assert np.array_equal(got, want), np.abs(got - want).max()
        # the rotation matters: the no-rotate path gives a different field
        assert np.abs(orc.dense(vol, units) - want).max() > 1e-3


def test_oracle_matches_compiled_reference(oracle_cls, ref_lib):
    """Same calls on the reference library and on the restatement, fresh random input."""
    from sift3d_b200 import capi
    from sift3d_b200.volumes import smooth_noise_volume
    vol = smooth_noise_volume((33, 41, 38), seed=11)
    units = (1.0, 0.5, 2.0)
    orc = oracle_cls()
    okp = orc.detect(vol, units)
    with capi.Sift3D(ref_lib) as s:
        kp = s.detect_keypoints(vol, units)
        assert len(kp) == len(okp) and len(kp) > 0
        for o in range(s.num_octaves()):
            for lv in range(-1, 5):
                assert np.array_equal(s.level_data("gpyr", o, lv).view(np.uint32),
                                      orc.level("gpyr", o, lv)[0].view(np.uint32)), (o, lv)
        for f in ("xd", "yd", "zd", "o", "s", "sd"):
            assert np.array_equal(kp[f], okp[f])
        d = s.extract_descriptors()
        od, _ = orc.describe(okp)
        rel = np.linalg.norm(d["hists"] - od, axis=1) / np.linalg.norm(d["hists"], axis=1)
        assert rel.max() <= 1e-6


def test_gauss_taps_values(oracle_cls):
    """Filter widths and sigmas of the default pyramid (from the reference)."""
    orc = oracle_cls()
    s = [1.6 * 2 ** (k / 3.0) for k in range(-1, 5)]
    first = np.sqrt(s[0] ** 2 - 1.15 ** 2)
    assert len(orc.gauss_taps(first)) == 5
    widths = [len(orc.gauss_taps(np.sqrt(s[i + 1] ** 2 - s[i] ** 2))) for i in range(5)]
    assert widths == [7, 9, 11, 13, 17]
    t = orc.gauss_taps(2.452547)
    assert abs(float(t.sum()) - 1.0) < 1e-6 and np.array_equal(t, t[::-1])


def test_blur_boundary_quirks(oracle_cls):
    """A.2 quirks: linear extrapolation at c in (-1,0) for octaves >= 1 (Q2) and the
    0.1-voxel offset of the right-hand mirror (Q3)."""
    orc = oracle_cls()
    taps = np.array([0.25, 0.5, 0.25], np.float32)
    line = np.arange(16, dtype=np.float32) ** 2
    vol = np.tile(line, (3, 3, 1)).copy()
    # units 2 along x => tap spacing 0.5 voxel; y/z units huge => spacing ~0 (identity-like)
    out = orc.blur(vol, taps, units=(2.0, 1e9, 1e9))
    c0 = np.float32(0.25) * (np.float32(1.5) * line[0] + np.float32(-0.5) * line[1])
    assert np.isfinite(out).all()
    # x = 0, d = +1 samples c = -0.5 -> lo = 0, frac = -0.5 (extrapolation), d = -1 samples 0.5
    want = np.float32(0.25) * (np.float32(0.5) * line[0] + np.float32(0.5) * line[1])
    want = np.float32(want + np.float32(0.5) * line[0])
    want = np.float32(want + c0)
    assert out[1, 1, 0] == pytest.approx(float(want), rel=1e-6)


def test_eig3_against_numpy(oracle_cls):
    import ctypes as C
    orc = oracle_cls()
    rng = np.random.default_rng(0)
    for _ in range(50):
        m = rng.standard_normal((3, 3))
        a = (m @ m.T).astype(np.float64)
        q = np.zeros(9)
        lam = np.zeros(3)
        orc.L.orc_eig3(a.ctypes.data, q.ctypes.data, lam.ctypes.data)
        w, v = np.linalg.eigh(a)
        assert np.allclose(lam, w, rtol=1e-12, atol=1e-12)
        q = q.reshape(3, 3)
        for k in range(3):
            assert abs(abs(q[:, k] @ v[:, k]) - 1.0) < 1e-9


def test_mesh_quirk(oracle_cls):
    """Q4: all 20 faces take the swap branch: v[0] is the vector of idx[1] and vice versa."""
    orc = oracle_cls()
    v, idx = orc.mesh()
    gr = 1.6180339887
    vert = np.array([[0, 1, gr], [0, -1, gr], [0, 1, -gr], [0, -1, -gr], [1, gr, 0], [-1, gr, 0],
                     [1, -gr, 0], [-1, -gr, 0], [gr, 0, 1], [-gr, 0, 1], [gr, 0, -1],
                     [-gr, 0, -1]], np.float32)
    vert /= np.linalg.norm(vert, axis=1, keepdims=True)
    for i in range(20):
        assert np.allclose(v[i, 0], vert[idx[i, 1]], atol=1e-6)
        assert np.allclose(v[i, 1], vert[idx[i, 0]], atol=1e-6)
        assert np.allclose(v[i, 2], vert[idx[i, 2]], atol=1e-6)


# ---------------------------------------------------------------- SURVEY.md 8f rows N1 / N3
def _ref_imutil():
    import oracle_api
    if not oracle_api.REF_IMUTIL.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    import ctypes as C
    return C.CDLL(str(oracle_api.REF_IMUTIL))


def _ref_inv_transform(L, vol, A, out_shape, interp):
    """The unmodified reference: init_Affine + Affine_set_mat + im_inv_transform."""
    import ctypes as C
    from sift3d_b200 import capi
    vol = np.ascontiguousarray(vol, np.float32)
    nc = vol.shape[3] if vol.ndim == 4 else 1
    src = capi.make_image(vol, (1.0, 1.0, 1.0), nc)
    aff = (C.c_char * 64)()                      # Affine: Tform (16 B) + Mat_rm (32 B)
    assert L.init_Affine(aff, 3) == 0
    Am = np.ascontiguousarray(A, np.float64).reshape(3, 4)
    m = capi.Mat_rm()
    m.data = Am.ctypes.data
    m.size = Am.nbytes
    m.num_cols, m.num_rows, m.static_mem, m.type = 4, 3, 1, 0   # SIFT3D_DOUBLE
    assert L.Affine_set_mat(C.byref(m), aff) == 0
    dnz, dny, dnx = out_shape
    out = np.zeros((dnz, dny, dnx) + ((nc,) if vol.ndim == 4 else ()), np.float32)
    dst = capi.make_image(out, (1.0, 1.0, 1.0), nc)
    L.im_inv_transform.argtypes = [C.c_void_p, C.POINTER(capi.Image), C.c_int, C.c_int,
                                   C.POINTER(capi.Image)]
    assert L.im_inv_transform(aff, C.byref(src), interp, 0, C.byref(dst)) == 0
    return out


AFFINES = {
    "identity": [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0],
    "scale": [0.5, 0, 0, 0, 0, 1.0 / 0.7, 0, 0, 0, 0, 1.25, 0],
    "rot_shift": [0.9553, -0.2955, 0.0, 3.1, 0.2955, 0.9553, 0.05, -2.4, -0.03, 0.02, 1.01, 1.7],
}


@pytest.mark.parametrize("name", list(AFFINES))
@pytest.mark.parametrize("interp", [0, 1])
def test_resample_restatement_equals_reference(oracle_cls, name, interp):
    L = _ref_imutil()
    rng = np.random.default_rng(3)
    vol = rng.random((14, 17, 19), dtype=np.float32)
    out_shape = (16, 15, 23)
    want = _ref_inv_transform(L, vol, AFFINES[name], out_shape, interp)
    got = oracle_cls().resample_affine(vol, AFFINES[name], out_shape, interp)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert (want != 0).mean() > 0.2
    if interp == 0:                                   # multi-channel, linear
        v4 = rng.random((6, 7, 8, 3), dtype=np.float32)
        want = _ref_inv_transform(L, v4, AFFINES[name], (7, 6, 9), 0)
        got = oracle_cls().resample_affine(v4, AFFINES[name], (7, 6, 9), 0)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_nn_match_restatement_equals_reference(oracle_cls, ref_lib):
    import ctypes as C
    from sift3d_b200 import capi
    rng = np.random.default_rng(5)
    a = rng.random((60, 771)).astype(np.float32)
    b = np.concatenate([a[:30] + 0.02 * rng.random((30, 771)).astype(np.float32), a[:5],
                        rng.random((25, 771)).astype(np.float32)])

    def store(rows):
        m = capi.Mat_rm()
        m.data = rows.ctypes.data
        m.size = rows.nbytes
        m.num_cols, m.num_rows, m.static_mem, m.type = 771, len(rows), 1, 1
        d = capi.SIFT3D_Descriptor_store()
        ref_lib.lib.init_SIFT3D_Descriptor_store(C.byref(d))
        ref_lib.lib.Mat_rm_to_SIFT3D_Descriptor_store.argtypes = [
            C.POINTER(capi.Mat_rm), C.POINTER(capi.SIFT3D_Descriptor_store)]
        assert ref_lib.lib.Mat_rm_to_SIFT3D_Descriptor_store(C.byref(m), C.byref(d)) == 0
        return d
    for thresh in (0.8, 0.95):
        d1, d2 = store(a), store(b)
        m = C.POINTER(C.c_int)()
        assert ref_lib.lib.SIFT3D_nn_match(C.byref(d1), C.byref(d2), C.c_float(thresh),
                                           C.byref(m)) == 0
        want = np.array([m[i] for i in range(len(a))])
        got = oracle_cls().nn_match(a[:, 3:], b[:, 3:], thresh)
        assert np.array_equal(got, want) and (want >= 0).sum() >= 20
