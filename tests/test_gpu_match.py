"""GPU parity of SIFT3D_nn_match: the device search must return exactly the
matches of the reference's CPU implementation (sift.c:2840-2969) -- argmin with first-index
tie-breaking, ratio test, forward-backward check -- on random, duplicated (exact ties, 0/0
ratio), ragged-size and real descriptor sets."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def make_store(lib, rows):
    from sift3d_b200 import capi
    rows = np.ascontiguousarray(rows, np.float32)
    m = capi.Mat_rm()
    m.data = rows.ctypes.data
    m.size = rows.nbytes
    m.num_cols, m.num_rows, m.static_mem, m.type = 771, len(rows), 1, 1
    d = capi.SIFT3D_Descriptor_store()
    lib.lib.init_SIFT3D_Descriptor_store(C.byref(d))
    lib.lib.Mat_rm_to_SIFT3D_Descriptor_store.argtypes = [
        C.POINTER(capi.Mat_rm), C.POINTER(capi.SIFT3D_Descriptor_store)]
    assert lib.lib.Mat_rm_to_SIFT3D_Descriptor_store(C.byref(m), C.byref(d)) == 0
    return d


def nn_match(lib, a, b, thresh=0.8):
    d1, d2 = make_store(lib, a), make_store(lib, b)
    m = C.POINTER(C.c_int)()
    assert lib.lib.SIFT3D_nn_match(C.byref(d1), C.byref(d2), C.c_float(thresh), C.byref(m)) == 0
    out = np.array([m[i] for i in range(len(a))])
    lib._libc.free(C.cast(m, C.c_void_p))
    lib.lib.cleanup_SIFT3D_Descriptor_store(C.byref(d1))
    lib.lib.cleanup_SIFT3D_Descriptor_store(C.byref(d2))
    return out


def test_recovers_permutation(b200_lib):
    rng = np.random.default_rng(0)
    n = 40
    rows = np.zeros((n, 771), np.float32)
    rows[:, :3] = rng.uniform(0, 50, (n, 3))
    rows[:, 3:] = rng.random((n, 768))
    perm = rng.permutation(n)
    rows2 = rows[perm].copy()
    rows2[:, 3:] += 1e-4 * rng.random((n, 768)).astype(np.float32)
    assert np.array_equal(nn_match(b200_lib, rows, rows2), np.argsort(perm))


@pytest.mark.parametrize("n1,n2,thresh", [(60, 55, 0.8), (1, 1, 0.8), (7, 1, 0.8), (130, 200, 0.9),
                                          (257, 129, 0.6), (64, 64, 0.99)])
def test_equals_reference_random(b200_lib, ref_lib, n1, n2, thresh):
    rng = np.random.default_rng(n1 * 1000 + n2)
    a = rng.random((n1, 771)).astype(np.float32)
    k = min(n1, n2) // 2
    b = np.concatenate([a[:k] + 0.02 * rng.random((k, 771)).astype(np.float32),
                        rng.random((n2 - k, 771)).astype(np.float32)])
    got, want = nn_match(b200_lib, a, b, thresh), nn_match(ref_lib, a, b, thresh)
    assert np.array_equal(got, want)
    if n1 >= 60:
        assert (got >= 0).sum() >= k // 2


def test_equals_reference_with_exact_ties(b200_lib, ref_lib):
    """Duplicated descriptors: equal SSDs (first index wins), zero SSDs (0/0 ratio)."""
    rng = np.random.default_rng(9)
    base = rng.random((50, 771)).astype(np.float32)
    a = np.concatenate([base, base[:10]])            # a[50+i] == a[i]
    b = np.concatenate([base[5:45], base[5:15], rng.random((30, 771)).astype(np.float32)])
    for thresh in (0.8, 1.5):
        assert np.array_equal(nn_match(b200_lib, a, b, thresh), nn_match(ref_lib, a, b, thresh))


def test_equals_reference_on_real_descriptors(b200_lib, ref_lib):
    """Descriptors of two overlapping crops of a synthetic volume, matched by both libraries."""
    from sift3d_b200 import capi
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume((72, 64, 64), seed=8)
    descs = []
    for sl in (slice(0, 56), slice(12, 72)):
        with capi.Sift3D(b200_lib) as s:
            kp = s.detect_keypoints(vol[sl])
            assert len(kp) > 20
            d = s.extract_descriptors()
            rows = np.zeros((len(d), 771), np.float32)
            rows[:, 0], rows[:, 1], rows[:, 2] = d["xd"], d["yd"], d["zd"]
            rows[:, 3:] = d["hists"]
            descs.append(rows)
    got = nn_match(b200_lib, descs[0], descs[1])
    want = nn_match(ref_lib, descs[0], descs[1])
    assert np.array_equal(got, want) and (got >= 0).sum() > 5
