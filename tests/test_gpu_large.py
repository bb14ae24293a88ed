"""GPU parity at the sizes bench.py actually times (BASELINE.json configs[1] and configs[2]).

The goldens (`tests/golden/blob{256,512}_large.npz`) come from the UNMODIFIED reference
(oracle/_ref) run once in the build container on `blob_volume(n, seed=1234)` -- the very
volume rank 0 of bench.py times (`python tests/golden/make_golden.py --large`; 512^3 needs
about 7 minutes of CPU there).  They hold the sha256 of the input, of every Gaussian and DoG
level of all octaves, the candidate count per (octave, level), the complete keypoint table
and every k-th descriptor; the 256^3 file also holds a sample of the reference's dense
descriptors (SIFT3D_extract_dense_descriptors, sift.c:2354).

Bar: pyramid levels bit-identical, keypoints identical incl. order, R to 1e-5, descriptors
within 1e-4 relative L2 (BASELINE.json north_star).
"""
import ctypes as C
import hashlib

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
DESC_TOL = 1e-4


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def load_large(n):
    z = np.load(GOLDEN / f"blob{n}_large.npz")
    return {k: z[k] for k in z.files}


def golden_volume(g):
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume(int(g["n"]), seed=int(g["seed"]))
    assert digest(vol) == str(g["input_sha256"]), \
        "blob_volume() is not bit-identical to the volume the golden was generated from"
    return vol


@pytest.mark.parametrize("n", [256, 512])
def test_detect_and_describe_match_reference_at_bench_size(b200_lib, n):
    from sift3d_b200 import capi
    g = load_large(n)
    vol = golden_volume(g)
    with capi.Sift3D(b200_lib) as s:
        kp = s.detect_keypoints(vol)
        assert s.num_octaves() == int(g["noct"])
        want = {}
        for row in g["level_sha256"]:
            w, o, lv, h = str(row).split(",")
            want[(w, int(o), int(lv))] = h
        bad = []
        for o in range(s.num_octaves()):
            for lv in range(-1, 5):
                if digest(s.level_data("gpyr", o, lv)) != want[("g", o, lv)]:
                    bad.append(("gpyr", o, lv))
            for lv in range(-1, 4):
                if digest(s.level_data("dog", o, lv)) != want[("d", o, lv)]:
                    bad.append(("dog", o, lv))
        assert not bad, f"pyramid levels not bit-identical to the reference: {bad}"
        ncand = b200_lib.lib.sift3d_b200_num_candidates(C.byref(s.s))
        assert ncand == int(g["candidates_per_level"].sum())
        assert len(kp) == len(g["kp_o"]), (len(kp), len(g["kp_o"]))
        assert np.array_equal(np.stack([kp["xd"], kp["yd"], kp["zd"]], 1),
                              g["kp_xyz"].astype(np.float64))
        assert np.array_equal(kp["o"], g["kp_o"]) and np.array_equal(kp["s"], g["kp_s"])
        assert np.array_equal(kp["sd"], g["kp_sd"])
        assert np.abs(kp["R"] - g["kp_R"]).max() <= 1e-5
        d = s.extract_descriptors()
        every = int(g["desc_every"])
        rel = np.linalg.norm(d["hists"][::every] - g["desc"], axis=1) / \
            np.linalg.norm(g["desc"], axis=1)
        assert rel.max() <= DESC_TOL, rel.max()
        # all descriptors: unit norm, and the sum of norms the reference produced
        nrm = np.linalg.norm(d["hists"].astype(np.float64), axis=1)
        assert abs(nrm.sum() - float(g["desc_norm_sum"])) <= 1e-6 * len(kp)
        f = 2.0 ** kp["o"]
        assert np.array_equal(d["xd"], kp["xd"] * f) and np.array_equal(d["zd"], kp["zd"] * f)


def test_dense_descriptors_match_reference_at_256(b200_lib):
    """BASELINE.json configs[2]: SIFT3D_extract_dense_descriptors (sift.c:2354-2424,
    dense_rotate = 0) on the 256^3 volume against a strided sample, one full plane and the
    per-plane channel sums of the reference's output."""
    from sift3d_b200 import capi
    g = load_large(256)
    vol = golden_volume(g)
    with capi.Sift3D(b200_lib) as s:
        dd = s.extract_dense_descriptors(vol)
    n = int(g["n"])
    assert dd.shape == (n, n, n, 12)
    scale = max(float(np.abs(g["dense_sub"]).max()), 1e-30)
    assert np.abs(dd[3::8, 3::8, 3::8] - g["dense_sub"]).max() / scale <= 1e-5
    assert np.abs(dd[n // 2, ::2, ::2] - g["dense_plane"]).max() / scale <= 1e-5
    sums = dd.astype(np.float64).sum(axis=(1, 2))
    ref = g["dense_plane_sums"]
    assert np.abs(sums - ref).max() <= 1e-5 * np.abs(ref).max()
