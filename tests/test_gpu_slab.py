"""GPU parity of the Z-slab tiled path (BASELINE.json configs[4] at sizes the CPU can check,
SURVEY.md D6): a volume tiled over N ranks must give, bit for bit, the pyramid, the keypoints
(order included) and the descriptors of the same volume processed whole -- and those are
checked against the oracle / the reference's golden fixtures elsewhere (test_gpu_parity.py).
The ranks run as host threads of this process on one GPU (in-process transport), which
exercises all of the tiling logic; the NCCL transport is covered by test_slab_nccl_two_gpus."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu


KP_FIELDS = ("xd", "yd", "zd", "sd", "o", "s", "R")


def same_kp(a, b):
    return len(a) == len(b) and all(np.array_equal(a[f], b[f]) for f in KP_FIELDS)


def whole(b200_lib, vol, units, levels=True):
    from sift3d_b200 import capi
    with capi.Sift3D(b200_lib) as s:
        kp = s.detect_keypoints(vol, units)
        desc = s.extract_descriptors() if len(kp) else np.zeros(0, capi.DESCRIPTOR_DTYPE)
        lv = {}
        if levels:
            for o in range(s.num_octaves()):
                for l in range(-1, 5):
                    lv[("gpyr", o, l)] = s.level_data("gpyr", o, l)
                for l in range(-1, 4):
                    lv[("dog", o, l)] = s.level_data("dog", o, l)
        return kp, desc, lv, s.num_octaves()


def tiled(b200_lib, vol, units, nranks, levels=True):
    import ctypes as C
    from sift3d_b200 import capi, slab
    zs = slab.split_planes(vol.shape[0], nranks)
    world = slab.LocalWorld(nranks)

    def work(rank, comm):
        with capi.Sift3D(b200_lib) as s:
            kp = slab.detect_slab(s, vol[zs[rank]:zs[rank + 1]], zs, comm, units)
            desc = s.extract_descriptors() if len(kp) else np.zeros(0, capi.DESCRIPTOR_DTYPE)
            lv = {}
            if levels:
                fetch = s.L.sift3d_b200_fetch_level
                fetch.argtypes = [C.POINTER(capi.SIFT3D), C.c_int, C.c_int, C.c_int, C.c_void_p]
                for o in range(s.num_octaves()):
                    info = slab.slab_info(s, o)
                    m = s.level_meta("gpyr", o, 0)
                    for which, n in (("gpyr", 5), ("dog", 4)):
                        for l in range(-1, n):
                            buf = np.zeros((max(info["hi"] - info["lo"], 0), m["ny"], m["nx"]),
                                           np.float32)
                            if buf.size:
                                assert fetch(C.byref(s.s), 0 if which == "gpyr" else 1, o, l,
                                             buf.ctypes.data) == 0
                            a, b = info["own0"] - info["lo"], info["own1"] - info["lo"]
                            lv[(which, o, l)] = (info["own0"], info["own1"], buf[a:b].copy())
            return kp, desc, lv
    try:
        return world.run(work), zs
    finally:
        world.close()


CASES = [
    # (nz, ny, nx), units, nranks
    ((96, 64, 64), (1.0, 1.0, 1.0), 2),     # fused blur at octave 0, generic below
    ((96, 64, 64), (1.0, 1.0, 1.0), 3),     # 32-plane slabs < 39-plane halo: multi-peer halos
    ((70, 33, 45), (1.0, 1.0, 1.0), 2),     # odd sizes: unaligned plane offsets, odd split
    ((150, 64, 72), (1.0, 1.0, 1.0), 4),    # uneven 38/38/37/37 split
    ((64, 40, 48), (0.7, 0.9, 1.3), 2),     # non-dyadic units: generic lerp path in z
    ((80, 48, 64), (1.0, 1.0, 2.0), 3),     # anisotropic z: smaller plane halos
]


@pytest.mark.parametrize("shape,units,nranks", CASES)
def test_tiled_equals_whole(b200_lib, shape, units, nranks):
    from sift3d_b200 import slab
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume(shape, seed=11 + nranks)
    kp_w, desc_w, lv_w, noct = whole(b200_lib, vol, units)
    parts, zs = tiled(b200_lib, vol, units, nranks)
    # pyramids: every owned plane bit-identical
    bad = []
    for r, (_, _, lv) in enumerate(parts):
        for key, (a, b, data) in lv.items():
            if not np.array_equal(data.view(np.uint32), lv_w[key][a:b].view(np.uint32)):
                bad.append((r,) + key)
    assert not bad, f"levels differ (rank, pyramid, octave, level): {bad[:10]}"
    kp, desc = slab.merge_ranks([p[0] for p in parts], [p[1] for p in parts])
    assert len(kp_w) > 0 and len(kp) == len(kp_w), (len(kp), len(kp_w))
    assert same_kp(kp, kp_w)
    # the descriptor histogram is fixed-point (order-independent): bit-identical too
    assert np.array_equal(desc["hists"], desc_w["hists"])
    for f in ("xd", "yd", "zd", "sd"):
        assert np.array_equal(desc[f], desc_w[f]), f
    # every rank reported only keypoints of its own planes
    own = slab.plan_octaves(zs, noct)
    for r, (k, _, _) in enumerate(parts):
        for o in range(noct):
            z = k["zd"][k["o"] == o]
            assert ((z >= own[r, o, 0]) & (z < own[r, o, 1])).all()


def test_tiled_matches_oracle(b200_lib, oracle_cls):
    """Independent of the whole-volume GPU path: tiled result vs the CPU oracle."""
    from sift3d_b200 import slab
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume((72, 40, 48), seed=5)
    parts, _ = tiled(b200_lib, vol, (1.0, 1.0, 1.0), 3, levels=False)
    kp, desc = slab.merge_ranks([p[0] for p in parts], [p[1] for p in parts])
    orc = oracle_cls()
    okp = orc.detect(vol)
    assert len(kp) == len(okp) and len(kp) > 0
    for f in ("xd", "yd", "zd", "o", "s"):
        assert np.array_equal(kp[f], okp[f]), f
    od, _ = orc.describe(okp)
    rel = np.linalg.norm(desc["hists"] - od, axis=1) / np.linalg.norm(od, axis=1)
    assert rel.max() <= 1e-4


def test_tiled_repeat_and_switch_back(b200_lib):
    """Same SIFT3D object: tiled twice (allocation reuse), then a whole volume again."""
    from sift3d_b200 import capi, slab
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume((64, 40, 64), seed=3)
    kp_w, _, _, _ = whole(b200_lib, vol, (1.0, 1.0, 1.0), levels=False)
    zs = slab.split_planes(64, 2)
    world = slab.LocalWorld(2)

    def work(rank, comm):
        with capi.Sift3D(b200_lib) as s:
            a = slab.detect_slab(s, vol[zs[rank]:zs[rank + 1]], zs, comm)
            b = slab.detect_slab(s, vol[zs[rank]:zs[rank + 1]], zs, comm)
            assert same_kp(a, b)
            return a
    try:
        parts = world.run(work)
    finally:
        world.close()
    assert same_kp(slab.merge_ranks(parts), kp_w)
    with capi.Sift3D(b200_lib) as s:
        w = slab.LocalWorld(1)
        try:
            one = slab.detect_slab(s, vol, [0, 64], w.comms[0])   # a single rank owns everything
            assert same_kp(one, kp_w)
            again = s.detect_keypoints(vol)                        # back to the whole-volume path
            assert same_kp(again, kp_w)
        finally:
            w.close()


def test_slab_nccl_two_gpus(built):
    """NCCL transport, one process per GPU (needs >= 2 GPUs; `gpurun --gpus 2`)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29651",
           str(REPO / "tools" / "slab_nccl_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "SLAB_NCCL_OK" in r.stdout
