"""Host logic of the Z-slab tiling (no GPU): plane split, per-octave ownership (C ABI vs its
Python twin), halo transfer plan (every send has its matching receive, in the same order, and
the received planes are exactly the halo), and the ordered merge of per-rank results over a
world_size-2 gloo group."""
import ctypes as C
import json
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import REPO


def test_split_and_plan(built):
    from sift3d_b200 import slab
    for nz in (8, 37, 70, 128, 1024):
        for nr in (1, 2, 3, 8):
            zs = slab.split_planes(nz, nr)
            assert zs[0] == 0 and zs[-1] == nz and len(zs) == nr + 1
            sizes = np.diff(zs)
            assert sizes.max() - sizes.min() <= 1
            noct = max(1, int(np.log2(nz)) - 3 + 1)
            own = slab.plan_octaves(zs, noct)
            assert np.array_equal(own, slab.plan_octaves_c(zs, noct))
            n = nz
            for o in range(noct):
                # the owned ranges tile [0, n) in rank order
                assert own[0, o, 0] == 0 and own[-1, o, 1] == n
                assert np.array_equal(own[1:, o, 0], own[:-1, o, 1])
                if o + 1 < noct:  # a rank owns dst plane z iff it owns src plane 2z
                    for r in range(nr):
                        mine = [z for z in range(n // 2) if own[r, o, 0] <= 2 * z < own[r, o, 1]]
                        a, b = own[r, o + 1]
                        assert mine == list(range(a, b))
                n //= 2
    with pytest.raises(ValueError):
        slab.plan_octaves_c([0, 5, 3, 9], 1)


def halo_plan(own, o, NZ, h, rank):
    from sift3d_b200.slab import _CudaLib
    L = _CudaLib.get()
    L.s3d_slab_halo_plan.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_int]
    own = np.ascontiguousarray(own, np.int32)
    out = np.zeros((64, 4), np.int32)
    n = L.s3d_slab_halo_plan(own.shape[0], own.shape[1], own.ctypes.data, o, NZ, h, rank,
                             out.ctypes.data, 64)
    assert n >= 0
    return [tuple(int(v) for v in row) for row in out[:n]]


@pytest.mark.parametrize("nz,nr,h", [(64, 2, 5), (70, 3, 39), (128, 8, 39), (40, 4, 12), (9, 8, 3)])
def test_halo_plan_is_consistent(built, nz, nr, h):
    from sift3d_b200 import slab
    zs = slab.split_planes(nz, nr)
    noct = max(1, int(np.log2(nz)) - 3 + 1)
    own = slab.plan_octaves(zs, noct)
    n = nz
    for o in range(noct):
        plans = [halo_plan(own, o, n, h, r) for r in range(nr)]
        for r in range(nr):
            a, b = own[r, o]
            got = set()
            for kind, peer, z0, z1 in plans[r]:
                assert z1 > z0 and peer != r
                if kind == 1:  # received planes belong to the peer and are unique
                    assert own[peer, o, 0] <= z0 and z1 <= own[peer, o, 1]
                    assert not (got & set(range(z0, z1)))
                    got |= set(range(z0, z1))
                else:          # sent planes are mine
                    assert a <= z0 and z1 <= b
            want = set() if b <= a else \
                (set(range(max(a - h, 0), a)) | set(range(b, min(b + h, n))))
            assert got == want, (o, r)
            for p in range(nr):  # pairwise: r's sends to p == p's receives from r, same order
                s = [(z0, z1) for k, q, z0, z1 in plans[r] if k == 0 and q == p]
                v = [(z0, z1) for k, q, z0, z1 in plans[p] if k == 1 and q == r]
                assert s == v
        n //= 2


def test_merge_ranks_order():
    from sift3d_b200 import capi, slab

    def mk(rows):
        k = np.zeros(len(rows), capi.KEYPOINT_DTYPE)
        for i, (o, s, z) in enumerate(rows):
            k["o"][i], k["s"][i], k["zd"][i] = o, s, z
        return k
    r0 = mk([(0, 0, 1), (0, 0, 5), (0, 1, 2), (1, 0, 1)])
    r1 = mk([(0, 0, 9), (0, 2, 8), (1, 0, 3), (1, 0, 4)])
    m = slab.merge_ranks([r0, r1])
    got = [(int(a), int(b), int(c)) for a, b, c in zip(m["o"], m["s"], m["zd"])]
    assert got == sorted(got) and len(got) == 8
    assert len(slab.merge_ranks([mk([]), mk([])])) == 0


WORKER = textwrap.dedent("""
    import json, sys
    sys.path.insert(0, {repo!r})
    import numpy as np
    import torch.distributed as dist
    from sift3d_b200 import capi, dist as sd, slab
    rank, world = sd.init_process_group("gloo")
    zs = slab.split_planes(37, world)
    own = slab.plan_octaves(zs, 2)
    # every rank fabricates the keypoints a tiled detect would return for its planes
    rows = [(o, s, z) for o in range(2) for s in range(3)
            for z in range(own[rank, o, 0], own[rank, o, 1]) if (z * 7 + s + o) % 3 == 0]
    k = np.zeros(len(rows), capi.KEYPOINT_DTYPE)
    for i, (o, s, z) in enumerate(rows):
        k["o"][i], k["s"][i], k["zd"][i] = o, s, z
    parts = [None] * world
    dist.all_gather_object(parts, k.tobytes())
    kps = [np.frombuffer(p, capi.KEYPOINT_DTYPE) for p in parts]
    m = slab.merge_ranks(kps)
    if rank == 0:
        print(json.dumps({{"n": len(m), "rows": [[int(a), int(b), int(c)] for a, b, c in
                                                zip(m["o"], m["s"], m["zd"])]}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=str(REPO)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29641", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    want = [[o, s, z] for o in range(2) for s in range(3) for z in range(37 >> o)
            if (z * 7 + s + o) % 3 == 0]
    assert out["rows"] == want
