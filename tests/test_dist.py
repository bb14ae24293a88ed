"""world_size-2 gloo test of the multi-GPU host logic (no GPU needed): volume sharding,
max-over-ranks timing and ordered gathering of per-volume counts."""
import os
import subprocess
import sys
import textwrap

import pytest

from conftest import REPO


def test_shard_volumes_partition():
    from sift3d_b200.dist import shard_volumes
    for n in (1, 2, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            if n < world:
                continue
            got = [shard_volumes(n, r, world) for r in range(world)]
            flat = [v for g in got for v in g]
            assert flat == list(range(n))
            assert max(len(g) for g in got) - min(len(g) for g in got) <= 1
    with pytest.raises(ValueError):
        shard_volumes(4, 4, 4)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, {repo!r})
    import torch, torch.distributed as dist
    from sift3d_b200 import dist as sd
    rank, world = sd.init_process_group("gloo")
    mine = sd.shard_volumes(5, rank, world)
    # pretend each volume v yields 10*v+1 keypoints and rank r needs (r+1)*3.5 ms
    counts = sd.gather_counts([10 * v + 1 for v in mine], 5)
    tmax = sd.max_over_ranks([(rank + 1) * 3.5, 1.0 - rank])
    if rank == 0:
        print(json.dumps({{"world": world, "mine": mine, "counts": counts, "tmax": tmax}}))
    dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(repo=str(REPO)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29631", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0, r.stdout + r.stderr
    import json
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["world"] == 2 and out["mine"] == [0, 1, 2]
    assert out["counts"] == [1, 11, 21, 31, 41]
    assert out["tmax"] == [7.0, 1.0]


def test_bench_reference_arm_other_ranks_exit_quietly():
    """`bench.py --impl reference` under torchrun: only rank 0 works and prints."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env,
                       timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_reference_arm_line_has_the_contract_keys(built):
    """`bench.py --impl reference` (rank 0): ONE JSON line with the driver's keys, the metric of
    BASELINE.json, a cpu_baseline describing this run and an e2e object without transfers."""
    import json
    import oracle_api
    if not oracle_api.REF_LIB.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    r = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-size", "48"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.loads((REPO / "BASELINE.json").read_text())
    assert d["impl"] == "reference" and d["metric"].split(",")[0] in base["metric"]
    for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and "workload" in d["config"]
