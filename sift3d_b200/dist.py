"""Multi-GPU plumbing: independent volumes sharded over ranks (BASELINE.json configs[3]).

The accelerated path has NO data-path collective: every rank owns whole volumes (one
`SIFT3D` object / engine per rank) and the ranks only exchange a few scalars -- timings
(max over ranks) and per-volume keypoint counts.  `torch.distributed` (NCCL on GPUs,
gloo in the CPU tests) is used for exactly that.
"""
from __future__ import annotations

import os
from typing import List, Sequence


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), \
        int(os.environ.get("LOCAL_RANK", "0"))


def shard_volumes(n_volumes: int, rank: int, world: int) -> List[int]:
    """Volume indices owned by `rank`: contiguous blocks, sizes differing by at most one
    (volume v of a batch goes to rank v*world//n for n >= world)."""
    if n_volumes <= 0 or world <= 0 or not (0 <= rank < world):
        raise ValueError("bad shard request")
    base, extra = divmod(n_volumes, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def init_process_group(backend: str, device=None):
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    if world > 1 and not dist.is_initialized():
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, world


def max_over_ranks(values: Sequence[float], device=None) -> List[float]:
    """Element-wise MAX over ranks (timing rule: a multi-GPU step takes as long as its
    slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]


def gather_counts(local_counts: Sequence[int], n_volumes: int, device=None) -> List[int]:
    """Per-volume result counts of the whole batch, in volume order, on every rank."""
    import torch
    import torch.distributed as dist
    rank, world, _ = env_rank_world()
    out = torch.zeros(n_volumes, dtype=torch.int64, device=device)
    mine = shard_volumes(n_volumes, rank, world)
    assert len(mine) == len(local_counts)
    for v, c in zip(mine, local_counts):
        out[v] = int(c)
    if dist.is_available() and dist.is_initialized() and world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return [int(v) for v in out.tolist()]
