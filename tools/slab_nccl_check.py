"""torchrun worker: Z-slab tiling over NCCL (one process per GPU) vs the whole volume on rank 0.
Prints SLAB_NCCL_OK on success.  Usage: torchrun --nproc-per-node N tools/slab_nccl_check.py"""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist
    from sift3d_b200 import capi, slab
    from sift3d_b200.volumes import blob_volume
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["SIFT3D_CUDA_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    shape = tuple(int(v) for v in os.environ.get("SLAB_SHAPE", "160,96,128").split(","))
    vol = blob_volume(shape, seed=21)
    zs = slab.split_planes(shape[0], world)
    comm = slab.nccl_comm(local)
    lib = capi.load_b200()
    with capi.Sift3D(lib) as s:
        kp = slab.detect_slab(s, vol[zs[rank]:zs[rank + 1]], zs, comm)
        desc = s.extract_descriptors() if len(kp) else np.zeros(0, capi.DESCRIPTOR_DTYPE)
    parts = [None] * world
    dist.all_gather_object(parts, (kp.tobytes(), desc.tobytes()))
    ok = True
    if rank == 0:
        kps = [np.frombuffer(p[0], capi.KEYPOINT_DTYPE) for p in parts]
        ds = [np.frombuffer(p[1], capi.DESCRIPTOR_DTYPE) for p in parts]
        mk, md = slab.merge_ranks(kps, ds)
        with capi.Sift3D(lib) as s:
            wk = s.detect_keypoints(vol)
            wd = s.extract_descriptors()
        same_kp = len(mk) == len(wk) and all(
            np.array_equal(mk[f], wk[f]) for f in ("xd", "yd", "zd", "sd", "o", "s", "R"))
        same_d = same_kp and np.array_equal(md["hists"], wd["hists"])
        print(f"ranks={world} shape={shape} keypoints tiled={len(mk)} whole={len(wk)} "
              f"kp_equal={same_kp} desc_equal={same_d} per_rank={[len(k) for k in kps]}")
        ok = same_kp and same_d and len(wk) > 0
        if ok:
            print("SLAB_NCCL_OK")
    comm.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
