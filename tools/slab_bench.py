"""torchrun worker: throughput of the Z-slab tiled path (BASELINE.json configs[4] shape class):
ONE volume of nx*ny*(nzl*world) voxels tiled over `world` GPUs, NCCL halo exchange.
Reports voxels/s (device-resident: the slab is already in HBM, results stay in HBM), timed with
CUDA events on the engine's stream, max over ranks.  Not the bench.py headline (that is
configs[1]/[3]); a measurement of the tiling overhead next to it.
Env: SLAB_NX, SLAB_NY, SLAB_NZL (planes per rank), SLAB_STEPS, SLAB_TRANSPORT=nccl|local"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    import torch.distributed as dist
    from sift3d_b200 import capi, slab
    from sift3d_b200.volumes import blob_volume_torch
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ["SIFT3D_CUDA_DEVICE"] = str(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank() if world > 1 else 0
    nx, ny, nzl = (int(os.environ.get(k, d)) for k, d in
                   (("SLAB_NX", "1024"), ("SLAB_NY", "1024"), ("SLAB_NZL", "256")))
    steps = int(os.environ.get("SLAB_STEPS", "3"))
    zs = [nzl * r for r in range(world + 1)]
    vol = blob_volume_torch((nzl, ny, nx), 1234 + rank, dev)
    host = vol.cpu().numpy()
    if world > 1:
        comm = slab.nccl_comm(local)
        lw = None
    else:
        lw = slab.LocalWorld(1, [local])
        comm = lw.comms[0]
    lib = capi.load_b200()
    cu = C.CDLL(str(capi.CUDA_LIB))
    lib.lib.sift3d_b200_engine.restype = C.c_void_p
    lib.lib.sift3d_b200_engine.argtypes = [C.POINTER(capi.SIFT3D)]
    for f, at in (("s3d_slab_image_from_device", [C.c_void_p, C.c_void_p]),
                  ("s3d_build_pyramid", [C.c_void_p]),
                  ("s3d_detect_extrema", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
                  ("s3d_assign_orientations", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
                  ("s3d_extract_descriptors_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
                  ("s3d_engine_set_stream", [C.c_void_p, C.c_void_p])):
        getattr(cu, f).argtypes = at
    cu.s3d_device_keypoints.argtypes = [C.c_void_p]
    cu.s3d_device_keypoints.restype = C.c_void_p
    s = capi.Sift3D(lib)
    kp = slab.detect_slab(s, host, zs, comm, copy=False)   # sizes everything, warms the kernels
    nkp0 = len(kp)
    eng = lib.lib.sift3d_b200_engine(C.byref(s.s))
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    cu.s3d_engine_set_stream(eng, C.c_void_p(stream.cuda_stream))
    desc = torch.empty(max(nkp0, 1) * 3104 * 2 + 4096, dtype=torch.uint8, device=dev)

    def step(split=None):
        nc, nk = C.c_int(0), C.c_int(0)
        rc = cu.s3d_slab_image_from_device(eng, vol.data_ptr())
        rc |= cu.s3d_build_pyramid(eng)
        if split is not None:
            split.record(stream)
        rc |= cu.s3d_detect_extrema(eng, s.s.peak_thresh, C.byref(nc))
        rc |= cu.s3d_assign_orientations(eng, s.s.corner_thresh, C.byref(nk))
        if nk.value > 0:
            rc |= cu.s3d_extract_descriptors_device(eng, cu.s3d_device_keypoints(eng), nk.value,
                                                    desc.data_ptr())
        if rc:
            raise RuntimeError("slab step failed")
        return nc.value, nk.value

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    barrier()
    a, b, m = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    a.record(stream)
    for i in range(steps):
        nc, nk = step(m if i == steps - 1 else None)
    b.record(stream)
    barrier()
    ms = a.elapsed_time(b) / steps
    t = torch.tensor([ms, float(nk), float(nc)], dtype=torch.float64, device=dev)
    tot = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    if rank == 0:
        nvox = nx * ny * nzl * world
        print("SLAB_BENCH " + json.dumps({
            "workload": f"one {nx}x{ny}x{nzl * world} float32 volume Z-slab tiled over {world} GPU(s)",
            "transport": "nccl" if world > 1 else "local", "n_gpus": world, "ms_per_step": float(t[0]),
            "voxels_per_s": nvox / (float(t[0]) * 1e-3), "keypoints": int(tot[1]),
            "candidates": int(tot[2]), "steps": steps}))
    s.close()
    comm.close()
    if lw is not None:
        lw.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
