"""Z-slab tiling of ONE volume over several GPUs (BASELINE.json configs[4]).

Host-side plumbing only: the plane split, the communicator handles of `libsift3d_cuda.so`
(NCCL: one process per GPU, id broadcast through torch.distributed; local: one host thread
per rank in this process), and the ordered merge of the ranks' keypoints/descriptors into
the reference's (octave, level, z, y, x) scan order (`detect_extrema`, sift.c:1154-1190).
Nothing here computes on voxels.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Sequence

import numpy as np

from . import capi


def split_planes(nz: int, nranks: int) -> List[int]:
    """Contiguous z ranges, sizes differing by at most one plane: rank r owns
    [split[r], split[r+1])."""
    if nz < 0 or nranks < 1:
        raise ValueError("bad split request")
    base, extra = divmod(nz, nranks)
    out = [0]
    for r in range(nranks):
        out.append(out[-1] + base + (1 if r < extra else 0))
    return out


def plan_octaves(zsplit: Sequence[int], num_octaves: int) -> np.ndarray:
    """own[r, o] = (first, end) plane of octave o owned by rank r: octave o+1 is the 2x
    decimation dst[z] = src[2z] (im_downsample_2x, imutil.c:1742-1768), so a rank owns the
    planes whose source plane it owns.  Pure-Python twin of `s3d_slab_plan`."""
    nr = len(zsplit) - 1
    own = np.zeros((nr, num_octaves, 2), np.int32)
    for r in range(nr):
        a, b, nz = zsplit[r], zsplit[r + 1], zsplit[-1]
        for o in range(num_octaves):
            own[r, o] = (min(a, nz), min(b, nz))
            a, b, nz = (a + 1) // 2, (b + 1) // 2, nz // 2
    return own


def merge_ranks(kps: Sequence[np.ndarray], descs: Sequence[np.ndarray] | None = None):
    """Concatenate per-rank results into the single-volume order.  Each rank's list is already
    in (o, s, z, y, x) order with z inside its own planes, and the planes of rank r precede
    those of rank r+1, so the global order is: for each (o, s), rank 0's run, rank 1's run, ..."""
    keys = []
    for r, k in enumerate(kps):
        for i in range(len(k)):
            keys.append((int(k["o"][i]), int(k["s"][i]), r, i))
    keys.sort()
    if not keys:
        kp = np.zeros(0, capi.KEYPOINT_DTYPE)
        return (kp, np.zeros(0, capi.DESCRIPTOR_DTYPE)) if descs is not None else kp
    kp = np.stack([kps[r][i] for (_, _, r, i) in keys])
    if descs is None:
        return kp
    return kp, np.stack([descs[r][i] for (_, _, r, i) in keys])


class _CudaLib:
    _inst = None

    def __init__(self):
        if not capi.CUDA_LIB.exists():
            raise FileNotFoundError(f"{capi.CUDA_LIB} missing: run __graft_entry__.build()")
        L = self.L = C.CDLL(str(capi.CUDA_LIB))
        L.s3d_local_world_create.argtypes = [C.c_int]
        L.s3d_local_world_create.restype = C.c_void_p
        L.s3d_local_world_destroy.argtypes = [C.c_void_p]
        L.s3d_local_world_destroy.restype = None
        L.s3d_comm_create_local.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int]
        L.s3d_nccl_unique_id.argtypes = [C.c_void_p]
        L.s3d_comm_create_nccl.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_void_p,
                                           C.c_int]
        L.s3d_comm_destroy.argtypes = [C.c_void_p]
        L.s3d_comm_destroy.restype = None
        L.s3d_slab_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.s3d_slab_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst.L


def plan_octaves_c(zsplit: Sequence[int], num_octaves: int) -> np.ndarray:
    """`s3d_slab_plan` of the C ABI (host logic; runs without a GPU)."""
    L = _CudaLib.get()
    nr = len(zsplit) - 1
    zs = np.asarray(zsplit, np.int32)
    own = np.zeros((nr, num_octaves, 2), np.int32)
    if L.s3d_slab_plan(nr, num_octaves, int(zs[-1]), zs.ctypes.data, own.ctypes.data) != 0:
        raise ValueError("s3d_slab_plan rejected the split")
    return own


class Comm:
    """Handle of one rank's communicator (`s3d_comm`)."""

    def __init__(self, handle, rank, nranks, world=None):
        self.h, self.rank, self.nranks, self._world = handle, rank, nranks, world

    def close(self):
        if self.h:
            _CudaLib.get().s3d_comm_destroy(self.h)
            self.h = None


class LocalWorld:
    """In-process transport: `nranks` host threads, one communicator each."""

    def __init__(self, nranks: int, devices: Sequence[int] | None = None):
        L = _CudaLib.get()
        self.n = nranks
        self.devices = list(devices) if devices is not None else [0] * nranks
        self.h = L.s3d_local_world_create(nranks)
        if not self.h:
            raise RuntimeError("s3d_local_world_create failed")
        self.comms = []
        for r in range(nranks):
            h = C.c_void_p()
            if L.s3d_comm_create_local(C.byref(h), self.h, r, self.devices[r]) != 0:
                raise RuntimeError("s3d_comm_create_local failed (no CUDA device?)")
            self.comms.append(Comm(h, r, nranks, self))

    def close(self):
        for c in self.comms:
            c.close()
        if self.h:
            _CudaLib.get().s3d_local_world_destroy(self.h)
            self.h = None

    def run(self, fn):
        """fn(rank, comm) on one thread per rank; returns the list of results (re-raises)."""
        out = [None] * self.n
        err = [None] * self.n

        def work(r):
            try:
                out[r] = fn(r, self.comms[r])
            except BaseException as ex:  # noqa: BLE001 - reported to the caller below
                err[r] = ex

        ts = [threading.Thread(target=work, args=(r,)) for r in range(self.n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for ex in err:
            if ex is not None:
                raise ex
        return out


def nccl_comm(device: int) -> Comm:
    """One process per GPU (torchrun): rank 0 makes the NCCL id, torch.distributed broadcasts
    it, every rank joins with `ncclCommInitRank` inside libsift3d_cuda.so."""
    import torch.distributed as dist
    L = _CudaLib.get()
    rank, world = dist.get_rank(), dist.get_world_size()
    ident = np.zeros(128, np.uint8)
    if rank == 0 and L.s3d_nccl_unique_id(ident.ctypes.data) != 0:
        raise RuntimeError("s3d_nccl_unique_id failed")
    box = [ident.tobytes()]
    dist.broadcast_object_list(box, src=0)
    ident = np.frombuffer(box[0], np.uint8).copy()
    h = C.c_void_p()
    if L.s3d_comm_create_nccl(C.byref(h), rank, world, ident.ctypes.data, device) != 0:
        raise RuntimeError("s3d_comm_create_nccl failed")
    return Comm(h, rank, world)


def detect_slab(s: "capi.Sift3D", slab_vol: np.ndarray, zsplit: Sequence[int], comm: Comm,
                units=(1.0, 1.0, 1.0), copy: bool = True) -> np.ndarray:
    """`SIFT3D_detect_keypoints_slab` (extension of sift.h for tiled volumes): `slab_vol` is
    this rank's planes [zsplit[rank], zsplit[rank+1]) as a [z][y][x] float32 array."""
    f = s.L.SIFT3D_detect_keypoints_slab
    f.argtypes = [C.POINTER(capi.SIFT3D), C.POINTER(capi.Image), C.c_void_p, C.c_void_p,
                  C.POINTER(capi.Keypoint_store)]
    f.restype = C.c_int
    vol = np.ascontiguousarray(slab_vol, np.float32)
    im = capi.make_image(vol, units)
    zs = np.asarray(zsplit, np.int32)
    rc = f(C.byref(s.s), C.byref(im), zs.ctypes.data, comm.h, C.byref(s.kp))
    if rc != 0:
        raise RuntimeError(f"SIFT3D_detect_keypoints_slab returned {rc}")
    return s.keypoints(copy)


def slab_info(s: "capi.Sift3D", o: int) -> dict:
    """Planes of octave o this rank owns / holds (`s3d_slab_info`)."""
    L = _CudaLib.get()
    eng = s.L.sift3d_b200_engine
    eng.argtypes = [C.POINTER(capi.SIFT3D)]
    eng.restype = C.c_void_p
    info = np.zeros(6, np.int32)
    if L.s3d_slab_info(eng(C.byref(s.s)), o, info.ctypes.data) != 0:
        raise RuntimeError("s3d_slab_info failed (not in slab mode?)")
    return dict(own0=int(info[0]), own1=int(info[1]), lo=int(info[2]), hi=int(info[3]),
                nz=int(info[4]), halo=int(info[5]))
