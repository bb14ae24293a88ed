#!/usr/bin/env python
"""bench.py -- voxels/s through pyramid + descriptors on B200 (BASELINE.json metric).

One "step" = SIFT3D_detect_keypoints + SIFT3D_extract_descriptors over one synthetic 512^3
float32 volume (BASELINE.json configs[1]; 7 octaves x 3 keypoint levels -- the octave count is
not settable in the reference).  The volume is `blob_volume(512, 1234 + rank)`
(numpy): rank 0 times the very volume the committed golden file
tests/golden/blob512_large.npz was generated from with the UNMODIFIED reference, and the run
asserts its candidate and keypoint counts against that file.

  value        device-resident throughput (volume already in HBM, results left in HBM), CUDA
               events on the launching stream, max over ranks.
  e2e          the same step through the drop-in C API (libsift3D.so) with HOST buffers: H2D of
               the volume from PAGEABLE (malloc) memory -- what a stock caller has after
               im_read -- and D2H of keypoints + descriptors inside the timed region;
               `e2e.pinned` is the same from a pinned source, beside it.
  roofline     the separable 3-D Gaussian (dominant pyramid kernel): 8 B/voxel algorithmic
               traffic / CUDA-event time of each launch against the measured HBM
               copy bandwidth (MEASURED_PEAKS.json); `traffic` = DRAM bytes per launch from
               this round's ncu capture (profiles/r02_ncu_blur_traffic.json, written by
               tools/ncu_blur_traffic.py).
  cpu_baseline the reference's own OpenMP CPU path (oracle/_ref, unmodified sources) on the SAME
               512^3 volume, one run, detect/describe split, all host cores (N = 1 only).
  dense        BASELINE.json configs[2]: SIFT3D_extract_dense_descriptors on a 256^3 volume
               through the C API (N = 1 only).
  slab         N > 1 only -- BASELINE.json configs[4]: ONE 2048 x 2048 x (128 N) volume Z-slab
               tiled over the N ranks with NCCL halo exchange (ms/step, voxels/s, halo bytes,
               time in exchanges), plus a parity leg: 512 x 512 x 1024 tiled over the N ranks
               against the whole-volume result of rank 0 (bit-identical keypoints/descriptors).

N > 1 (torchrun): independent volumes, one per GPU (BASELINE.json configs[3]); no data-path
collective, weak scaling; the only communication of that leg is the timing all-reduce.

`--impl reference` times the reference CPU implementation instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))  # oracle_api: only the CPU-baseline legs use it

METRIC = "voxels/sec through pyramid+descriptors"
UNIT = "voxels/s"
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent
SEED = 1234


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.gpu = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def hbm_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def blur_traffic():
    """DRAM bytes (read + write) per blur launch, mean over the six pyramid filters, from this
    round's `ncu --set full` capture (tools/ncu_blur_traffic.py -> profiles/r02_ncu_blur_traffic.json).
    Returns (bytes or None, per-filter list or None)."""
    p = REPO / "profiles" / "r02_ncu_blur_traffic.json"
    if not p.exists():
        return None, None
    try:
        d = json.loads(p.read_text())
        per = d["per_filter"]
        return float(np.mean([f["dram_bytes"] for f in per])), per
    except Exception:
        return None, None


def pyramid_filters():
    """sigma of the default pyramid's six octave-0 filters."""
    s = [1.6 * 2 ** (k / 3.0) for k in range(-1, 5)]
    return [float(np.sqrt(s[0] ** 2 - 1.15 ** 2))] + [float(np.sqrt(s[i + 1] ** 2 - s[i] ** 2))
                                                      for i in range(5)]


def gauss_taps(sigma):
    """init_Gauss_filter (imutil.c:3657-3710) restated for the bench's kernel-level timing."""
    hw = max(int(np.ceil(sigma * 3.0)), 1)
    x = (np.arange(2 * hw + 1, dtype=np.float64) - hw) / (sigma + np.finfo(np.float64).eps)
    k = np.exp(-0.5 * x * x).astype(np.float32)
    acc = np.float32(0)
    for v in k:
        acc = np.float32(acc + v)
    return (k / acc).astype(np.float32)


def sha256(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def golden_for(n, seed):
    """Counts the unmodified reference produced on blob_volume(n, seed) (tests/golden)."""
    p = REPO / "tests" / "golden" / f"blob{n}_large.npz"
    if not p.exists():
        return None
    z = np.load(p)
    if int(z["n"]) != n or int(z["seed"]) != seed:
        return None
    return {"candidates": int(z["candidates_per_level"].sum()), "keypoints": int(len(z["kp_o"])),
            "input_sha256": str(z["input_sha256"]), "file": f"tests/golden/{p.name}"}


# --------------------------------------------------------------------------- CPU reference legs
def cpu_reference_run(vol_path, threads, reps=1):
    """Time the reference's own CPU path (oracle/_ref; the oracle port if _ref is absent) on
    the volume stored at `vol_path` (.npy).  Returns a dict; runs in a subprocess (below)."""
    from sift3d_b200 import capi
    import oracle_api
    vol = np.load(vol_path)
    best = None
    for _ in range(reps):
        if oracle_api.REF_LIB.exists():
            with capi.Sift3D(oracle_api.load_reference()) as s:
                t0 = time.perf_counter()
                kp = s.detect_keypoints(vol, copy=False)
                nk = len(kp)
                t1 = time.perf_counter()
                if nk:
                    s.extract_descriptors(copy=False)
                t2 = time.perf_counter()
            kind = "reference"
        else:
            orc = oracle_api.Oracle()
            t0 = time.perf_counter()
            kp = orc.detect(vol)
            nk = len(kp)
            t1 = time.perf_counter()
            if nk:
                orc.describe(kp)
            t2 = time.perf_counter()
            kind = "port"
        r = dict(kind=kind, cores=threads, detect_s=t1 - t0, describe_s=t2 - t1, total_s=t2 - t0,
                 keypoints=int(nk), voxels=int(vol.size))
        if best is None or r["total_s"] < best["total_s"]:
            best = r
    return best


def cpu_reference_subprocess(vol_path, threads, reps=1, steps=1):
    """cpu_reference_run in a fresh process with OMP_NUM_THREADS=threads (libgomp reads it at
    load time); `steps` back-to-back runs, one result dict per run."""
    code = ("import sys, json; sys.path.insert(0, %r); import bench\n"
            "for i in range(%d):\n"
            "    print('RESULT ' + json.dumps(bench.cpu_reference_run(%r, %d, %d)), flush=True)\n"
            ) % (str(REPO), steps, str(vol_path), threads, reps)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_WAIT_POLICY="passive")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    out = [json.loads(ln[7:]) for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    if len(out) != steps:
        raise RuntimeError("reference subprocess failed: " + r.stderr[-500:])
    return out


def shm_dir():
    d = Path("/dev/shm")
    return d if d.is_dir() and os.access(d, os.W_OK) else Path(tempfile.gettempdir())


def save_volume(vol, tag):
    p = shm_dir() / f"s3d_bench_{os.getpid()}_{tag}.npy"
    np.save(p, vol)
    return p


def best_cpu_threads(ncpu):
    """The reference's OpenMP regions are many short loops between serial transposes; on a
    many-core host it is fastest well below the core count.  Calibrate on a small volume.
    Returns (threads, voxels/s of the calibration volume)."""
    from sift3d_b200.volumes import blob_volume
    p = save_volume(blob_volume(96, seed=SEED), "cal")
    try:
        cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
        best_t, best_v = cands[0], -1.0
        for t in cands:
            r = cpu_reference_subprocess(p, t)[0]
            v = r["voxels"] / r["total_s"]
            if v > best_v:
                best_t, best_v = t, v
    finally:
        p.unlink(missing_ok=True)
    return best_t, best_v


def run_reference_arm(args):
    """The reference's own CPU implementation on the host cores (rank 0 only).  A 512^3 step
    takes the reference minutes, so a step is the largest cube of the same generator whose
    (warmup + steps) runs fit the arm's time budget; with --steps 1 --warmup 0 that is the full
    512^3 volume.  The size is stated in config.workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sift3d_b200.volumes import blob_volume
    ncpu = os.cpu_count() or 1
    threads, cal_rate = best_cpu_threads(ncpu)
    nrun = args.warmup + args.steps
    if args.ref_size > 0:
        n = args.ref_size
    else:  # throughput drops with size (serial transposes grow): assume 70 % of the calibration rate
        per_step = args.ref_budget_s / max(nrun, 1)
        n = int((0.7 * cal_rate * per_step) ** (1.0 / 3.0)) // 8 * 8
        n = max(64, min(args.size, n))
    p = save_volume(blob_volume(n, seed=SEED), "ref")
    try:
        res = cpu_reference_subprocess(p, threads, steps=nrun)[args.warmup:]
    finally:
        p.unlink(missing_ok=True)
    ms = 1e3 * float(np.mean([r["total_s"] for r in res]))
    value = n ** 3 / (ms * 1e-3)
    sample = (f"{n}^3 blob volume (seed {SEED}), detect+describe, OMP threads={threads}; "
              f"detect {np.mean([r['detect_s'] for r in res]):.2f} s, "
              f"describe {np.mean([r['describe_s'] for r in res]):.2f} s, {res[0]['keypoints']} keypoints")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (f"configs[1] generator at {n}^3 per step: the largest cube whose "
                                f"{nrun} runs fit the {args.ref_budget_s:.0f} s budget of this arm "
                                f"(512^3 = one run of minutes: bench.py cpu_baseline), kpSift3D defaults"
                                if n != args.size else
                                f"single {n}^3 synthetic float32 volume (configs[1]), kpSift3D defaults"),
                   "size": n, "host_cores": ncpu, "omp_threads": threads,
                   "note": "OMP thread count calibrated for best reference throughput"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": res[0]["kind"],
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: keep this rank's threads (and so its first-touch host memory: the
    volume, the pinned staging ring) on the NUMA node its GPU hangs off -- what `numactl` would
    do in a deployment.  Best effort: returns a description, or None when the topology is not
    readable (then nothing is changed)."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id
        bdf = f"{torch.cuda.get_device_properties(local_rank).pci_domain_id:04x}:{bdf:02x}:" \
              f"{torch.cuda.get_device_properties(local_rank).pci_device_id:02x}.0"
        node = int(Path(f"/sys/bus/pci/devices/{bdf}/numa_node").read_text())
        if node < 0:
            return None
        cpus = set()
        for part in Path(f"/sys/devices/system/node/node{node}/cpulist").read_text().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception:
        return None


# --------------------------------------------------------------------------- device plumbing
def bind_cuda(capi):
    cu = C.CDLL(str(capi.CUDA_LIB))
    for f, at in (("s3d_image_from_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
                  ("s3d_slab_image_from_device", [C.c_void_p, C.c_void_p]),
                  ("s3d_build_pyramid", [C.c_void_p]),
                  ("s3d_detect_extrema", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
                  ("s3d_assign_orientations", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
                  ("s3d_extract_descriptors_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
                  ("s3d_engine_set_stream", [C.c_void_p, C.c_void_p]),
                  ("s3d_engine_launch_count", [C.c_void_p]),
                  ("s3d_engine_sync", [C.c_void_p]),
                  ("s3d_set_option", [C.c_void_p, C.c_char_p, C.c_int]),
                  ("s3d_dense_last_timing", [C.c_void_p, C.c_void_p]),
                  ("s3d_slab_stats", [C.c_void_p, C.c_void_p])):
        getattr(cu, f).argtypes = at
    cu.s3d_device_keypoints.argtypes = [C.c_void_p]
    cu.s3d_device_keypoints.restype = C.c_void_p
    cu.s3d_engine_launch_count.restype = C.c_longlong
    return cu


def engine_of(lib, s, capi):
    lib.lib.sift3d_b200_engine.restype = C.c_void_p
    lib.lib.sift3d_b200_engine.argtypes = [C.POINTER(capi.SIFT3D)]
    return lib.lib.sift3d_b200_engine(C.byref(s.s))


def dense_leg(lib, capi, cu, n):
    """BASELINE.json configs[2]: SIFT3D_extract_dense_descriptors (sift.c:2354) on n^3 through
    the C API, host buffers in and out (the 48 B/voxel result lands in the caller's Image)."""
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume(n, seed=SEED)
    out = {"workload": f"SIFT3D_extract_dense_descriptors, {n}^3 float32 volume (configs[2]), "
                       "dense_rotate=0, host buffers in and out"}
    with capi.Sift3D(lib) as s:
        im = capi.make_image(vol)
        res = capi.empty_image()
        times = []
        for rep in range(4):
            t0 = time.perf_counter()
            rc = s.L.SIFT3D_extract_dense_descriptors(C.byref(s.s), C.byref(im), C.byref(res))
            times.append(time.perf_counter() - t0)
            if rc != 0:
                raise RuntimeError("SIFT3D_extract_dense_descriptors failed")
        ms3 = (C.c_double * 3)()
        have = cu.s3d_dense_last_timing(engine_of(lib, s, capi), ms3) == 0
        arr = np.ctypeslib.as_array(res.data, shape=(res.nx * res.ny * res.nz * res.nc,))
        ok = bool(np.isfinite(arr[::997]).all())
        lib._libc.free(C.cast(res.data, C.c_void_p))
    ms = 1e3 * float(np.min(times[1:]))
    peak, _ = hbm_peak()
    out.update({"ms_e2e": round(ms, 3), "voxels_per_s_e2e": n ** 3 / (ms * 1e-3),
                "h2d_bytes": 4 * n ** 3, "d2h_bytes": 48 * n ** 3, "finite": ok})
    if have:
        dev = float(ms3[1])
        out.update({"ms_upload": round(float(ms3[0]), 3), "ms_kernels": round(dev, 3),
                    "ms_download": round(float(ms3[2]), 3),
                    "algorithmic_bytes": 52 * n ** 3,
                    "kernels_gbs": round(52 * n ** 3 / (dev * 1e-3) / 1e9, 1),
                    "kernels_frac_of_hbm": round(52 * n ** 3 / (dev * 1e-3) / 1e9 / peak, 4)})
    return out


def slab_leg(args, lib, capi, cu, dev, rank, world, local_rank, stream):
    """BASELINE.json configs[4]: ONE volume Z-slab tiled over the ranks, NCCL halo exchange."""
    import torch
    import torch.distributed as dist
    from sift3d_b200 import slab
    from sift3d_b200.volumes import blob_volume_torch
    comm = slab.nccl_comm(local_rank)
    out = {}

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # ---- parity leg: 512 x 512 x 1024 tiled over the ranks vs the whole volume on rank 0 ----
    pz, py, px = args.slab_parity
    whole_t = blob_volume_torch((pz, py, px), SEED + 77, dev) if rank == 0 else \
        torch.empty((pz, py, px), dtype=torch.float32, device=dev)
    dist.broadcast(whole_t, src=0)   # every rank slices the SAME bits
    zs = slab.split_planes(pz, world)
    mine = whole_t[zs[rank]:zs[rank + 1]].cpu().numpy()
    whole = whole_t.cpu().numpy() if rank == 0 else None
    del whole_t
    with capi.Sift3D(lib) as s:
        kp = slab.detect_slab(s, mine, zs, comm)
        d = s.extract_descriptors() if len(kp) else np.zeros(0, capi.DESCRIPTOR_DTYPE)
    del mine
    gathered = [None] * world if rank == 0 else None
    dist.gather_object((kp, d), gathered, dst=0)
    if rank == 0:
        kps, ds = [g[0] for g in gathered], [g[1] for g in gathered]
        mk, md = slab.merge_ranks(kps, ds) if sum(len(k) for k in kps) else (kps[0], ds[0])
        with capi.Sift3D(lib) as s:
            wk = s.detect_keypoints(whole)
            wd = s.extract_descriptors() if len(wk) else np.zeros(0, capi.DESCRIPTOR_DTYPE)
        same_kp = len(mk) == len(wk) and all(np.array_equal(mk[f], wk[f]) for f in
                                              ("xd", "yd", "zd", "sd", "o", "s", "R"))
        same_d = same_kp and np.array_equal(md["hists"], wd["hists"])
        out["parity"] = {"volume": f"{px}x{py}x{pz} tiled over {world} ranks vs whole volume on rank 0",
                         "keypoints": int(len(wk)), "keypoints_identical": bool(same_kp),
                         "descriptors_bit_identical": bool(same_d)}
        del gathered, kps, ds, mk, md, wk, wd
    del whole
    barrier()

    # ---- throughput leg: nx x ny x (nzl * world) --------------------------------------------
    nx, ny, nzl = args.slab_nx, args.slab_ny, args.slab_nzl
    zs = [nzl * r for r in range(world + 1)]
    vol = blob_volume_torch((nzl, ny, nx), SEED + rank, dev)
    host = vol.cpu().numpy()
    s = capi.Sift3D(lib)
    kp = slab.detect_slab(s, host, zs, comm, copy=False)   # sizes everything, warms the kernels
    del host
    eng = engine_of(lib, s, capi)
    cu.s3d_engine_set_stream(eng, C.c_void_p(stream.cuda_stream))
    desc = torch.empty(max(len(kp), 1) * 3104 * 2 + 4096, dtype=torch.uint8, device=dev)

    def step():
        nc, nk = C.c_int(0), C.c_int(0)
        rc = cu.s3d_slab_image_from_device(eng, vol.data_ptr())
        rc |= cu.s3d_build_pyramid(eng)
        rc |= cu.s3d_detect_extrema(eng, s.s.peak_thresh, C.byref(nc))
        rc |= cu.s3d_assign_orientations(eng, s.s.corner_thresh, C.byref(nk))
        if nk.value > 0:
            rc |= cu.s3d_extract_descriptors_device(eng, cu.s3d_device_keypoints(eng), nk.value,
                                                    desc.data_ptr())
        if rc:
            raise RuntimeError("slab step failed")
        return nc.value, nk.value

    for _ in range(2):
        step()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(args.slab_steps):
        nc, nk = step()
    b.record(stream)
    barrier()
    ms = a.elapsed_time(b) / args.slab_steps
    # one extra, untimed step with events around every exchange
    cu.s3d_set_option(eng, b"slab_timing", 1)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    cu.s3d_slab_image_from_device(eng, vol.data_ptr())
    ev[0].record(stream)
    cu.s3d_build_pyramid(eng)
    ev[1].record(stream)
    torch.cuda.synchronize()
    st = (C.c_double * 5)()
    cu.s3d_slab_stats(eng, st)
    cu.s3d_set_option(eng, b"slab_timing", 0)
    t = torch.tensor([ms, float(nk), float(nc), st[0], st[3], st[4], ev[0].elapsed_time(ev[1])],
                     dtype=torch.float64, device=dev)
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    s.close()
    comm.close()
    if rank == 0:
        nvox = nx * ny * nzl * world
        out.update({
            "workload": f"one {nx}x{ny}x{nzl * world} float32 volume Z-slab tiled over {world} GPUs "
                        f"({nzl} planes per rank), NCCL halo exchange (configs[4] shape per rank)",
            "ms_per_step": float(mx[0]), "voxels_per_s": nvox / (float(mx[0]) * 1e-3),
            "steps": args.slab_steps, "keypoints": int(sm[1]), "candidates": int(sm[2]),
            "halo_bytes_sent_per_step": float(sm[3]),
            "pyramid_ms_max": round(float(mx[6]), 3),
            "halo_exchange_ms_max": round(float(mx[4]), 3),
            "allreduce_ms_max": round(float(mx[5]), 3),
            "exchange_note": "CUDA-event intervals around ncclSend/ncclRecv groups and the two "
                             "all-reduces on the engine stream of one extra step (they include "
                             "waiting for the slower neighbour); the pyramid is the only stage "
                             "that communicates"})
    return out


# --------------------------------------------------------------------------- main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="edge of the synthetic volume")
    ap.add_argument("--ref-size", type=int, default=0,
                    help="edge of the CPU arm's volume (0 = the largest that fits --ref-budget-s)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    ap.add_argument("--cpu-sample", type=int, default=-1,
                    help="edge of the cpu_baseline volume (-1 = --size, the full workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true")
    ap.add_argument("--no-slab", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="N > 1: do not bind each rank to the NUMA node of its GPU")
    ap.add_argument("--dense-size", type=int, default=256)
    ap.add_argument("--slab-nx", type=int, default=2048)
    ap.add_argument("--slab-ny", type=int, default=2048)
    ap.add_argument("--slab-nzl", type=int, default=128, help="planes per rank of the slab leg")
    ap.add_argument("--slab-steps", type=int, default=3)
    ap.add_argument("--slab-parity", type=int, nargs=3, default=[1024, 512, 512],
                    metavar=("NZ", "NY", "NX"))
    ap.add_argument("--blur-reps", type=int, default=5)
    ap.add_argument("--opt", action="append", default=[],
                    help="engine option name=value for A/B runs (s3d_set_option), repeatable")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    from sift3d_b200 import capi
    from sift3d_b200.engine_api import Engine
    from sift3d_b200.volumes import blob_volume

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from sift3d_b200 import dist as sdist
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        sdist.init_process_group("nccl", device=dev)
    os.environ["SIFT3D_CUDA_DEVICE"] = str(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 and not args.no_numa_bind else None

    n = args.size
    nvox = n ** 3
    vol_host = blob_volume(n, seed=SEED + rank)        # pageable (malloc) memory, like im_read's
    vol_pinned_t = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
    vol_pinned_t.copy_(torch.from_numpy(vol_host))
    vol_pinned = vol_pinned_t.numpy()
    vol_dev = vol_pinned_t.to(dev)
    torch.cuda.synchronize()
    gold = golden_for(n, SEED) if rank == 0 else None
    golden_note = "no golden file for this size"
    if gold is not None and sha256(vol_host) != gold["input_sha256"]:
        raise SystemExit("bench.py: blob_volume() differs from the volume the golden was made from")

    lib = capi.load_b200()
    cu = bind_cuda(capi)
    s = capi.Sift3D(lib)
    # a real (non-NULL) stream: the engine launches on it and the events below bracket it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def e2e_step(src):
        # views of the caller-visible stores (what a C caller reads): no extra Python copy
        kp = s.detect_keypoints(src, copy=False)              # H2D inside
        d = s.extract_descriptors(copy=False) if len(kp) else None   # D2H of descriptors inside
        return len(kp), (0 if d is None else d.nbytes) + kp.nbytes

    # first call creates the engine, sizes the pyramid and warms every kernel
    nkp, d2h = e2e_step(vol_pinned)
    eng = engine_of(lib, s, capi)
    cu.s3d_engine_set_stream(eng, C.c_void_p(stream.cuda_stream))
    for kv in args.opt:
        k, v = kv.split("=")
        if cu.s3d_set_option(eng, k.encode(), int(v)) != 0:
            raise SystemExit(f"unknown engine option {k}")
    desc_dev = torch.empty(max(nkp, 1) * 3104 + 4096, dtype=torch.uint8, device=dev)

    def dev_step():
        nc, nk = C.c_int(0), C.c_int(0)
        rc = cu.s3d_image_from_device(eng, vol_dev.data_ptr(), n, n, n)
        rc |= cu.s3d_build_pyramid(eng)
        rc |= cu.s3d_detect_extrema(eng, s.s.peak_thresh, C.byref(nc))
        rc |= cu.s3d_assign_orientations(eng, s.s.corner_thresh, C.byref(nk))
        if nk.value > 0:
            if nk.value * 3104 > desc_dev.numel():
                raise RuntimeError("descriptor buffer too small")
            rc |= cu.s3d_extract_descriptors_device(eng, cu.s3d_device_keypoints(eng), nk.value,
                                                    desc_dev.data_ptr())
        if rc:
            raise RuntimeError("device step failed")
        return nc.value, nk.value

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ---------------------------------------------------
    for _ in range(args.warmup):
        ncand, nkp = dev_step()
    if gold is not None:   # the timed workload is the one the reference was run on: same counts
        if (ncand, nkp) != (gold["candidates"], gold["keypoints"]):
            raise SystemExit(f"bench.py: {ncand} candidates / {nkp} keypoints, the reference found "
                             f"{gold['candidates']} / {gold['keypoints']} ({gold['file']})")
        golden_note = (f"ok: {ncand} candidates / {nkp} keypoints = the unmodified reference on this "
                       f"volume ({gold['file']}; full parity: tests/test_gpu_large.py)")
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    l0 = cu.s3d_engine_launch_count(eng)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        ncand, nkp = dev_step()
    ev1.record(stream)
    barrier()
    launches = int(cu.s3d_engine_launch_count(eng) - l0)
    ms_dev_local = ev0.elapsed_time(ev1) / args.steps

    # ---- end-to-end timing through the C API with host buffers -----------------------------
    def time_e2e(src):
        for _ in range(2):
            e2e_step(src)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            _, nbytes = e2e_step(src)
        barrier()
        return 1e3 * (time.perf_counter() - t0) / args.steps, nbytes

    ms_e2e_local, d2h = time_e2e(vol_host)         # pageable source: the headline
    ms_e2e_pin_local, _ = time_e2e(vol_pinned)
    clocks = sampler.stop()

    # max over ranks (a multi-GPU step takes as long as its slowest rank), and every rank's own
    ms_dev, ms_e2e, ms_e2e_pin = sdist.max_over_ranks([ms_dev_local, ms_e2e_local, ms_e2e_pin_local],
                                                      device=dev)
    per_rank = None
    if dist is not None:
        mine = torch.tensor([ms_dev_local, ms_e2e_local, float(ncand), float(nkp)],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "ms_per_step": round(float(t[0]), 3), "ms_e2e": round(float(t[1]), 3),
                     "candidates": int(t[2]), "keypoints": int(t[3])} for r, t in enumerate(allr)]

    # ---- stage split of one device-resident step (untimed extra pass, rank 0, informational) --
    stages = None
    if rank == 0:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        nc_, nk_ = C.c_int(0), C.c_int(0)
        cu.s3d_image_from_device(eng, vol_dev.data_ptr(), n, n, n)
        evs[0].record(stream)
        cu.s3d_build_pyramid(eng)
        evs[1].record(stream)
        cu.s3d_detect_extrema(eng, s.s.peak_thresh, C.byref(nc_))
        evs[2].record(stream)
        cu.s3d_assign_orientations(eng, s.s.corner_thresh, C.byref(nk_))
        evs[3].record(stream)
        if nk_.value > 0:
            cu.s3d_extract_descriptors_device(eng, cu.s3d_device_keypoints(eng), nk_.value,
                                              desc_dev.data_ptr())
        evs[4].record(stream)
        torch.cuda.synchronize()
        names = ["pyramid+dog", "extrema", "orientation(+gradients)", "descriptors"]
        stages = {nm: round(evs[i].elapsed_time(evs[i + 1]), 3) for i, nm in enumerate(names)}

    # ---- roofline of the separable Gaussian (rank 0) ---------------------------------------
    roof = None
    if rank == 0:
        peak, peak_src = hbm_peak()
        e2 = Engine(local_rank)
        e2.set_stream(C.c_void_p(stream.cuda_stream))
        dst = torch.empty_like(vol_dev)
        per = []
        tot_ms, tot_bytes = 0.0, 0.0
        for sg in pyramid_filters():
            taps = gauss_taps(sg)
            for _ in range(3):
                e2.blur_device(vol_dev.data_ptr(), dst.data_ptr(), n, n, n, taps)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(args.blur_reps):
                e2.blur_device(vol_dev.data_ptr(), dst.data_ptr(), n, n, n, taps)
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.blur_reps
            gbs = 8.0 * nvox / (ms * 1e-3) / 1e9
            per.append({"width": int(len(taps)), "ms": round(ms, 4), "gbs": round(gbs, 1),
                        "frac": round(gbs / peak, 4)})
            tot_ms += ms
            tot_bytes += 8.0 * nvox
        ach = tot_bytes / (tot_ms * 1e-3) / 1e9
        traffic, traffic_per = blur_traffic()
        roof = {"bound": "hbm", "kernel": "separable 3-D Gaussian blur, k_blur_tma (TMA + mbarrier fed fused X/Y/Z pass; 6 octave-0 filters of the pyramid)",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": 8.0 * nvox,
                "per_filter": per}
        if traffic_per is not None:
            roof["traffic_per_filter"] = traffic_per
            roof["traffic_source"] = "profiles/r02_ncu_blur_traffic.json (ncu --set full, this round)"
        e2.close()
        del dst

    s.close()

    dense = None
    if rank == 0 and world == 1 and not args.no_dense:
        dense = dense_leg(lib, capi, cu, args.dense_size)

    slab = None
    if world > 1 and not args.no_slab:
        slab = slab_leg(args, lib, capi, cu, dev, rank, world, local_rank, stream)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        threads, _ = best_cpu_threads(ncpu)
        nc_ = n if args.cpu_sample < 0 else args.cpu_sample
        p = save_volume(vol_host if nc_ == n else blob_volume(nc_, seed=SEED), "cpu")
        try:
            r = cpu_reference_subprocess(p, threads)[0]
        finally:
            p.unlink(missing_ok=True)
        cpu = {"value": r["voxels"] / r["total_s"], "unit": UNIT, "cores": threads, "host_cores": ncpu,
               "kind": r["kind"],
               "sample": (f"the whole workload: the same {nc_}^3 volume (seed {SEED}), one run, "
                          if nc_ == n else f"{nc_}^3 blob volume (seed {SEED}), one run, ") +
                         f"detect+describe, OMP threads={threads}",
               "detect_s": round(r["detect_s"], 2), "describe_s": round(r["describe_s"], 2),
               "keypoints": r["keypoints"]}

    if rank == 0:
        value = nvox * world / (ms_dev * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"single {n}^3 synthetic float32 volume per GPU (configs[1]; "
                                   f"configs[3] at N>1), 7 octaves x 3 keypoint levels, kpSift3D defaults; "
                                   f"blob_volume({n}, seed {SEED} + rank)",
                       "candidates": ncand, "keypoints": nkp, "golden_check": golden_note,
                       "l2_note": f"inputs ({4 * nvox >> 20} MiB/level) larger than the 126 MB L2",
                       "parallelism": f"independent volumes x{world}, no data-path collective" +
                                      (f"; ranks bound to their GPU's NUMA node (rank 0: {numa})" if numa else "")},
            "e2e": {"value": nvox * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 4 * nvox, "d2h_bytes_per_step": int(d2h),
                    "source": "pageable (malloc) host volume, as a stock im_read caller has",
                    "host_copy": (f"staged through pinned slots by min(8, {os.cpu_count()} cores / {world} ranks) host "
                                  f"threads ({'polling' if world == 1 else 'poller-free'} chunk pipeline, "
                                  "csrc/host_pipe.h); descriptor chunks of 4096 keypoints alternate between two "
                                  "compute streams, their D2H copies run behind the next chunks"),
                    "pinned": {"value": nvox * world / (ms_e2e_pin * 1e-3), "ms_per_step": ms_e2e_pin}},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if stages is not None:
            out["stages_ms"] = stages
        if per_rank is not None:
            out["per_rank"] = per_rank
        if roof is not None:
            out["roofline"] = roof
        if cpu is not None:
            out["cpu_baseline"] = cpu
        if dense is not None:
            out["dense"] = dense
        if slab is not None:
            out["slab"] = slab
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
