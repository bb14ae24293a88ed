#!/usr/bin/env python
"""bench.py -- voxels/s through pyramid + descriptors on B200 (BASELINE.json metric).

One "step" = SIFT3D_detect_keypoints + SIFT3D_extract_descriptors over one synthetic
512^3 float32 volume (BASELINE.json configs[1]; 7 octaves x 3 keypoint levels -- the
octave count is not settable in the reference).

  value      : device-resident throughput (volume already in HBM, results left in HBM),
               timed with CUDA events on the launching stream, max over ranks.
  e2e        : the same step through the drop-in C API (libsift3D.so) with HOST buffers:
               H2D of the pinned volume and D2H of keypoints + descriptors inside the timed
               region.
  roofline   : the separable 3-D Gaussian (the dominant pyramid kernel): 8 B/voxel
               algorithmic traffic / CUDA-event time of each launch, against
               the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline: the reference's own OpenMP CPU path (oracle/_ref, unmodified sources) on
               a bounded sample of the same workload, all host cores.

N > 1 (torchrun): independent volumes, one per GPU (BASELINE.json configs[3]); no
data-path collective, weak scaling; the only communication is the timing all-reduce.

`--impl reference` times the reference CPU implementation instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))  # oracle_api: only the CPU-baseline legs use it

METRIC = "voxels/sec through pyramid+descriptors"
UNIT = "voxels/s"
HBM_FALLBACK_GBS = 6650.0  # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent


def blob_volume_torch(n, seed, device):
    from sift3d_b200.volumes import blob_volume_torch as gen
    return gen((n, n, n), seed, device)


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
    """dram bytes (read+write) of the widest blur launch from the committed ncu --set full
    capture (profiles/r01_ncu_final_blur_w17.csv); None if the capture is absent."""
    p = REPO / "profiles" / "r01_ncu_final_blur_w17.csv"
    if not p.exists():
        return None
    tot = 0.0
    for ln in p.read_text().splitlines():
        f = ln.split(",")
        if f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum") and len(f) >= 3:
            tot += float(f[2]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(f[1], 1.0)
    return tot or None


def pyramid_filters():
    """sigma and width of the default pyramid's filters."""
    s = [1.6 * 2 ** (k / 3.0) for k in range(-1, 5)]
    sig = [float(np.sqrt(s[0] ** 2 - 1.15 ** 2))] + [float(np.sqrt(s[i + 1] ** 2 - s[i] ** 2))
                                                     for i in range(5)]
    return sig


def gauss_taps(sigma):
    """init_Gauss_filter (imutil.c:3657-3710) restated for the bench's kernel-level timing."""
    hw = max(int(np.ceil(sigma * 3.0)), 1)
    x = (np.arange(2 * hw + 1, dtype=np.float64) - hw) / (sigma + np.finfo(np.float64).eps)
    k = np.exp(-0.5 * x * x).astype(np.float32)
    acc = np.float32(0)
    for v in k:
        acc = np.float32(acc + v)
    return (k / acc).astype(np.float32)


def cpu_reference_run(n_sample, seed, threads=None):
    """Time the reference's own CPU path (oracle/_ref) -- or the oracle port if _ref is
    absent -- on an n_sample^3 volume of the same generator.  Returns (voxels/s, info)."""
    from sift3d_b200 import capi
    from sift3d_b200.volumes import blob_volume
    vol = blob_volume(n_sample, seed=seed)
    cores = threads or os.cpu_count()
    import oracle_api
    if oracle_api.REF_LIB.exists():
        ref = oracle_api.load_reference()
        with capi.Sift3D(ref) as s:
            t0 = time.perf_counter()
            kp = s.detect_keypoints(vol)
            t1 = time.perf_counter()
            if len(kp):
                s.extract_descriptors()
            t2 = time.perf_counter()
        kind = "reference"
    else:
        from oracle_api import Oracle
        orc = Oracle()
        t0 = time.perf_counter()
        kp = orc.detect(vol)
        t1 = time.perf_counter()
        if len(kp):
            orc.describe(kp)
        t2 = time.perf_counter()
        kind = "port"
    return vol.size / (t2 - t0), dict(kind=kind, cores=cores, detect_s=t1 - t0, describe_s=t2 - t1,
                                      keypoints=int(len(kp)),
                                      sample=f"{n_sample}^3 blob volume (seed {seed}), detect+describe, "
                                             f"OMP threads={cores}")


def cpu_reference_subprocess(n_sample, seed, threads, reps=1):
    """Run cpu_reference_run in a fresh process with OMP_NUM_THREADS=threads (libgomp reads it
    at load time).  Returns (best voxels/s over reps, info)."""
    code = ("import sys, json; sys.path.insert(0, %r); import bench; best=None\n"
            "for i in range(%d):\n"
            "    v, info = bench.cpu_reference_run(%d, %d, threads=%d)\n"
            "    if best is None or v > best[0]: best = (v, info)\n"
            "print('RESULT ' + json.dumps([best[0], best[1]]))\n") % (str(REPO), reps, n_sample, seed, threads)
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OMP_WAIT_POLICY="passive")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    for ln in r.stdout.splitlines():
        if ln.startswith("RESULT "):
            v, info = json.loads(ln[7:])
            return v, info
    raise RuntimeError("reference subprocess failed: " + r.stderr[-500:])


def best_cpu_threads(ncpu):
    """The reference's OpenMP regions are many short loops between serial transposes; on a
    128-core host it is fastest well below the core count.  Calibrate on a small volume."""
    cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
    best_t, best_v = cands[0], -1.0
    for t in cands:
        v, _ = cpu_reference_subprocess(64, 1234, t)
        if v > best_v:
            best_t, best_v = t, v
    return best_t


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.ref_size
    ncpu = os.cpu_count() or 1
    threads = best_cpu_threads(ncpu)
    times, info = [], None
    for it in range(args.warmup + args.steps):
        v, info = cpu_reference_subprocess(n, 1234 + it % 4, threads)
        if it >= args.warmup:
            times.append(n ** 3 / v)
    ms = 1e3 * float(np.mean(times))
    value = n ** 3 / (ms * 1e-3)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"bounded sample of configs[1]: {n}^3 synthetic float32 volume per step "
                               "(the reference CPU path needs minutes for 512^3), kpSift3D defaults",
                   "host_cores": ncpu, "omp_threads": threads,
                   "note": "OMP thread count calibrated for best reference throughput"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": info["kind"],
                         "sample": info["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512, help="edge of the synthetic volume")
    ap.add_argument("--ref-size", type=int, default=128, help="edge of the CPU-arm sample volume")
    ap.add_argument("--cpu-sample", type=int, default=160, help="edge of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    n = args.size
    nvox = n ** 3
    vol_dev = blob_volume_torch(n, seed=1234 + rank, device=dev)   # per-rank independent volume
    vol_pinned = torch.empty((n, n, n), dtype=torch.float32, pin_memory=True)
    vol_pinned.copy_(vol_dev)
    vol_host = vol_pinned.numpy()
    torch.cuda.synchronize()

    lib = capi.load_b200()
    cu = C.CDLL(str(capi.CUDA_LIB))
    lib.lib.sift3d_b200_engine.restype = C.c_void_p
    lib.lib.sift3d_b200_engine.argtypes = [C.POINTER(capi.SIFT3D)]
    for f, at in (("s3d_image_from_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
                  ("s3d_build_pyramid", [C.c_void_p]),
                  ("s3d_detect_extrema", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
                  ("s3d_assign_orientations", [C.c_void_p, C.c_double, C.POINTER(C.c_int)]),
                  ("s3d_extract_descriptors_device", [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
                  ("s3d_engine_set_stream", [C.c_void_p, C.c_void_p]),
                  ("s3d_engine_launch_count", [C.c_void_p])):
        getattr(cu, f).argtypes = at
    cu.s3d_device_keypoints.argtypes = [C.c_void_p]
    cu.s3d_device_keypoints.restype = C.c_void_p
    cu.s3d_engine_launch_count.restype = C.c_longlong

    s = capi.Sift3D(lib)
    # a real (non-NULL) stream: the engine launches on it and the events below bracket it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def e2e_step():
        # views of the caller-visible stores (what a C caller reads): no extra Python copy
        kp = s.detect_keypoints(vol_host, copy=False)       # H2D inside (pinned source)
        d = s.extract_descriptors(copy=False) if len(kp) else None   # D2H of descriptors inside
        return len(kp), (0 if d is None else d.nbytes) + kp.nbytes

    # first call creates the engine, sizes the pyramid and warms every kernel
    nkp, d2h = e2e_step()
    eng = lib.lib.sift3d_b200_engine(C.byref(s.s))
    cu.s3d_engine_set_stream(eng, C.c_void_p(stream.cuda_stream))
    cu.s3d_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
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
    ms_dev = ev0.elapsed_time(ev1) / args.steps

    # ---- end-to-end timing through the C API with host buffers -----------------------------
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ee0.record(stream)
    for _ in range(args.steps):
        nkp_e, d2h = e2e_step()
    ee1.record(stream)
    barrier()
    ms_e2e = 1e3 * (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()

    # max over ranks
    ms_dev, ms_e2e = sdist.max_over_ranks([ms_dev, ms_e2e], device=dev)

    # ---- stage split of one device-resident step (untimed extra pass, rank 0, informational) --
    stages = None
    if rank == 0:
        cu.s3d_engine_sync.argtypes = [C.c_void_p]
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
        src = vol_dev
        dst = torch.empty_like(vol_dev)
        per = []
        tot_ms, tot_bytes = 0.0, 0.0
        for sg in pyramid_filters():
            taps = gauss_taps(sg)
            for _ in range(3):
                e2.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record(stream)
            for _ in range(args.blur_reps):
                e2.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps)
            b.record(stream)
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / args.blur_reps
            gbs = 8.0 * nvox / (ms * 1e-3) / 1e9
            per.append({"width": int(len(taps)), "ms": round(ms, 4), "gbs": round(gbs, 1),
                        "frac": round(gbs / peak, 4)})
            tot_ms += ms
            tot_bytes += 8.0 * nvox
        ach = tot_bytes / (tot_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "separable 3-D Gaussian blur (6 octave-0 filters of the pyramid)",
                "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                "traffic": blur_traffic(), "peak_source": peak_src, "algorithmic_bytes_per_launch": 8.0 * nvox,
                "per_filter": per}
        e2.close()
        del dst

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        threads = best_cpu_threads(ncpu)
        v, info = cpu_reference_subprocess(args.cpu_sample, 1234, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "host_cores": ncpu, "kind": info["kind"],
               "sample": info["sample"], "detect_s": round(info["detect_s"], 2),
               "describe_s": round(info["describe_s"], 2), "keypoints": info["keypoints"]}

    s.close()
    if rank == 0:
        value = nvox * world / (ms_dev * 1e-3)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"single {n}^3 synthetic float32 volume per GPU (configs[1]; "
                                   f"configs[3] at N>1), 7 octaves x 3 keypoint levels, kpSift3D defaults",
                       "candidates": ncand, "keypoints": nkp,
                       "l2_note": f"inputs ({4 * nvox >> 20} MiB/level) larger than the 126 MB L2",
                       "parallelism": f"independent volumes x{world}, no data-path collective"},
            "e2e": {"value": nvox * world / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": 4 * nvox, "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if stages is not None:
            out["stages_ms"] = stages
        if roof is not None:
            out["roofline"] = roof
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
