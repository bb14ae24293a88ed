#!/usr/bin/env python
"""Generate the golden fixtures from the UNMODIFIED reference (oracle/_ref).

Run in the build container (needs /root/reference for the real-data crop and
oracle/_ref built by oracle/build_ref.sh):

    python tests/golden/make_golden.py

Every fixture stores the input (or the recipe to regenerate it) and the outputs
of the reference's own SIFT3D_detect_keypoints / SIFT3D_extract_descriptors /
SIFT3D_extract_dense_descriptors on it.  Fixtures are small (about 4 MB in total).
"""
import gzip
import hashlib
import struct
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent.parent / "oracle"))
import oracle_api  # noqa: E402
from sift3d_b200 import capi  # noqa: E402
from sift3d_b200.volumes import blob_volume, smooth_noise_volume  # noqa: E402


def read_nii(path):
    """NIfTI-1 float32 reader restating read_nii's scaling (imutil/nifti.c:100-111)."""
    raw = gzip.open(path).read()
    dim = struct.unpack("<8h", raw[40:56])
    datatype = struct.unpack("<h", raw[70:72])[0]
    assert datatype == 16, datatype
    pixdim = struct.unpack("<8f", raw[76:108])
    off = int(struct.unpack("<f", raw[108:112])[0])
    slope, inter = struct.unpack("<ff", raw[112:120])
    nx, ny, nz = dim[1:4]
    v = np.frombuffer(raw, "<f4", nx * ny * nz, off).reshape(nz, ny, nx)
    if slope != 0:
        v = (v.astype(np.float64) * np.float64(slope) + np.float64(inter)).astype(np.float32)
    return np.ascontiguousarray(v), tuple(float(p) for p in pixdim[1:4])


def level_digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8)).hexdigest()


def run_case(ref, name, vol, units=(1.0, 1.0, 1.0), store_input=True, dense_crop=None, **params):
    out = {"units": np.asarray(units, np.float64),
           "params": np.asarray([params.get("peak_thresh", 0.1), params.get("corner_thresh", 0.4),
                                 params.get("sigma_n", 1.15), params.get("sigma0", 1.6),
                                 params.get("num_kp_levels", 3)], np.float64)}
    with capi.Sift3D(ref, **params) as s:
        kp = s.detect_keypoints(vol, units)
        out["noct"] = np.int32(s.num_octaves())
        K = s.s.gpyr.num_kp_levels
        digs = []
        for o in range(s.num_octaves()):
            for lv in range(-1, K + 2):
                digs.append(f"g,{o},{lv}," + level_digest(s.level_data("gpyr", o, lv)))
            for lv in range(-1, K + 1):
                digs.append(f"d,{o},{lv}," + level_digest(s.level_data("dog", o, lv)))
        out["level_sha256"] = np.asarray(digs)
        # one full level for localisation of a mismatch
        out["gpyr_0_0"] = s.level_data("gpyr", 0, 0).astype(np.float32) if vol.size <= 48 ** 3 else \
            np.zeros(0, np.float32)
        for f in ("xd", "yd", "zd", "sd", "o", "s", "R"):
            out["kp_" + f] = kp[f]
        if len(kp):
            d = s.extract_descriptors()
            out["desc"] = d["hists"].astype(np.float32)
            out["desc_coords"] = np.stack([d["xd"], d["yd"], d["zd"], d["sd"]], 1)
        if dense_crop is not None:
            sub = np.ascontiguousarray(vol[dense_crop])
            out["dense_in_slices"] = np.asarray([[sl.start, sl.stop] for sl in dense_crop])
            out["dense"] = s.extract_dense_descriptors(sub, units)
    if store_input:
        out["input"] = vol
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"{name}: {vol.shape} units={units} octaves={int(out['noct'])} keypoints={len(kp)}")


def run_dense_rotate(ref):
    """dense_rotate = 1 (sift.c:2521-2588) on two small volumes (the reference is scalar here)."""
    out = {}
    cases = {"iso": (blob_volume((48, 52, 44), seed=7)[4:22, 6:26, 3:25], (1.0, 1.0, 1.0)),
             "aniso": (smooth_noise_volume((18, 20, 22), seed=3), (1.0, 0.8, 1.3))}
    for key, (vol, units) in cases.items():
        vol = np.ascontiguousarray(vol, np.float32)
        with capi.Sift3D(ref) as s:
            s.s.dense_rotate = 1
            out[key + "_input"] = vol
            out[key + "_units"] = np.asarray(units, np.float64)
            out[key + "_dense"] = s.extract_dense_descriptors(vol, units)
        print(f"dense_rotate/{key}: {vol.shape} units={units}")
    np.savez_compressed(HERE / "dense_rotate.npz", **out)


def candidate_count(prev, cur, nxt, peak_thresh):
    """numpy restatement of detect_extrema's test (sift3d/sift.c:1153-1208) on one DoG triple:
    strict 6 face neighbours + the same voxel one level below and above, |v| > f32(peak*dogmax)."""
    dogmax = np.float32(np.abs(cur).max())
    thr = np.float32(np.float64(peak_thresh) * np.float64(dogmax))
    c = cur[1:-1, 1:-1, 1:-1]
    nb = [cur[1:-1, 1:-1, 2:], cur[1:-1, 1:-1, :-2], cur[1:-1, 2:, 1:-1], cur[1:-1, :-2, 1:-1],
          cur[:-2, 1:-1, 1:-1], cur[2:, 1:-1, 1:-1], prev[1:-1, 1:-1, 1:-1], nxt[1:-1, 1:-1, 1:-1]]
    lo = np.ones(c.shape, bool)
    hi = np.ones(c.shape, bool)
    for n in nb:
        lo &= n > c
        hi &= n < c
    return int((((c > thr) | (c < -thr)) & (lo | hi)).sum())


def run_large(ref, n, seed=1234, desc_every=25, with_dense=False):
    """BASELINE.json configs[1] (n = 512) / the configs[2] volume (n = 256): the exact volume
    bench.py times, through the UNMODIFIED reference.  Stores digests of every pyramid level,
    the candidate count, the full keypoint table and every `desc_every`-th descriptor."""
    import time
    t0 = time.time()
    vol = blob_volume(n, seed=seed)
    out = {"n": np.int32(n), "seed": np.int32(seed), "input_sha256": np.asarray(level_digest(vol)),
           "desc_every": np.int32(desc_every),
           "units": np.ones(3), "params": np.asarray([0.1, 0.4, 1.15, 1.6, 3], np.float64)}
    print(f"blob{n}: volume in {time.time() - t0:.0f} s", flush=True)
    with capi.Sift3D(ref) as s:
        t0 = time.time()
        kp = s.detect_keypoints(vol)
        out["ref_detect_s"] = np.float64(time.time() - t0)
        out["noct"] = np.int32(s.num_octaves())
        digs, ncand = [], []
        for o in range(s.num_octaves()):
            for lv in range(-1, 5):
                digs.append(f"g,{o},{lv}," + level_digest(s.level_data("gpyr", o, lv)))
            dog = [s.level_data("dog", o, lv) for lv in range(-1, 4)]
            for lv in range(-1, 4):
                digs.append(f"d,{o},{lv}," + level_digest(dog[lv + 1]))
            for lv in range(0, 3):
                ncand.append(candidate_count(dog[lv], dog[lv + 1], dog[lv + 2], 0.1))
            del dog
        out["level_sha256"] = np.asarray(digs)
        out["candidates_per_level"] = np.asarray(ncand, np.int64)
        out["kp_xyz"] = np.stack([kp["xd"], kp["yd"], kp["zd"]], 1).astype(np.int16)
        assert np.array_equal(out["kp_xyz"].astype(np.float64),
                              np.stack([kp["xd"], kp["yd"], kp["zd"]], 1))
        out["kp_o"] = kp["o"].astype(np.int8)
        out["kp_s"] = kp["s"].astype(np.int8)
        out["kp_sd"] = kp["sd"]
        out["kp_R"] = kp["R"]
        t0 = time.time()
        d = s.extract_descriptors()
        out["ref_describe_s"] = np.float64(time.time() - t0)
        out["desc"] = d["hists"][::desc_every].astype(np.float32)
        out["desc_norm_sum"] = np.float64(np.linalg.norm(d["hists"].astype(np.float64), axis=1).sum())
        print(f"blob{n}: octaves={int(out['noct'])} candidates={int(sum(ncand))} keypoints={len(kp)} "
              f"detect {float(out['ref_detect_s']):.0f} s describe {float(out['ref_describe_s']):.0f} s",
              flush=True)
        if with_dense:
            t0 = time.time()
            dd = s.extract_dense_descriptors(vol)
            out["ref_dense_s"] = np.float64(time.time() - t0)
            out["dense_sub"] = dd[3::8, 3::8, 3::8].copy()
            out["dense_plane_sums"] = dd.astype(np.float64).sum(axis=(1, 2))
            out["dense_plane"] = dd[n // 2, ::2, ::2].copy()
            print(f"blob{n}: dense {float(out['ref_dense_s']):.0f} s", flush=True)
    np.savez_compressed(HERE / f"blob{n}_large.npz", **out)


def main():
    ref = oracle_api.load_reference()
    if "--large" in sys.argv:
        # run once in the build container (about 2 + 8 minutes of CPU on 8 cores)
        sizes = [int(a) for a in sys.argv[sys.argv.index("--large") + 1:]] or [256, 512]
        for n in sizes:
            run_large(ref, n, desc_every=10 if n <= 256 else 50, with_dense=(n == 256))
        return
    if "--dense-rotate-only" in sys.argv:
        run_dense_rotate(ref)
        return
    run_dense_rotate(ref)
    # 1. synthetic blobs, isotropic, 3 octaves (input regenerated from the seed by the tests)
    run_case(ref, "blob48", blob_volume((48, 52, 44), seed=7),
             dense_crop=(slice(4, 24), slice(6, 24), slice(3, 19)))
    # 2. anisotropic, non-dyadic units: exercises the general tap spacing
    run_case(ref, "aniso40", smooth_noise_volume((40, 45, 37), seed=3), units=(0.7, 0.9, 1.3),
             dense_crop=(slice(0, 18), slice(0, 20), slice(0, 16)))
    # 3. dyadic non-unit units + non-default parameters
    run_case(ref, "units2_params", smooth_noise_volume((36, 40, 44), seed=5), units=(1.0, 1.0, 2.0),
             peak_thresh=0.05, corner_thresh=0.3, num_kp_levels=2, sigma0=1.8)
    # 4. real data: a 64^3 crop of the reference's examples/data/1.nii.gz
    nii = Path("/root/reference/examples/data/1.nii.gz")
    if nii.exists():
        vol, units = read_nii(nii)
        crop = np.ascontiguousarray(vol[60:124, 80:144, 60:124])
        run_case(ref, "real_crop64", crop, units=units)
    # 5. larger synthetic (4 octaves): only keypoints/descriptors/digests, input from seed
    run_case(ref, "blob96", blob_volume(96, seed=1234), store_input=False)


if __name__ == "__main__":
    main()
