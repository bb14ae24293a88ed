"""End-to-end drop-in test (BASELINE.json configs[0]): the reference's UNMODIFIED kpSift3D
(cli/kpSift3D.c, compiled from the reference tree into oracle/_ref by oracle/build_ref.sh)
runs on `examples/data/1.nii.gz` twice --

  * kpSift3D_stock: resolves libsift3D.so to THIS repo's library (B200 path),
  * kpSift3D_cpu:   bound to the reference's own libsift3D (CPU path),

both with the reference-built libimutil and this repo's zlib NIfTI reader preloaded (the
reference's read_nii is an error stub without nifticlib).  The CSV files the two runs write must
agree: identical keypoint rows (coordinates, octave, scale, orientation), descriptors within
1e-4 relative L2.  Known answer from the survey: 5 639 keypoints."""
import os
import subprocess

import numpy as np
import pytest

from conftest import REPO

pytestmark = pytest.mark.gpu
REF = REPO / "oracle" / "_ref"
LIB = REPO / "sift3d_b200" / "lib"


def run_cli(binary, out_dir, tag, lib_path):
    env = dict(os.environ, LD_PRELOAD=str(LIB / "libsift3d_nifti.so"),
               LD_LIBRARY_PATH=f"{lib_path}:{os.environ.get('LD_LIBRARY_PATH', '')}",
               OMP_NUM_THREADS=str(min(os.cpu_count() or 8, 32)))
    keys, desc = out_dir / f"keys_{tag}.csv", out_dir / f"desc_{tag}.csv"
    r = subprocess.run([str(REF / binary), str(REF / "data" / "1.nii.gz"), "--keys", str(keys),
                        "--desc", str(desc)], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return np.loadtxt(keys, delimiter=",", ndmin=2), np.loadtxt(desc, delimiter=",", ndmin=2)


def test_stock_kpsift3d_on_example_volume(built, tmp_path):
    if not (REF / "kpSift3D_stock").exists() or not (REF / "data" / "1.nii.gz").exists():
        pytest.skip("oracle/_ref CLI binaries not built (needs /root/reference at build time)")
    ldd = subprocess.run(["ldd", str(REF / "kpSift3D_stock")], capture_output=True, text=True,
                         env=dict(os.environ, LD_LIBRARY_PATH=str(LIB))).stdout
    assert str(LIB / "libsift3D.so") in ldd, ldd           # the stock binary loads OUR library
    k_gpu, d_gpu = run_cli("kpSift3D_stock", tmp_path, "gpu", LIB)
    k_cpu, d_cpu = run_cli("kpSift3D_cpu", tmp_path, "cpu", REF)
    assert len(k_cpu) == 5639                               # SURVEY.md 8d, C1 ground truth
    assert k_gpu.shape == k_cpu.shape and d_gpu.shape == d_cpu.shape == (5639, 771)
    # x, y, z, octave, scale: identical text; R (9 values, %f): equal to the printed precision
    assert np.array_equal(k_gpu[:, :5], k_cpu[:, :5])
    assert np.abs(k_gpu[:, 5:] - k_cpu[:, 5:]).max() <= 2e-6
    assert np.array_equal(d_gpu[:, :3], d_cpu[:, :3])
    num = np.linalg.norm(d_gpu[:, 3:] - d_cpu[:, 3:], axis=1)
    den = np.linalg.norm(d_cpu[:, 3:], axis=1)
    # %f keeps 6 decimals of values <= 0.2: quantisation alone is ~3e-5 relative
    assert (num / den).max() <= 1e-4 + 5e-5
