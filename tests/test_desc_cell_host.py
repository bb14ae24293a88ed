"""CPU check of the window enumeration of the cell-owner descriptor kernel (k_descriptor3).

The geometry code of the kernel (sift3d_b200/csrc/desc_cell_geom.cuh: cell bounding rows, row
scan, exact membership test) is `__host__ __device__`; tools/desc_cell_host_check.cu drives it
with host loops for 600 random keypoints -- rotations incl. exactly and nearly axis-aligned and
non-orthonormal ones, sub-voxel centres, anisotropic units, windows clipped by the volume
border -- and compares the set of voxels reached through  cells x rows x scanned intervals x
d3_member  with a brute-force sweep of the reference's tests (sphere, 0 <= vb < 4;
sift.c:1866-1881): identical sets, every voxel exactly once.  Needs nvcc (host compile), no GPU.
"""
import shutil
import subprocess

import pytest

from conftest import REPO


def test_cell_enumeration_visits_the_reference_voxels_exactly_once(tmp_path):
    nvcc = shutil.which("nvcc")
    if nvcc is None:
        pytest.skip("nvcc not available")
    exe = tmp_path / "desc_cell_host_check"
    r = subprocess.run([nvcc, "-O2", "-o", str(exe), str(REPO / "tools" / "desc_cell_host_check.cu")],
                       capture_output=True, text=True, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe), "600"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "all cases identical" in r.stdout
    # the scan is tight: under 5 % of the scanned voxels are rejected by the exact test
    scanned = float(r.stdout.split("scanned (x")[1].split(")")[0])
    assert scanned < 1.06, scanned
