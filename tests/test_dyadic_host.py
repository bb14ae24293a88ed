"""CPU check of the register-blocked dyadic FIR (sift3d_b200/csrc/blur_dyadic.cu).

The line arithmetic of those kernels is `__host__ __device__`; tools/dyadic_host_check.cu drives
the very same code with host loops (block phases run tid by tid) and compares every instantiated
(order, half width) against the oracle's apply_Sep_FIR_filter restatement, bit for bit, on odd
sizes that exercise the mirror ends, partial runs and partial x tiles.  Needs nvcc (to compile the
.cu for the host) but no GPU.
"""
import shutil
import subprocess

import pytest

from conftest import REPO


def test_dyadic_kernels_match_oracle_on_host(built, tmp_path):
    nvcc = shutil.which("nvcc")
    if nvcc is None:
        pytest.skip("nvcc not available")
    exe = tmp_path / "dyadic_host_check"
    ora = REPO / "oracle" / "_build"
    cmd = [nvcc, "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I", str(REPO / "include"),
           str(REPO / "tools" / "dyadic_host_check.cu"), str(ora / "liboracle.so"), "-o", str(exe),
           "-Xlinker", f"-rpath={ora}"]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "all bit-identical" in r.stdout
    assert r.stdout.count(" 0 / ") >= 20  # every case reports zero differing values
