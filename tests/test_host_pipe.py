"""CPU check of the host-thread pipeline that moves volumes between PAGEABLE host memory and the
device (sift3d_b200/csrc/host_pipe.h, used by engine.cu: pipe_transfer).

The pipeline is plain C++; tools/pipe_host_check.cpp drives it with a mock DMA engine (a thread
that executes the queued copies in order after random delays and completes the slot's "event",
the behaviour of cudaMemcpyAsync + cudaEventRecord / cudaEventQuery on one stream): uploads and
downloads of sizes around the chunk and ring boundaries with 2 ... 16 slots must deliver the
exact bytes, never refill a slot the other side still needs, and terminate.  Run with the team
sizes a 1-rank and an 8-rank job get, and with more threads than cores.  No GPU.
"""
import os
import shutil
import subprocess

import pytest

from conftest import REPO


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = tmp_path_factory.mktemp("pipe") / "pipe_host_check"
    r = subprocess.run([gxx, "-O2", "-std=c++17", "-pthread", "-Wall", "-Werror", "-I",
                        str(REPO / "sift3d_b200" / "csrc"), str(REPO / "tools" / "pipe_host_check.cpp"),
                        "-o", str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return out


@pytest.mark.parametrize("threads", [1, 4, 8, 24])
def test_chunk_pipeline_against_a_mock_dma_engine(exe, threads):
    env = dict(os.environ, S3D_COPY_THREADS=str(threads))
    r = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"workers {threads}, ok" in r.stdout


def test_team_size_follows_the_ranks_on_the_host(exe):
    """One team per process: min(8, cores / $LOCAL_WORLD_SIZE) threads, at least 2 (torchrun sets
    the variable; a rank on a shared host then takes the poller-free pipeline, engine.cu)."""
    cores = os.cpu_count() or 1
    for ranks in (1, 2, 8):
        env = {k: v for k, v in os.environ.items() if k != "S3D_COPY_THREADS"}
        env["LOCAL_WORLD_SIZE"] = str(ranks)
        r = subprocess.run([str(exe)], capture_output=True, text=True, env=env, timeout=600)
        want = max(2, min(8, cores // ranks))
        assert r.returncode == 0 and f"workers {want}, ok" in r.stdout, (ranks, r.stdout[-500:])
