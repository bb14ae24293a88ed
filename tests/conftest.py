import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "oracle"))  # oracle_api: test infrastructure, outside the package
GOLDEN = REPO / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAVE_GPU = _have_gpu()


def pytest_collection_modifyitems(config, items):
    if HAVE_GPU:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build (or reuse) the in-tree libraries + the oracle restatement."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def oracle_cls(built):
    from oracle_api import Oracle
    return Oracle


@pytest.fixture(scope="session")
def ref_lib(built):
    import oracle_api
    if not oracle_api.REF_LIB.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    return oracle_api.load_reference()


@pytest.fixture(scope="session")
def b200_lib(built):
    from sift3d_b200 import capi
    return capi.load_b200()


def load_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    d = {k: z[k] for k in z.files}
    if "input" not in d:
        from sift3d_b200.volumes import blob_volume
        assert name == "blob96"
        d["input"] = blob_volume(96, seed=1234)
    p = d["params"]
    d["kwargs"] = dict(peak_thresh=float(p[0]), corner_thresh=float(p[1]), sigma_n=float(p[2]),
                       sigma0=float(p[3]), num_kp_levels=int(p[4]))
    return d


GOLDEN_CASES = ["blob48", "aniso40", "units2_params", "real_crop64", "blob96"]
