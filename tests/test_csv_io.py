"""The text writer behind write_Keypoint_store / write_SIFT3D_Descriptor_store (SURVEY.md 8f N2;
sift3d_b200/host/csv_io.c): a parallel writer whose "%f" formatter works in exact integer
arithmetic.  The contract is the reference's write_Mat_rm (imutil.c:1343-1421): the same BYTES.

* the formatter against snprintf("%f") on millions of doubles (tools/csv_format_check.c);
* whole files against write_Mat_rm of the compiled reference (oracle/_ref), all three matrix
  types, .csv and .csv.gz (compared after decompression: deflate blocks may differ);
* the two store writers against the reference's, same stores in, same files out.
No GPU."""
import ctypes as C
import gzip
import shutil
import subprocess

import numpy as np
import pytest

from conftest import REPO


def test_format_f_is_printf_exact(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "csv_format_check"
    r = subprocess.run([gcc, "-O2", "-I", str(REPO / "include"), "-I", str(REPO / "sift3d_b200" / "host"),
                        str(REPO / "tools" / "csv_format_check.c"), str(REPO / "sift3d_b200" / "host" / "csv_io.c"),
                        "-lz", "-lm", "-fopenmp", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe), "1500000"], capture_output=True, text=True)
    assert r.returncode == 0 and " 0 mismatches" in r.stdout, r.stdout[-2000:]


def _mat(capi, a):
    a = np.ascontiguousarray(a)
    m = capi.Mat_rm()
    m.data = a.ctypes.data
    m.size = a.nbytes
    m.num_rows, m.num_cols = a.shape
    m.static_mem = 1
    m.type = {np.dtype(np.float64): 0, np.dtype(np.float32): 1, np.dtype(np.int32): 2}[a.dtype]
    return m, a


def _matrices():
    rng = np.random.default_rng(3)
    d = rng.standard_normal((301, 14)) * 10.0 ** rng.integers(-8, 9, (301, 14))
    d[0, :6] = [0.0, -0.0, 0.0078125, -2.5e-6, 1e15, -1e-300]
    f = (rng.random((257, 771)) * 0.0333).astype(np.float32)
    f[:, :3] = rng.integers(0, 512, (257, 3)).astype(np.float32)
    i = rng.integers(-2 ** 31, 2 ** 31 - 1, (64, 5)).astype(np.int32)
    return {"double": d, "float": f, "int": i, "one": np.array([[1.5]]), "wide": rng.random((1, 5000))}


@pytest.mark.parametrize("ext", [".csv", ".csv.gz"])
def test_files_equal_the_reference_writer(b200_lib, built, tmp_path, ext):
    import oracle_api
    from sift3d_b200 import capi
    if not oracle_api.REF_IMUTIL.exists():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    ref = C.CDLL(str(oracle_api.REF_IMUTIL))
    ours = b200_lib.lib
    for name, a in _matrices().items():
        m, keep = _mat(capi, a)
        p_ref, p_our = tmp_path / f"ref_{name}{ext}", tmp_path / "sub" / "dir" / f"our_{name}{ext}"
        assert ref.write_Mat_rm(str(p_ref).encode(), C.byref(m)) == 0
        assert ours.sift3d_b200_write_Mat_rm(str(p_our).encode(), C.byref(m)) == 0   # creates sub/dir
        rd = (lambda p: gzip.open(p, "rb").read()) if ext.endswith(".gz") else (lambda p: p.read_bytes())
        assert rd(p_our) == rd(p_ref), name


def test_store_writers_equal_the_reference(b200_lib, ref_lib, tmp_path):
    """write_SIFT3D_Descriptor_store / write_Keypoint_store (sift.c:3143-3230) of both libraries
    on the same stores."""
    from sift3d_b200 import capi
    rng = np.random.default_rng(5)
    rows = np.concatenate([rng.integers(0, 300, (500, 3)).astype(np.float32),
                           (rng.random((500, 768)) * 0.0333).astype(np.float32)], axis=1)
    out = {}
    for tag, lib in (("ours", b200_lib), ("ref", ref_lib)):
        L = lib.lib
        m, keep = _mat(capi, rows)
        d = capi.SIFT3D_Descriptor_store()
        L.init_SIFT3D_Descriptor_store(C.byref(d))
        assert L.Mat_rm_to_SIFT3D_Descriptor_store(C.byref(m), C.byref(d)) == 0
        p = tmp_path / f"desc_{tag}.csv"
        assert L.write_SIFT3D_Descriptor_store(str(p).encode(), C.byref(d)) == 0
        kp = capi.Keypoint_store()
        L.init_Keypoint_store(C.byref(kp))
        assert L.resize_Keypoint_store(C.byref(kp), 40) == 0
        for i in range(40):
            k = kp.buf[i]
            k.xd, k.yd, k.zd, k.sd, k.o, k.s = i * 1.25, i * 0.5, 3.0 * i, 1.6 * 2 ** (i % 3 / 3), i % 4, i % 3
            for j in range(9):
                k.r_data[j] = float(np.float32(np.sin(i * 9 + j)))
        pk = tmp_path / f"kp_{tag}.csv"
        assert L.write_Keypoint_store(str(pk).encode(), C.byref(kp)) == 0
        out[tag] = (p.read_bytes(), pk.read_bytes())
        L.cleanup_Keypoint_store(C.byref(kp))
        L.cleanup_SIFT3D_Descriptor_store(C.byref(d))
    assert out["ours"][0] == out["ref"][0]
    assert out["ours"][1] == out["ref"][1]
    assert out["ours"][0].count(b"\n") == 500 and out["ours"][1].count(b"\n") == 40
