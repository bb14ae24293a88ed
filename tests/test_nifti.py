"""NIfTI-1 reader/writer (read_nii / write_nii of imutil/nifti.c:51-221) on
zlib alone.  nifticlib is absent here, so the checker is an independent parse of the NIfTI-1
layout in numpy (gzip + the 348-byte header) with the reference's scaling rule
`(float)((double)v * slope + inter)` (nifti.c:105-111); the real fixture is the reference's own
example volume when /root/reference is present."""
import ctypes as C
import gzip
import struct
from pathlib import Path

import numpy as np
import pytest

from conftest import REPO

NII_LIB = REPO / "sift3d_b200" / "lib" / "libsift3d_nifti.so"
DTYPES = {2: "u1", 4: "i2", 8: "i4", 16: "f4", 64: "f8", 256: "i1", 512: "u2", 768: "u4",
          1024: "i8", 1280: "u8"}


@pytest.fixture(scope="module")
def nii(built):
    from sift3d_b200 import capi
    L = C.CDLL(str(NII_LIB))
    L.read_nii.argtypes = [C.c_char_p, C.POINTER(capi.Image)]
    L.write_nii.argtypes = [C.c_char_p, C.POINTER(capi.Image)]
    return L


def py_parse(path):
    """Independent NIfTI-1 parse -> ([z][y][x][c] float32, units)."""
    raw = Path(path).read_bytes()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    end = "<" if struct.unpack("<i", raw[:4])[0] == 348 else ">"
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype = struct.unpack(end + "h", raw[70:72])[0]
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset, slope, inter = struct.unpack(end + "3f", raw[108:120])
    rank = max([i for i in range(1, dim[0] + 1) if dim[i] > 1] or [0])
    nx, ny, nz, nt = [dim[i] if i <= dim[0] else 1 for i in range(1, 5)]
    nc = nt if rank == 4 else 1
    n = nx * ny * nz * nc
    a = np.frombuffer(raw, np.dtype(end + DTYPES[datatype]), n, int(vox_offset))
    slope = 1.0 if slope == 0 else slope
    v = (a.astype(np.float64) * np.float64(np.float32(slope)) + np.float64(np.float32(inter)))
    v = v.astype(np.float32).reshape(nc, nz, ny, nx).transpose(1, 2, 3, 0)
    units = tuple(1.0 if (p == 0 or not np.isfinite(p)) else float(np.float32(p))
                  for p in pixdim[1:4])
    return np.ascontiguousarray(v), units


def read(nii, path):
    from sift3d_b200 import capi
    im = capi.empty_image()
    assert nii.read_nii(str(path).encode(), C.byref(im)) == 0
    n = im.nx * im.ny * im.nz * im.nc
    arr = np.ctypeslib.as_array(im.data, shape=(n,)).reshape(im.nz, im.ny, im.nx, im.nc).copy()
    assert (im.xs, im.ys, im.zs) == (im.nc, im.nc * im.nx, im.nc * im.nx * im.ny)
    C.CDLL(None).free(C.cast(im.data, C.c_void_p))
    return arr, (im.ux, im.uy, im.uz)


def make_file(path, arr_czyx, datatype, pixdim=(1.0, 1.0, 1.0), slope=1.0, inter=0.0, end="<",
              ext_bytes=0):
    nc, nz, ny, nx = arr_czyx.shape
    h = bytearray(348)
    struct.pack_into(end + "i", h, 0, 348)
    dim = [4 if nc > 1 else 3, nx, ny, nz, nc if nc > 1 else 1, 1, 1, 1]
    struct.pack_into(end + "8h", h, 40, *dim)
    struct.pack_into(end + "hh", h, 70, datatype, 8 * np.dtype(DTYPES[datatype]).itemsize)
    struct.pack_into(end + "8f", h, 76, 1.0, *pixdim, 1.0, 1.0, 1.0, 1.0)
    struct.pack_into(end + "3f", h, 108, 352.0 + ext_bytes, slope, inter)
    h[344:348] = b"n+1\0"
    body = bytes(h) + b"\0" * (4 + ext_bytes) + \
        arr_czyx.astype(np.dtype(end + DTYPES[datatype])).tobytes()
    if str(path).endswith(".gz"):
        body = gzip.compress(body, 1)
    Path(path).write_bytes(body)


@pytest.mark.parametrize("datatype", sorted(DTYPES))
@pytest.mark.parametrize("end", ["<", ">"])
def test_read_every_datatype_and_byte_order(nii, tmp_path, datatype, end):
    rng = np.random.default_rng(datatype)
    kind = np.dtype(DTYPES[datatype])
    shape = (1, 5, 6, 7)
    a = (rng.random(shape) * 100 - (0 if kind.kind == "u" else 50)).astype(kind)
    p = tmp_path / "t.nii.gz"
    make_file(p, a, datatype, pixdim=(0.5, 0.75, 2.0), slope=0.2922, inter=-1.5, end=end)
    got, units = read(nii, p)
    want, wunits = py_parse(p)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert units == wunits == (0.5, 0.75, 2.0)


def test_read_channels_extensions_zero_slope_plain_file(nii, tmp_path):
    rng = np.random.default_rng(1)
    a = rng.random((3, 4, 5, 6)).astype(np.float32)            # 3 channels, planar in the file
    p = tmp_path / "c.nii"                                      # not gzipped
    make_file(p, a, 16, pixdim=(1.0, 0.0, 3.0), slope=0.0, inter=0.0, ext_bytes=32)
    got, units = read(nii, p)
    assert got.shape == (4, 5, 6, 3) and units == (1.0, 1.0, 3.0)   # pixdim 0 -> 1, slope 0 -> 1
    assert np.array_equal(got, a.transpose(1, 2, 3, 0))
    from sift3d_b200 import capi
    im = capi.empty_image()
    assert nii.read_nii(str(tmp_path / "missing.nii").encode(), C.byref(im)) == -1
    (tmp_path / "bad.nii").write_bytes(b"\0" * 400)
    assert nii.read_nii(str(tmp_path / "bad.nii").encode(), C.byref(im)) == -1


@pytest.mark.parametrize("name,nc", [("w.nii", 1), ("w.nii.gz", 1), ("w4.nii.gz", 12)])
def test_write_then_read_round_trip(nii, tmp_path, name, nc):
    from sift3d_b200 import capi
    rng = np.random.default_rng(2)
    vol = rng.standard_normal((7, 8, 9) + ((nc,) if nc > 1 else ())).astype(np.float32)
    im = capi.make_image(vol, (0.8, 1.1, 2.5), nc)
    p = tmp_path / name
    assert nii.write_nii(str(p).encode(), C.byref(im)) == 0
    want, wunits = py_parse(p)                     # an independent reader understands the file
    got, units = read(nii, p)
    ref = vol.reshape(7, 8, 9, nc)
    assert np.array_equal(want, ref) and np.array_equal(got, ref)
    f32 = tuple(float(np.float32(u)) for u in (0.8, 1.1, 2.5))
    assert units == wunits == f32
    raw = p.read_bytes()
    hdr = gzip.decompress(raw)[:352] if name.endswith(".gz") else raw[:352]
    assert struct.unpack("<i", hdr[:4])[0] == 348 and hdr[344:348] == b"n+1\0"
    assert struct.unpack("<8h", hdr[40:56])[:5] == ((4 if nc > 1 else 3), 9, 8, 7, nc)


def test_reads_the_reference_example_volume(nii):
    p = Path("/root/reference/examples/data/1.nii.gz")
    if not p.exists():
        pytest.skip("reference tree not present")
    got, units = read(nii, p)
    want, wunits = py_parse(p)
    assert got.shape == (181, 217, 181, 1) and units == wunits == (1.0, 1.0, 1.0)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert got.max() > 0


def test_read_nii_rejects_hostile_headers(nii, tmp_path):
    """The header comes from an untrusted file: a non-finite or out-of-range vox_offset and a
    voxel count beyond the reference's own `int` size arithmetic (imutil.c:1533) are refused
    instead of driving the skip loop / the allocation."""
    from sift3d_b200 import capi
    vol = np.arange(4 * 5 * 6, dtype=np.float32).reshape(4, 5, 6)
    good = tmp_path / "good.nii"
    im = capi.make_image(vol)
    assert nii.write_nii(str(good).encode(), C.byref(im)) == 0
    raw = bytearray(good.read_bytes())

    def attempt(patch):
        b = bytearray(raw)
        patch(b)
        p = tmp_path / "bad.nii"
        p.write_bytes(bytes(b))
        out = capi.empty_image()
        rc = nii.read_nii(str(p).encode(), C.byref(out))
        if rc == 0 and out.data:
            C.CDLL(None).free(C.cast(out.data, C.c_void_p))
        return rc

    assert attempt(lambda b: None) == 0
    assert attempt(lambda b: b.__setitem__(slice(108, 112), struct.pack("<f", float("inf")))) != 0
    assert attempt(lambda b: b.__setitem__(slice(108, 112), struct.pack("<f", float("nan")))) != 0
    assert attempt(lambda b: b.__setitem__(slice(108, 112), struct.pack("<f", 350.0))) != 0
    assert attempt(lambda b: b.__setitem__(slice(108, 112), struct.pack("<f", 3e9))) != 0
    assert attempt(lambda b: b.__setitem__(slice(42, 48), struct.pack("<3h", 32767, 32767, 32767))) != 0
