"""Per-CTA timing of the fused blur + exhaustive kernel parity sweep (prints every failure)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from sift3d_b200.engine_api import Engine
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent / 'oracle'))
from oracle_api import Oracle
from bench import gauss_taps, pyramid_filters
e = Engine(0)
e.L.s3d_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
e.L.s3d_debug_read.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
orc = Oracle()
rng = np.random.default_rng(3)
if "parity" in sys.argv:
    for shape in [(64, 64, 64), (37, 45, 70), (40, 32, 64), (33, 70, 132)]:
        vol = rng.random(shape, dtype=np.float32)
        for sg in pyramid_filters():
            taps = gauss_taps(sg)
            want = orc.blur(vol, taps)
            got = e.blur(vol, taps, mode=0)
            bad = np.argwhere(got.view(np.uint32) != want.view(np.uint32))
            msg = "ok" if len(bad) == 0 else f"{len(bad)} mismatches; z range {bad[:,0].min()}..{bad[:,0].max()} y {bad[:,1].min()}..{bad[:,1].max()} x {bad[:,2].min()}..{bad[:,2].max()} max|d| {np.abs(got-want).max():.3g}"
            print(f"shape {shape} w={len(taps):2d}: {msg}")
if "timing" in sys.argv:
    n = 512
    e.L.s3d_set_option(e.h, b"blur_dbg", 1)
    src = torch.rand((n, n, n), device="cuda"); dst = torch.empty_like(src)
    for which, flags in ((0, 0), (5, 0), (5, 1), (5, 3), (5, 7)):
        e.L.s3d_set_option(e.h, b"blur_flags", flags)
        print("flags", flags)
        taps = gauss_taps(pyramid_filters()[which])
        for _ in range(2):
            e.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps)
        buf = np.zeros((148, 4), np.int64)
        e.L.s3d_debug_read(e.h, buf.ctypes.data, buf.nbytes)
        dur = buf[:, 1] - buf[:, 0]
        per = dur / np.maximum(buf[:, 3], 1)
        order = np.argsort(dur)
        print(f"w={len(taps)}: CTA cycles min {dur.min()} med {int(np.median(dur))} max {dur.max()}; cycles/step min {per.min():.0f} med {np.median(per):.0f} max {per.max():.0f}")
        print("  slowest CTAs (id, cycles, smid, steps):", [(int(i), int(dur[i]), int(buf[i,2]), int(buf[i,3])) for i in order[-8:]])
        print("  fastest CTAs:", [(int(i), int(dur[i]), int(buf[i,2]), int(buf[i,3])) for i in order[:6]])
        tx = np.array([((int(i * 65536 / 148) // 512) % 8) for i in range(148)]); ty = np.array([((int(i * 65536 / 148) // 512) // 8) for i in range(148)])
        for name, sel in (("tx=0", tx == 0), ("tx=7", tx == 7), ("ty=0", ty == 0), ("ty=15", ty == 15), ("interior", (tx > 0) & (tx < 7) & (ty > 0) & (ty < 15))):
            print(f"   {name}: median cycles/step {np.median(per[sel]):.0f}")
