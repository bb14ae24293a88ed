#!/usr/bin/env python
"""First-contact GPU check: detailed parity report + rough stage timings (not a bench)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from sift3d_b200 import capi  # noqa: E402
from sift3d_b200.engine_api import Engine  # noqa: E402
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent / 'oracle'))
from oracle_api import Oracle  # noqa: E402
from sift3d_b200.volumes import blob_volume, noise_volume  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    big = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    lib = capi.load_b200()
    orc = Oracle()
    eng = Engine()
    rng = np.random.default_rng(0)
    # 1. kernel-level blur parity
    vol = rng.random((37, 45, 70), dtype=np.float32)
    for units in [(1, 1, 1), (2, 2, 2), (1.0, 2.0, 0.7)]:
        for sg in (0.5387, 1.2263, 2.4525):
            taps = orc.gauss_taps(sg)
            want = orc.blur(vol, taps, units)
            got = eng.blur(vol, taps, units, mode=1)
            print(f"blur generic units={units} w={len(taps)} exact={np.array_equal(got.view(np.uint32), want.view(np.uint32))} maxabs={np.abs(got-want).max():.3g}")
    # 2. full pipeline parity vs oracle
    for name, v in (("blob", blob_volume((n, n + 8, n - 8), seed=5)), ("noise", noise_volume((40, 48, 56), 1))):
        okp = orc.detect(v)
        with capi.Sift3D(lib) as s:
            t = time.time()
            kp = s.detect_keypoints(v)
            dt = time.time() - t
            bad = []
            for o in range(s.num_octaves()):
                for lv in range(-1, 5):
                    a = s.level_data("gpyr", o, lv)
                    b, _ = orc.level("gpyr", o, lv)
                    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                        bad.append(("g", o, lv, float(np.abs(a - b).max())))
                for lv in range(-1, 4):
                    a = s.level_data("dog", o, lv)
                    b, _ = orc.level("dog", o, lv)
                    if not np.array_equal(a.view(np.uint32), b.view(np.uint32)):
                        bad.append(("d", o, lv, float(np.abs(a - b).max())))
            ncand = lib.lib.sift3d_b200_num_candidates(C.byref(s.s))
            print(f"[{name}] detect {dt*1e3:.1f} ms octaves={s.num_octaves()} level mismatches={bad[:6]} cand gpu/orc={ncand}/{len(orc.candidates())} kp gpu/orc={len(kp)}/{len(okp)}")
            same = len(kp) == len(okp) and all(np.array_equal(kp[f], okp[f]) for f in ("xd", "yd", "zd", "o", "s", "sd"))
            print(f"[{name}] keypoints identical={same}", "max|dR|=%.3g" % np.abs(kp["R"] - okp["R"]).max() if same else "")
            if not same:
                sg = set(zip(kp["o"], kp["s"], kp["zd"], kp["yd"], kp["xd"]))
                so = set(zip(okp["o"], okp["s"], okp["zd"], okp["yd"], okp["xd"]))
                print("   only gpu:", sorted(sg - so)[:5], " only oracle:", sorted(so - sg)[:5])
            if len(kp):
                t = time.time()
                d = s.extract_descriptors()
                dt = time.time() - t
                od, oc = orc.describe(okp) if same else (None, None)
                if same:
                    rel = np.linalg.norm(d["hists"] - od, axis=1) / np.linalg.norm(od, axis=1)
                    print(f"[{name}] descriptors {dt*1e3:.1f} ms rel-L2 max={rel.max():.3g} mean={rel.mean():.3g} coords eq={np.array_equal(np.stack([d['xd'],d['yd'],d['zd'],d['sd']],1), oc)}")
            sub = np.ascontiguousarray(v[:20, :18, :16])
            dd = s.extract_dense_descriptors(sub)
            do = orc.dense(sub)
            print(f"[{name}] dense max rel err={np.abs(dd-do).max()/np.abs(do).max():.3g}")
    # 3. timings at a larger size (wall clock incl. PCIe; rough)
    v = blob_volume(big, seed=1234)
    with capi.Sift3D(lib) as s:
        for rep in range(2):
            t0 = time.time()
            kp = s.detect_keypoints(v)
            t1 = time.time()
            d = s.extract_descriptors()
            t2 = time.time()
            ncand = lib.lib.sift3d_b200_num_candidates(C.byref(s.s))
            print(f"{big}^3 rep{rep}: detect {1e3*(t1-t0):.1f} ms, describe {1e3*(t2-t1):.1f} ms, cand={ncand} kp={len(kp)} -> {v.size/(t2-t0)/1e6:.1f} Mvox/s")
    eng.close()


if __name__ == "__main__":
    main()
