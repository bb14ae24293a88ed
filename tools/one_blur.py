import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, '/root/repo')
import numpy as np
from sift3d_b200.engine_api import Engine
from bench import gauss_taps
e = Engine(0)
vol = np.random.default_rng(1).random((40, 64, 64), dtype=np.float32)
taps = gauss_taps(float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
got = e.blur(vol, taps)
print("ok", got.sum())
