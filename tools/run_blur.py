"""Run the fused blur a few times at n^3 for one filter width (ncu target)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from sift3d_b200.engine_api import Engine
from bench import gauss_taps, pyramid_filters
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
which = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
src = torch.rand((n, n, n), device="cuda")
dst = torch.empty_like(src)
e = Engine(0)
taps = gauss_taps(pyramid_filters()[which])
for _ in range(reps):
    e.blur_device(src.data_ptr(), dst.data_ptr(), n, n, n, taps, mode=mode)
e.sync()
print("done", len(taps))
