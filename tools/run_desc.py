"""Detect + describe once at n^3 (ncu target for the keypoint kernels)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
vol = blob_volume(n, seed=1234)
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    for _ in range(2):
        kp = s.detect_keypoints(vol)
        d = s.extract_descriptors()
print("keypoints", len(kp))
