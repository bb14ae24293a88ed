#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/r2_run2_pytest.txt
python - <<'PY' 2>&1 | tee gpurun_out/r2_run2_err.txt
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sift3d_b200 import capi
from sift3d_b200.volumes import blob_volume
z = np.load('tests/golden/blob256_large.npz')
vol = blob_volume(256, 1234)
lib = capi.load_b200()
with capi.Sift3D(lib) as s:
    kp = s.detect_keypoints(vol); d = s.extract_descriptors()
    e = int(z['desc_every'])
    rel = np.linalg.norm(d['hists'][::e] - z['desc'], axis=1) / np.linalg.norm(z['desc'], axis=1)
    print('k3 vs reference 256^3: max', rel.max(), 'mean', rel.mean(), 'p99', np.quantile(rel, 0.99))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_descriptor3 -s 1 -c 1 -o gpurun_out/r2_desc3 \
    python tools/run_desc.py 192 > gpurun_out/r2_ncu_desc3.log 2>&1
tail -3 gpurun_out/r2_ncu_desc3.log
