#!/bin/bash
cd "$GRAFT_REPO_ROOT" || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fm.py -q > gpurun_out/c2_fm.log 2>&1; echo "fm rc=$?"
tail -15 gpurun_out/c2_fm.log
PVIO_B200_TUNE_LIB=$PWD/tools/_variants/libpvio_stamps.so timeout 300 python tools/solve_stamps.py 2>&1 | tee gpurun_out/c2_stamps.log
timeout 300 python - <<'P' 2>&1 | tee gpurun_out/c2_fm_timing.log
import time, numpy as np, cv2
from synthetic import synth
from pvio_b200 import klt
from pvio_b200.bundle_adjustor import BundleAdjustor
ba = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=16, max_obs=64)
for n, of in [(400, 0.2), (400, 0.5), (150, 0.1)]:
    p, q = synth.make_fm_matches(652, n, of)
    for _ in range(3): klt.find_fundamental_mask(ba, p, q)
    t = time.perf_counter()
    for _ in range(50): m = klt.find_fundamental_mask(ba, p, q, return_info=True)
    tg = (time.perf_counter() - t) / 50
    t = time.perf_counter()
    for _ in range(50): cv2.findFundamentalMat(p, q, cv2.FM_RANSAC, 1.0, 0.99)
    tc = (time.perf_counter() - t) / 50
    print(n, of, "gpu ms", tg * 1e3, "cv2 ms", tc * 1e3, m[2])
P
