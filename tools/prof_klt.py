"""Profiling target for the KLT kernels: one warm pair, then one uncached pair (upload + CLAHE + pyramids + LK) and one
pair whose frames are both in the device-side pyramid cache (LK only).
    ncu --set full --clock-control none -k regex:'klt|pyr|clahe' -o gpurun_out/klt python tools/prof_klt.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
from pvio_b200 import klt

ba = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=64, max_obs=256)
prev, nxt, pts, _ = synth.make_klt_pair()
klt.track_keypoints(ba, prev, nxt, pts, clahe_clip=6.0)
klt.track_keypoints(ba, prev, nxt, pts, clahe_clip=6.0, prev_id=11, next_id=12)
short = len(sys.argv) > 1 and sys.argv[1] == "short"       # under ncu --set full: one pair of each kind is enough
for rep in range(1 if short else 3):
    ba.timer_start()
    klt.track_keypoints(ba, None, None, pts, clahe_clip=6.0, prev_id=11, next_id=12, shape=prev.shape)
    ms = ba.timer_stop()
t = time.perf_counter()
for _ in range(1 if short else 20):
    klt.track_keypoints(ba, None, nxt, pts, clahe_clip=6.0, prev_id=12, next_id=0 + 13 + _, shape=prev.shape)
print(f"both frames cached: {ms * 1e3:.1f} us device per pair ({len(pts)} points); prev cached, next uploaded: {(time.perf_counter() - t) / (1 if short else 20) * 1e3:.3f} ms per pair e2e")
ba.close()
