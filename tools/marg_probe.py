"""Profiling target: marginalize_frame + single-window solves on the cfg3 window (for an ncu launch list)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
w, s, _ = synth.make_cfg3()
b = BundleAdjustor(max_windows=1, max_frames=9, max_landmarks=320, max_obs=2560)
for _ in range(3):
    b.marginalize_frame(w, s, 0)
t = time.time()
for _ in range(5):
    b.marginalize_frame(w, s, 0)
print("marg ms", (time.time() - t) * 200)
w2, s2, _ = synth.make_cfg2()
for ww, ss in ((w, s), (w2, s2)):
    bb = BundleAdjustor(max_windows=1, max_frames=10, max_landmarks=512, max_obs=4608)
    bb.solve(ww, ss, max_iterations=10)
    t = time.time()
    for _ in range(5):
        _, sm = bb.solve(ww, ss, max_iterations=10)
    print("solve ms", (time.time() - t) * 200, sm['iterations'], sm['solve_seconds'] * 1e3)
    bb.close()
b.close()
