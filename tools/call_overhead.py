"""Wall time of BundleAdjustor.solve per call against the device time of the same call (percentiles), cfg3 / cfg4 in both orders."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

def run(name, maker):
    w, s, _ = maker()
    b = BundleAdjustor(max_windows=1, max_frames=w.N, max_landmarks=512, max_obs=4608)
    for _ in range(3):
        b.solve(w, s, max_iterations=10)
    wall, dev = [], []
    for _ in range(60):
        t = time.perf_counter()
        _, sm = b.solve(w, s, max_iterations=10)
        wall.append((time.perf_counter() - t) * 1e3); dev.append(sm["solve_seconds"] * 1e3)
    wall, dev = np.array(wall), np.array(dev)
    print(name, "wall ms p10/p50/p90/max", np.round(np.percentile(wall, [10, 50, 90, 100]), 3), "device p50/max",
          np.round(np.percentile(dev, [50, 100]), 3), "first five wall", np.round(wall[:5], 2))
    b.close()

for name, maker in (("cfg3", synth.make_cfg3), ("cfg4", synth.make_cfg4), ("cfg3", synth.make_cfg3), ("cfg2", synth.make_cfg2)):
    run(name, maker)
