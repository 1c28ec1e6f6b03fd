"""Latency probe of the inertial single-window path (cfg3: 9 frames, 8 IMU factors, 120-dim prior)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
ba = BundleAdjustor(max_windows=1, max_frames=12, max_landmarks=640, max_obs=6000)
w, st, _ = synth.make_cfg3()
for _ in range(2):
    o, sm = ba.solve(w, st, max_iterations=10)
t = time.perf_counter()
for _ in range(5):
    o, sm = ba.solve(w, st, max_iterations=10)
print("cfg3 solve ms", (time.perf_counter() - t) / 5 * 1e3, sm["iterations"], sm["solve_seconds"] * 1e3)
if len(sys.argv) > 1 and sys.argv[1] == "cfg4":
    w4, st4, _ = synth.make_cfg4()
    for _ in range(3):
        o, sm = ba.solve(w4, st4, max_iterations=10)
    print("cfg4 solve", sm["iterations"], sm["solve_seconds"] * 1e3)
