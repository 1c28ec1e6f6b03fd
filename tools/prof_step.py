"""Profiling driver: a few batched GN steps at a given batch size (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pvio_b200 import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
n = int(sys.argv[1]); steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=n, max_frames=10, max_landmarks=512, max_obs=4608)
ba.batch_set(0, w, st); ba.batch_replicate(n); ba.batch_upload(n)
for _ in range(steps):
    ba.batch_gn_step(n, 1e-8, apply=False)
ba.sync()
