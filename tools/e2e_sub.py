import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
W = 4096
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=W, max_frames=10, max_landmarks=512, max_obs=4608)
ba.batch_set(0, w, st); ba.batch_replicate(W)
stride = 15 * w.N + w.M
dx = np.zeros((W, stride)); costs = np.zeros((W, 2))
for _ in range(3): ba.batch_gn_step_host(W, stride, dx=dx, costs=costs)
t = time.perf_counter()
K = 10
for _ in range(K): ba.batch_gn_step_host(W, stride, dx=dx, costs=costs)
ms = (time.perf_counter() - t) / K * 1e3
print(f"sub={os.environ.get('PVIO_B200_SUB','512')}: e2e {ms:.3f} ms per step, {W/ms*1e3:.0f} window-iters/s")
