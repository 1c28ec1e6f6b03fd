"""Scratch timing of the batched GN step (not the bench contract; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvio_b200 import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=W, max_frames=10, max_landmarks=512, max_obs=4608)
ba.batch_set(0, w, st)
t = time.time(); ba.batch_replicate(W); print('replicate s', time.time() - t)
ba.timer_start(); ba.batch_upload(W); print('upload ms', ba.timer_stop())
for n in ([1, 8, 64, 512, W] if W >= 512 else [1, W]):
    for _ in range(3):
        ba.batch_gn_step(n, 1e-8, apply=False)
    ba.sync()
    ba.timer_start()
    K = 10
    for _ in range(K):
        ba.batch_gn_step(n, 1e-8, apply=False)
    ms = ba.timer_stop() / K
    lin = ba.last_kernel_ms()
    bytes_alg = 16 * w.K + 16 * w.M + 64 * w.N + 4 * (48 * 48 + 48) + 4 * w.M
    print(f"n={n}: {ms*1e3:.1f} us/step, {n/ms*1e3:.0f} window-iters/s, lin kernel {lin*1e3:.1f} us, "
          f"alg GB/s (whole step) {bytes_alg*n/ms/1e6:.1f}, lin-only {bytes_alg*n/max(lin,1e-9)/1e6:.1f}")
t = time.time()
dx, costs = ba.batch_gn_step_host(W, 15 * w.N + w.M)
print('e2e host step s', time.time() - t, costs[0])
