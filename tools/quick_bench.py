"""Scratch timing of the batched GN step (not the bench contract; see bench.py)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
short = len(sys.argv) > 2 and sys.argv[2] == "short"
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=W, max_frames=10, max_landmarks=512, max_obs=4608)
ba.batch_set(0, w, st)
t = time.time(); ba.batch_replicate(W); print('replicate s', time.time() - t)
ba.timer_start(); ba.batch_upload(W); print('upload ms', ba.timer_stop())
for n in ([W] if short else [1, 8, 64, 512, W] if W >= 512 else [1, W]):
    for _ in range(3):
        ba.batch_gn_step(n, 1e-8, apply=False)
    ba.sync()
    ba.timer_start()
    K = 10
    for _ in range(K):
        ba.batch_gn_step(n, 1e-8, apply=False)
    ms = ba.timer_stop() / K
    lin = ba.last_kernel_ms(1)
    lin_a, lin_s = ba.last_kernel_ms(2), ba.last_kernel_ms(3)
    ba.last_kernel_ms(-1)
    bytes_alg = 16 * w.K + 16 * w.M + 64 * w.N + 4 * (48 * 48 + 48) + 4 * w.M
    print(f"n={n}: {ms*1e3:.1f} us/step, {n/ms*1e3:.0f} window-iters/s, stage {lin*1e3:.1f} us (linearise {lin_a*1e3:.1f} + schur {lin_s*1e3:.1f}), "
          f"alg GB/s (whole step) {bytes_alg*n/ms/1e6:.1f}, lin-only {bytes_alg*n/max(lin,1e-9)/1e6:.1f}")
if short:
    sys.exit(0)
t = time.time()
dx, costs = ba.batch_gn_step_host(W, 15 * w.N + w.M)
print('e2e host step s', time.time() - t, costs[0])
ba.batch_upload(W); ba.sync()
for n in ([1, W] if W > 1 else [1]):
    ba.batch_upload(n); ba.batch_solve(n, max_iterations=10); ba.sync()
    ba.batch_upload(n); ba.sync()
    ba.timer_start(); ba.batch_solve(n, max_iterations=10); ms = ba.timer_stop()
    fr, rh, sm = ba.batch_download_state(n, w.N, w.M)
    its = sum(x['iterations'] for x in sm)
    print(f"batch_solve n={n}: {ms:.3f} ms, {its} window-iterations -> {its/ms*1e3:.0f} window-iters/s, final cost {sm[0]['final_cost']:.4f} it {sm[0]['iterations']}")
fr = np.zeros((W, w.N * 16)); rh = np.zeros((W, w.M))
ba.batch_solve_host(W, w.N, w.M, frames=fr, rho=rh)
t = time.time()
_, _, sm = ba.batch_solve_host(W, w.N, w.M, frames=fr, rho=rh)
dt = time.time() - t
its = sum(x.iterations for x in sm)
print(f"e2e batch_solve_host: {dt*1e3:.2f} ms, {its} window-iterations -> {its/dt:.0f} window-iters/s")
# e2e scaling: < 1024 windows take the single-shot path (upload, solve, download back to back), >= 1024 the pipelined one
for n in (1000, 2048, W):
    fr = np.zeros((n, w.N * 16)); rh = np.zeros((n, w.M))
    ba.batch_solve_host(n, w.N, w.M, frames=fr, rho=rh)
    t = time.time()
    for _ in range(3):
        _, _, sm = ba.batch_solve_host(n, w.N, w.M, frames=fr, rho=rh)
    dt = (time.time() - t) / 3
    print(f"e2e batch_solve_host n={n}: {dt*1e3:.2f} ms -> {sum(x.iterations for x in sm)/dt:.0f} window-iters/s")
    ba.batch_upload(n); ba.sync()
    t = time.time(); ba.batch_upload(n); ba.sync(); up = time.time() - t
    t = time.time(); ba.batch_solve(n, max_iterations=10); ba.sync(); so = time.time() - t
    t = time.time(); ba.batch_download_state(n, w.N, w.M); dn = time.time() - t
    print(f"   parts: upload {up*1e3:.2f} ms, solve {so*1e3:.2f} ms, download+scatter {dn*1e3:.2f} ms")
