"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvio_b200 import klt
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
ba = BundleAdjustor(max_windows=80, max_frames=8, max_landmarks=96, max_obs=800)
w, st, _ = synth.make_cfg2(N=6, M=70, staggered=True)
print('gn_step cfg2', np.linalg.norm(ba.gn_step(w, st)['dx']))
for i in range(80):
    ba.batch_set(i, w, st)
ba.batch_upload(80); ba.batch_gn_step(80, 1e-8, apply=True); dx, c = ba.batch_download(80, 15 * 6 + 70)
print('batch 80 (one CTA per window + lean solve)', np.linalg.norm(dx[79]), c[0])
w3, st3, _ = synth.make_cfg3(N=6, M=60)
print('gn_step cfg3', np.linalg.norm(ba.gn_step(w3, st3)['dx']))
w4, st4, _ = synth.make_cfg4(N=6, M=40, tracks_per_plane=20)
print('gn_step cfg4', np.linalg.norm(ba.gn_step(w4, st4)['dx']))
s, summ = ba.solve(w3, st3, max_iterations=3)
print('solve', summ['iterations'], summ['final_cost'])
S, e = ba.marginalize_frame(w3, st3, 0)
print('marg', np.linalg.norm(S), np.linalg.norm(e))
print('reproj err', ba.compute_reprojection_error(w, st))
prev, nxt, pts, _ = synth.make_klt_pair(size=(160, 120), n_points=20)
p, s_, e_ = klt.track_keypoints(ba, prev, nxt, pts)
print('klt', int(s_.sum()))
# PnP, IMU pre-integration, triangulation, Lie self-test, batched device-side solve
from pvio_b200 import pnp, imu, triangulate as tri
d = synth.make_pnp()
_, ps = pnp.visual_inertial_pnp(ba, d['frame'], d['last'], d['imu'], d['pts'], d['zs'], d['cam_q'], d['cam_p'], d['imu_q'], d['imu_p'], d['W'], True)
print('pnp', ps['iterations'])
_, _, truth = synth.make_cfg3(N=4, M=30)
print('imu', np.linalg.norm(imu.preintegrate(ba, truth.imu_factors, truth.imu_noise)[:, :11]))
P = np.array([np.c_[np.eye(3), -np.array([0.3 * i, 0, 0])] for i in range(4)])      # identity rotation, camera at (0.3 i, 0, 0)
X = np.array([0.2, 0.1, 5.0])
zz = np.array([(P[f] @ np.r_[X, 1])[:2] / (P[f] @ np.r_[X, 1])[2] for f in range(4)])
print('tri', tri.triangulate(ba, P, [0, 4], [0, 1, 2, 3], zz)[0])
print('lie', np.linalg.norm(ba.selftest_lie(np.array([[0.1, 0.2, 0.3], [1e-9, 0, 0], [3.14159, 0, 0]]))))
ba.batch_set(0, w, st); ba.batch_replicate(80); ba.batch_upload(80); ba.batch_solve(80, max_iterations=3)
fr, rh, sm = ba.batch_download_state(80, w.N, w.M)
print('batch solve', sm[0]['iterations'], sm[79]['final_cost'])
ba.close()
