"""Small end-to-end pass over every kernel for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvio_b200 import synth, klt
from pvio_b200.bundle_adjustor import BundleAdjustor
ba = BundleAdjustor(max_windows=80, max_frames=8, max_landmarks=96, max_obs=800)
w, st, _ = synth.make_cfg2(N=6, M=70, staggered=True)
print('gn_step cfg2', np.linalg.norm(ba.gn_step(w, st)['dx']))
for i in range(80):
    ba.batch_set(i, w, st)
ba.batch_upload(80); ba.batch_gn_step(80, 1e-8, apply=True); dx, c = ba.batch_download(80, 15 * 6 + 70)
print('batch 80 (tpl kernels + lean solve)', np.linalg.norm(dx[79]), c[0])
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
ba.close()
