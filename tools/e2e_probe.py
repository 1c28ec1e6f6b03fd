import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
W=4096
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=W, max_frames=10, max_landmarks=512, max_obs=4608)
ba.batch_set(0, w, st); ba.batch_replicate(W)
stride=15*w.N+w.M
dx=np.zeros((W,stride)); costs=np.zeros((W,2))
for _ in range(2): ba.batch_gn_step_host(W, stride, 1e-8, dx, costs)
for k in range(3):
    t=time.perf_counter(); ba.batch_gn_step_host(W, stride, 1e-8, dx, costs); print('host step ms', (time.perf_counter()-t)*1e3)
ba.timer_start(); ba.batch_upload(W); print('upload only ms', ba.timer_stop())
t=time.perf_counter(); ba.batch_download(W, stride); print('download+scatter ms', (time.perf_counter()-t)*1e3)
ba.timer_start(); ba.batch_gn_step(W,1e-8,False); print('step ms', ba.timer_stop())
