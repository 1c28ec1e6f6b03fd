"""Profiling target: the device-side trust-region loop over W cfg2 windows (pvio_b200_batch_solve), once warm, once measured.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/solve_launches.csv python tools/prof_solve.py 4096"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=W, max_frames=10, max_landmarks=512, max_obs=4608)
ba.batch_set(0, w, st)
ba.batch_replicate(W)
for _ in range(2):
    ba.batch_upload(W)
    ba.batch_solve(W)
    ba.sync()
print("done", ba.kernel_launches)
ba.close()
