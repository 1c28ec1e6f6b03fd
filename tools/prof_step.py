"""Profiling target: W cfg2 windows (default 592 = 4 per SM), a few warm-up GN steps, then 2 steps.
    ncu --set full --clock-control none --import-source on -k regex:'lin_obs|schur_kernel|update_obs|solve_kernel' \
        -s <4 x warm-up launches> -c 4 -o gpurun_out/prof python tools/prof_step.py 592"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

W = int(sys.argv[1]) if len(sys.argv) > 1 else 592
hetero = len(sys.argv) > 2 and sys.argv[2] == "hetero"
w, st, _ = synth.make_cfg2()
ba = BundleAdjustor(max_windows=W, max_frames=10, max_landmarks=512, max_obs=4608)
if hetero:
    kinds = [synth.make_cfg2(N=10, M=500, seed=41)[:2], synth.make_cfg2(N=10, M=400, staggered=True, seed=42)[:2],
             synth.make_cfg2(N=8, M=300, seed=43)[:2], synth.make_cfg2(N=9, M=350, staggered=True, seed=44)[:2]]
    for i in range(W):
        ba.batch_set(i, *kinds[i % len(kinds)])
else:
    ba.batch_set(0, w, st)
    ba.batch_replicate(W)
ba.batch_upload(W)
for _ in range(3):
    ba.batch_gn_step(W, 1e-8, apply=False)
ba.sync()
for _ in range(2):
    ba.batch_gn_step(W, 1e-8, apply=False)
ba.sync()
print("done", ba.kernel_launches)
ba.close()
