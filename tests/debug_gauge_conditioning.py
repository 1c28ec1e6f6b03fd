"""Diagnostic script (not collected by pytest; run by hand on a GPU box): where the dx difference of the gauge-prior
window lives -- conditioning of the reduced system, eigen-direction of the error.  Lives under tests/ because it
uses the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor
from oracle import ba_oracle as bo
np.set_printoptions(precision=3, linewidth=220)
w, st, _ = synth.make_cfg3(prior='gauge', N=6, M=120)
ba = BundleAdjustor(max_windows=1, max_frames=8, max_landmarks=128, max_obs=1024)
out = ba.gn_step(w, st, want_system=True)
ref = bo.gn_step(w, st, mu=0.0, schur=True)
ref8 = bo.gn_step(w, st, schur=True)
P = 15 * w.N
dH = out['Hred'] - ref['Hred']; sc = np.sqrt(np.abs(np.diag(ref['Hred'])))
rel = np.abs(dH) / np.outer(sc, sc)
print('max rel H err', rel.max(), np.unravel_index(rel.argmax(), rel.shape))
for f in range(w.N):
    print('frame', f, 'diag-block rel err', rel[15*f:15*f+15, 15*f:15*f+15].max(), 'g rel', np.abs(out['gred'][15*f:15*f+15]-ref['gred'][15*f:15*f+15]).max()/np.abs(ref['gred']).max())
d = out['dx'] - ref8['dx']
for f in range(w.N):
    print('dx frame', f, out['dx'][15*f:15*f+15], '\n   ref   ', ref8['dx'][15*f:15*f+15])
# solve the GPU system on CPU
free = ref['free'][:P]
A = out['Hred'] + np.diag(ref8['reg'][:P]); 
idx = np.where(free)[0]
x = np.zeros(P); x[idx] = -np.linalg.solve(A[np.ix_(idx, idx)], out['gred'][idx])
print('GPU system solved on CPU vs GPU dx', np.linalg.norm(x - out['dx'][:P]) / np.linalg.norm(x), ' vs oracle', np.linalg.norm(x - ref8['dx'][:P]) / np.linalg.norm(x))

os.makedirs('gpurun_out', exist_ok=True); np.savez('gpurun_out/gauge_dump.npz', H=out['Hred'], g=out['gred'], dx=out['dx'], cost=out['cost'])
