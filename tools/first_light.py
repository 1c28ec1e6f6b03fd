"""First run of a new pipeline on the GPU: one small window, every intermediate against the oracle, loud output."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import ba_oracle as bo
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor


def rel(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def check(name, w, st, ba):
    ref = bo.gn_step(w, st, schur=True)
    out = ba.gn_step(w, st, mu=1e-8, want_system=True)
    P = 15 * w.N
    H0 = bo.gn_step(w, st, mu=0.0, schur=True)
    free = ref['free'][:P]
    hs = np.sqrt(np.abs(np.diag(H0['Hred'])))[free]
    A, B = out['Hred'][np.ix_(free, free)], H0['Hred'][np.ix_(free, free)]
    print(f"{name}: dx pose {rel(out['dx'][:P], ref['dx'][:P]):.2e} lm {rel(out['dx'][P:], ref['dx'][P:]):.2e} "
          f"cost {out['cost']:.6f}/{ref['cost']:.6f} new {out['new_cost']:.6f} Hred {np.max(np.abs(A - B) / np.outer(hs, hs)):.2e} "
          f"gred {rel(out['gred'][free], H0['gred'][free]) if 'gred' in H0 else -1:.2e}", flush=True)


ba = BundleAdjustor(max_windows=160, max_frames=12, max_landmarks=640, max_obs=6000)
check("cfg2 small", *synth.make_cfg2(N=5, M=40)[:2], ba)
check("cfg2 full", *synth.make_cfg2()[:2], ba)
check("cfg2b", *synth.make_cfg2(staggered=True)[:2], ba)
w, st, _ = synth.make_cfg2()
W = 160
ba.batch_set(0, w, st); ba.batch_replicate(W); ba.batch_upload(W); ba.batch_gn_step(W, 1e-8)
dx, costs = ba.batch_download(W, 15 * w.N + w.M)
ref = bo.gn_step(w, st, schur=True)
print("batch 160:", rel(dx[0], ref['dx']), rel(dx[W - 1], ref['dx']), np.array_equal(dx[0], dx[W - 1]), costs[0], ref['cost'], flush=True)
ref_state, ref_sum = bo.solve(w, st, max_iter=6)
out, summ = ba.solve(w, st, max_iterations=6)
print("solve:", summ['iterations'], ref_sum['iterations'], summ['final_cost'], ref_sum['final_cost'], summ['accepted_steps'],
      np.linalg.norm(out.p - ref_state.p), summ['solve_seconds'], flush=True)
check("cfg3", *synth.make_cfg3()[:2], ba)
check("cfg4", *synth.make_cfg4()[:2], ba)
w3, s3, _ = synth.make_cfg3(N=6, M=100)
ref_state, ref_sum = bo.solve(w3, s3, max_iter=6)
out, summ = ba.solve(w3, s3, max_iterations=6)
print("solve cfg3:", summ['iterations'], ref_sum['iterations'], summ['final_cost'], ref_sum['final_cost'], summ['accepted_steps'], flush=True)
ba.close()
