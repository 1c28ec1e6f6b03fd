"""Generates tests/golden/ba_golden.npz: small seeded windows (BASELINE configs 2, 2b, 3, 4 in
miniature) together with the fp64 oracle's Gauss-Newton step, cost, reduced system and the
marginalisation information.  The reference cannot produce these vectors itself (no Ceres /
Eigen in the build container, no tests in the reference); the oracle that does is pinned by
tests/test_oracle_functors.py.  Run: python tests/golden/make_ba_golden.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from synthetic import synth
from oracle import ba_oracle as bo

CASES = {
    "cfg2": lambda: synth.make_cfg2(N=6, M=64, seed=648),
    "cfg2b": lambda: synth.make_cfg2(N=7, M=70, seed=648, staggered=True),
    "cfg3": lambda: synth.make_cfg3(N=6, M=80, seed=649),
    "cfg4": lambda: synth.make_cfg4(N=6, M=60, seed=650, tracks_per_plane=20),
}
out = {}
for name, mk in CASES.items():
    w, st, _ = mk()
    ref = bo.gn_step(w, st, schur=True)
    out[name + "_dx"] = ref["dx"]
    out[name + "_cost"] = np.array(ref["cost"])
    out[name + "_newcost"] = np.array(bo.total_cost(w, bo.apply_step(w, st, ref["dx"])))
    out[name + "_gred"] = ref["gred"]
    if w.use_inertial and w.n_planes == 0:
        S, e, Hm, bm = bo.marginalize(w, st, 0)
        out[name + "_margH"], out[name + "_margb"] = Hm, bm
    valid, quality = bo.landmark_postpass(w, st)
    out[name + "_valid"], out[name + "_quality"] = valid, quality
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ba_golden.npz"), **out)
print("wrote", sorted(out))
