"""Writes the synthetic windows as the flat binaries oracle/ceres_ref/ceres_dump.cpp reads (see its header comment).
Run where Ceres + PVIO are available; the outputs of ceres_dump go to tests/golden/ceres/<name>.bin and are then
compared by tests/test_ceres_golden.py.  No reference code runs here."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from synthetic import synth  # noqa: E402

CASES = {"cfg2": lambda: synth.make_cfg2(N=6, M=80), "cfg2_full": synth.make_cfg2,
         "cfg3": lambda: synth.make_cfg3(N=6, M=100), "cfg3_full": synth.make_cfg3}


def export(name, out_dir, iters=8):
    w, st, truth = CASES[name]()
    with open(os.path.join(out_dir, name + ".bin"), "wb") as f:
        n_imu = w.n_imu if w.use_inertial else 0
        n_prior = w.n_prior if w.use_inertial else 0
        f.write(struct.pack("<7i", w.N, w.M, w.K, int(w.use_inertial), n_imu, n_prior, iters))
        d = lambda *a: f.write(np.asarray(np.concatenate([np.ravel(x) for x in a]), dtype="<f8").tobytes())
        for i in range(w.N):
            d(st.q[i], st.p[i], st.v[i], st.bg[i], st.ba[i], [float(w.frame_fixed[i])])
        d(w.cam_q_cs, w.cam_p_cs, w.imu_q_cs, w.imu_p_cs)
        d([w.K_fx, w.K_fy, 0.0, 0.0], w.sqrt_inv_cov)
        for l in range(w.M):
            b0, b1 = int(w.lm_obs_begin[l]), int(w.lm_obs_begin[l + 1])
            f.write(struct.pack("<i", int(w.lm_anchor[l]))); d(w.lm_z_ref[l], [st.rho[l]]); f.write(struct.pack("<i", b1 - b0))
            for k in range(b0, b1):
                f.write(struct.pack("<i", int(w.obs_frame[k]))); d(w.obs_z[k])
        for n in range(n_imu):
            samples, t_end, _, _ = truth.imu_factors[n]
            f.write(struct.pack("<i", int(w.imu_frame_j[n]))); d(*truth.imu_noise)
            f.write(struct.pack("<i", len(samples))); d(samples, [t_end])
        if n_prior:
            for i in range(n_prior):
                f.write(struct.pack("<i", int(w.prior_frames[i])))
            d(w.prior_S, w.prior_e)


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/windows"
    os.makedirs(out, exist_ok=True)
    for name in CASES:
        export(name, out)
    print("wrote", sorted(os.listdir(out)))
