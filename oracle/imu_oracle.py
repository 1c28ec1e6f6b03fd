"""TEST INFRASTRUCTURE (CPU oracle) -- IMU pre-integration, a restatement of
pvio/src/pvio/estimation/preintegrator.cpp:24-100 (reset / increment / integrate / compute_sqrt_inv_cov)
in NumPy fp64.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Parity pin: the reference has no test for this path ("parity unpinned" in the sense of DESIGN.md 2);
self-consistency is pinned in tests/test_oracle_functors.py: the bias Jacobians against finite differences
of integrate() with respect to the biases, and sqrt_inv_cov^T sqrt_inv_cov = cov^-1.

Error-state order (preintegrator.h): ES_Q 0, ES_P 3, ES_V 6, ES_BG 9, ES_BA 12."""
import numpy as np

from . import lie


def integrate(samples, t_end, bg, ba, cov_w, cov_a, cov_bg, cov_ba):
    """samples: [K][7] rows (t, w xyz, a xyz); the last sample is integrated up to t_end
    (preintegrator.cpp:85-98).  cov_*: 3x3 noise matrices.  Returns the dict of the factor's inputs."""
    samples = np.asarray(samples, dtype=np.float64).reshape(-1, 7)
    bg, ba = np.asarray(bg, dtype=np.float64), np.asarray(ba, dtype=np.float64)
    dq = np.array([0.0, 0.0, 0.0, 1.0])                       # :25-30 reset
    dp, dv, T = np.zeros(3), np.zeros(3), 0.0
    cov = np.zeros((15, 15))
    dq_dbg, dp_dbg, dp_dba, dv_dbg, dv_dba = (np.zeros((3, 3)) for _ in range(5))
    ts = list(samples[:, 0]) + [float(t_end)]
    for i in range(len(samples)):
        dt = ts[i + 1] - ts[i]                                # :89-92
        w = samples[i, 1:4] - bg                              # :42-43
        a = samples[i, 4:7] - ba
        Rd = lie.qmat(dq)
        Ri_t = lie.qmat(lie.qconj(lie.expmap(w * dt)))
        Jr = lie.right_jacobian(w * dt)
        A = np.eye(9)                                         # :46-51
        A[0:3, 0:3] = Ri_t
        A[6:9, 0:3] = -dt * Rd @ lie.hat(a)
        A[3:6, 0:3] = -0.5 * dt * dt * Rd @ lie.hat(a)
        A[3:6, 6:9] = dt * np.eye(3)
        B = np.zeros((9, 6))                                  # :53-57
        B[0:3, 0:3] = dt * Jr
        B[6:9, 3:6] = dt * Rd
        B[3:6, 3:6] = 0.5 * dt * dt * Rd
        inv_dt = 1.0 / max(dt, 1.0e-7)                        # :60
        Q = np.zeros((6, 6))
        Q[0:3, 0:3] = cov_w * inv_dt
        Q[3:6, 3:6] = cov_a * inv_dt
        cov[0:9, 0:9] = A @ cov[0:9, 0:9] @ A.T + B @ Q @ B.T  # :65
        cov[9:12, 9:12] += cov_bg * dt                        # :66-67
        cov[12:15, 12:15] += cov_ba * dt
        dp_dbg = dp_dbg + dt * dv_dbg - 0.5 * dt * dt * Rd @ lie.hat(a) @ dq_dbg   # :71-75, in this order
        dp_dba = dp_dba + dt * dv_dba - 0.5 * dt * dt * Rd
        dv_dbg = dv_dbg - dt * Rd @ lie.hat(a) @ dq_dbg
        dv_dba = dv_dba - dt * Rd
        dq_dbg = Ri_t @ dq_dbg - dt * Jr
        T += dt                                               # :78-81
        dp = dp + dt * dv + 0.5 * dt * dt * (Rd @ a)
        dv = dv + dt * (Rd @ a)
        dq = lie.qnormalized(lie.qmul(dq, lie.expmap(w * dt)))
    sqrt_inv_cov = np.linalg.cholesky(np.linalg.inv(cov)).T   # :100-102  LLT(cov^-1).matrixL()^T
    return dict(dt=T, dq=dq, dp=dp, dv=dv, cov=cov, sqrt_inv_cov=sqrt_inv_cov,
                dq_dbg=dq_dbg, dp_dbg=dp_dbg, dp_dba=dp_dba, dv_dbg=dv_dbg, dv_dba=dv_dba)


def record(out, bg, ba):
    """the [288] record of include/pvio_b200.h (PVIO_B200_IMU_* offsets)."""
    rec = np.zeros(288)
    rec[0] = out['dt']
    rec[1:5], rec[5:8], rec[8:11] = out['dq'], out['dp'], out['dv']
    rec[11:236] = out['sqrt_inv_cov'].reshape(225)
    for k, o in (('dq_dbg', 236), ('dp_dbg', 245), ('dp_dba', 254), ('dv_dbg', 263), ('dv_dba', 272)):
        rec[o:o + 9] = out[k].reshape(9)
    rec[281:284], rec[284:287] = bg, ba
    return rec
