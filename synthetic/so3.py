"""Small vectorised SO(3)/quaternion helpers for host-side input synthesis and the
host mirror of PVIO's PreIntegrator (estimation/preintegrator.cpp:39-100).
Quaternions are (x, y, z, w), Hamilton product -- Eigen's convention, which the
reference uses throughout (pvio/include/pvio/pvio.h:28-40).
Not part of the GPU hot path; the device restatement lives in csrc/ba_device.cuh.
"""
import numpy as np


def hat(w):
    w = np.asarray(w, dtype=np.float64)
    o = np.zeros(w.shape[:-1] + (3, 3))
    o[..., 0, 1], o[..., 0, 2] = -w[..., 2], w[..., 1]
    o[..., 1, 0], o[..., 1, 2] = w[..., 2], -w[..., 0]
    o[..., 2, 0], o[..., 2, 1] = -w[..., 1], w[..., 0]
    return o


def qmul(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def qconj(q):
    q = np.asarray(q, dtype=np.float64)
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def qmat(q):
    q = np.asarray(q, dtype=np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - z * w)
    R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w)
    R[..., 2, 1] = 2 * (y * z + x * w)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def qrot(q, v):
    return np.einsum('...ij,...j->...i', qmat(q), np.asarray(v, dtype=np.float64))


def qexp(w):
    w = np.asarray(w, dtype=np.float64)
    a = np.linalg.norm(w, axis=-1, keepdims=True)
    k = np.where(a > 1e-12, np.sin(0.5 * a) / np.where(a > 1e-12, a, 1.0), 0.5)
    return np.concatenate([k * w, np.cos(0.5 * a)], axis=-1)


def qnormalize(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def mat2quat(R):
    """Rotation matrix -> quaternion (x,y,z,w), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    q = q / np.linalg.norm(q)
    return q if q[3] >= 0 else -q


def right_jacobian(w):
    w = np.asarray(w, dtype=np.float64)
    a = np.linalg.norm(w)
    hw = hat(w)
    if a < 1e-5:
        c, s = 0.5 - a * a / 24.0, 1.0 / 6.0 - a * a / 120.0
    else:
        c, s = (1 - np.cos(a)) / (a * a), (a - np.sin(a)) / (a ** 3)
    return np.eye(3) - c * hw + s * hw @ hw


class PreIntegrator:
    """Host mirror of estimation/preintegrator.cpp:39-100 (integrate / increment /
    compute_sqrt_inv_cov).  Produces the per-factor inputs of the IMU kernel."""

    def __init__(self, cov_w, cov_a, cov_bg, cov_ba):
        self.cov_w, self.cov_a, self.cov_bg, self.cov_ba = [np.asarray(c, dtype=np.float64) * np.eye(3)
                                                            for c in (cov_w, cov_a, cov_bg, cov_ba)]
        self.data = []          # list of (t, w[3], a[3])

    def integrate(self, t_end, bg, ba):
        dq = np.array([0., 0., 0., 1.])
        dp, dv = np.zeros(3), np.zeros(3)
        cov = np.zeros((15, 15))
        dq_dbg, dp_dbg, dp_dba = np.zeros((3, 3)), np.zeros((3, 3)), np.zeros((3, 3))
        dv_dbg, dv_dba = np.zeros((3, 3)), np.zeros((3, 3))
        T = 0.0
        ts = [d[0] for d in self.data] + [t_end]
        for i, (t, w_m, a_m) in enumerate(self.data):
            dt = ts[i + 1] - t
            w, a = w_m - bg, a_m - ba
            Rd = qmat(dq)
            Rinc_t = qmat(qconj(qexp(w * dt)))
            A = np.eye(9)
            A[0:3, 0:3] = Rinc_t
            A[6:9, 0:3] = -dt * Rd @ hat(a)
            A[3:6, 0:3] = -0.5 * dt * dt * Rd @ hat(a)
            A[3:6, 6:9] = dt * np.eye(3)
            B = np.zeros((9, 6))
            B[0:3, 0:3] = dt * right_jacobian(w * dt)
            B[6:9, 3:6] = dt * Rd
            B[3:6, 3:6] = 0.5 * dt * dt * Rd
            inv_dt = 1.0 / max(dt, 1.0e-7)
            Q = np.zeros((6, 6))
            Q[0:3, 0:3] = self.cov_w * inv_dt
            Q[3:6, 3:6] = self.cov_a * inv_dt
            cov[0:9, 0:9] = A @ cov[0:9, 0:9] @ A.T + B @ Q @ B.T
            cov[9:12, 9:12] += self.cov_bg * dt
            cov[12:15, 12:15] += self.cov_ba * dt
            dp_dbg = dp_dbg + dt * dv_dbg - 0.5 * dt * dt * Rd @ hat(a) @ dq_dbg
            dp_dba = dp_dba + dt * dv_dba - 0.5 * dt * dt * Rd
            dv_dbg = dv_dbg - dt * Rd @ hat(a) @ dq_dbg
            dv_dba = dv_dba - dt * Rd
            dq_dbg = Rinc_t @ dq_dbg - dt * right_jacobian(w * dt)
            T += dt
            dp = dp + dt * dv + 0.5 * dt * dt * (Rd @ a)
            dv = dv + dt * (Rd @ a)
            dq = qnormalize(qmul(dq, qexp(w * dt)))
        L = np.linalg.cholesky(np.linalg.inv(cov))
        return dict(dt=T, dq=dq, dp=dp, dv=dv, cov=cov, sqrt_inv_cov=L.T.copy(),
                    dq_dbg=dq_dbg, dp_dbg=dp_dbg, dp_dba=dp_dba, dv_dbg=dv_dbg, dv_dba=dv_dba)
