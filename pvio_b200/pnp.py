"""Host mirror of visual_inertial_pnp (pvio/src/pvio/estimation/pnp.cpp:32-100) over the C-ABI:
one kernel launch runs the whole single-frame trust-region solve on the GPU."""
import ctypes as C

import numpy as np

from . import _lib


def imu_record(imu):
    """PreIntegrator outputs (dict as pvio_b200.so3.PreIntegrator.integrate returns) -> [288] record."""
    rec = np.zeros(_lib.IMU_STRIDE)
    rec[0] = imu['dt']
    rec[1:5], rec[5:8], rec[8:11] = imu['dq'], imu['dp'], imu['dv']
    rec[11:236] = np.asarray(imu['sqrt_inv_cov']).reshape(225)
    for k, o in (('dq_dbg', 236), ('dp_dbg', 245), ('dp_dba', 254), ('dv_dbg', 263), ('dv_dba', 272)):
        rec[o:o + 9] = np.asarray(imu[k]).reshape(9)
    return rec


def visual_inertial_pnp(ba, frame, last_frame, imu, points, z, cam_q, cam_p, imu_q, imu_p, sqrt_inv_cov,
                        use_inertial=True, max_iterations=10, initial_radius=0.0, cauchy_a=1.0):
    """ba: a BundleAdjustor (device handle).  frame / last_frame: 16-vectors (q xyzw, p, v, bg, ba).
    Returns (frame', summary dict)."""
    pb = _lib.CPnpProblem()
    pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
    zz = np.ascontiguousarray(z, dtype=np.float64).reshape(-1, 2)
    last = np.ascontiguousarray(last_frame, dtype=np.float64)
    out = np.array(frame, dtype=np.float64, copy=True)
    rec = imu_record(imu) if use_inertial else np.zeros(_lib.IMU_STRIDE)
    pb.n_points, pb.use_inertial = len(pts), 1 if use_inertial else 0
    pb.points, pb.z = _lib._ptr(pts, C.c_double), _lib._ptr(zz, C.c_double)
    pb.cam_q_cs[:], pb.cam_p_cs[:] = list(cam_q), list(cam_p)
    pb.imu_q_cs[:], pb.imu_p_cs[:] = list(imu_q), list(imu_p)
    pb.sqrt_inv_cov[:] = list(np.asarray(sqrt_inv_cov, dtype=np.float64).reshape(4))
    pb.cauchy_a = cauchy_a
    pb.last_frame, pb.imu_data = _lib._ptr(last, C.c_double), _lib._ptr(rec, C.c_double)
    opt = _lib.COptions(max_iterations, 1e6, 1, 0, initial_radius)
    sm = _lib.CSummary()
    ba._ck(ba.lib.pvio_b200_pnp_solve(ba.h, C.byref(pb), _lib._ptr(out, C.c_double), C.byref(opt), C.byref(sm)))
    return out, {k: getattr(sm, k) for k, _ in _lib.CSummary._fields_}
