"""Host-side mirror of PreIntegrator::integrate (pvio/src/pvio/estimation/preintegrator.cpp:85-98) over
the C ABI: batches of independent IMU factors, one warp each on the device (csrc/imu.cu)."""
import ctypes as C

import numpy as np

from . import _lib


def preintegrate(ba, factors, noise_cov):
    """ba: a BundleAdjustor (device handle).  factors: list of (samples [K][7], t_end, bg[3], ba[3]).
    noise_cov: (cov_w, cov_a, cov_bg, cov_ba) 3x3 each.  Returns records [n][288] (PVIO_B200_IMU_* layout)."""
    n = len(factors)
    begin = np.zeros(n + 1, dtype=np.int32)
    for i, f in enumerate(factors):
        begin[i + 1] = begin[i] + len(f[0])
    samples = np.ascontiguousarray(np.concatenate([np.asarray(f[0], dtype=np.float64).reshape(-1, 7) for f in factors]))
    t_end = np.array([f[1] for f in factors], dtype=np.float64)
    bias = np.ascontiguousarray(np.array([np.concatenate([f[2], f[3]]) for f in factors], dtype=np.float64))
    noise = np.ascontiguousarray(np.array([np.asarray(c, dtype=np.float64).reshape(9) for c in noise_cov]))
    rec = np.zeros((n, _lib.IMU_STRIDE))
    fn = ba.lib.pvio_b200_preintegrate
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    ba._ck(fn(ba.h, n, _lib._ptr(begin, C.c_int32), _lib._ptr(samples, C.c_double), _lib._ptr(t_end, C.c_double),
              _lib._ptr(bias, C.c_double), _lib._ptr(noise, C.c_double), _lib._ptr(rec, C.c_double)))
    return rec
