"""ctypes loader for oracle/_build/libba_oracle.so (the C restatement, CPU baseline).
TEST / BASELINE INFRASTRUCTURE ONLY -- see oracle/ba_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libba_oracle.so")
_lib = None


def load(rebuild=False):
    global _lib
    if _lib is not None and not rebuild:
        return _lib
    src = os.path.join(_HERE, "ba_oracle.c")
    if rebuild or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if rebuild else []), check=True)
    _lib = C.CDLL(_SO)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _args(win, st):
    assert not win.use_inertial, "the C restatement covers reprojection-only windows (BASELINE config 2)"
    frames = np.ascontiguousarray(np.concatenate([st.q, st.p, st.v, st.bg, st.ba], axis=1))
    keep = dict(fixed=np.ascontiguousarray(win.frame_fixed, np.uint8), cq=np.ascontiguousarray(win.cam_q_cs, np.float64),
                cp=np.ascontiguousarray(win.cam_p_cs, np.float64), W=np.ascontiguousarray(win.sqrt_inv_cov, np.float64).reshape(4),
                an=np.ascontiguousarray(win.lm_anchor, np.int32), zr=np.ascontiguousarray(win.lm_z_ref, np.float64),
                ob=np.ascontiguousarray(win.lm_obs_begin, np.int32), of=np.ascontiguousarray(win.obs_frame, np.int32),
                oz=np.ascontiguousarray(win.obs_z, np.float64), fr=frames, rho=np.ascontiguousarray(st.rho, np.float64))
    a = [win.N, win.M, _p(keep['fixed'], C.c_uint8), _p(keep['cq'], C.c_double), _p(keep['cp'], C.c_double),
         _p(keep['W'], C.c_double), C.c_double(win.cauchy_a), _p(keep['an'], C.c_int32), _p(keep['zr'], C.c_double),
         _p(keep['ob'], C.c_int32), _p(keep['of'], C.c_int32), _p(keep['oz'], C.c_double), _p(keep['fr'], C.c_double),
         _p(keep['rho'], C.c_double)]
    return a, keep


def gn_step(win, st, mu=1e-8):
    lib = load()
    a, keep = _args(win, st)
    dx = np.zeros(15 * win.N + win.M)
    c0, c1 = C.c_double(), C.c_double()
    rc = lib.ba_oracle_gn_step(*a, C.c_double(mu), _p(dx, C.c_double), C.byref(c0), C.byref(c1))
    assert rc == 0
    return dict(dx=dx, cost=c0.value, new_cost=c1.value)


def gn_step_batch(win, st, n_windows, n_threads=0, mu=1e-8):
    """n_windows independent copies over OpenMP threads; returns (dx, costs, threads_used)."""
    lib = load()
    a, keep = _args(win, st)
    dx = np.zeros((n_windows, 15 * win.N + win.M))
    costs = np.zeros((n_windows, 2))
    used = lib.ba_oracle_gn_step_batch(n_windows, n_threads, *a, C.c_double(mu), _p(dx, C.c_double), _p(costs, C.c_double))
    return dx, costs, used
