"""tcgen05 3xTF32 SYRK building block (csrc/tc_syrk.cuh) against an fp64 product of the same fp32 data."""
import ctypes as C
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("K", [8, 128, 200, 450])
def test_syrk_3xtf32(K):
    from pvio_b200 import _lib
    from pvio_b200.bundle_adjustor import BundleAdjustor
    ba = BundleAdjustor(max_windows=1, max_frames=10, max_landmarks=64, max_obs=512)
    lib = _lib.load()
    lib.pvio_b200_selftest_syrk.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_double)]
    rng = np.random.default_rng(648 + K)
    A = (rng.standard_normal((K, 64)) * np.exp(rng.uniform(-6, 6, size=(K, 64)))).astype(np.float32)
    A[:, 61:] = 0.0
    D = np.zeros((64, 64))
    rc = lib.pvio_b200_selftest_syrk(ba.h, A.ctypes.data_as(C.POINTER(C.c_float)), K, D.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0, lib.pvio_b200_last_error(ba.h)
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    # error relative to the magnitude sum |a_m|.|a_n| (what a dot product's rounding is measured against)
    mag = np.abs(A).astype(np.float64).T @ np.abs(A).astype(np.float64) + 1e-300
    err = np.max(np.abs(D - ref) / mag)
    print('K', K, 'max err / mag', err, 'diag bias', np.mean(np.diag(D - ref)[:61] / np.diag(ref)[:61]))
    assert err < 2e-6, err          # 3xTF32: ~2^-21; one TF32 pass would be ~1e-3
