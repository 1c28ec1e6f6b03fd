import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pvio_b200 import _lib
from pvio_b200.bundle_adjustor import BundleAdjustor
ba = BundleAdjustor(max_windows=1, max_frames=10, max_landmarks=64, max_obs=512)
lib = _lib.load()
f = lib.pvio_b200_selftest_syrk_raw
f.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float), C.c_int]
np.set_printoptions(linewidth=250, precision=3, suppress=True)
rows = np.r_[0:16, 32:48, 64:80, 96:112]
def run(A, mode):
    A = np.ascontiguousarray(A, dtype=np.float32)
    out = np.zeros(128 * 64 + 4, dtype=np.float32)
    rc = f(ba.h, A.ctypes.data_as(C.POINTER(C.c_float)), A.shape[0], out.ctypes.data_as(C.POINTER(C.c_float)), mode)
    assert rc == 0
    return out[:128 * 64].reshape(128, 64)[rows].astype(np.float64)
rng = np.random.default_rng(0)
for K in (8, 16, 32, 64, 128):
    A = rng.standard_normal((K, 64)).astype(np.float32)
    ref = A.astype(np.float64).T @ A.astype(np.float64)
    mag = np.abs(A).astype(np.float64).T @ np.abs(A).astype(np.float64)
    D = run(A, 0)
    e = (D - ref)
    dg = np.diag(e) / np.diag(ref)
    # fp32 sequential RN accumulation for comparison
    acc = np.zeros((64, 64), dtype=np.float32)
    for k in range(K):
        acc = acc + np.outer(A[k], A[k]).astype(np.float32)
    e32 = acc.astype(np.float64) - ref
    print(f"K={K:4d}  3xTF32: max|e|/mag {np.max(np.abs(e)/mag):.2e}  diag rel err mean {dg.mean():+.2e} rms {np.sqrt((dg**2).mean()):.2e}"
          f"   | fp32 RN loop: max|e|/mag {np.max(np.abs(e32)/mag):.2e} diag mean {np.mean(np.diag(e32)/np.diag(ref)):+.2e} rms {np.sqrt(np.mean((np.diag(e32)/np.diag(ref))**2)):.2e}")
