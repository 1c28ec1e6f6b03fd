"""IMU pre-integration kernel (csrc/imu.cu) against the NumPy restatement of preintegrator.cpp
(oracle/imu_oracle.py) on the same samples.  fp64 on both sides: deltas / Jacobians to 1e-11,
sqrt_inv_cov (a Cholesky factor of an ill-conditioned 15x15 inverse) to 1e-7 relative."""
import numpy as np
import pytest

from oracle import imu_oracle
from pvio_b200 import imu
from pvio_b200.bundle_adjustor import BundleAdjustor

pytestmark = pytest.mark.gpu

# config/euroc.yaml:23-42 (continuous-time noise densities squared)
COV = (np.eye(3) * 2.8791e-8, np.eye(3) * 4.0e-6, np.eye(3) * 3.7608e-10, np.eye(3) * 9.0e-6)


def _factor(rng, K, hz=200.0, ragged=False):
    t = np.cumsum(np.r_[0.0, rng.uniform(0.5, 1.5, K - 1) / hz]) if ragged else np.arange(K) / hz
    w = rng.normal(0, 0.3, (K, 3)) + np.array([0.1, -0.2, 0.05])
    a = rng.normal(0, 1.0, (K, 3)) + np.array([0.0, 0.0, 9.81])
    samples = np.c_[t, w, a]
    t_end = t[-1] + rng.uniform(0.1, 1.0) / hz
    return samples, t_end, rng.normal(0, 1e-2, 3), rng.normal(0, 5e-2, 3)


def _compare(rec, samples, t_end, bg, ba):
    ref = imu_oracle.record(imu_oracle.integrate(samples, t_end, bg, ba, *COV), bg, ba)
    assert np.allclose(rec[:11], ref[:11], rtol=0, atol=1e-11)                 # dt, dq, dp, dv
    assert np.allclose(rec[236:281], ref[236:281], rtol=1e-10, atol=1e-12)     # bias Jacobians
    assert np.array_equal(rec[281:287], ref[281:287])                          # linearisation point
    U, U0 = rec[11:236].reshape(15, 15), ref[11:236].reshape(15, 15)
    assert np.all(np.tril(U, -1) == 0.0)                                       # matrixL().transpose() is upper triangular
    assert np.max(np.abs(U - U0)) <= 1e-7 * np.max(np.abs(U0))
    # the information matrix itself
    P, P0 = U.T @ U, U0.T @ U0
    assert np.max(np.abs(P - P0)) <= 1e-7 * np.max(np.abs(P0))


def test_preintegrate_matches_oracle():
    rng = np.random.default_rng(648)
    ba = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=8, max_obs=16)
    factors = [_factor(rng, K, ragged=(i % 2 == 1)) for i, K in enumerate([2, 3, 10, 20, 21, 40, 100, 7, 33])]
    rec = imu.preintegrate(ba, factors, COV)
    assert rec.shape == (len(factors), 288)
    for r, f in zip(rec, factors):
        _compare(r, *f)


def test_preintegrate_batch_is_independent_of_order():
    rng = np.random.default_rng(3)
    ba = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=8, max_obs=16)
    factors = [_factor(rng, int(rng.integers(5, 50))) for _ in range(300)]
    rec = imu.preintegrate(ba, factors, COV)
    perm = rng.permutation(len(factors))
    rec2 = imu.preintegrate(ba, [factors[i] for i in perm], COV)
    assert np.array_equal(rec[perm], rec2)                                      # bit-identical, whatever the warp / CTA
    _compare(rec[17], *factors[17])


def test_preintegrated_records_feed_the_ba_kernels():
    """cfg3 with its IMU records replaced by device-integrated ones: same GN step as with the host records."""
    from oracle import ba_oracle as bo
    from synthetic import synth
    w, st, truth = synth.make_cfg3(N=6, M=100)
    ba = BundleAdjustor(max_windows=1, max_frames=8, max_landmarks=128, max_obs=1024)
    rec = imu.preintegrate(ba, truth.imu_factors, truth.imu_noise)
    assert np.allclose(rec[:, 1:5], w.imu_dq, rtol=0, atol=1e-11) and np.allclose(rec[:, 5:8], w.imu_dp, rtol=0, atol=1e-11)
    w.imu_dt, w.imu_dq, w.imu_dp, w.imu_dv = rec[:, 0].copy(), rec[:, 1:5].copy(), rec[:, 5:8].copy(), rec[:, 8:11].copy()
    w.imu_sqrt_inv_cov = rec[:, 11:236].reshape(-1, 15, 15).copy()
    for k, o in (('dq_dbg', 236), ('dp_dbg', 245), ('dp_dba', 254), ('dv_dbg', 263), ('dv_dba', 272)):
        setattr(w, 'imu_' + k, rec[:, o:o + 9].reshape(-1, 3, 3).copy())
    ref = bo.gn_step(w, st, schur=True)
    out = ba.gn_step(w, st, mu=1e-8)
    assert np.linalg.norm(out['dx'] - ref['dx']) < 1e-5 * np.linalg.norm(ref['dx'])


def test_preintegrate_rejects_empty_factor():
    ba = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=8, max_obs=16)
    with pytest.raises(Exception):
        imu.preintegrate(ba, [(np.zeros((0, 7)), 0.1, np.zeros(3), np.zeros(3))], COV)
