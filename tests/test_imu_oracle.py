"""CPU checks of oracle/imu_oracle.py (no reference test exists for preintegrator.cpp): the bias Jacobians
against central differences of integrate() in the biases, the information factor, and agreement with the
package's own host mirror (pvio_b200.so3.PreIntegrator, written independently for the synthetic windows)."""
import numpy as np

from oracle import imu_oracle, lie
from synthetic import so3

COV = (np.eye(3) * 2.8791e-8, np.eye(3) * 4.0e-6, np.eye(3) * 3.7608e-10, np.eye(3) * 9.0e-6)


def _samples(rng, K=25, hz=200.0):
    t = np.arange(K) / hz
    return np.c_[t, rng.normal(0, 0.3, (K, 3)) + 0.1, rng.normal(0, 1.0, (K, 3)) + np.array([0, 0, 9.81])], t[-1] + 0.7 / hz


def test_bias_jacobians_match_finite_differences():
    rng = np.random.default_rng(0)
    s, t_end = _samples(rng)
    bg, ba = rng.normal(0, 1e-2, 3), rng.normal(0, 5e-2, 3)
    out = imu_oracle.integrate(s, t_end, bg, ba, *COV)
    eps = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = eps
        pg, mg = imu_oracle.integrate(s, t_end, bg + e, ba, *COV), imu_oracle.integrate(s, t_end, bg - e, ba, *COV)
        pa, ma = imu_oracle.integrate(s, t_end, bg, ba + e, *COV), imu_oracle.integrate(s, t_end, bg, ba - e, *COV)
        assert np.allclose((pg['dp'] - mg['dp']) / (2 * eps), out['dp_dbg'][:, k], atol=1e-7)
        assert np.allclose((pg['dv'] - mg['dv']) / (2 * eps), out['dv_dbg'][:, k], atol=1e-7)
        assert np.allclose((pa['dp'] - ma['dp']) / (2 * eps), out['dp_dba'][:, k], atol=1e-7)
        assert np.allclose((pa['dv'] - ma['dv']) / (2 * eps), out['dv_dba'][:, k], atol=1e-7)
        # dq(bg + e) = dq * exp(dq_dbg e): preintegration_error_cost.h:61
        dth = lie.logmap(lie.qmul(lie.qconj(mg['dq']), pg['dq'])) / (2 * eps)
        assert np.allclose(dth, out['dq_dbg'][:, k], atol=1e-6)


def test_information_factor_and_host_mirror():
    rng = np.random.default_rng(1)
    s, t_end = _samples(rng, K=40)
    bg, ba = rng.normal(0, 1e-2, 3), rng.normal(0, 5e-2, 3)
    out = imu_oracle.integrate(s, t_end, bg, ba, *COV)
    U = out['sqrt_inv_cov']
    assert np.all(np.tril(U, -1) == 0)
    assert np.allclose(U.T @ U @ out['cov'], np.eye(15), atol=1e-6)
    pre = so3.PreIntegrator(COV[0][0, 0], COV[1][0, 0], COV[2][0, 0], COV[3][0, 0])
    pre.data = [(r[0], r[1:4], r[4:7]) for r in s]
    mir = pre.integrate(t_end, bg, ba)
    for k in ('dq', 'dp', 'dv', 'dq_dbg', 'dp_dbg', 'dp_dba', 'dv_dbg', 'dv_dba', 'cov'):
        assert np.allclose(out[k], mir[k], rtol=1e-12, atol=1e-15), k
