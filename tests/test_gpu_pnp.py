"""GPU parity of pvio_b200_pnp_solve (whole single-frame solve in one kernel) against the fp64 oracle
restatement of pvio/src/pvio/estimation/pnp.cpp:32-100."""
import numpy as np
import pytest

from oracle import pnp_oracle as po
from pvio_b200 import pnp
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ba():
    b = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=16, max_obs=64)
    yield b
    b.close()


@pytest.mark.parametrize("inertial,npts,r0", [(True, 150, 0.0), (False, 150, 0.0), (True, 40, 3.0), (False, 8, 0.0)])
def test_pnp_matches_oracle(ba, inertial, npts, r0):
    d = synth.make_pnp(use_inertial=inertial, n_points=npts, seed=651 + npts)
    kw = dict(radius0=r0) if r0 > 0 else {}
    xr, sr = po.pnp_solve(d['frame'], d['last'], d['imu'], d['pts'], d['zs'], d['cam_q'], d['cam_p'], d['imu_q'],
                          d['imu_p'], d['W'], inertial, **kw)
    xg, sg = pnp.visual_inertial_pnp(ba, d['frame'], d['last'], d['imu'], d['pts'], d['zs'], d['cam_q'], d['cam_p'],
                                     d['imu_q'], d['imu_p'], d['W'], inertial, initial_radius=r0)
    print(sr['iterations'], sg['iterations'], sr['final_cost'], sg['final_cost'], sg['solve_seconds'])
    assert sg['iterations'] == sr['iterations'] and sg['accepted_steps'] == sum(sr['accepted'])
    assert abs(sg['final_cost'] - sr['final_cost']) <= 1e-9 * sr['final_cost']
    assert abs(sg['initial_cost'] - sr['initial_cost']) <= 1e-9 * sr['initial_cost']
    move = np.linalg.norm(xr - d['frame'])
    assert np.linalg.norm(xg - xr) < 1e-8 * max(move, 1e-3)
    if not inertial:
        assert np.array_equal(xg[7:], d['frame'][7:])          # v, bg, ba are not parameters (pnp.cpp:46-54)


def test_pnp_empty_points_inertial_only(ba):
    d = synth.make_pnp(use_inertial=True, n_points=5)
    xr, sr = po.pnp_solve(d['frame'], d['last'], d['imu'], d['pts'][:0], d['zs'][:0], d['cam_q'], d['cam_p'],
                          d['imu_q'], d['imu_p'], d['W'], True)
    xg, sg = pnp.visual_inertial_pnp(ba, d['frame'], d['last'], d['imu'], d['pts'][:0], d['zs'][:0], d['cam_q'],
                                     d['cam_p'], d['imu_q'], d['imu_p'], d['W'], True)
    assert sg['iterations'] == sr['iterations']
    assert np.linalg.norm(xg - xr) < 1e-8
