"""Flat SoA "window" container: the host-side image of what PVIO's shim gathers from
Map / Frame / Track and what the C-ABI in include/pvio_b200.h receives.

Field <-> reference mapping (paths relative to /root/reference/pvio/src/pvio):
  N, frame_fixed          map/map.h frame_num(), Frame::flag(FF_FIX_POSE)  (bundle_adjustor.cpp:75-82)
  cam_q_cs/p_cs, imu_*    Frame::camera / Frame::imu ExtrinsicParams        (estimation/state.h:38-41)
  sqrt_inv_cov, K_fx/fy   Frame::sqrt_inv_cov, Frame::K                      (core/core.cpp:112-116)
  lm_anchor, lm_z_ref     Track::first_frame()/first_keypoint()              (map/track.h:61-63)
  obs_frame, obs_z        Track::keypoint_map() minus the anchor observation (bundle_adjustor.cpp:149)
  rho (State)             Track::landmark.inv_depth                          (estimation/state.h:71-82)
  imu_*                   Frame::preintegration.{delta,jacobian}             (estimation/preintegrator.h:29-44)
  prior_*                 MarginalizationErrorCost members                   (ceres/marginalization_error_cost.h:96-105)
  plane_*, pt_*           Plane::parameter, plane tracks' keypoint_map()     (map/plane.h:31-46)
Observations are sorted by landmark, then by frame index; landmarks are in the
reference's first-visit order over frames x keypoints (bundle_adjustor.cpp:92-103).
Keypoints are NORMALISED image coordinates (remove_k applied, map/frame.cpp:85).
"""
from dataclasses import dataclass, field, replace
import numpy as np


def _z(shape, dt=np.float64):
    return np.zeros(shape, dtype=dt)


@dataclass
class Window:
    N: int = 0
    M: int = 0
    K: int = 0
    use_inertial: bool = False
    cauchy_a: float = 1.0
    frame_fixed: np.ndarray = field(default_factory=lambda: _z(0, np.uint8))
    cam_q_cs: np.ndarray = field(default_factory=lambda: np.array([0., 0., 0., 1.]))
    cam_p_cs: np.ndarray = field(default_factory=lambda: _z(3))
    imu_q_cs: np.ndarray = field(default_factory=lambda: np.array([0., 0., 0., 1.]))
    imu_p_cs: np.ndarray = field(default_factory=lambda: _z(3))
    sqrt_inv_cov: np.ndarray = field(default_factory=lambda: np.eye(2))
    K_fx: float = 1.0
    K_fy: float = 1.0
    # landmarks / observations
    lm_anchor: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    lm_z_ref: np.ndarray = field(default_factory=lambda: _z((0, 2)))
    lm_obs_begin: np.ndarray = field(default_factory=lambda: _z(1, np.int32))
    lm_in_victim: np.ndarray = field(default_factory=lambda: _z(0, np.uint8))
    obs_frame: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    obs_z: np.ndarray = field(default_factory=lambda: _z((0, 2)))
    # IMU factors
    n_imu: int = 0
    imu_frame_i: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    imu_frame_j: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    imu_dt: np.ndarray = field(default_factory=lambda: _z(0))
    imu_dq: np.ndarray = field(default_factory=lambda: _z((0, 4)))
    imu_dp: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_dv: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_sqrt_inv_cov: np.ndarray = field(default_factory=lambda: _z((0, 15, 15)))
    imu_dq_dbg: np.ndarray = field(default_factory=lambda: _z((0, 3, 3)))
    imu_dp_dbg: np.ndarray = field(default_factory=lambda: _z((0, 3, 3)))
    imu_dp_dba: np.ndarray = field(default_factory=lambda: _z((0, 3, 3)))
    imu_dv_dbg: np.ndarray = field(default_factory=lambda: _z((0, 3, 3)))
    imu_dv_dba: np.ndarray = field(default_factory=lambda: _z((0, 3, 3)))
    imu_bg0: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    imu_ba0: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    # marginalisation prior
    n_prior: int = 0
    prior_frames: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    prior_S: np.ndarray = field(default_factory=lambda: _z((0, 0)))
    prior_e: np.ndarray = field(default_factory=lambda: _z(0))
    prior_q0: np.ndarray = field(default_factory=lambda: _z((0, 4)))
    prior_p0: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    prior_v0: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    prior_bg0: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    prior_ba0: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    # planes (only planes with >= 20 tracks reach the window, bundle_adjustor.cpp:165)
    n_planes: int = 0
    plane_normal: np.ndarray = field(default_factory=lambda: _z((0, 3)))
    plane_distance: np.ndarray = field(default_factory=lambda: _z(0))
    plane_sqrt_inv_cov: float = 100.0   # sqrt(1 / plane_distance_cov), config.cpp:24-26
    n_ptracks: int = 0
    pt_plane: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    pt_obs_begin: np.ndarray = field(default_factory=lambda: _z(1, np.int32))
    pt_obs_frame: np.ndarray = field(default_factory=lambda: _z(0, np.int32))
    pt_obs_z: np.ndarray = field(default_factory=lambda: _z((0, 2)))

    def with_bias_lin_point(self, st):
        """Quirk Q1: bg_i_0 / ba_i_0 alias frame_i->motion (preintegration_error_cost.h:57-58)."""
        if self.n_imu == 0:
            return self
        return replace(self, imu_bg0=st.bg[self.imu_frame_i].copy(),
                       imu_ba0=st.ba[self.imu_frame_i].copy())

    def validate(self):
        assert self.frame_fixed.shape == (self.N,)
        assert self.lm_anchor.shape == (self.M,) and self.lm_z_ref.shape == (self.M, 2)
        assert self.lm_obs_begin.shape == (self.M + 1,) and int(self.lm_obs_begin[-1]) == self.K
        assert self.obs_frame.shape == (self.K,) and self.obs_z.shape == (self.K, 2)
        if self.lm_in_victim.shape != (self.M,):
            self.lm_in_victim = np.zeros(self.M, dtype=np.uint8)
        assert self.imu_dt.shape == (self.n_imu,)
        assert self.prior_S.shape == (15 * self.n_prior, 15 * self.n_prior)
        assert self.pt_obs_begin.shape == (self.n_ptracks + 1,)
        return self


@dataclass
class State:
    q: np.ndarray      # [N,4] (x,y,z,w)
    p: np.ndarray      # [N,3]
    v: np.ndarray      # [N,3]
    bg: np.ndarray     # [N,3]
    ba: np.ndarray     # [N,3]
    rho: np.ndarray    # [M]  inverse depths

    def copy(self):
        return State(self.q.copy(), self.p.copy(), self.v.copy(), self.bg.copy(),
                     self.ba.copy(), self.rho.copy())

    @staticmethod
    def zeros(N, M):
        q = np.zeros((N, 4))
        q[:, 3] = 1.0
        return State(q, _z((N, 3)), _z((N, 3)), _z((N, 3)), _z((N, 3)), np.ones(M))
