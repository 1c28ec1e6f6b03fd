"""Pins the NumPy LK restatement (oracle/klt_oracle.py) against the reference's own KLT
implementation, cv2.calcOpticalFlowPyrLK with PVIO's arguments (opencv_image.cpp:103)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
from oracle import klt_oracle as ko
from synthetic import synth


def _cv(prev, nxt, pts, init):
    p1, st, err = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2).copy(), init.reshape(-1, 1, 2).copy(),
                                           winSize=(21, 21), maxLevel=3,
                                           criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                           flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    return p1.reshape(-1, 2), st.ravel(), err.ravel()


def test_pyramid_and_scharr_match_cv2():
    prev, _, _, _ = synth.make_klt_pair(size=(160, 120), n_points=10)
    assert np.array_equal(ko.pyr_down(prev), cv2.pyrDown(prev))
    odd = prev[:119, :157]
    assert np.array_equal(ko.pyr_down(odd), cv2.pyrDown(odd))
    d = ko.scharr_deriv(prev)
    dx = cv2.Scharr(prev, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101)
    dy = cv2.Scharr(prev, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101)
    assert np.array_equal(d[:, :, 0], dx) and np.array_equal(d[:, :, 1], dy)


@pytest.mark.parametrize("size,n", [((320, 240), 120), ((752, 480), 200)])
def test_lk_matches_cv2(size, n):
    prev, nxt, pts, truth = synth.make_klt_pair(size=size, n_points=n)
    p_cv, st_cv, err_cv = _cv(prev, nxt, pts, pts)
    p_or, st_or, err_or = ko.calc_optical_flow_pyr_lk(prev, nxt, pts, pts)
    assert np.array_equal(st_or, st_cv)
    ok = st_cv == 1
    assert ok.sum() > 0.8 * len(pts)
    assert np.max(np.abs(p_or[ok] - p_cv[ok])) < 1e-2
    assert np.max(np.abs(err_or[ok] - err_cv[ok])) < 1e-2
    # and the tracker actually tracks: close to the true warp
    assert np.median(np.linalg.norm(p_cv[ok] - truth[ok], axis=1)) < 0.1


def test_lk_border_points_and_lost_tracks():
    """Points whose windows leave the image at coarse levels, a point in a flat region
    (min-eigenvalue rejection) and a guess far outside the image (status 0)."""
    prev, nxt, pts, _ = synth.make_klt_pair(size=(320, 240), n_points=60)
    prev = prev.copy(); nxt = nxt.copy()
    prev[100:140, 150:190] = 128; nxt[100:140, 150:190] = 128         # flat patch
    extra = np.array([[170.0, 120.0], [3.0, 4.0], [316.0, 236.0], [22.0, 21.0]], dtype=np.float32)
    pts = np.concatenate([pts, extra])
    init = pts.copy()
    init[5] = [900.0, 50.0]                                            # guess outside the image
    p_cv, st_cv, _ = _cv(prev, nxt, pts, init)
    p_or, st_or, _ = ko.calc_optical_flow_pyr_lk(prev, nxt, pts, init)
    assert np.array_equal(st_or, st_cv)
    assert st_cv[len(pts) - 4] == 0 and st_cv[5] == 0
    ok = st_cv == 1
    assert np.max(np.abs(p_or[ok] - p_cv[ok])) < 1e-2
