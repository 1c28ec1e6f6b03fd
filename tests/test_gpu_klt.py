"""GPU KLT parity: pvio_b200_klt_track vs (a) committed golden vectors produced by the reference's
KLT implementation (cv2) and (b) cv2 / the NumPy oracle on seeded pairs.  Status flags (track
indices) must be bit-exact; positions within 1e-2 px (OpenCV accumulates the 2x2 system in
float SIMD order, the kernel in exact integers)."""
import os
import numpy as np
import pytest

from oracle import klt_oracle as ko
from pvio_b200 import klt
from synthetic import synth
from pvio_b200.bundle_adjustor import BundleAdjustor

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "klt_golden.npz")


@pytest.fixture(scope="module")
def ba():
    b = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=16, max_obs=64)
    yield b
    b.close()


def test_klt_golden(ba):
    g = np.load(GOLD)
    nxt, st, err = klt.track_keypoints(ba, g["prev"], g["next"], g["pts"], g["init"], raw=True)
    assert np.array_equal(st, g["cv_status"])
    ok = st == 1
    assert np.max(np.abs(nxt[ok] - g["cv_next"][ok])) < 1e-2
    assert np.max(np.abs(err[ok] - g["cv_err"][ok])) < 1e-2


@pytest.mark.parametrize("size,n", [((320, 240), 70), ((752, 480), 500)])
def test_klt_vs_oracle_and_cv2(ba, size, n):
    prev, nxt_img, pts, truth = synth.make_klt_pair(size=size, n_points=n)
    nxt, st, err = klt.track_keypoints(ba, prev, nxt_img, pts, raw=True)
    if n <= 100:
        p_or, st_or, err_or = ko.calc_optical_flow_pyr_lk(prev, nxt_img, pts, pts)
        assert np.array_equal(st, st_or)
        assert np.max(np.abs(nxt[st == 1] - p_or[st == 1])) < 2e-3
    cv2 = pytest.importorskip("cv2")
    p1, s1, e1 = cv2.calcOpticalFlowPyrLK(prev, nxt_img, pts.reshape(-1, 1, 2).copy(), pts.reshape(-1, 1, 2).copy(),
                                          winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                          flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    assert np.array_equal(st, s1.ravel())
    ok = st == 1
    assert np.max(np.abs(nxt[ok] - p1.reshape(-1, 2)[ok])) < 1e-2
    assert np.median(np.linalg.norm(nxt[ok] - truth[ok], axis=1)) < 0.1


def test_klt_border_rule_and_empty(ba):
    prev, nxt_img, pts, _ = synth.make_klt_pair(size=(320, 240), n_points=30)
    pts = np.concatenate([pts, np.array([[10.0, 100.0]], dtype=np.float32)])
    nxt, st, _ = klt.track_keypoints(ba, prev, nxt_img, pts)
    assert st[-1] == 0                      # inside 20 px of the border: opencv_image.cpp:106
    nxt0, st0, _ = klt.track_keypoints(ba, prev, nxt_img, np.zeros((0, 2), dtype=np.float32))
    assert len(nxt0) == 0 and len(st0) == 0


def test_clahe_bit_exact_with_oracle_and_cv2(ba):
    """Device CLAHE (opencv_image.cpp:138-143) against the cv2-pinned oracle and cv2 itself: every pixel equal."""
    from oracle import clahe_oracle
    rng = np.random.default_rng(5)
    for shape in [(480, 752), (512, 512), (64, 96)]:
        img = np.clip(np.linspace(20, 230, shape[1])[None, :] + rng.normal(0, 25, shape), 0, 255).astype(np.uint8)
        out = klt.clahe(ba, img)
        assert np.array_equal(out, clahe_oracle.clahe(img))
        try:
            import cv2
            assert np.array_equal(out, cv2.createCLAHE(6.0, (8, 8)).apply(img))
        except ImportError:
            pass
    flat = np.full((64, 64), 200, dtype=np.uint8)
    assert np.array_equal(klt.clahe(ba, flat), clahe_oracle.clahe(flat))


def test_track_from_raw_frames_equals_track_of_equalised_frames(ba):
    """pvio_b200_klt_track_raw = CLAHE on the device + the same tracker: identical to tracking host-equalised frames."""
    from oracle import clahe_oracle
    prev, nxt, pts, _ = synth.make_klt_pair(size=(752, 480), n_points=300)
    a = klt.track_keypoints(ba, prev, nxt, pts, clahe_clip=6.0)
    b = klt.track_keypoints(ba, clahe_oracle.clahe(prev), clahe_oracle.clahe(nxt), pts)
    assert np.array_equal(a[1], b[1]) and np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def test_klt_pyramid_cache_matches_uncached_calls(ba):
    """pvio_b200_klt_track_cached keeps the finished pyramids of the last two frames (prev of a call = next of the call
    before, core/feature_tracker.cpp:92): a three-frame sequence through the cache -- the second call passes no pixels
    for its prev frame -- gives bit for bit what the self-contained calls give, border rule included."""
    a, b, pts, _ = synth.make_klt_pair(size=(320, 240), n_points=60, seed=5)
    _, c, _, _ = synth.make_klt_pair(size=(320, 240), n_points=60, seed=6)
    ref1 = klt.track_keypoints(ba, a, b, pts)
    ref2 = klt.track_keypoints(ba, b, c, ref1[0])
    l0 = ba.kernel_launches
    got1 = klt.track_keypoints(ba, a, b, pts, prev_id=101, next_id=102)
    l1 = ba.kernel_launches
    got2 = klt.track_keypoints(ba, None, c, got1[0], prev_id=102, next_id=103)      # frame 102 is cached: no pixels needed
    l2 = ba.kernel_launches
    for r, g in ((ref1, got1), (ref2, got2)):
        assert np.array_equal(r[1], g[1]) and np.array_equal(r[0], g[0]) and np.array_equal(r[2], g[2])
    assert (l2 - l1) < (l1 - l0)            # the second call built one pyramid, not two
    # CLAHE on the device goes through the same cache
    r3 = klt.track_keypoints(ba, a, b, pts, clahe_clip=6.0)
    g3 = klt.track_keypoints(ba, a, b, pts, clahe_clip=6.0, prev_id=7, next_id=8)
    g4 = klt.track_keypoints(ba, None, None, pts, clahe_clip=6.0, prev_id=7, next_id=8, shape=a.shape)      # both cached
    assert np.array_equal(r3[0], g3[0]) and np.array_equal(r3[1], g3[1]) and np.array_equal(g3[0], g4[0])


def test_klt_small_image_pyramid_depth_matches_opencv(ba):
    """cv::buildOpticalFlowPyramid halves first and drops the new level when it is not larger than the window: a
    160 x 120 frame with maxLevel 3 has levels 0..2 (80 x 60, 40 x 30; 20 x 15 is dropped).  One level too many changes
    the coarse initialisation and with it positions and status flags."""
    cv2 = pytest.importorskip("cv2")
    prev, nxt_img, pts, _ = synth.make_klt_pair(size=(160, 120), n_points=25, max_shift=6.0, seed=9)
    nxt, st, err = klt.track_keypoints(ba, prev, nxt_img, pts, raw=True)
    p1, s1, e1 = cv2.calcOpticalFlowPyrLK(prev, nxt_img, pts.reshape(-1, 1, 2).copy(), pts.reshape(-1, 1, 2).copy(),
                                          winSize=(21, 21), maxLevel=3,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                          flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    assert np.array_equal(st, s1.ravel())
    ok = st == 1
    assert ok.sum() >= 3 and np.max(np.abs(nxt[ok] - p1.reshape(-1, 2)[ok])) < 1e-2
