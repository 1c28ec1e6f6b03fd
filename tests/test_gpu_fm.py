"""GPU parity of the F-matrix outlier rejection (pvio_b200_find_fundamental_mask / pvio_b200_track_keypoints,
opencv_image.cpp:88-136) through the C-ABI: against the cv2-pinned oracle (same winning iteration / model / iteration
count, masks equal) and against cv2.findFundamentalMat itself, with OpenCV's own sample schedule and with injected ones;
the LMedS branch OpenCV takes below 15 matches; the whole track_keypoints against the cv2 pipeline of the reference."""
import os

import numpy as np
import pytest

from oracle import fm_oracle as fo
from pvio_b200 import klt
from pvio_b200.bundle_adjustor import BundleAdjustor
from synthetic import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fm_golden.npz")
SCENES = [(150, 0.1), (300, 0.3), (400, 0.5), (60, 0.2), (15, 0.1), (20, 0.0), (200, 0.7), (1000, 0.4)]


@pytest.fixture(scope="module")
def ba():
    b = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=16, max_obs=64)
    yield b
    b.close()


def _assert_masks(p, q, m_dev, m_ref, F_ref, thr2=1.0):
    """Equal, except matches whose error lies within 1e-6 relative of the threshold in the oracle (rounding knife edge)."""
    if np.array_equal(m_dev, m_ref):
        return
    err = fo.compute_error(p, q, F_ref).astype(np.float64)
    bad = np.nonzero(m_dev != m_ref)[0]
    assert np.all(np.abs(err[bad] - thr2) < 1e-6 * thr2), (bad, err[bad])


def test_fm_golden_cv2_vectors(ba):
    g = np.load(GOLD)
    for i in range(int(g["count"])):
        m = klt.find_fundamental_mask(ba, g[f"p{i}"], g[f"q{i}"])[0]
        assert np.array_equal(m, g[f"cv_mask{i}"]), i


def test_fm_ransac_matches_oracle_and_cv2(ba):
    cv2 = pytest.importorskip("cv2")
    for seed in range(8):
        for n, of in SCENES:
            p, q = synth.make_fm_matches(seed, n, of, planar=(seed % 7 == 3))
            m, F, info = klt.find_fundamental_mask(ba, p, q, return_info=True)
            mo, Fo, io = fo.find_fundamental_mask(p, q, null_space=fo.null_space_householder, return_info=True)
            assert info == io, (seed, n, of)
            _assert_masks(p, q, m, mo, Fo)
            assert np.allclose(F, Fo, rtol=1e-6, atol=1e-9 * np.abs(Fo).max())
            _, mc = cv2.findFundamentalMat(p, q, cv2.FM_RANSAC, 1.0, 0.99)
            _assert_masks(p, q, m, mc.ravel(), Fo)


def test_fm_injected_schedule(ba):
    r = np.random.default_rng(1)
    for seed, n, of in [(11, 300, 0.3), (12, 80, 0.5), (13, 500, 0.1)]:
        p, q = synth.make_fm_matches(seed, n, of)
        sch = np.stack([r.choice(n, 7, replace=False) for _ in range(1000)]).astype(np.int32)
        m, F, info = klt.find_fundamental_mask(ba, p, q, schedule=sch, return_info=True)
        mo, Fo, io = fo.find_fundamental_mask(p, q, schedule=sch, null_space=fo.null_space_householder, return_info=True)
        assert info == io
        _assert_masks(p, q, m, mo, Fo)
        # a short schedule ends the run where it ends; a row of -1 is getSubset giving up
        sch2 = sch[:40].copy()
        sch2[25:] = -1
        m2, _, info2 = klt.find_fundamental_mask(ba, p, q, schedule=sch2, return_info=True)
        mo2, Fo2, io2 = fo.find_fundamental_mask(p, q, schedule=sch2, null_space=fo.null_space_householder, return_info=True)
        assert info2 == io2 and info2["iterations"] <= 25
        _assert_masks(p, q, m2, mo2, Fo2)
    bad = sch.copy()
    bad[3, 2] = 10 ** 6
    with pytest.raises(RuntimeError):
        klt.find_fundamental_mask(ba, p, q, schedule=bad)


def test_fm_other_threshold_and_confidence(ba):
    cv2 = pytest.importorskip("cv2")
    p, q = synth.make_fm_matches(21, 250, 0.3, noise=0.6)
    for thr, conf in [(0.5, 0.99), (2.0, 0.999), (3.0, 0.9)]:
        m, F, info = klt.find_fundamental_mask(ba, p, q, threshold=thr, confidence=conf, return_info=True)
        mo, Fo, io = fo.find_fundamental_mask(p, q, thr, conf, null_space=fo.null_space_householder, return_info=True)
        assert info == io
        _assert_masks(p, q, m, mo, Fo, thr * thr)
        _, mc = cv2.findFundamentalMat(p, q, cv2.FM_RANSAC, thr, conf)
        _assert_masks(p, q, m, mc.ravel(), Fo, thr * thr)


def test_fm_lmeds_branch_and_small_inputs(ba):
    cv2 = pytest.importorskip("cv2")
    for seed in range(10):                       # n == 14: pinned against cv2 (see tests/test_fm_oracle.py)
        p, q = synth.make_fm_matches(2000 + seed, 14, 0.15 if seed % 2 else 0.3)
        m, F, info = klt.find_fundamental_mask(ba, p, q, return_info=True)
        mo, Fo, io = fo.find_fundamental_mask(p, q, null_space=fo.null_space_householder, return_info=True)
        assert info == io and info["method"] == "lmeds" and info["iterations"] == 300
        assert np.array_equal(m, mo)
        _, mc = cv2.findFundamentalMat(p, q, cv2.FM_RANSAC, 1.0, 0.99)
        assert np.array_equal(m, mc.ravel())
    for n in range(8, 14):                       # rounding noise picks the model inside OpenCV itself: structure only
        p, q = synth.make_fm_matches(3000 + n, n, 0.1)
        m, F, info = klt.find_fundamental_mask(ba, p, q, return_info=True)
        assert info["method"] == "lmeds" and info["iterations"] == 300 and info["best_iteration"] >= 0
        err = fo.compute_error(p, q, F)
        med = float(np.sort(err)[n // 2])
        sigma = max(2.5 * 1.4826 * (1 + 5.0 / (n - 7)) * np.sqrt(med), 0.001)
        assert np.array_equal(m, (err <= np.float32(sigma * sigma)).astype(np.uint8))
        assert m.sum() >= 7
    p, q = synth.make_fm_matches(1, 7, 0.0)
    m, F, info = klt.find_fundamental_mask(ba, p, q, return_info=True)
    assert m.tolist() == [1] * 7 and info["method"] == "7point"
    Fo = fo.run_7point(p, q, fo.null_space_householder)[0]
    assert np.allclose(F, Fo, rtol=1e-6, atol=1e-9 * np.abs(Fo).max())
    m, F = klt.find_fundamental_mask(ba, p[:5], q[:5])
    assert m.tolist() == [0] * 5
    m, F = klt.find_fundamental_mask(ba, p[:0], q[:0])
    assert len(m) == 0


def test_track_keypoints_whole_call_matches_the_cv2_pipeline(ba):
    """OpenCvImage::track_keypoints as the reference runs it (opencv_image.cpp:88-136) with cv2: LK, border, F-RANSAC."""
    cv2 = pytest.importorskip("cv2")
    for seed, size, n in [(648, (752, 480), 400), (7, (512, 512), 150)]:
        prev, nxt_img, pts, _ = synth.make_klt_pair(seed=seed, size=size, n_points=n)
        r = np.random.default_rng(seed)
        init = pts.copy()
        bad = r.choice(n, n // 10, replace=False)          # a tenth of the guesses start 6 px off: LK locks onto other texture
        init[bad] += r.uniform(-6, 6, (len(bad), 2)).astype(np.float32)
        nx, st = klt.track_keypoints_ransac(ba, prev, nxt_img, pts, init, prev_id=10 * seed + 1, next_id=10 * seed + 2)
        p1, s1, _ = cv2.calcOpticalFlowPyrLK(prev, nxt_img, pts.reshape(-1, 1, 2).copy(), init.reshape(-1, 1, 2).copy(),
                                             winSize=(21, 21), maxLevel=3,
                                             criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                             flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        p1, s1 = p1.reshape(-1, 2), s1.ravel().copy()
        w, h = size
        s1[(p1[:, 0] < 20) | (p1[:, 0] >= w - 20) | (p1[:, 1] < 20) | (p1[:, 1] >= h - 20)] = 0
        # the F-matrix stage of the reference on OUR tracked positions (LK positions agree to 1e-2 px, not bit for bit)
        st_lk = klt.track_keypoints(ba, prev, nxt_img, pts, init, prev_id=10 * seed + 1, next_id=10 * seed + 2)[1]
        assert np.array_equal(st_lk, s1)
        l = np.nonzero(st_lk)[0]
        _, mc = cv2.findFundamentalMat(pts[l], nx[l], cv2.FM_RANSAC, 1.0, 0.99)
        expect = st_lk.copy()
        expect[l[mc.ravel() == 0]] = 0
        assert np.array_equal(st, expect)
        assert 0 < expect.sum() <= st_lk.sum()
