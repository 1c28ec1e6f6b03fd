"""OpenCvImage::detect_keypoints on the device (csrc/detect.cu) against cv2 + the restated Poisson-disk filter."""
import numpy as np
import pytest

from oracle import detect_oracle as D
from synthetic import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ba():
    from pvio_b200.bundle_adjustor import BundleAdjustor
    b = BundleAdjustor(max_windows=1, max_frames=4, max_landmarks=8, max_obs=16)
    yield b
    b.close()


def _frames():
    import cv2
    out = []
    for seed in (648, 649, 650):
        prev, nxt, pts, _ = synth.make_klt_pair(seed=seed)
        for im in (prev, nxt):
            out.append(cv2.createCLAHE(6.0, (8, 8)).apply(im))
    return out


def test_gftt_corner_set_equals_opencv(ba):
    """Parity of the detector proper: the SET of corners cv2.goodFeaturesToTrack(useHarrisDetector=True) selects
    (the responses agree to rounding, not bit for bit: see csrc/detect.cu)."""
    from pvio_b200.detect import detect_keypoints
    for img in _frames():
        ref = D.gftt(img)
        _, got = detect_keypoints(ba, img, want_gftt=True)
        a, b = set(map(tuple, ref.astype(int))), set(map(tuple, got.astype(int)))
        assert len(ref) > 300 and a == b, (len(a), len(b), len(a & b))
        same_order = np.array_equal(ref.astype(int), got.astype(int))
        print("corners", len(ref), "identical order:", same_order)


def test_detect_keypoints_matches_the_reference_pipeline(ba):
    """Whole call: GFTT -> Poisson-disk filter against the frame's existing keypoints -> 20-px border (indices exact)."""
    from pvio_b200.detect import detect_keypoints
    rng = np.random.default_rng(3)
    for img in _frames()[:3]:
        existing = rng.uniform(30, 400, size=(120, 2))
        got, corners = detect_keypoints(ba, img, existing, keypoint_distance=25.0, want_gftt=True)
        ref = D.detect_keypoints(img, existing, 25.0, corners=corners)            # same corners: the sequential part exactly
        assert np.array_equal(got, ref)
        ref_cv = D.detect_keypoints(img, existing, 25.0)                          # corners from cv2: equal as sets
        assert set(map(tuple, got)) == set(map(tuple, ref_cv))
        assert len(got) > 50


def test_detect_on_the_frame_cached_by_the_tracker_with_device_clahe(ba):
    """The frame the tracker uploaded (raw, equalised on the device) is detected on without another upload."""
    import cv2
    from pvio_b200 import klt
    from pvio_b200.detect import detect_keypoints
    prev, nxt, pts, _ = synth.make_klt_pair()
    klt.track_keypoints(ba, prev, nxt, pts, clahe_clip=6.0, prev_id=101, next_id=102)
    have = pts[::4]                                                                # a quarter of the grid is occupied
    got = detect_keypoints(ba, None, have, keypoint_distance=25.0, clahe_clip=6.0, frame_id=102, shape=nxt.shape)
    ref = D.detect_keypoints(cv2.createCLAHE(6.0, (8, 8)).apply(nxt), have, 25.0)
    assert set(map(tuple, got)) == set(map(tuple, ref)) and len(got) > 20
    raw = detect_keypoints(ba, nxt, have, keypoint_distance=25.0, clahe_clip=6.0)  # not cached: upload + device CLAHE
    assert np.array_equal(raw, got)
