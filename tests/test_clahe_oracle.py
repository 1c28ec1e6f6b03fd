"""oracle/clahe_oracle.py pinned against OpenCV (cv2.createCLAHE(6, (8, 8)), opencv_image.cpp:139): bit for bit."""
import numpy as np
import pytest

from oracle import clahe_oracle

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(480, 752), (512, 512), (64, 96)])
def test_clahe_oracle_is_bit_exact_with_cv2(shape):
    rng = np.random.default_rng(shape[0])
    c = cv2.createCLAHE(6.0, (8, 8))
    imgs = [rng.integers(0, 256, shape, dtype=np.uint8),
            np.clip(128 + 400 * cv2.GaussianBlur(rng.standard_normal(shape).astype(np.float32), (0, 0), 5), 0, 255).astype(np.uint8),
            np.clip(np.linspace(0, 255, shape[1])[None, :] + rng.normal(0, 3, shape), 0, 255).astype(np.uint8),
            np.full(shape, 77, dtype=np.uint8)]
    for img in imgs:
        assert np.array_equal(c.apply(img), clahe_oracle.clahe(img))
