"""TEST INFRASTRUCTURE (CPU oracle) -- CLAHE as pvio-extra applies it before tracking
(pvio-extra/src/pvio/extra/opencv_image.cpp:138-143,178: cv::createCLAHE(6.0, Size(8, 8))->apply on the 8-bit
level-0 image).  The algorithm lives in OpenCV (imgproc/src/clahe.cpp, unpinned version, 4.13.0 in this image);
this file restates it in NumPy with OpenCV's float32 operation order and is PINNED against cv2 bit for bit
(tests/test_clahe_oracle.py).  Only tests/, smoke() and bench.py's cpu_baseline leg may import it.
Restricted, like the kernel, to image sizes that are multiples of the tile grid (752x480 and 512x512 of the
reference's datasets are): OpenCV pads other sizes with BORDER_REFLECT_101 first."""
import numpy as np


def clahe(img, clip_limit=6.0, tiles=(8, 8)):
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    tx, ty = tiles
    assert W % tx == 0 and H % ty == 0, "image size must be a multiple of the tile grid"
    tw, th = W // tx, H // ty
    area = tw * th
    clip = max(int(clip_limit * area / 256), 1)                     # clahe.cpp: clipLimit scaled to the tile, >= 1
    lut_scale = np.float32(255.0) / np.float32(area)
    luts = np.zeros((ty, tx, 256), dtype=np.uint8)
    for j in range(ty):
        for i in range(tx):
            h = np.bincount(img[j * th:(j + 1) * th, i * tw:(i + 1) * tw].ravel(), minlength=256).astype(np.int64)
            clipped = int(np.sum(np.maximum(h - clip, 0)))
            h = np.minimum(h, clip)
            batch, residual = clipped // 256, clipped % 256
            h += batch
            if residual:
                step = max(256 // residual, 1)
                idx = np.arange(0, 256, step)[:residual]
                h[idx] += 1
            cs = np.cumsum(h).astype(np.float32) * lut_scale         # float32 product, then cvRound (half to even)
            luts[j, i] = np.clip(np.rint(cs), 0, 255).astype(np.uint8)
    inv_tw, inv_th = np.float32(1.0) / np.float32(tw), np.float32(1.0) / np.float32(th)
    xs = np.arange(W, dtype=np.float32) * inv_tw - np.float32(0.5)
    x1 = np.floor(xs).astype(np.int32)
    xa = (xs - x1.astype(np.float32)).astype(np.float32)
    xa1 = np.float32(1.0) - xa
    x2 = np.minimum(x1 + 1, tx - 1)
    x1 = np.maximum(x1, 0)
    ys = np.arange(H, dtype=np.float32) * inv_th - np.float32(0.5)
    y1 = np.floor(ys).astype(np.int32)
    ya = (ys - y1.astype(np.float32)).astype(np.float32)
    ya1 = np.float32(1.0) - ya
    y2 = np.minimum(y1 + 1, ty - 1)
    y1 = np.maximum(y1, 0)
    v = img.astype(np.int64)
    Y1, Y2 = y1[:, None], y2[:, None]
    X1, X2 = x1[None, :], x2[None, :]
    l11 = luts[Y1, X1, v].astype(np.float32); l12 = luts[Y1, X2, v].astype(np.float32)
    l21 = luts[Y2, X1, v].astype(np.float32); l22 = luts[Y2, X2, v].astype(np.float32)
    top = (l11 * xa1[None, :] + l12 * xa[None, :]).astype(np.float32)
    bot = (l21 * xa1[None, :] + l22 * xa[None, :]).astype(np.float32)
    res = (top * ya1[:, None] + bot * ya[:, None]).astype(np.float32)
    return np.clip(np.rint(res), 0, 255).astype(np.uint8)
