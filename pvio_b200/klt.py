"""Host mirror of OpenCvImage::track_keypoints (pvio-extra/src/pvio/extra/opencv_image.cpp:88-136)
over the C-ABI: pyramidal LK on the GPU (pvio_b200_klt_track) followed by the reference's 20-px
border rejection (:106).  The F-matrix RANSAC of :121-129 is a SURVEY 8(f) "next" row and is not
part of this call."""
import ctypes as C

import numpy as np

from . import _lib


def track_keypoints(ba, prev_img, next_img, curr_keypoints, next_keypoints=None, max_level=3, max_iter=30,
                    eps=0.01, border=20, raw=False):
    """ba: a BundleAdjustor (owns the device handle).  Images: uint8 [h, w].  Keypoints in pixels.
    Returns (next_keypoints float32 [n,2], status uint8 [n], err float32 [n])."""
    prev_img = np.ascontiguousarray(prev_img, dtype=np.uint8)
    next_img = np.ascontiguousarray(next_img, dtype=np.uint8)
    h, w = prev_img.shape
    cur = np.ascontiguousarray(curr_keypoints, dtype=np.float32).reshape(-1, 2)
    nxt = cur.copy() if next_keypoints is None or len(next_keypoints) == 0 else \
        np.array(next_keypoints, dtype=np.float32, copy=True).reshape(-1, 2)
    n = len(cur)
    status = np.zeros(max(n, 1), dtype=np.uint8)
    err = np.zeros(max(n, 1), dtype=np.float32)
    rc = ba.lib.pvio_b200_klt_track(ba.h, _lib._ptr(prev_img, C.c_uint8), _lib._ptr(next_img, C.c_uint8), w, h, w,
                                    _lib._ptr(cur, C.c_float), _lib._ptr(nxt, C.c_float), _lib._ptr(status, C.c_uint8),
                                    _lib._ptr(err, C.c_float), n, max_level, max_iter, eps)
    ba._ck(rc)
    status, err = status[:n], err[:n]
    if not raw:
        out = (nxt[:, 0] < border) | (nxt[:, 0] >= w - border) | (nxt[:, 1] < border) | (nxt[:, 1] >= h - border)
        status = np.where(out, 0, status).astype(np.uint8)      # opencv_image.cpp:106-108
    return nxt, status, err
