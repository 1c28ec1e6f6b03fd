"""Host mirror of OpenCvImage::detect_keypoints (pvio-extra/src/pvio/extra/opencv_image.cpp:54-86) over the C-ABI:
Harris-GFTT on the device, PVIO's Poisson-disk filter against the frame's existing keypoints, 20-pixel border."""
import ctypes as C

import numpy as np

from . import _lib


def detect_keypoints(ba, image, keypoints=(), keypoint_distance=25.0, clahe_clip=0.0, clahe_tiles=(8, 8), frame_id=0,
                     shape=None, want_gftt=False, max_out=2048):
    """ba: a BundleAdjustor (owns the device handle).  image: uint8 [h, w] (None if frame_id is in the KLT pyramid
    cache; `shape` then names the frame size).  keypoints: the frame's existing keypoints [n, 2] (pixels).
    Returns the NEW keypoints [m, 2] in the reference's order (and the GFTT corners before the filter if want_gftt)."""
    img = None if image is None else np.ascontiguousarray(image, dtype=np.uint8)
    h, w = img.shape if img is not None else shape
    ex = np.ascontiguousarray(keypoints, dtype=np.float64).reshape(-1, 2)
    out = np.zeros((max_out, 2))
    gftt = np.zeros((1000, 2), dtype=np.float32)
    n, ng = C.c_int(), C.c_int()
    fn = ba.lib.pvio_b200_detect_keypoints
    fn.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                   C.POINTER(C.c_double), C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int),
                   C.POINTER(C.c_float), C.POINTER(C.c_int)]
    ba._ck(fn(ba.h, int(frame_id), None if img is None else _lib._ptr(img, C.c_uint8), w, h, w, float(clahe_clip),
              clahe_tiles[0], clahe_tiles[1], _lib._ptr(ex, C.c_double) if len(ex) else None, len(ex), float(keypoint_distance),
              max_out, _lib._ptr(out, C.c_double), C.byref(n), _lib._ptr(gftt, C.c_float), C.byref(ng)))
    return (out[:n.value].copy(), gftt[:ng.value].copy()) if want_gftt else out[:n.value].copy()
