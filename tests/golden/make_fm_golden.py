"""Writes tests/golden/fm_golden.npz: cv2.findFundamentalMat(FM_RANSAC, 1.0, 0.99) masks (the reference's call,
opencv_image.cpp:123) on seeded scenes.  Run here (cv2 4.13.0): python tests/golden/make_fm_golden.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from synthetic import synth  # noqa: E402

out = {}
cases = [(100, 300, 0.2, False), (101, 400, 0.5, False), (102, 150, 0.1, True), (103, 40, 0.25, False), (104, 15, 0.1, False),
         (105, 14, 0.2, False), (106, 250, 0.65, False), (107, 500, 0.35, False)]
for i, (seed, n, of, planar) in enumerate(cases):
    p, q = synth.make_fm_matches(seed, n, of, planar=planar)
    F, mask = cv2.findFundamentalMat(p, q, cv2.FM_RANSAC, 1.0, 0.99)
    out[f"p{i}"], out[f"q{i}"], out[f"cv_mask{i}"], out[f"cv_F{i}"] = p, q, mask.ravel(), F
out["count"] = np.int32(len(cases))
out["cv_version"] = np.array(cv2.__version__)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fm_golden.npz"), **out)
print("wrote", len(cases), "cases, cv2", cv2.__version__)
