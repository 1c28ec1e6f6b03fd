"""Generates tests/golden/klt_golden.npz: a small image pair, keypoints and the output of the
reference's KLT implementation (cv2.calcOpticalFlowPyrLK with PVIO's arguments,
pvio-extra/src/pvio/extra/opencv_image.cpp:103).  Run in the build container (cv2 4.13.0)."""
import os, sys
import numpy as np
import cv2
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from synthetic import synth

prev, nxt, pts, truth = synth.make_klt_pair(seed=648, size=(256, 192), n_points=48)
prev = prev.copy(); nxt = nxt.copy()
prev[80:112, 120:152] = 128; nxt[80:112, 120:152] = 128     # flat patch -> min-eigenvalue rejection
pts = np.concatenate([pts, np.array([[136.0, 96.0], [3.0, 4.0], [252.0, 188.0], [22.0, 21.0]], dtype=np.float32)])
init = pts.copy(); init[5] = [900.0, 50.0]
p1, st, err = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.reshape(-1, 1, 2).copy(), init.reshape(-1, 1, 2).copy(),
                                       winSize=(21, 21), maxLevel=3,
                                       criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01),
                                       flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "klt_golden.npz"), prev=prev, next=nxt,
                    pts=pts, init=init, cv_next=p1.reshape(-1, 2), cv_status=st.ravel(), cv_err=err.ravel(),
                    cv_version=cv2.__version__)
print("status ok:", int(st.sum()), "of", len(pts))
