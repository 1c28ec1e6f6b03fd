"""TEST INFRASTRUCTURE ONLY.  CPU statement of OpenCvImage::detect_keypoints
(/root/reference/pvio-extra/src/pvio/extra/opencv_image.cpp:54-86):
  * the corner detector IS the reference's own dependency, cv2 (GFTTDetector::create(1000, 1e-3, 20, 3, true), :183 --
    the same parameters through cv2.goodFeaturesToTrack, whose order is decreasing response, which makes the std::sort
    of :63-65 a no-op apart from exact ties);
  * PoissonDiskFilter<2> is restated from /root/reference/pvio/src/pvio/utility/poisson_disk_filter.h:32-113 including
    the cell walk of test_point (:92-107), which advances BEFORE it looks: the first cell of the 5 x 5 block is never
    examined, one cell past its end is.  The restatement is pinned by a brute-force property in
    tests/test_detect_oracle.py (with at most one preset point per cell the filter equals "accept iff no accepted /
    preset point is closer than the radius": the skipped cell lies a full cell away in both directions, i.e. >= radius).
  * :71-79: 20-pixel border applied AFTER the filter (rejected border points still block their neighbourhood)."""
import math

import numpy as np


def gftt(image):
    import cv2
    c = cv2.goodFeaturesToTrack(image, 1000, 1.0e-3, 20, blockSize=3, useHarrisDetector=True, k=0.04)
    return np.zeros((0, 2), dtype=np.float32) if c is None else c.reshape(-1, 2)


class PoissonDiskFilter:
    def __init__(self, radius):
        self.r2 = radius * radius
        self.gs = radius / math.sqrt(2.0)
        self.span = int(math.ceil(math.sqrt(2.0)))
        self.points, self.grid = [], {}

    def index(self, p):
        return (int(math.floor(p[0] / self.gs)), int(math.floor(p[1] / self.gs)))

    def preset(self, p):
        self.grid[self.index(p)] = len(self.points)        # :44-48: the cell remembers the LAST point put into it
        self.points.append((float(p[0]), float(p[1])))

    def test(self, p):
        ix, iy = self.index(p)
        bx, by, ex, ey = ix - self.span, iy - self.span, ix + self.span, iy + self.span
        cx, cy = bx, by
        while cy <= ey:                                     # :98-106
            cx += 1
            if cx > ex:
                cx, cy = bx, cy + 1
            j = self.grid.get((cx, cy))
            if j is not None:
                q = self.points[j]
                if (p[0] - q[0]) ** 2 + (p[1] - q[1]) ** 2 < self.r2:
                    return False
        return True

    def insert(self, p):
        if not self.test(p):
            return False
        self.preset(p)
        return True


def detect_keypoints(image, existing, keypoint_distance, corners=None):
    """image: the PREPROCESSED (CLAHE) frame.  Returns the new keypoints [m, 2] float64 in the reference's order."""
    corners = gftt(image) if corners is None else corners
    f = PoissonDiskFilter(keypoint_distance)
    for p in np.asarray(existing, dtype=np.float64).reshape(-1, 2):
        f.preset(p)
    h, w = image.shape
    out = []
    for c in corners:
        p = (float(c[0]), float(c[1]))
        if f.insert(p) and not (p[0] < 20 or p[1] < 20 or p[0] >= w - 20 or p[1] >= h - 20):
            out.append(p)
    return np.array(out, dtype=np.float64).reshape(-1, 2)
