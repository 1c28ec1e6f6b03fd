"""TEST INFRASTRUCTURE (CPU oracle) -- multi-view DLT triangulation, a restatement of
pvio/src/pvio/geometry/stereo.h:67-75 (triangulate_point, N views) and :104-128 (triangulate_point_scored)
as Track::triangulate uses them (pvio/src/pvio/map/track.cpp:83-106).  NumPy fp64, SVD like the reference
(Eigen JacobiSVD).  Parity unpinned in the reference (no test); the known-answer check in
tests/test_tri_oracle.py is exact recovery of noise-free points."""
import numpy as np

from . import lie


def projection_matrix(q_wc, p_wc):
    """track.cpp:90-95: P = [R | T], R = q^-1, T = -R p for the CAMERA pose (q, p) in the world."""
    R = lie.qmat(lie.qconj(q_wc))
    return np.c_[R, -R @ np.asarray(p_wc, dtype=np.float64)]


def triangulate_scored(Ps, points):
    """stereo.h:104-128.  Returns (has_parallax, p[3], score)."""
    A = np.zeros((2 * len(points), 4))
    for i, (P, z) in enumerate(zip(Ps, points)):              # stereo.h:69-72
        A[2 * i] = z[0] * P[2] - P[0]
        A[2 * i + 1] = z[1] * P[2] - P[1]
    q = np.linalg.svd(A)[2][3]                                # matrixV().col(3)
    ok, score = True, 0.0
    for P, z in zip(Ps, points):
        qi = P @ q
        if not (qi[2] * q[3] > 0):
            ok = False
        with np.errstate(divide='ignore', invalid='ignore'):
            if not (qi[2] / q[3] < 100):
                ok = False
            score += np.sum((qi[:2] / qi[2] - z) ** 2)
    score /= len(points)
    with np.errstate(divide='ignore', invalid='ignore'):
        p = q[:3] / q[3] if ok else q[:3] / np.linalg.norm(q[:3])
    return ok, p, score
