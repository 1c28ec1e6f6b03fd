"""TEST INFRASTRUCTURE (oracle): CPU restatement of the F-matrix outlier rejection of
OpenCvImage::track_keypoints (pvio-extra/src/pvio/extra/opencv_image.cpp:121-129):

    findFundamentalMat(p, q, cv::FM_RANSAC, 1.0, 0.99, mask)

The algorithm lives in a third-party dependency that is NOT in /root/reference: OpenCV (calib3d:
fundam.cpp `findFundamentalMat`, `FMEstimatorCallback`, `run7Point`; ptsetreg.cpp
`RANSACPointSetRegistrator::run`, `LMeDSPointSetRegistrator::run`, `getSubset`, `RANSACUpdateNumIters`;
core: `cv::RNG`, `cv::solveCubic`).  The installed build is cv2 4.13.0; this file restates the published
algorithm and is PINNED against `cv2.findFundamentalMat` itself (tests/test_fm_oracle.py: inlier masks equal
on every seeded scene, iteration for iteration because the sample schedule of `cv::RNG((uint64)-1)` is
restated too).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

What OpenCV does, as restated here:
  * n < 7: nothing.  n == 7: the 7-point models, mask all ones.
  * n >= 15: RANSAC, model size 7, up to 1000 iterations, each: draw 7 distinct indices with RNG.uniform
    (redraw a duplicate), reject the subset if its LAST point is collinear with two earlier ones in either image
    (and redraw, up to 10000 attempts), 7-point solve (1..3 models, in solveCubic's root order), inliers =
    (float) max(d1^2 s1, d2^2 s2) <= (float) threshold^2; a model with MORE inliers than the best so far (and >= 7)
    replaces it and shrinks the iteration count to RANSACUpdateNumIters(confidence, outlier ratio, 7, niters).
    No refit on the inliers: the mask of the best model is the result.
  * 8 <= n < 15 (track_keypoints calls from 8 survivors up): LMedS with 300 iterations (outlier ratio 0.45),
    score = the median of the errors (element count/2 of the sorted errors), then the inliers of the best model
    under sigma = 2.5 * 1.4826 * (1 + 5 / (n - 7)) * sqrt(median), at least 0.001.
The null space of the 7 x 9 system is taken from numpy's SVD; OpenCV's Jacobi SVD returns another basis of the same
plane, so the cubic and its roots differ but the set of models does not (the model ORDER inside one sample may, which
only matters when two models of one sample tie for a new best count)."""
import math

import numpy as np

FLT_EPSILON = float(np.finfo(np.float32).eps)
DBL_EPSILON = float(np.finfo(np.float64).eps)
DBL_MIN = float(np.finfo(np.float64).tiny)
RNG_COEFF = 4164903690


class CvRNG:
    """cv::RNG: multiply-with-carry, `state = (uint32) state * 4164903690 + (state >> 32)`; uniform(a, b) = next() % (b - a) + a."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def have_collinear_points(pts, count):
    """fundam.cpp haveCollinearPoints: only the LAST of `count` points is tested against the pairs before it."""
    i = count - 1
    for j in range(i):
        dx1 = float(pts[j, 0]) - float(pts[i, 0])
        dy1 = float(pts[j, 1]) - float(pts[i, 1])
        for k in range(j):
            dx2 = float(pts[k, 0]) - float(pts[i, 0])
            dy2 = float(pts[k, 1]) - float(pts[i, 1])
            if abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2)):
                return True
    return False


def get_subset(m1, m2, rng, model_points=7, max_attempts=10000):
    """ptsetreg.cpp getSubset.  Returns the index list or None."""
    count = len(m1)
    for _ in range(max_attempts):
        idx = []
        for _i in range(model_points):
            v = rng.uniform(0, count)
            while v in idx:
                v = rng.uniform(0, count)
            idx.append(v)
        if not have_collinear_points(m1[idx], model_points) and not have_collinear_points(m2[idx], model_points):
            return idx
    return None


def sample_schedule(m1, m2, n_iters, max_attempts=10000, seed=0xFFFFFFFFFFFFFFFF):
    """The first n_iters subsets `RANSACPointSetRegistrator::run` / the LMedS run would draw (independent of the
    models: every iteration consumes exactly one getSubset).  int32 [n_iters][7]; a row of -1 = getSubset failed."""
    rng = CvRNG(seed)
    out = np.full((n_iters, 7), -1, np.int32)
    for it in range(n_iters):
        idx = get_subset(m1, m2, rng, 7, max_attempts)
        if idx is None:
            break
        out[it] = idx
    return out


def solve_cubic(c):
    """cv::solveCubic for 4 double coefficients; returns the list of roots in OpenCV's order."""
    a0, a1, a2, a3 = (float(v) for v in c)
    if a0 == 0:
        if a1 == 0:
            if a2 == 0:
                return []
            return [-a3 / a2]
        d = a2 * a2 - 4 * a1 * a3
        if d < 0:
            return []
        d = math.sqrt(d)
        q1 = (-a2 + d) * 0.5
        q2 = (a2 + d) * -0.5
        if abs(q1) > abs(q2):
            x0, x1 = q1 / a1, a3 / q1
        else:
            x0, x1 = q2 / a1, a3 / q2
        return [x0, x1] if d > 0 else [x0]
    a0 = 1.0 / a0
    a1 *= a0
    a2 *= a0
    a3 *= a0
    Q = (a1 * a1 - 3 * a2) * (1.0 / 9)
    R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1.0 / 54)
    Qcubed = Q * Q * Q
    d = Qcubed - R * R
    if d > 0:
        theta = math.acos(R / math.sqrt(Qcubed))
        sqrtQ = math.sqrt(Q)
        t0 = -2 * sqrtQ
        t1 = theta * (1.0 / 3)
        t2 = a1 * (1.0 / 3)
        return [t0 * math.cos(t1) - t2, t0 * math.cos(t1 + (2.0 * math.pi / 3)) - t2,
                t0 * math.cos(t1 + (4.0 * math.pi / 3)) - t2]
    if d == 0:
        if R >= 0:
            x0 = -2 * R ** (1.0 / 3) - a1 / 3
            x1 = R ** (1.0 / 3) - a1 / 3
        else:
            x0 = 2 * (-R) ** (1.0 / 3) - a1 / 3
            x1 = -((-R) ** (1.0 / 3)) - a1 / 3
        return [x0] if x0 == x1 else [x0, x1]
    d = math.sqrt(-d)
    e = (d + abs(R)) ** (1.0 / 3)
    if R > 0:
        e = -e
    return [(e + Q / e) - a1 * (1.0 / 3)]


def null_space_cubic(f1, f2):
    """det(lambda f1d + f2) with f1d = f1 - f2: coefficients c[0..3] of lambda^3..1 (fundam.cpp run7Point)."""
    f1 = f1 - f2
    t0 = f2[4] * f2[8] - f2[5] * f2[7]
    t1 = f2[3] * f2[8] - f2[5] * f2[6]
    t2 = f2[3] * f2[7] - f2[4] * f2[6]
    c = np.zeros(4)
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2
    c[2] = (f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
            f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
            f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
            f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) + f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]))
    t0 = f1[4] * f1[8] - f1[5] * f1[7]
    t1 = f1[3] * f1[8] - f1[5] * f1[6]
    t2 = f1[3] * f1[7] - f1[4] * f1[6]
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2
    c[1] = (f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
            f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
            f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
            f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) + f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]))
    return f1, c


def null_space_svd(A):
    """Basis (f1, f2) of the null space of the 7 x 9 system: the last two right singular vectors."""
    _, _, vt = np.linalg.svd(A, full_matrices=True)
    return vt[7].copy(), vt[8].copy()


def null_space_householder(A):
    """The same plane from a Householder QR of A^T (what the device does: csrc/fmat.cu): the last two columns of Q."""
    R = A.T.astype(np.float64).copy()          # 9 x 7
    vs = []
    for k in range(7):
        x = R[k:, k].copy()
        alpha = -math.copysign(np.linalg.norm(x), x[0] if x[0] != 0 else 1.0)
        v = x.copy()
        v[0] -= alpha
        nv = float(v @ v)
        vs.append((v, nv))
        if nv > 0:
            R[k:, k:] -= np.outer(v, (2.0 / nv) * (v @ R[k:, k:]))
    out = []
    for col in (7, 8):
        e = np.zeros(9)
        e[col] = 1.0
        for k in reversed(range(7)):
            v, nv = vs[k]
            if nv > 0:
                e[k:] -= v * ((2.0 / nv) * (v @ e[k:]))
        out.append(e)
    return out[0], out[1]


def run_7point(m1, m2, null_space=null_space_svd):
    """fundam.cpp run7Point (4.x: Hartley normalisation of the 7 points first).  m1, m2: float32 [7][2].
    Returns the list of 3x3 models (row-major, F33 scaled to 1 when it is not ~0)."""
    p1 = m1.astype(np.float64)
    p2 = m2.astype(np.float64)
    c1 = p1.sum(0) * (1.0 / 7)
    c2 = p2.sum(0) * (1.0 / 7)
    s1 = np.sqrt(((p1 - c1) ** 2).sum(1)).sum() * (1.0 / 7)
    s2 = np.sqrt(((p2 - c2) ** 2).sum(1)).sum() * (1.0 / 7)
    if s1 < FLT_EPSILON or s2 < FLT_EPSILON:
        return []
    s1 = math.sqrt(2.0) / s1
    s2 = math.sqrt(2.0) / s2
    x0 = (p1[:, 0] - c1[0]) * s1
    y0 = (p1[:, 1] - c1[1]) * s1
    x1 = (p2[:, 0] - c2[0]) * s2
    y1 = (p2[:, 1] - c2[1]) * s2
    A = np.stack([x1 * x0, x1 * y0, x1, y1 * x0, y1 * y0, y1, x0, y0, np.ones(7)], axis=1)
    f1, f2 = null_space(A)
    f1, c = null_space_cubic(f1, f2)
    roots = solve_cubic(c)
    T1 = np.array([[s1, 0, -s1 * c1[0]], [0, s1, -s1 * c1[1]], [0, 0, 1.0]])
    T2 = np.array([[s2, 0, -s2 * c2[0]], [0, s2, -s2 * c2[1]], [0, 0, 1.0]])
    models = []
    for r in roots:
        lam, mu = r, 1.0
        s = f1[8] * r + f2[8]
        f = np.zeros(9)
        if abs(s) > DBL_EPSILON:
            mu = 1.0 / s
            lam *= mu
            f[8] = 1.0
        else:
            f[8] = 0.0
        f[:8] = f1[:8] * lam + f2[:8] * mu
        F = T2.T @ f.reshape(3, 3) @ T1
        if abs(F[2, 2]) > FLT_EPSILON:
            F = F * (1.0 / F[2, 2])
        models.append(F)
    return models


def compute_error(m1, m2, F):
    """FMEstimatorCallback::computeError: (float) max(d1^2 s1, d2^2 s2), points read as float, arithmetic in double."""
    F = F.reshape(9)
    x1 = m1[:, 0].astype(np.float64)
    y1 = m1[:, 1].astype(np.float64)
    x2 = m2[:, 0].astype(np.float64)
    y2 = m2[:, 1].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = F[0] * x1 + F[1] * y1 + F[2]
        b = F[3] * x1 + F[4] * y1 + F[5]
        c = F[6] * x1 + F[7] * y1 + F[8]
        s2 = 1.0 / (a * a + b * b)
        d2 = x2 * a + y2 * b + c
        a = F[0] * x2 + F[3] * y2 + F[6]
        b = F[1] * x2 + F[4] * y2 + F[7]
        c = F[2] * x2 + F[5] * y2 + F[8]
        s1 = 1.0 / (a * a + b * b)
        d1 = x1 * a + y1 * b + c
        return np.maximum(d1 * d1 * s1, d2 * d2 * s2).astype(np.float32)


def ransac_update_num_iters(p, ep, model_points, max_iters):
    """ptsetreg.cpp RANSACUpdateNumIters (cvRound = round half to even)."""
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, DBL_MIN)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < DBL_MIN:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def find_fundamental_mask(p, q, threshold=1.0, confidence=0.99, max_iters=1000, schedule=None,
                          null_space=null_space_svd, return_info=False):
    """cv::findFundamentalMat(p, q, FM_RANSAC, threshold, confidence, mask) -> (mask uint8 [n], F or None).
    schedule: int32 [>= iterations][7] sample indices to use instead of cv::RNG's (the injected schedule of the
    device path); None = OpenCV's own."""
    m1 = np.ascontiguousarray(p, np.float32).reshape(-1, 2)
    m2 = np.ascontiguousarray(q, np.float32).reshape(-1, 2)
    n = len(m1)
    info = {"iterations": 0, "method": None, "best_iteration": -1, "best_model": -1}
    if n < 7:
        return (np.zeros(n, np.uint8), None, info) if return_info else (np.zeros(n, np.uint8), None)
    if threshold <= 0:
        threshold = 3.0
    if confidence < DBL_EPSILON or confidence > 1 - DBL_EPSILON:
        confidence = 0.99
    if n == 7:
        models = run_7point(m1, m2, null_space)
        out = (np.ones(n, np.uint8), models[0] if models else None)
        return out + (info,) if return_info else out
    ransac = n >= 15
    info["method"] = "ransac" if ransac else "lmeds"
    niters = max(max_iters, 1) if ransac else ransac_update_num_iters(confidence, 0.45, 7, max_iters)
    if schedule is None:
        schedule = sample_schedule(m1, m2, niters, 10000 if ransac else 1000)
    best_mask = np.zeros(n, np.uint8)
    best_F = None
    max_good = 0
    min_median = float("inf")
    t = np.float32(threshold * threshold)
    it = 0
    while it < niters:
        idx = schedule[it]
        if idx[0] < 0:                      # getSubset failed: iteration 0 -> no result, later -> stop
            break
        models = run_7point(m1[idx], m2[idx], null_space)
        for k, F in enumerate(models):
            err = compute_error(m1, m2, F)
            if ransac:
                mask = (err <= t).astype(np.uint8)
                good = int(mask.sum())
                if good > max(max_good, 6):
                    best_mask, best_F, max_good = mask, F, good
                    info["best_iteration"], info["best_model"] = it, k
                    niters = ransac_update_num_iters(confidence, (n - good) / n, 7, niters)
            else:
                med = float(np.sort(err)[n // 2])
                if med < min_median:
                    min_median, best_F = med, F
                    info["best_iteration"], info["best_model"] = it, k
        it += 1
    info["iterations"] = it
    if not ransac and best_F is not None:
        sigma = 2.5 * 1.4826 * (1 + 5.0 / (n - 7)) * math.sqrt(min_median)
        sigma = max(sigma, 0.001)
        best_mask = (compute_error(m1, m2, best_F) <= np.float32(sigma * sigma)).astype(np.uint8)
    out = (best_mask, best_F)
    return out + (info,) if return_info else out
