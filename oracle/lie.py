"""TEST INFRASTRUCTURE ONLY -- fp64 NumPy restatement of PVIO's Lie-group helpers.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this package.  The product path (pvio_b200/) never does.

Follows (reference paths relative to /root/reference):
  * pvio/src/pvio/geometry/lie_algebra.h:25-42   hat / expmap / logmap
  * pvio/src/pvio/geometry/lie_algebra.cpp:22-59 right_jacobian (Taylor guards)
  * Eigen conventions used by the reference: quaternion coefficient order
    (x, y, z, w), Hamilton product, q*v == rotate v by q, AngleAxisd(q) returns an
    angle in [0, pi] (Eigen/src/Geometry/AngleAxis.h, operator=(QuaternionBase)).
Eigen itself is not in /root/reference (find_package, pvio/depends/CMakeLists.txt:25);
its documented behaviour is restated here.
"""
import numpy as np

EPS = np.finfo(np.float64).eps
_ROOT2_EPS = np.sqrt(EPS)
_ROOT4_EPS = np.sqrt(_ROOT2_EPS)
_QDRT720 = np.sqrt(np.sqrt(720.0))
_QDRT5040 = np.sqrt(np.sqrt(5040.0))
_SQRT24 = np.sqrt(24.0)
_SQRT120 = np.sqrt(120.0)


def hat(w):
    """lie_algebra.h:25-30"""
    w = np.asarray(w, dtype=np.float64)
    return np.array([[0.0, -w[2], w[1]],
                     [w[2], 0.0, -w[0]],
                     [-w[1], w[0], 0.0]])


def qmul(a, b):
    """Hamilton product, (x,y,z,w) storage (Eigen::Quaternion::operator*)."""
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz])


def qconj(q):
    return np.array([-q[0], -q[1], -q[2], q[3]])


def qnormalized(q):
    q = np.asarray(q, dtype=np.float64)
    return q / np.sqrt(np.dot(q, q))


def qmat(q):
    """Eigen::Quaternion::toRotationMatrix."""
    x, y, z, w = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def qrot(q, v):
    """q * v (rotate)."""
    return qmat(q) @ np.asarray(v, dtype=np.float64)


def expmap(w):
    """lie_algebra.h:32-37: AngleAxisd(|w|, w.stableNormalized()) -> quaternion."""
    w = np.asarray(w, dtype=np.float64)
    m = np.max(np.abs(w))
    if m > 0:
        z = np.sum((w / m) ** 2)
        axis = w / (np.sqrt(z) * m)
    else:
        axis = w.copy()
    angle = np.sqrt(np.dot(w, w))
    s = np.sin(0.5 * angle)
    return np.array([s * axis[0], s * axis[1], s * axis[2], np.cos(0.5 * angle)])


def logmap(q):
    """lie_algebra.h:39-42: AngleAxisd(q); angle*axis, angle in [0, pi]."""
    q = np.asarray(q, dtype=np.float64)
    v = q[:3]
    n = np.sqrt(np.dot(v, v))
    if n == 0.0:
        return np.zeros(3)
    angle = 2.0 * np.arctan2(n, abs(q[3]))
    if q[3] < 0:
        n = -n
    return angle * v / n


def right_jacobian(w):
    """lie_algebra.cpp:22-59."""
    w = np.asarray(w, dtype=np.float64)
    angle = np.sqrt(np.dot(w, w))
    cangle, sangle = np.cos(angle), np.sin(angle)
    angle2 = angle * angle
    if angle > _ROOT4_EPS * _QDRT720:
        cos_term = (1 - cangle) / angle2
    else:
        cos_term = 0.5
        if angle > _ROOT2_EPS * _SQRT24:
            cos_term -= angle2 / 24.0
    if angle > _ROOT4_EPS * _QDRT5040:
        sin_term = (angle - sangle) / (angle * angle2)
    else:
        sin_term = 1.0 / 6.0
        if angle > _ROOT2_EPS * _SQRT120:
            sin_term -= angle2 / 120.0
    hw = hat(w)
    return np.eye(3) - cos_term * hw + sin_term * hw @ hw


def quat_plus(q, d):
    """quaternion_parameterization.h:28-31: normalize(q (x) exp(d))."""
    return qnormalized(qmul(q, expmap(d)))
