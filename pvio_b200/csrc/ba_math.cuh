// Device restatement of PVIO's Lie-group helpers (fp64) and small dense helpers.
//   geometry/lie_algebra.h:25-42   hat / expmap / logmap   (Eigen AngleAxis semantics)
//   geometry/lie_algebra.cpp:22-59 right_jacobian with the same Taylor branches
//   estimation/ceres/quaternion_parameterization.h:28-31  Plus
// Quaternions are (x,y,z,w), Hamilton product.  Matrices are row-major 3x3 (double[9]).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pvio {

__device__ __forceinline__ void quat_to_mat(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double o[4]) {
    const double ax = a[0], ay = a[1], az = a[2], aw = a[3];
    const double bx = b[0], by = b[1], bz = b[2], bw = b[3];
    o[0] = aw * bx + ax * bw + ay * bz - az * by;
    o[1] = aw * by + ay * bw + az * bx - ax * bz;
    o[2] = aw * bz + az * bw + ax * by - ay * bx;
    o[3] = aw * bw - ax * bx - ay * by - az * bz;
}

__device__ __forceinline__ void quat_conj(const double q[4], double o[4]) {
    o[0] = -q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = q[3];
}

// lie_algebra.h:32-37 (AngleAxisd(|w|, w.stableNormalized()))
__device__ __forceinline__ void expmap(const double w[3], double q[4]) {
    const double m = fmax(fabs(w[0]), fmax(fabs(w[1]), fabs(w[2])));
    double ax = w[0], ay = w[1], az = w[2];
    if (m > 0) {
        const double a0 = w[0] / m, a1 = w[1] / m, a2 = w[2] / m;
        const double z = a0 * a0 + a1 * a1 + a2 * a2;
        const double k = 1.0 / (sqrt(z) * m);
        ax = w[0] * k; ay = w[1] * k; az = w[2] * k;
    }
    const double angle = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double s, c;
    sincos(0.5 * angle, &s, &c);
    q[0] = s * ax; q[1] = s * ay; q[2] = s * az; q[3] = c;
}

// lie_algebra.h:39-42 (AngleAxisd(q): angle in [0, pi])
__device__ __forceinline__ void logmap(const double q[4], double w[3]) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (n == 0.0) { w[0] = w[1] = w[2] = 0.0; return; }
    const double angle = 2.0 * atan2(n, fabs(q[3]));
    if (q[3] < 0) n = -n;
    const double k = angle / n;
    w[0] = k * q[0]; w[1] = k * q[1]; w[2] = k * q[2];
}

// quaternion_parameterization.h:28-31
__device__ __forceinline__ void quat_plus(const double q[4], const double d[3], double o[4]) {
    double e[4], t[4];
    expmap(d, e);
    quat_mul(q, e, t);
    const double n = 1.0 / sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3]);
    o[0] = t[0] * n; o[1] = t[1] * n; o[2] = t[2] * n; o[3] = t[3] * n;
}

__device__ __forceinline__ void hat(const double w[3], double H[9]) {
    H[0] = 0;     H[1] = -w[2]; H[2] = w[1];
    H[3] = w[2];  H[4] = 0;     H[5] = -w[0];
    H[6] = -w[1]; H[7] = w[0];  H[8] = 0;
}

__device__ __forceinline__ void mat3_mul(const double A[9], const double B[9], double C[9]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

__device__ __forceinline__ void mat3_mul_tn(const double A[9], const double B[9], double C[9]) {  // A^T B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}

__device__ __forceinline__ void mat3_vec(const double A[9], const double v[3], double o[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
}

__device__ __forceinline__ void mat3_tvec(const double A[9], const double v[3], double o[3]) {  // A^T v
#pragma unroll
    for (int i = 0; i < 3; ++i) o[i] = A[i] * v[0] + A[3 + i] * v[1] + A[6 + i] * v[2];
}

// lie_algebra.cpp:22-59
__device__ __forceinline__ void right_jacobian(const double w[3], double J[9]) {
    const double root2_eps = 1.4901161193847656e-08;   // sqrt(DBL_EPSILON)
    const double root4_eps = 1.220703125e-04;          // sqrt(root2_eps)
    const double qdrt720 = 5.180044732550419, qdrt5040 = 8.425701449380466;
    const double sqrt24 = 4.898979485566356, sqrt120 = 10.954451150103322;
    const double angle = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    double sangle, cangle;
    sincos(angle, &sangle, &cangle);
    const double angle2 = angle * angle;
    double cos_term, sin_term;
    if (angle > root4_eps * qdrt720) {
        cos_term = (1 - cangle) / angle2;
    } else {
        cos_term = 0.5;
        if (angle > root2_eps * sqrt24) cos_term -= angle2 / 24.0;
    }
    if (angle > root4_eps * qdrt5040) {
        sin_term = (angle - sangle) / (angle * angle2);
    } else {
        sin_term = 1.0 / 6.0;
        if (angle > root2_eps * sqrt120) sin_term -= angle2 / 120.0;
    }
    double H[9], H2[9];
    hat(w, H);
    mat3_mul(H, H, H2);
#pragma unroll
    for (int i = 0; i < 9; ++i) J[i] = -cos_term * H[i] + sin_term * H2[i];
    J[0] += 1.0; J[4] += 1.0; J[8] += 1.0;
}

// general 3x3 inverse (Eigen's .inverse() for fixed 3x3 uses the cofactor formula)
__device__ __forceinline__ void mat3_inv(const double A[9], double I[9]) {
    const double c0 = A[4] * A[8] - A[5] * A[7];
    const double c1 = A[5] * A[6] - A[3] * A[8];
    const double c2 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c0 + A[1] * c1 + A[2] * c2;
    const double id = 1.0 / det;
    I[0] = c0 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c1 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c2 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi (fp64).  V columns = eigenvectors.
// Stands in for Eigen::SelfAdjointEigenSolver<matrix<3>> in
// augmented_plane_distance_error_cost.h:90; only V diag(f(lambda)) V^T is used, which does
// not depend on eigenvector sign/order.
__device__ __forceinline__ void sym3_eig(const double Ain[9], double lam[3], double V[9]) {
    double A[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { A[i] = Ain[i]; V[i] = 0.0; }
    V[0] = V[4] = V[8] = 1.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        if (off < 1e-300) break;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int p = (k == 2) ? 1 : 0, q = (k == 0) ? 1 : 2;   // (0,1) (0,2) (1,2)
            const double apq = A[3 * p + q];
            if (fabs(apq) < 1e-320) continue;
            const double app = A[3 * p + p], aqq = A[3 * q + q];
            const double theta = (aqq - app) / (2.0 * apq);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            // A <- J^T A J
            for (int r = 0; r < 3; ++r) {
                const double arp = A[3 * r + p], arq = A[3 * r + q];
                A[3 * r + p] = c * arp - s * arq;
                A[3 * r + q] = s * arp + c * arq;
            }
            for (int r = 0; r < 3; ++r) {
                const double apr = A[3 * p + r], aqr = A[3 * q + r];
                A[3 * p + r] = c * apr - s * aqr;
                A[3 * q + r] = s * apr + c * aqr;
            }
            for (int r = 0; r < 3; ++r) {
                const double vrp = V[3 * r + p], vrq = V[3 * r + q];
                V[3 * r + p] = c * vrp - s * vrq;
                V[3 * r + q] = s * vrp + c * vrq;
            }
        }
    }
    lam[0] = A[0]; lam[1] = A[4]; lam[2] = A[8];
}

}  // namespace pvio
