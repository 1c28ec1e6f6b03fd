/* TEST / BASELINE INFRASTRUCTURE ONLY -- dependency-free fp64 C restatement of PVIO's sliding-window BA on
 * the CPU, the way the reference + Ceres perform it.  It is (a) checked against the NumPy oracle
 * (oracle/ba_oracle.py) in tests/ and (b) the timed "reference CPU path" of bench.py (cpu_baseline.kind =
 * "port": Ceres / Eigen are not installed, so the reference's own binary cannot be built here).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load it.
 *
 *   ReprojectionErrorCost::Evaluate             estimation/ceres/reprojection_error_cost.h:40-120
 *   PreIntegrationErrorCost::Evaluate           estimation/ceres/preintegration_error_cost.h:40-160
 *   MarginalizationErrorCost::Evaluate          estimation/ceres/marginalization_error_cost.h:53-94
 *   AugmentedPlaneDistanceErrorCost::Evaluate   estimation/ceres/augmented_plane_distance_error_cost.h:53-136
 *   CauchyLoss(1.0) corrector                   estimation/bundle_adjustor.cpp:58,192 (ceres corrector.cc)
 *   Plus                                        estimation/ceres/quaternion_parameterization.h:28-31
 *   problem assembly                            estimation/bundle_adjustor.cpp:75-242
 *   ceres::Solve: SPARSE_SCHUR (landmarks eliminated, dense Cholesky of the reduced camera system), Jacobi scaling,
 *   TRADITIONAL_DOGLEG trust region             estimation/ceres/solver_options.h:26-33 (Ceres 1.14 minimiser logic)
 *   BundleAdjustor::marginalize_frame           estimation/bundle_adjustor.cpp:348-599
 * Single-threaded per window like the reference (num_threads = 1, solver_options.h:31); the batch entry points spread
 * INDEPENDENT windows over POSIX threads (the image has no libgomp).  The window / state / options / summary
 * structures are the C-ABI's (include/pvio_b200.h): the same gathered arrays feed both sides.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include "../include/pvio_b200.h"

typedef pvio_b200_window Win;

/* ------------------------------------------------------------------ small algebra (row-major 3x3) */
static void q2m(const double *q, double *R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
static void qmul(const double *a, const double *b, double *o) {
    const double t[4] = { a[3]*b[0] + a[0]*b[3] + a[1]*b[2] - a[2]*b[1], a[3]*b[1] + a[1]*b[3] + a[2]*b[0] - a[0]*b[2],
                          a[3]*b[2] + a[2]*b[3] + a[0]*b[1] - a[1]*b[0], a[3]*b[3] - a[0]*b[0] - a[1]*b[1] - a[2]*b[2] };
    memcpy(o, t, 32);
}
static void qconj(const double *q, double *o) { o[0] = -q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = q[3]; }
static void mv(const double *A, const double *v, double *o) { double t[3]; for (int i = 0; i < 3; ++i) t[i] = A[3*i]*v[0] + A[3*i+1]*v[1] + A[3*i+2]*v[2]; memcpy(o, t, 24); }
static void mtv(const double *A, const double *v, double *o) { double t[3]; for (int i = 0; i < 3; ++i) t[i] = A[i]*v[0] + A[3+i]*v[1] + A[6+i]*v[2]; memcpy(o, t, 24); }
static void mm(const double *A, const double *B, double *C) {
    double t[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[3*i+j] = A[3*i]*B[j] + A[3*i+1]*B[3+j] + A[3*i+2]*B[6+j];
    memcpy(C, t, 72);
}
static void mtm(const double *A, const double *B, double *C) {            /* A^T B */
    double t[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[3*i+j] = A[i]*B[j] + A[3+i]*B[3+j] + A[6+i]*B[6+j];
    memcpy(C, t, 72);
}
static void mmt(const double *A, const double *B, double *C) {            /* A B^T */
    double t[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) t[3*i+j] = A[3*i]*B[3*j] + A[3*i+1]*B[3*j+1] + A[3*i+2]*B[3*j+2];
    memcpy(C, t, 72);
}
static void hat(const double *w, double *H) {
    H[0] = 0; H[1] = -w[2]; H[2] = w[1]; H[3] = w[2]; H[4] = 0; H[5] = -w[0]; H[6] = -w[1]; H[7] = w[0]; H[8] = 0;
}
static void inv3(const double *A, double *I) {
    const double c0 = A[4]*A[8] - A[5]*A[7], c1 = A[5]*A[6] - A[3]*A[8], c2 = A[3]*A[7] - A[4]*A[6];
    const double id = 1.0 / (A[0]*c0 + A[1]*c1 + A[2]*c2);
    const double t[9] = { c0*id, (A[2]*A[7] - A[1]*A[8])*id, (A[1]*A[5] - A[2]*A[4])*id,
                          c1*id, (A[0]*A[8] - A[2]*A[6])*id, (A[2]*A[3] - A[0]*A[5])*id,
                          c2*id, (A[1]*A[6] - A[0]*A[7])*id, (A[0]*A[4] - A[1]*A[3])*id };
    memcpy(I, t, 72);
}
/* geometry/lie_algebra.h:32-42 (Eigen AngleAxis semantics), lie_algebra.cpp:22-59 */
static void expmap(const double *w, double *q) {
    const double a = sqrt(w[0]*w[0] + w[1]*w[1] + w[2]*w[2]);
    if (a > 0) { const double s = sin(0.5 * a) / a; q[0] = s*w[0]; q[1] = s*w[1]; q[2] = s*w[2]; q[3] = cos(0.5 * a); }
    else { q[0] = q[1] = q[2] = 0; q[3] = 1; }
}
static void logmap(const double *q, double *w) {
    double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2]);
    if (n == 0.0) { w[0] = w[1] = w[2] = 0; return; }
    const double angle = 2.0 * atan2(n, fabs(q[3]));
    if (q[3] < 0) n = -n;
    const double k = angle / n;
    w[0] = k*q[0]; w[1] = k*q[1]; w[2] = k*q[2];
}
static void right_jacobian(const double *w, double *J) {
    const double root2_eps = 1.4901161193847656e-08, root4_eps = 1.220703125e-04;
    const double qdrt720 = 5.180044732550419, qdrt5040 = 8.425701449380466, sqrt24 = 4.898979485566356, sqrt120 = 10.954451150103322;
    const double angle = sqrt(w[0]*w[0] + w[1]*w[1] + w[2]*w[2]), angle2 = angle * angle;
    double cos_term, sin_term;
    if (angle > root4_eps * qdrt720) cos_term = (1 - cos(angle)) / angle2;
    else { cos_term = 0.5; if (angle > root2_eps * sqrt24) cos_term -= angle2 / 24.0; }
    if (angle > root4_eps * qdrt5040) sin_term = (angle - sin(angle)) / (angle * angle2);
    else { sin_term = 1.0 / 6.0; if (angle > root2_eps * sqrt120) sin_term -= angle2 / 120.0; }
    double H[9], H2[9];
    hat(w, H); mm(H, H, H2);
    for (int i = 0; i < 9; ++i) J[i] = -cos_term * H[i] + sin_term * H2[i];
    J[0] += 1; J[4] += 1; J[8] += 1;
}
static void quat_plus(const double *q, const double *w, double *o) {
    double e[4], t[4];
    expmap(w, e); qmul(q, e, t);
    const double n = 1.0 / sqrt(t[0]*t[0] + t[1]*t[1] + t[2]*t[2] + t[3]*t[3]);
    for (int i = 0; i < 4; ++i) o[i] = t[i] * n;
}
/* symmetric 3x3 eigen-decomposition (cyclic Jacobi); only V f(lambda) V^T is used */
static void eig3(const double *Ain, double *lam, double *V) {
    double A[9]; memcpy(A, Ain, 72);
    for (int i = 0; i < 9; ++i) V[i] = 0; V[0] = V[4] = V[8] = 1;
    for (int sweep = 0; sweep < 30; ++sweep) {
        if (A[1]*A[1] + A[2]*A[2] + A[5]*A[5] < 1e-300) break;
        for (int k = 0; k < 3; ++k) {
            const int p = (k == 2) ? 1 : 0, q = (k == 0) ? 1 : 2;
            const double apq = A[3*p+q];
            if (fabs(apq) < 1e-320) continue;
            const double theta = (A[3*q+q] - A[3*p+p]) / (2.0 * apq);
            const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0)), c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int r = 0; r < 3; ++r) { const double a = A[3*r+p], b = A[3*r+q]; A[3*r+p] = c*a - s*b; A[3*r+q] = s*a + c*b; }
            for (int r = 0; r < 3; ++r) { const double a = A[3*p+r], b = A[3*q+r]; A[3*p+r] = c*a - s*b; A[3*q+r] = s*a + c*b; }
            for (int r = 0; r < 3; ++r) { const double a = V[3*r+p], b = V[3*r+q]; V[3*r+p] = c*a - s*b; V[3*r+q] = s*a + c*b; }
        }
    }
    lam[0] = A[0]; lam[1] = A[4]; lam[2] = A[8];
}

/* ------------------------------------------------------------------ cost functors */
typedef struct { double R[9], p[3]; } Pose;

/* reprojection_error_cost.h:40-120.  J: 2x13 = [q_tgt(3) p_tgt(3) q_ref(3) p_ref(3) rho] */
static void reproj(const Pose *tg, const Pose *rf, double rho, const double *zt, const double *zr,
                   const double *Rcs, const double *pcs, const double *W, double *r, double *J) {
    double yref[3] = {zr[0] / rho, zr[1] / rho, 1.0 / rho}, yrc[3], x[3], d[3], ytc[3], yt[3], t[3];
    mv(Rcs, yref, yrc); for (int i = 0; i < 3; ++i) yrc[i] += pcs[i];
    mv(rf->R, yrc, x); for (int i = 0; i < 3; ++i) { x[i] += rf->p[i]; d[i] = x[i] - tg->p[i]; }
    mtv(tg->R, d, ytc);
    for (int i = 0; i < 3; ++i) t[i] = ytc[i] - pcs[i];
    mtv(Rcs, t, yt);
    const double u0 = yt[0] / yt[2] - zt[0], u1 = yt[1] / yt[2] - zt[1];
    r[0] = W[0] * u0 + W[1] * u1; r[1] = W[2] * u0 + W[3] * u1;
    if (!J) return;
    const double iz = 1.0 / yt[2];
    const double dp[6] = {iz, 0, -yt[0] * iz * iz, 0, iz, -yt[1] * iz * iz};
    double A[6], Dtc[6], Dx[6], Drc[6];
    for (int j = 0; j < 3; ++j) { A[j] = W[0] * dp[j] + W[1] * dp[3 + j]; A[3 + j] = W[2] * dp[j] + W[3] * dp[3 + j]; }
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) {           /* Dtc = A Rcs^T :75, Dx = Dtc R_tgt^T :79, Drc = Dx R_ref :86 */
        Dtc[3*i+j] = A[3*i]*Rcs[3*j] + A[3*i+1]*Rcs[3*j+1] + A[3*i+2]*Rcs[3*j+2];
    }
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) Dx[3*i+j] = Dtc[3*i]*tg->R[3*j] + Dtc[3*i+1]*tg->R[3*j+1] + Dtc[3*i+2]*tg->R[3*j+2];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) Drc[3*i+j] = Dx[3*i]*rf->R[j] + Dx[3*i+1]*rf->R[3+j] + Dx[3*i+2]*rf->R[6+j];
    for (int i = 0; i < 2; ++i) {
        const double a0 = Dtc[3*i], a1 = Dtc[3*i+1], a2 = Dtc[3*i+2];       /* Dtc hat(ytc) :95 */
        J[13*i+0] = a1 * ytc[2] - a2 * ytc[1]; J[13*i+1] = a2 * ytc[0] - a0 * ytc[2]; J[13*i+2] = a0 * ytc[1] - a1 * ytc[0];
        for (int j = 0; j < 3; ++j) { J[13*i+3+j] = -Dx[3*i+j]; J[13*i+9+j] = Dx[3*i+j]; }   /* :100 :109 */
        const double b0 = Drc[3*i], b1 = Drc[3*i+1], b2 = Drc[3*i+2];       /* -Drc hat(yrc) :104 */
        J[13*i+6] = -(b1 * yrc[2] - b2 * yrc[1]); J[13*i+7] = -(b2 * yrc[0] - b0 * yrc[2]); J[13*i+8] = -(b0 * yrc[1] - b1 * yrc[0]);
    }
    double c[3]; mv(Rcs, yref, c);
    for (int i = 0; i < 2; ++i) J[13*i+12] = -(Drc[3*i]*c[0] + Drc[3*i+1]*c[1] + Drc[3*i+2]*c[2]) / rho;   /* :113 */
}

/* preintegration_error_cost.h:40-160; fi/fj frame states [16], rec the IMU record, (bg0, ba0) the bias linearisation
 * point.  Outputs the WHITENED r[15] and J[15][30] (local coordinates [th p v bg ba] of i, then j); J may be NULL. */
static void imu_eval(const double *fi, const double *fj, const double *rec, const double *bg0, const double *ba0,
                     const double *imu_q, const double *imu_p, double *rw, double *Jw) {
    static const double g[3] = {0.0, 0.0, -9.80665};
    const double *qic = fi, *pic = fi + 4, *vi = fi + 7, *bgi = fi + 10, *bai = fi + 13;
    const double *qjc = fj, *pjc = fj + 4, *vj = fj + 7, *bgj = fj + 10, *baj = fj + 13;
    const double dt = rec[0], *dq = rec + 1, *dp = rec + 5, *dv = rec + 8, *Wm = rec + 11;
    const double *dq_dbg = rec + 236, *dp_dbg = rec + 245, *dp_dba = rec + 254, *dv_dbg = rec + 263, *dv_dba = rec + 272;
    double dbg[3], dba[3], r[15], J[450];
    for (int k = 0; k < 3; ++k) { dbg[k] = bgi[k] - bg0[k]; dba[k] = bai[k] - ba0[k]; }
    double qi[4], qj[4], Rci[9], Rcj[9], t3[3], pi[3], pj[3];
    qmul(qic, imu_q, qi); qmul(qjc, imu_q, qj); q2m(qic, Rci); q2m(qjc, Rcj);
    mv(Rci, imu_p, t3); for (int k = 0; k < 3; ++k) pi[k] = pic[k] + t3[k];
    mv(Rcj, imu_p, t3); for (int k = 0; k < 3; ++k) pj[k] = pjc[k] + t3[k];
    double w3[3], e4[4], dqc[4], c1[4], c2[4], c3[4], c4[4];
    mv(dq_dbg, dbg, w3); expmap(w3, e4); qmul(dq, e4, dqc); qconj(dqc, c1); qconj(qi, c2); qmul(c1, c2, c3); qmul(c3, qj, c4);
    logmap(c4, r);                                                                   /* :79 */
    double Ri[9], a3[3], b3[3], c_[3], d_[3];
    q2m(qi, Ri);
    for (int k = 0; k < 3; ++k) a3[k] = pj[k] - pi[k] - dt * vi[k] - 0.5 * dt * dt * g[k];
    mtv(Ri, a3, b3); mv(dp_dbg, dbg, c_); mv(dp_dba, dba, d_);
    for (int k = 0; k < 3; ++k) r[3 + k] = b3[k] - (dp[k] + c_[k] + d_[k]);          /* :80 */
    for (int k = 0; k < 3; ++k) a3[k] = vj[k] - vi[k] - dt * g[k];
    mtv(Ri, a3, b3); mv(dv_dbg, dbg, c_); mv(dv_dba, dba, d_);
    for (int k = 0; k < 3; ++k) r[6 + k] = b3[k] - (dv[k] + c_[k] + d_[k]);          /* :81 */
    for (int k = 0; k < 3; ++k) { r[9 + k] = bgj[k] - bgi[k]; r[12 + k] = baj[k] - bai[k]; }
    for (int i = 0; i < 15; ++i) { double s = 0; for (int k = 0; k < 15; ++k) s += Wm[15*i+k] * r[k]; rw[i] = s; }
    if (!Jw) return;
    memset(J, 0, sizeof(J));
    double Jr[9], Jri[9], Rimu[9], Rj[9], M1[9], M2[9], H3[9], Rit[9];
    right_jacobian(r, Jr); inv3(Jr, Jri); q2m(imu_q, Rimu); q2m(qj, Rj);
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) Rit[3*a+b] = Ri[3*b+a];
#define PUT(row0, col0, M, sgn) for (int a_ = 0; a_ < 3; ++a_) for (int b_ = 0; b_ < 3; ++b_) J[((row0) + a_) * 30 + (col0) + b_] = (sgn) * (M)[3 * a_ + b_];
    mtm(Rj, Rci, M1); mm(Jri, M1, M2); PUT(0, 0, M2, -1.0);                          /* :86-93 */
    for (int k = 0; k < 3; ++k) a3[k] = pj[k] - pic[k] - dt * vi[k] - 0.5 * dt * dt * g[k];
    mtv(Rci, a3, b3); hat(b3, H3); mtm(Rimu, H3, M1); PUT(3, 0, M1, 1.0);
    for (int k = 0; k < 3; ++k) a3[k] = vj[k] - vi[k] - dt * g[k];
    mtv(Rci, a3, b3); hat(b3, H3); mtm(Rimu, H3, M1); PUT(6, 0, M1, 1.0);
    PUT(3, 3, Rit, -1.0); PUT(3, 6, Rit, -dt); PUT(6, 6, Rit, -1.0);                 /* :94-106 */
    {   double er[4], erc[4], Rer[9], Jr2[9];                                        /* :107-115 */
        expmap(r, er); qconj(er, erc); q2m(erc, Rer); right_jacobian(w3, Jr2);
        mm(Jri, Rer, M1); mm(M1, Jr2, M2); mm(M2, dq_dbg, M1); PUT(0, 9, M1, -1.0);
    }
    PUT(3, 9, dp_dbg, -1.0); PUT(6, 9, dv_dbg, -1.0);
    for (int k = 0; k < 3; ++k) J[(9 + k) * 30 + 9 + k] = -1.0;
    PUT(3, 12, dp_dba, -1.0); PUT(6, 12, dv_dba, -1.0);                              /* :116-123 */
    for (int k = 0; k < 3; ++k) J[(12 + k) * 30 + 12 + k] = -1.0;
    mmt(Jri, Rimu, M1); PUT(0, 15, M1, 1.0);                                         /* :124-130 */
    hat(imu_p, H3); mm(Rit, Rcj, M1); mm(M1, H3, M2); PUT(3, 15, M2, -1.0);
    PUT(3, 18, Rit, 1.0); PUT(6, 21, Rit, 1.0);                                      /* :131-154 */
    for (int k = 0; k < 3; ++k) { J[(9 + k) * 30 + 24 + k] = 1.0; J[(12 + k) * 30 + 27 + k] = 1.0; }
#undef PUT
    for (int i = 0; i < 15; ++i) for (int c = 0; c < 30; ++c) { double s = 0; for (int k = 0; k < 15; ++k) s += Wm[15*i+k] * J[30*k+c]; Jw[30*i+c] = s; }
}

/* augmented_plane_distance_error_cost.h:53-136.  fr[K] frame indices, z [K][2]; J [6K] (may be NULL) */
static double plane_eval(int K, const int32_t *fr, const double *z, const double *frames, const double *cam_q, const double *cam_p,
                         const double *pl, double sic, double *J) {
    double A[2 * PVIO_B200_MAX_FRAMES + 1][3], b[2 * PVIO_B200_MAX_FRAMES + 1], Rsw_l[PVIO_B200_MAX_FRAMES][9];
    double Rcs[9], tc[3];
    q2m(cam_q, Rcs); mtv(Rcs, cam_p, tc);
    for (int i = 0; i < K; ++i) {
        const double *fs = frames + 16 * fr[i];
        double R[9], *Rsw = Rsw_l[i], Tsw[3];
        q2m(fs, R);
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += Rcs[3*k+a] * R[3*c+k]; Rsw[3*a+c] = s; }   /* :68 */
        mv(Rsw, fs + 4, Tsw); for (int k = 0; k < 3; ++k) Tsw[k] = -Tsw[k] - tc[k];                                       /* :69 */
        const double u = z[2*i], v = z[2*i+1];
        for (int k = 0; k < 3; ++k) { A[2*i][k] = u * Rsw[6+k] - Rsw[k]; A[2*i+1][k] = v * Rsw[6+k] - Rsw[3+k]; }
        b[2*i] = u * Tsw[2] - Tsw[0]; b[2*i+1] = v * Tsw[2] - Tsw[1];
    }
    for (int k = 0; k < 3; ++k) A[2*K][k] = pl[k];
    b[2*K] = pl[3];
    double ATA[9] = {0}, ATb[3] = {0};
    for (int i = 0; i < 2 * K + 1; ++i) for (int a = 0; a < 3; ++a) { ATb[a] += A[i][a] * b[i]; for (int c = 0; c < 3; ++c) ATA[3*a+c] += A[i][a] * A[i][c]; }
    double lam[3], V[9], Ainv[9], x[3];
    eig3(ATA, lam, V);                                                               /* :90 */
    for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) { double s = 0; for (int k = 0; k < 3; ++k) s += V[3*a+k] * (lam[k] > 1e-8 ? 1.0 / lam[k] : 0.0) * V[3*c+k]; Ainv[3*a+c] = s; }
    mv(Ainv, ATb, x); for (int k = 0; k < 3; ++k) x[k] = -x[k];                       /* :94 */
    const double r = (pl[0]*x[0] + pl[1]*x[1] + pl[2]*x[2] - pl[3]) * sic;            /* :96 :133 */
    if (!J) return r;
    for (int i = 0; i < K; ++i) {
        const double *fs = frames + 16 * fr[i];
        const double u = z[2*i], v = z[2*i+1];
        const double Jb[2][3] = {{-1.0, 0.0, u}, {0.0, -1.0, v}};
        double R[9], M[9] = {0};
        q2m(fs, R);
        for (int rr = 0; rr < 2; ++rr) {
            const double *Ar = A[2*i+rr];
            const double sres = b[2*i+rr] + Ar[0]*x[0] + Ar[1]*x[1] + Ar[2]*x[2];
            double AiA[3], dxdA[9], cj[3], H3[9], dAdq[9], T9[9];
            for (int k = 0; k < 3; ++k) AiA[k] = Ar[0]*Ainv[k] + Ar[1]*Ainv[3+k] + Ar[2]*Ainv[6+k];
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) dxdA[3*a+c] = sres * Ainv[3*a+c] + AiA[a] * x[c];   /* :105 */
            mv(Rcs, Jb[rr], cj); hat(cj, H3); mm(R, H3, dAdq); mm(dxdA, dAdq, T9);                                  /* :107 */
            for (int k = 0; k < 9; ++k) M[k] += T9[k];
        }
        double pb[3], H3[9], AtJb[9], T1[9], T2[9], T3[9];
        mtv(R, fs + 4, pb); hat(pb, H3);                                                                            /* :110 */
        for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) AtJb[3*a+c] = A[2*i][a] * Jb[0][c] + A[2*i+1][a] * Jb[1][c];
        mm(Ainv, AtJb, T1); mmt(T1, Rcs, T2); mm(T2, H3, T3);
        for (int k = 0; k < 9; ++k) M[k] += T3[k];
        mm(T1, Rsw_l[i], T2);                                                                                       /* :117 */
        for (int k = 0; k < 3; ++k) {
            J[6*i+k] = (pl[0]*M[k] + pl[1]*M[3+k] + pl[2]*M[6+k]) * sic;
            J[6*i+3+k] = (pl[0]*T2[k] + pl[1]*T2[3+k] + pl[2]*T2[6+k]) * sic;
        }
    }
    return r;
}

/* marginalization_error_cost.h:53-94: r = S r0 + e, J = S E with E = blockdiag(Jr^-1, I, I, I, I) */
static void prior_eval(const Win *w, const double *frames, double *r, double *r0, double *Jri /* [n][9] or NULL */) {
    const int n = w->n_prior, d = 15 * n;
    for (int i = 0; i < n; ++i) {
        const double *fs = frames + 16 * w->prior_frames[i], *x0 = w->prior_state0 + 16 * i;
        double c[4], dq[4];
        qconj(x0, c); qmul(c, fs, dq); logmap(dq, r0 + 15 * i);
        for (int k = 0; k < 12; ++k) r0[15 * i + 3 + k] = fs[4 + k] - x0[4 + k];
        if (Jri) { double Jr[9]; right_jacobian(r0 + 15 * i, Jr); inv3(Jr, Jri + 9 * i); }
    }
    for (int i = 0; i < d; ++i) { double s = w->prior_e[i]; const double *Si = w->prior_S + (size_t)i * d; for (int k = 0; k < d; ++k) s += Si[k] * r0[k]; r[i] = s; }
}

/* ------------------------------------------------------------------ normal equations (pose block dense, landmarks 1x1) */
typedef struct {
    int N, M, stride, P;          /* stride 15 (inertial) or 6, P = stride N */
    double *H, *g;                /* [P][P], [P]  pose part of J^T J, J^T r (unreduced) */
    double *Hpl;                  /* [M][P] */
    double *Hll, *gl;             /* [M] */
    double cost;
} Normal;

static void normal_alloc(Normal *n, int N, int M, int inertial) {
    n->N = N; n->M = M; n->stride = inertial ? 15 : 6; n->P = n->stride * N;
    n->H = (double *)malloc(sizeof(double) * ((size_t)n->P * n->P + n->P + (size_t)M * n->P + 2 * (size_t)M + 8));
    n->g = n->H + (size_t)n->P * n->P; n->Hpl = n->g + n->P; n->Hll = n->Hpl + (size_t)M * n->P; n->gl = n->Hll + M;
}
static void normal_free(Normal *n) { free(n->H); }

static const double *lin_bias(const Win *w, const double *lin_frames, int n, int which /*0 bg, 1 ba*/) {
    if (lin_frames) return lin_frames + 16 * w->imu_frame_i[n] + (which ? 13 : 10);     /* quirk Q1: aliases frame_i->motion */
    return w->imu_data + (size_t)n * PVIO_B200_IMU_STRIDE + (which ? PVIO_B200_IMU_BA0 : PVIO_B200_IMU_BG0);
}

/* total cost (candidate evaluation): residuals only */
static double window_cost(const Win *w, const double *frames, const double *rho, const double *lin_frames) {
    const int N = w->n_frames, M = w->n_landmarks;
    const double b = w->cauchy_a * w->cauchy_a;
    double Rcs[9]; q2m(w->cam_q_cs, Rcs);
    Pose P[PVIO_B200_MAX_FRAMES];
    for (int f = 0; f < N; ++f) { q2m(frames + 16 * f, P[f].R); memcpy(P[f].p, frames + 16 * f + 4, 24); }
    double cost = 0;
    if (w->use_inertial && w->n_prior > 0) {
        const int d = 15 * w->n_prior;
        double *r = (double *)malloc(sizeof(double) * 2 * d);
        prior_eval(w, frames, r, r + d, 0);
        double s = 0; for (int i = 0; i < d; ++i) s += r[i] * r[i];
        cost += 0.5 * s; free(r);
    }
    for (int l = 0; l < M; ++l) for (int k = w->lm_obs_begin[l]; k < w->lm_obs_begin[l + 1]; ++k) {
        double r[2];
        reproj(&P[w->obs_frame[k]], &P[w->lm_anchor[l]], rho[l], w->obs_z + 2 * k, w->lm_z_ref + 2 * l, Rcs, w->cam_p_cs, w->sqrt_inv_cov, r, 0);
        cost += 0.5 * b * log(1.0 + (r[0]*r[0] + r[1]*r[1]) / b);
    }
    for (int t = 0; t < w->n_plane_tracks; ++t) {
        const int b0 = w->pt_obs_begin[t], K = w->pt_obs_begin[t + 1] - b0;
        const double r = plane_eval(K, w->pt_obs_frame + b0, w->pt_obs_z + 2 * b0, frames, w->cam_q_cs, w->cam_p_cs,
                                    w->plane_param + 4 * w->pt_plane[t], w->plane_sqrt_inv_cov, 0);
        cost += 0.5 * b * log(1.0 + r * r / b);
    }
    if (w->use_inertial) for (int n = 0; n < w->n_imu; ++n) {
        double r[15];
        imu_eval(frames + 16 * w->imu_frame_i[n], frames + 16 * w->imu_frame_j[n], w->imu_data + (size_t)n * PVIO_B200_IMU_STRIDE,
                 lin_bias(w, lin_frames, n, 0), lin_bias(w, lin_frames, n, 1), w->imu_q_cs, w->imu_p_cs, r, 0);
        double s = 0; for (int i = 0; i < 15; ++i) s += r[i] * r[i];
        cost += 0.5 * s;
    }
    return cost;
}

/* every residual block, loss-corrected, in the order bundle_adjustor.cpp adds them (:126-242) */
static void assemble(const Win *w, const double *frames, const double *rho, const double *lin_frames, Normal *nm) {
    const int N = nm->N, M = nm->M, st = nm->stride, P = nm->P;
    const double b = w->cauchy_a * w->cauchy_a;
    memset(nm->H, 0, sizeof(double) * ((size_t)P * P + P + (size_t)M * P + 2 * (size_t)M));
    double cost = 0, Rcs[9];
    q2m(w->cam_q_cs, Rcs);
    Pose Ps[PVIO_B200_MAX_FRAMES];
    for (int f = 0; f < N; ++f) { q2m(frames + 16 * f, Ps[f].R); memcpy(Ps[f].p, frames + 16 * f + 4, 24); }
    double *H = nm->H, *g = nm->g;
    if (w->use_inertial && w->n_prior > 0) {                                          /* prior :126-139, no loss */
        const int n = w->n_prior, d = 15 * n;
        double *r = (double *)malloc(sizeof(double) * (2 * (size_t)d + 9 * (size_t)n + (size_t)d * d)), *r0 = r + d, *Jri = r0 + d, *J = Jri + 9 * n;
        prior_eval(w, frames, r, r0, Jri);
        for (int i = 0; i < d; ++i) cost += 0.5 * r[i] * r[i];
        for (int i = 0; i < d; ++i) for (int c = 0; c < d; ++c) {                     /* J = S E */
            const int fc = c / 15, cc = c % 15;
            double s;
            if (cc < 3) { s = 0; for (int k = 0; k < 3; ++k) s += w->prior_S[(size_t)i * d + 15 * fc + k] * Jri[9 * fc + 3 * k + cc]; }
            else s = w->prior_S[(size_t)i * d + c];
            J[(size_t)i * d + c] = s;
        }
        for (int a = 0; a < d; ++a) {
            const int ga = 15 * w->prior_frames[a / 15] + a % 15;
            double s = 0; for (int i = 0; i < d; ++i) s += J[(size_t)i * d + a] * r[i];
            g[ga] += s;
            for (int c = 0; c < d; ++c) {
                const int gc = 15 * w->prior_frames[c / 15] + c % 15;
                double h = 0; for (int i = 0; i < d; ++i) h += J[(size_t)i * d + a] * J[(size_t)i * d + c];
                H[(size_t)ga * P + gc] += h;
            }
        }
        free(r);
    }
    for (int l = 0; l < M; ++l) {                                                      /* reprojection :142-161, CauchyLoss */
        const int a = w->lm_anchor[l];
        double *h = nm->Hpl + (size_t)l * P;
        for (int k = w->lm_obs_begin[l]; k < w->lm_obs_begin[l + 1]; ++k) {
            const int t = w->obs_frame[k];
            double r[2], J[26];
            reproj(&Ps[t], &Ps[a], rho[l], w->obs_z + 2 * k, w->lm_z_ref + 2 * l, Rcs, w->cam_p_cs, w->sqrt_inv_cov, r, J);
            const double s = r[0]*r[0] + r[1]*r[1], tt = 1.0 + s / b, sc = sqrt(1.0 / tt);   /* corrector */
            cost += 0.5 * b * log(tt);
            r[0] *= sc; r[1] *= sc; for (int i = 0; i < 26; ++i) J[i] *= sc;
            const int col[2] = {st * t, st * a};
            for (int bi = 0; bi < 2; ++bi) for (int i = 0; i < 6; ++i) {
                const double j0 = J[6 * bi + i], j1 = J[13 + 6 * bi + i];
                g[col[bi] + i] += j0 * r[0] + j1 * r[1];
                h[col[bi] + i] += j0 * J[12] + j1 * J[25];
                for (int bj = 0; bj < 2; ++bj) for (int j = 0; j < 6; ++j)
                    H[(size_t)(col[bi] + i) * P + col[bj] + j] += j0 * J[6 * bj + j] + j1 * J[13 + 6 * bj + j];
            }
            nm->Hll[l] += J[12] * J[12] + J[25] * J[25];
            nm->gl[l] += J[12] * r[0] + J[25] * r[1];
        }
    }
    for (int t = 0; t < w->n_plane_tracks; ++t) {                                      /* plane :162-196, CauchyLoss */
        const int b0 = w->pt_obs_begin[t], K = w->pt_obs_begin[t + 1] - b0;
        double J[6 * PVIO_B200_MAX_FRAMES];
        double r = plane_eval(K, w->pt_obs_frame + b0, w->pt_obs_z + 2 * b0, frames, w->cam_q_cs, w->cam_p_cs,
                              w->plane_param + 4 * w->pt_plane[t], w->plane_sqrt_inv_cov, J);
        const double tt = 1.0 + r * r / b, sc = sqrt(1.0 / tt);
        cost += 0.5 * b * log(tt);
        r *= sc; for (int i = 0; i < 6 * K; ++i) J[i] *= sc;
        for (int i = 0; i < 6 * K; ++i) {
            const int gi = st * w->pt_obs_frame[b0 + i / 6] + i % 6;
            g[gi] += J[i] * r;
            for (int j = 0; j < 6 * K; ++j) H[(size_t)gi * P + st * w->pt_obs_frame[b0 + j / 6] + j % 6] += J[i] * J[j];
        }
    }
    if (w->use_inertial) for (int n = 0; n < w->n_imu; ++n) {                          /* IMU :220-242, no loss */
        const int fi = w->imu_frame_i[n], fj = w->imu_frame_j[n];
        double r[15], J[450];
        imu_eval(frames + 16 * fi, frames + 16 * fj, w->imu_data + (size_t)n * PVIO_B200_IMU_STRIDE, lin_bias(w, lin_frames, n, 0),
                 lin_bias(w, lin_frames, n, 1), w->imu_q_cs, w->imu_p_cs, r, J);
        for (int i = 0; i < 15; ++i) cost += 0.5 * r[i] * r[i];
        for (int a = 0; a < 30; ++a) {
            const int ga = a < 15 ? 15 * fi + a : 15 * fj + a - 15;
            double s = 0; for (int k = 0; k < 15; ++k) s += J[30 * k + a] * r[k];
            g[ga] += s;
            for (int c = 0; c < 30; ++c) {
                const int gc = c < 15 ? 15 * fi + c : 15 * fj + c - 15;
                double h = 0; for (int k = 0; k < 15; ++k) h += J[30 * k + a] * J[30 * k + c];
                H[(size_t)ga * P + gc] += h;
            }
        }
    }
    nm->cost = cost;
}

static int pose_free(const Win *w, int stride, int i) { return !(w->frame_fixed && w->frame_fixed[i / stride] && (i % stride) < 6); }
static int lm_free(const Win *w, int l) { return w->lm_obs_begin[l + 1] > w->lm_obs_begin[l]; }

static double clampd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* Solve (H + reg) x = g on the free coordinates by landmark elimination + dense Cholesky (what SPARSE_SCHUR does).
 * regp [P], regl [M]: the regulariser added to the diagonal.  xp [P], xl [M].  Returns 0, or -1 if not positive definite. */
static int schur_solve(const Win *w, const Normal *nm, const double *regp, const double *regl, double *xp, double *xl) {
    const int P = nm->P, M = nm->M, st = nm->stride;
    double *A = (double *)malloc(sizeof(double) * ((size_t)P * P + P + M)), *gr = A + (size_t)P * P, *wl = gr + P;
    memcpy(A, nm->H, sizeof(double) * (size_t)P * P);
    memcpy(gr, nm->g, sizeof(double) * P);
    for (int l = 0; l < M; ++l) {
        wl[l] = 0;
        if (!lm_free(w, l)) continue;
        const double hw = 1.0 / (nm->Hll[l] + regl[l]);
        wl[l] = hw;
        const double *h = nm->Hpl + (size_t)l * P;
        for (int i = 0; i < P; ++i) {
            if (h[i] == 0.0) continue;
            const double hi = h[i] * hw;
            gr[i] -= hi * nm->gl[l];
            double *Ai = A + (size_t)i * P;
            for (int j = 0; j < P; ++j) Ai[j] -= hi * h[j];
        }
    }
    for (int i = 0; i < P; ++i) A[(size_t)i * P + i] += regp[i];
    for (int i = 0; i < P; ++i) if (!pose_free(w, st, i)) {
        for (int j = 0; j < P; ++j) { A[(size_t)i * P + j] = 0; A[(size_t)j * P + i] = 0; }
        A[(size_t)i * P + i] = 1; gr[i] = 0;
    }
    int rc = 0;
    for (int k = 0; k < P && rc == 0; ++k) {                                           /* dense Cholesky */
        double d = A[(size_t)k * P + k];
        for (int m = 0; m < k; ++m) d -= A[(size_t)k * P + m] * A[(size_t)k * P + m];
        if (!(d > 0) || !isfinite(d)) { rc = -1; break; }
        d = sqrt(d); A[(size_t)k * P + k] = d;
        for (int i = k + 1; i < P; ++i) {
            double s = A[(size_t)i * P + k];
            const double *Ai = A + (size_t)i * P, *Ak = A + (size_t)k * P;
            for (int m = 0; m < k; ++m) s -= Ai[m] * Ak[m];
            A[(size_t)i * P + k] = s / d;
        }
    }
    if (rc == 0) {
        for (int i = 0; i < P; ++i) { double s = gr[i]; for (int m = 0; m < i; ++m) s -= A[(size_t)i * P + m] * xp[m]; xp[i] = s / A[(size_t)i * P + i]; }
        for (int i = P - 1; i >= 0; --i) { double s = xp[i]; for (int m = i + 1; m < P; ++m) s -= A[(size_t)m * P + i] * xp[m]; xp[i] = s / A[(size_t)i * P + i]; }
        for (int l = 0; l < M; ++l) {
            if (!lm_free(w, l)) { xl[l] = 0; continue; }
            const double *h = nm->Hpl + (size_t)l * P;
            double s = nm->gl[l];
            for (int i = 0; i < P; ++i) s -= h[i] * xp[i];
            xl[l] = s * wl[l];
        }
    }
    free(A);
    return rc;
}

/* v^T H v over the full (unreduced) system; vp [P] (0 on fixed coordinates), vl [M] */
static double quad_form(const Normal *nm, const double *vp, const double *vl) {
    const int P = nm->P, M = nm->M;
    double s = 0;
    for (int i = 0; i < P; ++i) { if (vp[i] == 0.0) continue; const double *Hi = nm->H + (size_t)i * P; double t = 0; for (int j = 0; j < P; ++j) t += Hi[j] * vp[j]; s += vp[i] * t; }
    for (int l = 0; l < M; ++l) { if (vl[l] == 0.0) continue; const double *h = nm->Hpl + (size_t)l * P; double t = 0; for (int i = 0; i < P; ++i) t += h[i] * vp[i]; s += 2.0 * vl[l] * t + nm->Hll[l] * vl[l] * vl[l]; }
    return s;
}

static void apply_step(const Win *w, int stride, const double *frames, const double *rho, const double *dxp, const double *dxl,
                       double *frames_o, double *rho_o) {
    for (int f = 0; f < w->n_frames; ++f) {
        const double *d = dxp + stride * f;
        quat_plus(frames + 16 * f, d, frames_o + 16 * f);
        for (int k = 0; k < 3; ++k) frames_o[16 * f + 4 + k] = frames[16 * f + 4 + k] + d[3 + k];
        for (int k = 0; k < 9; ++k) frames_o[16 * f + 7 + k] = frames[16 * f + 7 + k] + (stride == 15 ? d[6 + k] : 0.0);
    }
    for (int l = 0; l < w->n_landmarks; ++l) rho_o[l] = rho[l] + dxl[l];
}

/* ------------------------------------------------------------------ entry points */
/* One regularised GN iteration at s (not modified): dx [15 N + M] in the ABI's layout, cost, candidate cost. */
int ba_oracle_gn_step(const Win *w, const pvio_b200_state *s, double mu, double *dx, double *cost_out, double *new_cost_out) {
    const int N = w->n_frames, M = w->n_landmarks;
    Normal nm;
    normal_alloc(&nm, N, M, w->use_inertial);
    assemble(w, s->frames, s->inv_depth, 0, &nm);
    const int P = nm.P, st = nm.stride;
    double *regp = (double *)calloc((size_t)2 * P + 3 * (size_t)M + 16 * (size_t)N, 8), *xp = regp + P, *regl = xp + P, *xl = regl + M, *rc_ = xl + M, *fc = rc_ + M;
    for (int i = 0; i < P; ++i) { const double h = nm.H[(size_t)i * P + i], sc = 1.0 / (1.0 + sqrt(h > 0 ? h : 0)); regp[i] = mu * clampd(sc * sc * h, 1e-6, 1e32) / (sc * sc); }
    for (int l = 0; l < M; ++l) { const double h = nm.Hll[l], sc = 1.0 / (1.0 + sqrt(h > 0 ? h : 0)); regl[l] = mu * clampd(sc * sc * h, 1e-6, 1e32) / (sc * sc); }
    const int rc = schur_solve(w, &nm, regp, regl, xp, xl);
    memset(dx, 0, sizeof(double) * (15 * (size_t)N + M));
    if (rc == 0) {
        for (int i = 0; i < P; ++i) { xp[i] = -xp[i]; dx[15 * (i / st) + i % st] = xp[i]; }
        for (int l = 0; l < M; ++l) { xl[l] = -xl[l]; dx[15 * N + l] = xl[l]; }
        apply_step(w, st, s->frames, s->inv_depth, xp, xl, fc, rc_);
        if (new_cost_out) *new_cost_out = window_cost(w, fc, rc_, 0);
    }
    if (cost_out) *cost_out = nm.cost;
    free(regp); normal_free(&nm);
    return rc;
}

/* ceres::Solve as PVIO configures it: the trust_region() / solve() of oracle/ba_oracle.py with the linear algebra done by
 * landmark elimination.  s is updated in place.  alias_bias: quirk Q1 (see ba_oracle.py solve()). */
int ba_oracle_solve(const Win *w, pvio_b200_state *s, const pvio_b200_options *opt, pvio_b200_summary *sm) {
    const int N = w->n_frames, M = w->n_landmarks, max_iter = opt ? opt->max_iterations : 10;
    const int alias = (!opt || opt->alias_bias) && w->use_inertial;
    double radius = (opt && opt->initial_trust_region_radius > 0) ? opt->initial_trust_region_radius : 1e4, mu = 1e-8;
    Normal nm;
    normal_alloc(&nm, N, M, w->use_inertial);
    const int P = nm.P, st = nm.stride, D = P + M;
    double *buf = (double *)calloc((size_t)10 * D + 4 * 16 * (size_t)N + 2 * (size_t)M, 8);
    double *scale = buf, *diag = scale + D, *grad = diag + D, *gn = grad + D, *step = gn + D, *reg = step + D, *x = reg + D, *tmp = x + D, *tmp2 = tmp + D, *dxv = tmp2 + D;
    double *fx = dxv + D, *fc = fx + 16 * N, *lin = fc + 16 * N, *rx = lin + 16 * N + 16 * N, *rcand = rx + M;
    memcpy(fx, s->frames, sizeof(double) * 16 * N); memcpy(rx, s->inv_depth, sizeof(double) * M);
    memcpy(lin, fx, sizeof(double) * 16 * N);
#define ISFREE(i) ((i) < P ? pose_free(w, st, (i)) : lm_free(w, (i) - P))
#define GVEC(i) ((i) < P ? nm.g[(i)] : nm.gl[(i) - P])
#define HDIAG(i) ((i) < P ? nm.H[(size_t)(i) * P + (i)] : nm.Hll[(i) - P])
    assemble(w, fx, rx, alias ? lin : 0, &nm);
    double cost = nm.cost;
    for (int i = 0; i < D; ++i) { const double h = HDIAG(i); scale[i] = 1.0 / (1.0 + sqrt(h > 0 ? h : 0)); }
    pvio_b200_summary S;
    memset(&S, 0, sizeof(S));
    S.initial_cost = cost; S.termination = PVIO_B200_TERM_NO_CONVERGENCE; S.usable = 1;
    double gmax = 0;
    for (int i = 0; i < D; ++i) if (ISFREE(i) && fabs(GVEC(i)) > gmax) gmax = fabs(GVEC(i));
    int it = 0, reuse = 0, rc = 0;
    double alpha = 0, gn_norm = 0, g_norm = 0;
    if (gmax <= 1e-10) { S.termination = PVIO_B200_TERM_CONVERGENCE; goto done; }
    for (;;) {
        if (it >= max_iter) break;
        ++it;
        if (!reuse) {
            for (int i = 0; i < D; ++i) {
                if (!ISFREE(i)) { diag[i] = 1; grad[i] = 0; tmp[i] = 0; continue; }
                diag[i] = sqrt(clampd(scale[i] * scale[i] * HDIAG(i), 1e-6, 1e32));
                grad[i] = GVEC(i) * scale[i] / diag[i];
                tmp[i] = scale[i] * grad[i] / diag[i];                                 /* unscaled direction of sg */
            }
            double g2 = 0; for (int i = 0; i < D; ++i) g2 += grad[i] * grad[i];
            alpha = g2 / quad_form(&nm, tmp, tmp + P);
            for (;;) {
                for (int i = 0; i < D; ++i) reg[i] = mu * diag[i] * diag[i] / (scale[i] * scale[i]);
                rc = schur_solve(w, &nm, reg, reg + P, x, x + P);
                int ok = rc == 0;
                if (ok) for (int i = 0; i < D; ++i) if (!isfinite(x[i])) ok = 0;
                if (ok) break;
                mu *= 10.0;
                if (mu > 1.0) { S.termination = PVIO_B200_TERM_FAILURE; S.usable = 0; goto done; }
            }
            for (int i = 0; i < D; ++i) gn[i] = ISFREE(i) ? -diag[i] * x[i] / scale[i] : 0.0;
        }
        gn_norm = 0; g_norm = 0;
        for (int i = 0; i < D; ++i) { gn_norm += gn[i] * gn[i]; g_norm += grad[i] * grad[i]; }
        gn_norm = sqrt(gn_norm); g_norm = sqrt(g_norm);
        double step_norm;
        if (gn_norm <= radius) { memcpy(step, gn, sizeof(double) * D); step_norm = gn_norm; }
        else if (g_norm * alpha >= radius) { for (int i = 0; i < D; ++i) step[i] = -(radius / g_norm) * grad[i]; step_norm = radius; }
        else {
            double gdg = 0; for (int i = 0; i < D; ++i) gdg += grad[i] * gn[i];
            const double b_dot_a = -alpha * gdg, a2 = (alpha * g_norm) * (alpha * g_norm), bma2 = a2 - 2.0 * b_dot_a + gn_norm * gn_norm;
            const double c = b_dot_a - a2, d = sqrt(c * c + bma2 * (radius * radius - a2));
            const double beta = c <= 0 ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
            for (int i = 0; i < D; ++i) step[i] = (-alpha * (1.0 - beta)) * grad[i] + beta * gn[i];
            step_norm = radius;
        }
        double sg = 0;
        for (int i = 0; i < D; ++i) { dxv[i] = ISFREE(i) ? step[i] / diag[i] * scale[i] : 0.0; sg += dxv[i] * GVEC(i); }
        const double model_change = -(sg + 0.5 * quad_form(&nm, dxv, dxv + P));
        if (model_change < 0) { radius *= 0.5; reuse = 1; continue; }                   /* invalid step */
        apply_step(w, st, fx, rx, dxv, dxv + P, fc, rcand);
        const double cand_cost = window_cost(w, fc, rcand, alias ? lin : 0);
        double xn = 0, dn = 0;                                                          /* ambient norms (q has 4 coordinates) */
        for (int f = 0; f < N; ++f) {
            const int lo = (w->frame_fixed && w->frame_fixed[f]) ? 7 : 0, hi = w->use_inertial ? 16 : 7;
            for (int k = lo; k < hi; ++k) { xn += fx[16 * f + k] * fx[16 * f + k]; const double e = fc[16 * f + k] - fx[16 * f + k]; dn += e * e; }
        }
        for (int l = 0; l < M; ++l) { xn += rx[l] * rx[l]; const double e = rcand[l] - rx[l]; dn += e * e; }
        if (sqrt(dn) <= 1e-8 * (sqrt(xn) + 1e-8)) { S.termination = PVIO_B200_TERM_CONVERGENCE; break; }
        if (fabs(cost - cand_cost) <= 1e-6 * cost) { S.termination = PVIO_B200_TERM_CONVERGENCE; break; }
        const double rel = (cost - cand_cost) / model_change;
        if (rel > 1e-3) {
            memcpy(fx, fc, sizeof(double) * 16 * N); memcpy(rx, rcand, sizeof(double) * M);
            ++S.accepted_steps;
            if (alias) memcpy(lin, fx, sizeof(double) * 16 * N);
            assemble(w, fx, rx, alias ? lin : 0, &nm);
            cost = nm.cost;
            if (rel < 0.25) radius *= 0.5;
            if (rel > 0.75) radius = radius > 3.0 * step_norm ? radius : 3.0 * step_norm;
            mu = 2.0 * mu / 10.0 > 1e-8 ? 2.0 * mu / 10.0 : 1e-8;
            reuse = 0;
            gmax = 0;
            for (int i = 0; i < D; ++i) if (ISFREE(i) && fabs(GVEC(i)) > gmax) gmax = fabs(GVEC(i));
            if (gmax <= 1e-10) { S.termination = PVIO_B200_TERM_CONVERGENCE; break; }
        } else { radius *= 0.5; reuse = 1; }
        if (radius <= 1e-32) { S.termination = PVIO_B200_TERM_CONVERGENCE; break; }
    }
done:
    S.iterations = it; S.final_cost = cost; S.final_radius = radius; S.final_mu = mu;
    memcpy(s->frames, fx, sizeof(double) * 16 * N); memcpy(s->inv_depth, rx, sizeof(double) * M);
    if (sm) *sm = S;
    free(buf); normal_free(&nm);
#undef ISFREE
#undef GVEC
#undef HDIAG
    return 0;
}

/* symmetric eigen-decomposition: Householder tridiagonalisation + implicit QL (EISPACK tred2 / tql2).
 * a [n][n] in: matrix, out: eigenvectors as columns; d [n] eigenvalues; e [n] scratch. */
static void tred2(double *a, int n, double *d, double *e) {
    for (int i = n - 1; i > 0; --i) {
        const int l = i - 1;
        double h = 0, scale = 0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += fabs(a[(size_t)i * n + k]);
            if (scale == 0.0) e[i] = a[(size_t)i * n + l];
            else {
                for (int k = 0; k <= l; ++k) { a[(size_t)i * n + k] /= scale; h += a[(size_t)i * n + k] * a[(size_t)i * n + k]; }
                double f = a[(size_t)i * n + l], g = f >= 0 ? -sqrt(h) : sqrt(h);
                e[i] = scale * g; h -= f * g; a[(size_t)i * n + l] = f - g; f = 0;
                for (int j = 0; j <= l; ++j) {
                    a[(size_t)j * n + i] = a[(size_t)i * n + j] / h; g = 0;
                    for (int k = 0; k <= j; ++k) g += a[(size_t)j * n + k] * a[(size_t)i * n + k];
                    for (int k = j + 1; k <= l; ++k) g += a[(size_t)k * n + j] * a[(size_t)i * n + k];
                    e[j] = g / h; f += e[j] * a[(size_t)i * n + j];
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = a[(size_t)i * n + j]; e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) a[(size_t)j * n + k] -= f * e[k] + g * a[(size_t)i * n + k];
                }
            }
        } else e[i] = a[(size_t)i * n + l];
        d[i] = h;
    }
    d[0] = 0; e[0] = 0;
    for (int i = 0; i < n; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0) for (int j = 0; j <= l; ++j) {
            double g = 0;
            for (int k = 0; k <= l; ++k) g += a[(size_t)i * n + k] * a[(size_t)k * n + j];
            for (int k = 0; k <= l; ++k) a[(size_t)k * n + j] -= g * a[(size_t)k * n + i];
        }
        d[i] = a[(size_t)i * n + i]; a[(size_t)i * n + i] = 1;
        for (int j = 0; j <= l; ++j) a[(size_t)j * n + i] = a[(size_t)i * n + j] = 0;
    }
}
static int tql2(double *d, double *e, int n, double *z) {
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) { const double dd = fabs(d[m]) + fabs(d[m + 1]); if (fabs(e[m]) <= 2.3e-16 * dd) break; }
            if (m != l) {
                if (iter++ == 60) return -1;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]), r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
                double s = 1, c = 1, p = 0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i], b = c * e[i];
                    e[i + 1] = r = hypot(f, g);
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0; break; }
                    s = f / r; c = g / r; g = d[i + 1] - p; r = (d[i] - g) * s + 2.0 * c * b; d[i + 1] = g + (p = s * r); g = c * r - b;
                    for (int k = 0; k < n; ++k) { f = z[(size_t)k * n + i + 1]; z[(size_t)k * n + i + 1] = s * z[(size_t)k * n + i] + c * f; z[(size_t)k * n + i] = c * z[(size_t)k * n + i] - s * f; }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0;
            }
        } while (m != l);
    }
    return 0;
}

/* BundleAdjustor::marginalize_frame (bundle_adjustor.cpp:348-599): S [dk][dk], e [dk], optional Hk, bk; dk = 15 (N - 1) */
int ba_oracle_marginalize(const Win *w, const pvio_b200_state *s, int index, double *S_out, double *e_out, double *Hk_out, double *bk_out) {
    const int N = w->n_frames, M = w->n_landmarks, n = 15 * N, dk = n - 15;
    const double *frames = s->frames, *rho = s->inv_depth;
    double *H = (double *)calloc((size_t)n * n + n + (size_t)dk * dk + dk + 2 * (size_t)dk, 8), *b = H + (size_t)n * n, *Hk = b + n, *bk = Hk + (size_t)dk * dk, *lam = bk + dk, *ev = lam + dk;
    double Rcs[9]; q2m(w->cam_q_cs, Rcs);
    Pose Ps[PVIO_B200_MAX_FRAMES];
    for (int f = 0; f < N; ++f) { q2m(frames + 16 * f, Ps[f].R); memcpy(Ps[f].p, frames + 16 * f + 4, 24); }
    if (w->n_prior > 0) {                                                               /* prior :369-413 */
        const int np = w->n_prior, d = 15 * np;
        double *r = (double *)malloc(sizeof(double) * (2 * (size_t)d + 9 * (size_t)np + (size_t)d * d)), *r0 = r + d, *Jri = r0 + d, *J = Jri + 9 * np;
        prior_eval(w, frames, r, r0, Jri);
        for (int i = 0; i < d; ++i) for (int c = 0; c < d; ++c) {
            const int fc = c / 15, cc = c % 15;
            double sv;
            if (cc < 3) { sv = 0; for (int k = 0; k < 3; ++k) sv += w->prior_S[(size_t)i * d + 15 * fc + k] * Jri[9 * fc + 3 * k + cc]; }
            else sv = w->prior_S[(size_t)i * d + c];
            J[(size_t)i * d + c] = sv;
        }
        for (int a = 0; a < d; ++a) {
            const int ga = 15 * w->prior_frames[a / 15] + a % 15;
            double sv = 0; for (int i = 0; i < d; ++i) sv += J[(size_t)i * d + a] * r[i];
            b[ga] += sv;
            for (int c = 0; c < d; ++c) { double h = 0; for (int i = 0; i < d; ++i) h += J[(size_t)i * d + a] * J[(size_t)i * d + c]; H[(size_t)ga * n + 15 * w->prior_frames[c / 15] + c % 15] += h; }
        }
        free(r);
    }
    for (int m = 0; m < w->n_imu; ++m) {                                                /* IMU factors adjacent to the victim :416-450 */
        const int fi = w->imu_frame_i[m], fj = w->imu_frame_j[m];
        if (fj != index && fj != index + 1) continue;
        double r[15], J[450];
        imu_eval(frames + 16 * fi, frames + 16 * fj, w->imu_data + (size_t)m * PVIO_B200_IMU_STRIDE, frames + 16 * fi + 10, frames + 16 * fi + 13,
                 w->imu_q_cs, w->imu_p_cs, r, J);                                       /* the functor reads the bias point from the parameter: dbg = dba = 0 */
        for (int a = 0; a < 30; ++a) {
            const int ga = a < 15 ? 15 * fi + a : 15 * fj + a - 15;
            double sv = 0; for (int k = 0; k < 15; ++k) sv += J[30 * k + a] * r[k];
            b[ga] += sv;
            for (int c = 0; c < 30; ++c) { double h = 0; for (int k = 0; k < 15; ++k) h += J[30 * k + a] * J[30 * k + c]; H[(size_t)ga * n + (c < 15 ? 15 * fi + c : 15 * fj + c - 15)] += h; }
        }
    }
    for (int l = 0; l < M; ++l) {                                                       /* reprojection :453-533 + landmark Schur :536-545 */
        if (!w->lm_in_victim || !w->lm_in_victim[l]) continue;
        const int a = w->lm_anchor[l];
        double mat = 0, vec = 0, h[PVIO_B200_MAX_FRAMES][6];
        unsigned touched = 0;
        memset(h, 0, sizeof(h));
        for (int k = w->lm_obs_begin[l]; k < w->lm_obs_begin[l + 1]; ++k) {
            const int t = w->obs_frame[k];
            double r[2], J[26];
            reproj(&Ps[t], &Ps[a], rho[l], w->obs_z + 2 * k, w->lm_z_ref + 2 * l, Rcs, w->cam_p_cs, w->sqrt_inv_cov, r, J);
            const int col[2] = {15 * t, 15 * a}, fr[2] = {t, a};
            for (int bi = 0; bi < 2; ++bi) for (int i = 0; i < 6; ++i) {
                const double j0 = J[6 * bi + i], j1 = J[13 + 6 * bi + i];
                b[col[bi] + i] += j0 * r[0] + j1 * r[1];
                h[fr[bi]][i] += j0 * J[12] + j1 * J[25];
                for (int bj = 0; bj < 2; ++bj) for (int j = 0; j < 6; ++j) H[(size_t)(col[bi] + i) * n + col[bj] + j] += j0 * J[6 * bj + j] + j1 * J[13 + 6 * bj + j];
            }
            touched |= (1u << t) | (1u << a);
            mat += J[12] * J[12] + J[25] * J[25]; vec += J[12] * r[0] + J[25] * r[1];
        }
        if (mat == 0.0 || !isfinite(1.0 / mat)) continue;
        const double inv = 1.0 / mat;
        for (int fi = 0; fi < N; ++fi) if ((touched >> fi) & 1u) for (int i = 0; i < 6; ++i) {
            b[15 * fi + i] -= h[fi][i] * inv * vec;
            for (int fj = 0; fj < N; ++fj) if ((touched >> fj) & 1u) for (int j = 0; j < 6; ++j) H[(size_t)(15 * fi + i) * n + 15 * fj + j] -= h[fi][i] * h[fj][j] * inv;
        }
    }
    /* frame Schur :547-581 with the explicit 15x15 inverse (Gauss-Jordan, partial pivoting) */
    double Mx[15][30];
    const int v0 = 15 * index;
    for (int i = 0; i < 15; ++i) for (int j = 0; j < 30; ++j) Mx[i][j] = j < 15 ? H[(size_t)(v0 + i) * n + v0 + j] : (j - 15 == i ? 1.0 : 0.0);
    for (int c = 0; c < 15; ++c) {
        int p = c; for (int i = c + 1; i < 15; ++i) if (fabs(Mx[i][c]) > fabs(Mx[p][c])) p = i;
        if (p != c) for (int j = 0; j < 30; ++j) { const double t = Mx[c][j]; Mx[c][j] = Mx[p][j]; Mx[p][j] = t; }
        const double ip = 1.0 / Mx[c][c];
        for (int j = 0; j < 30; ++j) Mx[c][j] *= ip;
        for (int i = 0; i < 15; ++i) if (i != c) { const double f = Mx[i][c]; if (f != 0.0) for (int j = 0; j < 30; ++j) Mx[i][j] -= f * Mx[c][j]; }
    }
    for (int i = 0; i < dk; ++i) {
        const int gi = i < v0 ? i : i + 15;
        double wrow[15];
        for (int k = 0; k < 15; ++k) { double t = 0; for (int m = 0; m < 15; ++m) t += H[(size_t)gi * n + v0 + m] * Mx[m][15 + k]; wrow[k] = t; }
        for (int j = 0; j < dk; ++j) {
            const int gj = j < v0 ? j : j + 15;
            double acc = 0; for (int k = 0; k < 15; ++k) acc += wrow[k] * H[(size_t)(v0 + k) * n + gj];
            Hk[(size_t)i * dk + j] = H[(size_t)gi * n + gj] - acc;
        }
        double acc = 0; for (int k = 0; k < 15; ++k) acc += wrow[k] * b[v0 + k];
        bk[i] = b[gi] - acc;
    }
    if (index > 0 && index < N - 1) for (int i = v0; i < dk; ++i) for (int j = 0; j < v0; ++j) Hk[(size_t)i * dk + j] = Hk[(size_t)j * dk + i];   /* :572-577 */
    if (Hk_out) memcpy(Hk_out, Hk, sizeof(double) * (size_t)dk * dk);
    if (bk_out) memcpy(bk_out, bk, sizeof(double) * dk);
    int rc = 0;
    if (S_out || e_out) {                                                               /* eigen factorisation :583-590 */
        double *V = (double *)malloc(sizeof(double) * (size_t)dk * dk);
        memcpy(V, Hk, sizeof(double) * (size_t)dk * dk);
        tred2(V, dk, lam, ev);
        rc = tql2(lam, ev, dk, V);
        for (int i = 0; i < dk; ++i) {
            const int pos = lam[i] > 1e-8;
            const double sl = pos ? sqrt(lam[i]) : 0.0, il = pos ? sqrt(1.0 / lam[i]) : 0.0;
            double dot = 0;
            for (int k = 0; k < dk; ++k) { const double v = V[(size_t)k * dk + i]; if (S_out) S_out[(size_t)i * dk + k] = sl * v; dot += v * bk[k]; }
            if (e_out) e_out[i] = il * dot;
        }
        free(V);
    }
    free(H);
    return rc;
}

/* ------------------------------------------------------------------ thread pool over independent windows */
/* CPUs this process may actually use: scheduler affinity, capped by the cgroup CPU quota */
int ba_oracle_usable_cores(void) {
    cpu_set_t set;
    int n = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (sched_getaffinity(0, sizeof(set), &set) == 0) { const int c = CPU_COUNT(&set); if (c > 0) n = c; }
    FILE *f = fopen("/sys/fs/cgroup/cpu.max", "r");
    if (f) {
        char q[64]; long long period = 0;
        if (fscanf(f, "%63s %lld", q, &period) == 2 && strcmp(q, "max") != 0 && period > 0) {
            const long long quota = atoll(q);
            const int c = (int)((quota + period - 1) / period);
            if (c > 0 && c < n) n = c;
        }
        fclose(f);
    }
    return n < 1 ? 1 : n;
}

typedef struct {
    int kind;                      /* 0 gn_step, 1 solve, 2 marginalize */
    int n_windows;
    const Win *w; const pvio_b200_state *s; const pvio_b200_options *opt;
    double mu; int index;
    double *dx_all, *costs;        /* kind 0: [n][15 N + M], [n][2] */
    int64_t *iters;                /* kind 1: [n] iterations executed */
    double *final_costs;           /* kind 1: [n] */
    volatile int next;
} BatchJob;

static void *batch_worker(void *arg) {
    BatchJob *j = (BatchJob *)arg;
    const int N = j->w->n_frames, M = j->w->n_landmarks;
    double *fr = (double *)malloc(sizeof(double) * (16 * (size_t)N + M + 4 * 225 * (size_t)N * N + 64)), *rh = fr + 16 * N, *scr = rh + M;
    for (;;) {
        const int wi = __sync_fetch_and_add(&j->next, 1);
        if (wi >= j->n_windows) break;
        if (j->kind == 0) {
            ba_oracle_gn_step(j->w, j->s, j->mu, j->dx_all + (size_t)wi * (15 * N + M), j->costs + 2 * wi, j->costs + 2 * wi + 1);
        } else if (j->kind == 1) {
            memcpy(fr, j->s->frames, sizeof(double) * 16 * N); memcpy(rh, j->s->inv_depth, sizeof(double) * M);
            pvio_b200_state st = {fr, rh};
            pvio_b200_summary sm;
            ba_oracle_solve(j->w, &st, j->opt, &sm);
            if (j->iters) j->iters[wi] = sm.iterations;
            if (j->final_costs) j->final_costs[wi] = sm.final_cost;
        } else {
            const int dk = 15 * (N - 1);
            ba_oracle_marginalize(j->w, j->s, j->index, scr, scr + (size_t)dk * dk, 0, 0);
        }
    }
    free(fr);
    return 0;
}

/* n_windows independent copies of the same problem over n_threads POSIX threads (<= 0: every usable core).  Returns the
 * number of threads used. */
int ba_oracle_batch(int kind, int n_windows, int n_threads, const Win *w, const pvio_b200_state *s, const pvio_b200_options *opt,
                    double mu, int index, double *dx_all, double *costs, int64_t *iters, double *final_costs) {
    if (n_threads <= 0) n_threads = ba_oracle_usable_cores();
    if (n_threads > n_windows) n_threads = n_windows;
    if (n_threads < 1) n_threads = 1;
    BatchJob j = {kind, n_windows, w, s, opt, mu, index, dx_all, costs, iters, final_costs, 0};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
    free(th);
    return n_threads;
}
