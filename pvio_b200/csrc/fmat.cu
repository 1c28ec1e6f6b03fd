// The F-matrix outlier rejection of OpenCvImage::track_keypoints (pvio-extra/src/pvio/extra/opencv_image.cpp:121-129):
//
//     findFundamentalMat(p, q, cv::FM_RANSAC, 1.0, 0.99, mask)
//
// OpenCV (absent from /root/reference; the published algorithm of calib3d/fundam.cpp + ptsetreg.cpp, restated and pinned
// against cv2 4.13.0 in oracle/fm_oracle.py) runs a SERIAL loop: draw 7 matches, 7-point solve (1..3 models), count the
// inliers of each model over all matches, keep a model with more inliers than the best so far and shrink the iteration
// budget from the new outlier ratio.  Nothing in an iteration depends on the previous ones except that budget.
//
// Here: SPECULATIVE evaluation + serial-equivalent replay.
//   host     the sample schedule: cv::RNG((uint64)-1) and getSubset (7 distinct indices, the collinearity test of the
//            last point) restated, so the schedule IS OpenCV's -- or the schedule the caller injects
//   device   fm_hypotheses_kernel: one warp per iteration of the schedule, ALL max_iters of them at once (1000 warps):
//            lane 0 solves the 7-point problem (Hartley normalisation, null space of the 7 x 9 system by a Householder
//            QR of its transpose, the cubic det(lambda f1 + (1 - lambda) f2) = 0 with cv::solveCubic's branches and
//            root order, de-normalisation), every lane then takes matches lane, lane + 32, ...: the symmetric epipolar
//            error of computeError in double, rounded to float, compared with (float) threshold^2; inlier bit masks by
//            ballot, counts by popc.  For 8 <= n < 15 OpenCV switches to LMedS: the same kernel returns the median error
//            (element n / 2 in the order of the float bit patterns, as std::nth_element on int sees them)
//   host     the replay: walk the iterations in order with OpenCV's acceptance rule and RANSACUpdateNumIters (the
//            host's own log / pow, as OpenCV's) and stop where the serial loop would have stopped; the mask of the
//            winning (iteration, model) is the result.  12 KB of counts come back, then 128 bytes of mask.
// The result is the one the serial loop produces for that schedule, whatever the inlier ratio; the device does at most
// 1000 x 3 x n error evaluations (0.4 GFLOP at n = 400) in one wave.
//
// Compiled with -fmad=false: OpenCV's scalar double code is not contracted.
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "api_internal.h"

namespace pvio {

struct FmState {
    float *d_pts = nullptr;        // [2][cap_n][2]  p then q
    int *d_sched = nullptr;        // [cap_s][7]
    int *d_counts = nullptr;       // [cap_s][4]  nmodels, count / median bits of each model
    double *d_F = nullptr;         // [cap_s][3][9]
    unsigned *d_masks = nullptr;   // [cap_s][3][cap_w]
    float *h_pts = nullptr;        // pinned
    int *h_sched = nullptr, *h_counts = nullptr;
    double *h_F = nullptr;
    unsigned *h_mask = nullptr;
    int cap_n = 0, cap_s = 0, cap_w = 0;
};

void fm_free(Handle *h) {
    FmState *f = h->fm;
    if (!f) return;
    cudaFree(f->d_pts); cudaFree(f->d_sched); cudaFree(f->d_counts); cudaFree(f->d_F); cudaFree(f->d_masks);
    if (f->h_pts) cudaFreeHost(f->h_pts);
    if (f->h_sched) cudaFreeHost(f->h_sched);
    if (f->h_counts) cudaFreeHost(f->h_counts);
    if (f->h_F) cudaFreeHost(f->h_F);
    if (f->h_mask) cudaFreeHost(f->h_mask);
    delete f;
    h->fm = nullptr;
}

namespace {

// ---- host: OpenCV's sample schedule --------------------------------------------------------------------------------
struct CvRng {                       // cv::RNG (core/operations.hpp): multiply-with-carry
    uint64_t state;
    explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffULL) {}
    unsigned next() {
        state = (uint64_t)(unsigned)state * 4164903690U + (unsigned)(state >> 32);
        return (unsigned)state;
    }
    int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// fundam.cpp haveCollinearPoints: the LAST of `count` points against every pair before it
bool last_point_collinear(const float *pts, const int *idx, int count) {
    const int i = count - 1;
    const float xi = pts[2 * idx[i]], yi = pts[2 * idx[i] + 1];
    for (int j = 0; j < i; ++j) {
        const double dx1 = pts[2 * idx[j]] - xi, dy1 = pts[2 * idx[j] + 1] - yi;
        for (int k = 0; k < j; ++k) {
            const double dx2 = pts[2 * idx[k]] - xi, dy2 = pts[2 * idx[k] + 1] - yi;
            if (std::fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2)))
                return true;
        }
    }
    return false;
}

// ptsetreg.cpp getSubset for every iteration the loop could run: iteration `it` always consumes exactly one subset.
// Returns the number of iterations that have a subset (getSubset giving up ends the loop).
int cv_schedule(const float *p, const float *q, int n, int iters, int max_attempts, int *sched) {
    CvRng rng((uint64_t)-1);
    for (int it = 0; it < iters; ++it) {
        int *idx = sched + 7 * it;
        int attempt = 0;
        for (; attempt < max_attempts; ++attempt) {
            for (int i = 0; i < 7; ++i) {
                int v;
                bool dup;
                do {
                    v = rng.uniform(0, n);
                    dup = false;
                    for (int j = 0; j < i; ++j) dup |= idx[j] == v;
                } while (dup);
                idx[i] = v;
            }
            if (!last_point_collinear(p, idx, 7) && !last_point_collinear(q, idx, 7)) break;
        }
        if (attempt == max_attempts) return it;
    }
    return iters;
}

// ptsetreg.cpp RANSACUpdateNumIters
int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = std::fmax(p, 0.); p = std::fmin(p, 1.);
    ep = std::fmax(ep, 0.); ep = std::fmin(ep, 1.);
    double num = std::fmax(1. - p, DBL_MIN);
    double denom = 1. - std::pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = std::log(num);
    denom = std::log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)std::nearbyint(num / denom);   // cvRound
}

// ---- device ---------------------------------------------------------------------------------------------------------
// cv::solveCubic (core/mathfuncs.cpp) for double coefficients; roots in OpenCV's order
__device__ int solve_cubic(double a0, double a1, double a2, double a3, double *x) {
    const double kPi = 3.1415926535897932384626433832795;
    if (a0 == 0) {
        if (a1 == 0) {
            if (a2 == 0) return a3 == 0 ? -1 : 0;
            x[0] = -a3 / a2;
            return 1;
        }
        double d = a2 * a2 - 4 * a1 * a3;
        if (d < 0) return 0;
        d = sqrt(d);
        const double q1 = (-a2 + d) * 0.5, q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) { x[0] = q1 / a1; x[1] = a3 / q1; }
        else { x[0] = q2 / a1; x[1] = a3 / q2; }
        return d > 0 ? 2 : 1;
    }
    a0 = 1. / a0;
    a1 *= a0; a2 *= a0; a3 *= a0;
    const double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    const double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    const double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    if (d > 0) {
        const double theta = acos(R / sqrt(Qcubed));
        const double sqrtQ = sqrt(Q);
        const double t0 = -2 * sqrtQ, t1 = theta * (1. / 3), t2 = a1 * (1. / 3);
        x[0] = t0 * cos(t1) - t2;
        x[1] = t0 * cos(t1 + (2. * kPi / 3)) - t2;
        x[2] = t0 * cos(t1 + (4. * kPi / 3)) - t2;
        return 3;
    }
    if (d == 0) {
        if (R >= 0) { x[0] = -2 * pow(R, 1. / 3) - a1 / 3; x[1] = pow(R, 1. / 3) - a1 / 3; }
        else { x[0] = 2 * pow(-R, 1. / 3) - a1 / 3; x[1] = -pow(-R, 1. / 3) - a1 / 3; }
        return x[0] == x[1] ? 1 : 2;
    }
    d = sqrt(-d);
    double e = pow(d + fabs(R), 1. / 3);
    if (R > 0) e = -e;
    x[0] = (e + Q / e) - a1 * (1. / 3);
    return 1;
}

// fundam.cpp run7Point on the 7 matches idx[0..6]; models row-major into F[k][9].  Returns the model count (<= 0: none).
__device__ int seven_point(const float2 *p, const float2 *q, const int *idx, double (*F)[9]) {
    double px[7], py[7], qx[7], qy[7];
    double c1x = 0, c1y = 0, c2x = 0, c2y = 0;
    for (int i = 0; i < 7; ++i) {
        const float2 a = p[idx[i]], b = q[idx[i]];
        px[i] = a.x; py[i] = a.y; qx[i] = b.x; qy[i] = b.y;
        c1x += px[i]; c1y += py[i]; c2x += qx[i]; c2y += qy[i];
    }
    const double t = 1. / 7;
    c1x *= t; c1y *= t; c2x *= t; c2y *= t;
    double scale1 = 0, scale2 = 0;
    for (int i = 0; i < 7; ++i) {
        const double ax = px[i] - c1x, ay = py[i] - c1y, bx = qx[i] - c2x, by = qy[i] - c2y;
        scale1 += sqrt(ax * ax + ay * ay);
        scale2 += sqrt(bx * bx + by * by);
    }
    scale1 *= t; scale2 *= t;
    if (scale1 < (double)FLT_EPSILON || scale2 < (double)FLT_EPSILON) return 0;
    scale1 = sqrt(2.) / scale1;
    scale2 = sqrt(2.) / scale2;

    // M = A^T (9 x 7): column i is the equation (m2_i, 1)^T F (m1_i, 1) = 0 of match i
    double M[9][7];
    for (int i = 0; i < 7; ++i) {
        const double x0 = (px[i] - c1x) * scale1, y0 = (py[i] - c1y) * scale1;
        const double x1 = (qx[i] - c2x) * scale2, y1 = (qy[i] - c2y) * scale2;
        M[0][i] = x1 * x0; M[1][i] = x1 * y0; M[2][i] = x1;
        M[3][i] = y1 * x0; M[4][i] = y1 * y0; M[5][i] = y1;
        M[6][i] = x0; M[7][i] = y0; M[8][i] = 1;
    }
    // Householder QR of M; the null space of A is spanned by the last two columns of Q = H_0 ... H_6
    double V[7][9], beta[7];
    for (int k = 0; k < 7; ++k) {
        double nrm = 0;
        for (int r = k; r < 9; ++r) nrm += M[r][k] * M[r][k];
        nrm = sqrt(nrm);
        const double x0 = M[k][k];
        const double alpha = x0 >= 0 ? -nrm : nrm;
        double vv = 0;
        for (int r = k; r < 9; ++r) {
            const double v = r == k ? x0 - alpha : M[r][k];
            V[k][r] = v;
            vv += v * v;
        }
        beta[k] = vv > 0 ? 2. / vv : 0.;
        for (int c = k; c < 7; ++c) {
            double s = 0;
            for (int r = k; r < 9; ++r) s += V[k][r] * M[r][c];
            s *= beta[k];
            for (int r = k; r < 9; ++r) M[r][c] -= s * V[k][r];
        }
    }
    double f1[9], f2[9];
    for (int r = 0; r < 9; ++r) { f1[r] = r == 7; f2[r] = r == 8; }
    for (int k = 6; k >= 0; --k) {
        double s1 = 0, s2 = 0;
        for (int r = k; r < 9; ++r) { s1 += V[k][r] * f1[r]; s2 += V[k][r] * f2[r]; }
        s1 *= beta[k]; s2 *= beta[k];
        for (int r = k; r < 9; ++r) { f1[r] -= s1 * V[k][r]; f2[r] -= s2 * V[k][r]; }
    }

    // f ~ lambda f1 + (1 - lambda) f2, det f = 0: the cubic in lambda (run7Point's expansion)
    for (int i = 0; i < 9; ++i) f1[i] -= f2[i];
    double t0 = f2[4] * f2[8] - f2[5] * f2[7];
    double t1 = f2[3] * f2[8] - f2[5] * f2[6];
    double t2 = f2[3] * f2[7] - f2[4] * f2[6];
    double c[4];
    c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
    c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
           f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) + f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
           f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) + f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
           f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) + f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    t0 = f1[4] * f1[8] - f1[5] * f1[7];
    t1 = f1[3] * f1[8] - f1[5] * f1[6];
    t2 = f1[3] * f1[7] - f1[4] * f1[6];
    c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
           f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) + f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
           f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) + f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
           f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) + f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    double roots[3] = {0, 0, 0};
    const int n = solve_cubic(c[0], c[1], c[2], c[3], roots);
    if (n < 1 || n > 3) return n;

    for (int k = 0; k < n; ++k) {
        double lambda = roots[k], mu = 1.;
        const double s = f1[8] * roots[k] + f2[8];
        double f[9];
        if (fabs(s) > DBL_EPSILON) { mu = 1. / s; lambda *= mu; f[8] = 1.; }
        else f[8] = 0.;
        for (int i = 0; i < 8; ++i) f[i] = f1[i] * lambda + f2[i] * mu;
        // F = T2^T f T1, T = [s 0 -s cx; 0 s -s cy; 0 0 1]
        double G[9];                                    // G = f T1
        for (int r = 0; r < 3; ++r) {
            G[3 * r + 0] = f[3 * r + 0] * scale1;
            G[3 * r + 1] = f[3 * r + 1] * scale1;
            G[3 * r + 2] = f[3 * r + 0] * (-scale1 * c1x) + f[3 * r + 1] * (-scale1 * c1y) + f[3 * r + 2];
        }
        double *o = F[k];
        for (int cc = 0; cc < 3; ++cc) {
            o[0 + cc] = scale2 * G[0 + cc];
            o[3 + cc] = scale2 * G[3 + cc];
            o[6 + cc] = (-scale2 * c2x) * G[0 + cc] + (-scale2 * c2y) * G[3 + cc] + G[6 + cc];
        }
        if (fabs(o[8]) > (double)FLT_EPSILON) {
            const double inv = 1. / o[8];
            for (int i = 0; i < 9; ++i) o[i] *= inv;
        }
    }
    return n;
}

// FMEstimatorCallback::computeError for one match
__device__ __forceinline__ float fm_error(const double *F, float2 m1, float2 m2) {
    double a = F[0] * m1.x + F[1] * m1.y + F[2];
    double b = F[3] * m1.x + F[4] * m1.y + F[5];
    double c = F[6] * m1.x + F[7] * m1.y + F[8];
    const double s2 = 1. / (a * a + b * b);
    const double d2 = m2.x * a + m2.y * b + c;
    a = F[0] * m2.x + F[3] * m2.y + F[6];
    b = F[1] * m2.x + F[4] * m2.y + F[7];
    c = F[2] * m2.x + F[5] * m2.y + F[8];
    const double s1 = 1. / (a * a + b * b);
    const double d1 = m1.x * a + m1.y * b + c;
    return (float)fmax(d1 * d1 * s1, d2 * d2 * s2);        // std::max(a, b): b if a < b (NaN handling as OpenCV's is not reproduced)
}

constexpr int kWarpsPerCta = 4;

// One warp per iteration of the schedule.  lmeds == 0: counts[it] = {nmodels, inliers of model 0, 1, 2}, masks[it][k][nw];
// lmeds == 1 (n < 32): counts[it] = {nmodels, float bits of the median error of model 0, 1, 2}.
__global__ void __launch_bounds__(32 * kWarpsPerCta) fm_hypotheses_kernel(const float2 *p, const float2 *q, int n, const int *sched,
                                                                          int iters, float thr2, int lmeds, int nw, int *counts,
                                                                          double *Fout, unsigned *masks) {
    __shared__ double sF[kWarpsPerCta][3][9];
    __shared__ int sN[kWarpsPerCta];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int it = blockIdx.x * kWarpsPerCta + warp;
    if (it >= iters) return;
    if (lane == 0) {
        int idx[7];
        for (int i = 0; i < 7; ++i) idx[i] = sched[7 * it + i];
        int nm = 0;
        if (idx[0] >= 0) nm = seven_point(p, q, idx, sF[warp]);
        sN[warp] = nm;
        counts[4 * it] = nm;
        for (int k = 0; k < 3; ++k)
            for (int i = 0; i < 9; ++i) Fout[(size_t)(3 * it + k) * 9 + i] = k < nm ? sF[warp][k][i] : 0.;
    }
    __syncwarp();
    const int nm = sN[warp];
    for (int k = 0; k < 3; ++k) {
        if (k >= nm) {
            if (lane == 0) counts[4 * it + 1 + k] = lmeds ? 0x7f800000 : 0;
            continue;
        }
        double F[9];
        for (int i = 0; i < 9; ++i) F[i] = sF[warp][k][i];
        if (!lmeds) {
            int good = 0;
            for (int w = 0; w < nw; ++w) {
                const int i = 32 * w + lane;
                bool in = false;
                if (i < n) in = fm_error(F, p[i], q[i]) <= thr2;
                const unsigned b = __ballot_sync(0xffffffffu, in);
                good += __popc(b);
                if (lane == 0) masks[((size_t)(3 * it + k)) * nw + w] = b;
            }
            if (lane == 0) counts[4 * it + 1 + k] = good;
        } else {
            // std::nth_element(int bit patterns)[n / 2]: rank by (bits, index)
            const int mine = lane < n ? __float_as_int(fm_error(F, p[lane], q[lane])) : 0x7fffffff;
            int rank = 0;
            for (int j = 0; j < 32; ++j) {
                const int other = __shfl_sync(0xffffffffu, mine, j);
                rank += (other < mine) || (other == mine && j < lane);
            }
            if (lane < n && rank == n / 2) counts[4 * it + 1 + k] = mine;
        }
    }
}

// mask of one model under one threshold (LMedS: the inliers of the best model under sigma; n == 7: not needed)
__global__ void fm_mask_kernel(const float2 *p, const float2 *q, int n, const double *Fm, float thr2, unsigned *mask) {
    double F[9];
    for (int i = 0; i < 9; ++i) F[i] = Fm[i];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (i < n) in = fm_error(F, p[i], q[i]) <= thr2;
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0) mask[i >> 5] = b;
}

int fm_reserve(Handle *h, int n, int s) {
    if (!h->fm) h->fm = new FmState();
    FmState *f = h->fm;
    if (n > f->cap_n || s > f->cap_s) {
        const int cn = std::max(n, std::max(f->cap_n, 1024)), cs = std::max(s, std::max(f->cap_s, 1000));
        const int cw = (cn + 31) / 32;
        CK(h, cudaStreamSynchronize(h->stream));
        cudaFree(f->d_pts); cudaFree(f->d_sched); cudaFree(f->d_counts); cudaFree(f->d_F); cudaFree(f->d_masks);
        if (f->h_pts) cudaFreeHost(f->h_pts);
        if (f->h_sched) cudaFreeHost(f->h_sched);
        if (f->h_counts) cudaFreeHost(f->h_counts);
        if (f->h_F) cudaFreeHost(f->h_F);
        if (f->h_mask) cudaFreeHost(f->h_mask);
        f->d_pts = nullptr; f->d_sched = nullptr; f->d_counts = nullptr; f->d_F = nullptr; f->d_masks = nullptr;
        f->h_pts = nullptr; f->h_sched = nullptr; f->h_counts = nullptr; f->h_F = nullptr; f->h_mask = nullptr;
        f->cap_n = f->cap_s = f->cap_w = 0;
        CK(h, cudaMalloc(&f->d_pts, sizeof(float) * 4 * cn));
        CK(h, cudaMalloc(&f->d_sched, sizeof(int) * 7 * cs));
        CK(h, cudaMalloc(&f->d_counts, sizeof(int) * 4 * cs));
        CK(h, cudaMalloc(&f->d_F, sizeof(double) * 27 * cs));
        CK(h, cudaMalloc(&f->d_masks, sizeof(unsigned) * 3 * (size_t)cs * cw));
        CK(h, cudaMallocHost(&f->h_pts, sizeof(float) * 4 * cn));
        CK(h, cudaMallocHost(&f->h_sched, sizeof(int) * 7 * cs));
        CK(h, cudaMallocHost(&f->h_counts, sizeof(int) * 4 * cs));
        CK(h, cudaMallocHost(&f->h_F, sizeof(double) * 27));
        CK(h, cudaMallocHost(&f->h_mask, sizeof(unsigned) * cw));
        f->cap_n = cn; f->cap_s = cs; f->cap_w = cw;
    }
    return 0;
}

}  // namespace

// host-only: the subsets OpenCV would draw (exported as pvio_b200_fm_sample_schedule)
int fm_cv_schedule(int n, const float *p, const float *q, int iters, int32_t *schedule) {
    return cv_schedule(p, q, n, iters, n >= 15 ? 10000 : 1000, schedule);
}

int fm_ransac_impl(Handle *h, int n, const float *p, const float *q, double threshold, double confidence, int max_iters,
                   const int32_t *schedule, int n_schedule, uint8_t *mask, double *F_out, int32_t *info) {
    int inf[4] = {0, -1, -1, 0};              // iterations run, winning iteration, winning model, method (1 RANSAC, 2 LMedS, 3 seven points)
    auto finish = [&](int rc) { if (info) std::memcpy(info, inf, sizeof(inf)); return rc; };
    if (F_out) std::memset(F_out, 0, 9 * sizeof(double));
    if (n < 7) {                              // findFundamentalMat returns an empty matrix and leaves the mask alone
        if (n > 0) std::memset(mask, 0, n);
        return finish(0);
    }
    if (threshold <= 0) threshold = 3;
    if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
    if (max_iters < 1) max_iters = 1;
    const bool seven = n == 7, ransac = n >= 15;
    const int budget = seven ? 1 : ransac ? max_iters : ransac_update_num_iters(confidence, 0.45, 7, max_iters);
    TRY(fm_reserve(h, n, budget));
    FmState *f = h->fm;
    std::memcpy(f->h_pts, p, sizeof(float) * 2 * n);
    std::memcpy(f->h_pts + 2 * n, q, sizeof(float) * 2 * n);
    int have = budget;                        // iterations that have a subset
    if (seven) {
        for (int i = 0; i < 7; ++i) f->h_sched[i] = i;
    } else if (schedule) {
        if (n_schedule < 1) return fail(h, PVIO_B200_EINVAL, "find_fundamental_mask: empty schedule");
        have = std::min(budget, n_schedule);
        for (int it = 0; it < have; ++it) {
            if (schedule[7 * it] < 0) { have = it; break; }          // a row starting with -1: getSubset gave up there
            for (int i = 0; i < 7; ++i) {
                const int v = schedule[7 * it + i];
                if (v < 0 || v >= n) return fail(h, PVIO_B200_EINVAL, "find_fundamental_mask: schedule index out of range");
                f->h_sched[7 * it + i] = v;
            }
        }
    } else {
        have = cv_schedule(p, q, n, budget, ransac ? 10000 : 1000, f->h_sched);
    }
    if (have == 0) {                          // getSubset failed at iteration 0: run() returns false, the mask stays as created
        std::memset(mask, 0, n);
        return finish(0);
    }
    cudaStream_t st = h->stream;
    const int nw = (n + 31) / 32;
    float2 *dp = (float2 *)f->d_pts, *dq = (float2 *)(f->d_pts + 2 * n);
    CK(h, cudaMemcpyAsync(f->d_pts, f->h_pts, sizeof(float) * 4 * n, cudaMemcpyHostToDevice, st));
    CK(h, cudaMemcpyAsync(f->d_sched, f->h_sched, sizeof(int) * 7 * have, cudaMemcpyHostToDevice, st));
    const float thr2 = (float)(threshold * threshold);
    fm_hypotheses_kernel<<<(have + kWarpsPerCta - 1) / kWarpsPerCta, 32 * kWarpsPerCta, 0, st>>>(
        dp, dq, n, f->d_sched, have, thr2, (!ransac && !seven) ? 1 : 0, nw, f->d_counts, f->d_F, f->d_masks);
    h->launches++;
    CK(h, cudaGetLastError());
    CK(h, cudaMemcpyAsync(f->h_counts, f->d_counts, sizeof(int) * 4 * have, cudaMemcpyDeviceToHost, st));
    CK(h, cudaStreamSynchronize(st));

    int best_it = -1, best_k = -1, it = 0;
    if (seven) {
        inf[3] = 3;
        std::memset(mask, 1, n);              // fundam.cpp: the mask of the 7-point / 8-point call is all ones
        if (f->h_counts[0] >= 1) { best_it = 0; best_k = 0; }
        it = 1;
    } else if (ransac) {
        inf[3] = 1;
        int niters = budget, max_good = 0;
        for (; it < niters && it < have; ++it) {
            const int nm = f->h_counts[4 * it];
            for (int k = 0; k < nm && k < 3; ++k) {
                const int good = f->h_counts[4 * it + 1 + k];
                if (good > std::max(max_good, 6)) {
                    max_good = good; best_it = it; best_k = k;
                    niters = ransac_update_num_iters(confidence, (double)(n - good) / n, 7, niters);
                }
            }
        }
    } else {
        inf[3] = 2;
        double min_median = DBL_MAX;
        for (; it < have; ++it) {
            const int nm = f->h_counts[4 * it];
            for (int k = 0; k < nm && k < 3; ++k) {
                float med;
                std::memcpy(&med, &f->h_counts[4 * it + 1 + k], 4);
                if ((double)med < min_median) { min_median = med; best_it = it; best_k = k; }
            }
        }
        if (best_it >= 0) {
            double sigma = 2.5 * 1.4826 * (1 + 5. / (n - 7)) * std::sqrt(min_median);
            sigma = std::fmax(sigma, 0.001);
            fm_mask_kernel<<<(n + 31) / 32, 32, 0, st>>>(dp, dq, n, f->d_F + (size_t)(3 * best_it + best_k) * 9, (float)(sigma * sigma),
                                                        f->d_masks);
            h->launches++;
            CK(h, cudaGetLastError());
        }
    }
    inf[0] = it; inf[1] = best_it; inf[2] = best_k;
    if (best_it < 0) {
        if (!seven) std::memset(mask, 0, n);
        return finish(0);
    }
    CK(h, cudaMemcpyAsync(f->h_F, f->d_F + (size_t)(3 * best_it + best_k) * 9, sizeof(double) * 9, cudaMemcpyDeviceToHost, st));
    if (!seven) {
        const unsigned *src = ransac ? f->d_masks + (size_t)(3 * best_it + best_k) * nw : f->d_masks;
        CK(h, cudaMemcpyAsync(f->h_mask, src, sizeof(unsigned) * nw, cudaMemcpyDeviceToHost, st));
    }
    CK(h, cudaStreamSynchronize(st));
    if (!seven)
        for (int i = 0; i < n; ++i) mask[i] = (f->h_mask[i >> 5] >> (i & 31)) & 1u;
    if (F_out) std::memcpy(F_out, f->h_F, 9 * sizeof(double));
    return finish(0);
}

}  // namespace pvio
