// Reduced-system kernel: assembles the dense pose/velocity/bias system of one window and
// solves it.  One CTA per window.
//
//   * vision part: block-lower-triangular Schur-reduced system from lin_schur_kernel in xi
//     coordinates, mapped to the reference's local coordinates delta = [dtheta dp] with
//     H_delta[f,g] = T_f^T H_xi[f,g] T_g  (T_f = [[R_f,0],[-[p_f]x R_f,-I]])
//   * PreIntegrationErrorCost::Evaluate      estimation/ceres/preintegration_error_cost.h:40-160
//   * MarginalizationErrorCost::Evaluate     estimation/ceres/marginalization_error_cost.h:53-94
//   * AugmentedPlaneDistanceErrorCost::Evaluate  .../augmented_plane_distance_error_cost.h:53-136
//     (+ CauchyLoss corrector, bundle_adjustor.cpp:192)
//   * constant blocks (FF_FIX_POSE, bundle_adjustor.cpp:79-82), Jacobi scaling and the
//     mu*diag regulariser of ceres' dogleg Gauss-Newton solve, dense Cholesky, and the
//     scalars of the trust-region step.
// Everything here is fp64: it is O(D^3/6) flops on a D <= 210 system, off the HBM-bound path.
#pragma once
#include "ba_math.cuh"
#include "ba_types.h"
#include "ba_lin.cuh"
#include "ba_tr.cuh"

namespace pvio {

struct SolveArgs {
    const WinHdr *hdr;
    const WinConst *cst;
    const double *frames;      // [W][Ncap][16]
    WinCtrl *ctrl;
    const double *Hred;        // from lin_schur_kernel
    const double *Hdd;
    const double *gdir;
    const double *gred;
    const double *cost_vis;
    // IMU
    const int32_t *imu_idx;    // [W][Ncap][2] (frame_i, frame_j)
    const double *imu_data;    // [W][Ncap][kImuStride]
    int alias_bias;            // 1: dbg = dba = 0 at the linearisation point (quirk Q1)
    // prior
    const int32_t *prior_frames;   // [W][Ncap]
    const double *prior_S;         // [W][dcap*dcap], dcap = 15*Ncap
    const double *prior_L;         // [W][dcap*dcap]  Lambda = S^T S
    const double *prior_e;         // [W][dcap]
    const double *prior_x0;        // [W][Ncap][16]
    // planes
    const double *plane_param;     // [W][Pcap][4]
    const int32_t *pt_plane;       // [W][Tcap]
    const int32_t *pt_begin;       // [W][Tcap+1]
    const int32_t *pt_frame;       // [W][Ocap]
    const float *pt_z;             // [W][Ocap][2]  (observations are stored fp32 like the reprojection table)
    int Pcap, Tcap, Ocap;
    double *pt_J;                  // [W][Tcap][6 Ncap + 2] scratch: corrected Jacobian rows of the plane tracks
    // scaling / outputs
    double *pose_scale;        // [W][15*Ncap]
    double *dx_pose;           // [W][Ncap][15]
    double *v_pose;            // [W][Ncap][15] scaled steepest-descent direction S^2 g / clamp(S^2 diag H)
    double *Hfull;             // optional [W][(15 Ncap)^2] dump of the reduced system (delta coords, before regularisation)
    double *gfull;             // optional [W][15 Ncap]
    int Ncap;
    int compute_scale;         // ignored when loop != 0 (then: WinCtrl::have_scale == 0)
    int w0;
    double mu_override;
    int loop;                  // 1: iteration of the device-side trust-region loop (ba_tr.cuh): obey the window's flags
    LinBufs bufs;
};

// The dense system is stored as packed lower-triangular 4x4 TILES (tile (I,J), J <= I, at
// I(I+1)/2 + J, 16 doubles row-major) so that the factorisation works on register-sized blocks.
// A tile occupies kTP = 18 doubles: the same element of consecutive tiles then lies 4 banks apart, so a warp whose
// lanes work on consecutive tiles reads 16 bytes per lane without bank conflicts (at a pitch of 16 every lane hit the
// same four banks: the factorisation was shared-memory-bandwidth bound, 5.7 K cycles per block column at D = 135).
constexpr int kTP = 18;
__device__ __forceinline__ int tile_off(int I, int J) { return (I * (I + 1) / 2 + J) * kTP; }
__device__ __forceinline__ int tri(int i, int j) {   // element (i,j), j <= i
    return tile_off(i >> 2, j >> 2) + ((i & 3) << 2) + (j & 3);
}

// ---- IMU factor: raw residual and Jacobian (before whitening); J is [15][30] row-major.
__device__ inline void imu_factor_raw(const double *fi, const double *fj, const double *rec,
                                      const WinConst &wc, int alias_bias, double *r, double *J,
                                      const double *bias0 = nullptr) {
    // J == nullptr: residual only (candidate-cost evaluations)
    // bias0 (bg, ba: 6 doubles): the bias linearisation point to use instead of the record's (candidate cost under Q1:
    // the CURRENT state's biases while fi is the candidate)
    const double g[3] = {0.0, 0.0, -9.80665};                 // preintegration_error_cost.h:41
    if (J) for (int i = 0; i < 450; ++i) J[i] = 0.0;
    const double *qic = fi, *pic = fi + 4, *vi = fi + 7, *bgi = fi + 10, *bai = fi + 13;
    const double *qjc = fj, *pjc = fj + 4, *vj = fj + 7, *bgj = fj + 10, *baj = fj + 13;
    const double dt = rec[0];
    const double *dq = rec + 1, *dp = rec + 5, *dv = rec + 8;
    const double *dq_dbg = rec + 236, *dp_dbg = rec + 245, *dp_dba = rec + 254, *dv_dbg = rec + 263, *dv_dba = rec + 272;
    double dbg[3], dba[3];
    for (int k = 0; k < 3; ++k) {
        dbg[k] = alias_bias ? 0.0 : bgi[k] - (bias0 ? bias0[k] : rec[281 + k]);    // :69-70 (Q1: bg_i_0 aliases the parameter)
        dba[k] = alias_bias ? 0.0 : bai[k] - (bias0 ? bias0[3 + k] : rec[284 + k]);
    }
    double qi[4], qj[4], Rci[9], Rcj[9], t3[3], pi[3], pj[3];
    quat_mul(qic, wc.imu_q, qi);                                // :60
    quat_mul(qjc, wc.imu_q, qj);                                // :62
    quat_to_mat(qic, Rci);
    quat_to_mat(qjc, Rcj);
    mat3_vec(Rci, wc.imu_p, t3);
    for (int k = 0; k < 3; ++k) pi[k] = pic[k] + t3[k];         // :61
    mat3_vec(Rcj, wc.imu_p, t3);
    for (int k = 0; k < 3; ++k) pj[k] = pjc[k] + t3[k];         // :63
    // r_q = log( (dq exp(dq_dbg dbg))^-1 qi^-1 qj )            // :79
    double w3[3], e4[4], dqc[4], c1[4], c2[4], c3[4], c4[4];
    mat3_vec(dq_dbg, dbg, w3);
    expmap(w3, e4);
    quat_mul(dq, e4, dqc);
    quat_conj(dqc, c1);
    quat_conj(qi, c2);
    quat_mul(c1, c2, c3);
    quat_mul(c3, qj, c4);
    logmap(c4, r + 0);
    double Ri[9];
    quat_to_mat(qi, Ri);
    double a3[3], b3[3];
    for (int k = 0; k < 3; ++k) a3[k] = pj[k] - pi[k] - dt * vi[k] - 0.5 * dt * dt * g[k];
    mat3_tvec(Ri, a3, b3);                                      // qi^-1 * (.)
    double c_[3], d_[3];
    mat3_vec(dp_dbg, dbg, c_);
    mat3_vec(dp_dba, dba, d_);
    for (int k = 0; k < 3; ++k) r[3 + k] = b3[k] - (dp[k] + c_[k] + d_[k]);   // :80
    for (int k = 0; k < 3; ++k) a3[k] = vj[k] - vi[k] - dt * g[k];
    mat3_tvec(Ri, a3, b3);
    mat3_vec(dv_dbg, dbg, c_);
    mat3_vec(dv_dba, dba, d_);
    for (int k = 0; k < 3; ++k) r[6 + k] = b3[k] - (dv[k] + c_[k] + d_[k]);   // :81
    for (int k = 0; k < 3; ++k) { r[9 + k] = bgj[k] - bgi[k]; r[12 + k] = baj[k] - bai[k]; }  // :82-83
    if (!J) return;

    double Jr[9], Jri[9], Rimu[9], Rj[9];
    right_jacobian(r, Jr);
    mat3_inv(Jr, Jri);
    quat_to_mat(wc.imu_q, Rimu);
    quat_to_mat(qj, Rj);
    double M1[9], M2[9], H3[9];
#define PUT(row0, col0, M, sgn)                                                      \
    for (int a_ = 0; a_ < 3; ++a_)                                                   \
        for (int b_ = 0; b_ < 3; ++b_) J[((row0) + a_) * 30 + (col0) + b_] = (sgn) * (M)[3 * a_ + b_];
    // d/dq_i :86-93
    mat3_mul_tn(Rj, Rci, M1);               // qj^-1 * q_center_i
    mat3_mul(Jri, M1, M2);
    PUT(0, 0, M2, -1.0);
    for (int k = 0; k < 3; ++k) a3[k] = pj[k] - pic[k] - dt * vi[k] - 0.5 * dt * dt * g[k];
    mat3_tvec(Rci, a3, b3);
    hat(b3, H3);
    mat3_mul_tn(Rimu, H3, M1);
    PUT(3, 0, M1, 1.0);
    for (int k = 0; k < 3; ++k) a3[k] = vj[k] - vi[k] - dt * g[k];
    mat3_tvec(Rci, a3, b3);
    hat(b3, H3);
    mat3_mul_tn(Rimu, H3, M1);
    PUT(6, 0, M1, 1.0);
    // d/dp_i :94-99, d/dv_i :100-106   (-qi^-1)
    double Rit[9];
    for (int a_ = 0; a_ < 3; ++a_) for (int b_ = 0; b_ < 3; ++b_) Rit[3 * a_ + b_] = Ri[3 * b_ + a_];
    PUT(3, 3, Rit, -1.0);
    PUT(3, 6, Rit, -dt);
    PUT(6, 6, Rit, -1.0);
    // d/dbg_i :107-115
    {
        double er[4], erc[4], Rer[9], Jr2[9];
        expmap(r, er);
        quat_conj(er, erc);
        quat_to_mat(erc, Rer);
        right_jacobian(w3, Jr2);
        mat3_mul(Jri, Rer, M1);
        mat3_mul(M1, Jr2, M2);
        mat3_mul(M2, dq_dbg, M1);
        PUT(0, 9, M1, -1.0);
    }
    PUT(3, 9, dp_dbg, -1.0);
    PUT(6, 9, dv_dbg, -1.0);
    for (int k = 0; k < 3; ++k) J[(9 + k) * 30 + 9 + k] = -1.0;
    // d/dba_i :116-123
    PUT(3, 12, dp_dba, -1.0);
    PUT(6, 12, dv_dba, -1.0);
    for (int k = 0; k < 3; ++k) J[(12 + k) * 30 + 12 + k] = -1.0;
    // d/dq_j :124-130
    mat3_mul_tn(Jri, Rimu, M1);             // careful: Jri * Rimu^T
    for (int a_ = 0; a_ < 3; ++a_)
        for (int b_ = 0; b_ < 3; ++b_) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += Jri[3 * a_ + k] * Rimu[3 * b_ + k];
            M1[3 * a_ + b_] = s;
        }
    PUT(0, 15, M1, 1.0);
    hat(wc.imu_p, H3);
    mat3_mul(Rit, Rcj, M1);
    mat3_mul(M1, H3, M2);
    PUT(3, 15, M2, -1.0);
    // d/dp_j, dv_j, dbg_j, dba_j :131-154
    PUT(3, 18, Rit, 1.0);
    PUT(6, 21, Rit, 1.0);
    for (int k = 0; k < 3; ++k) { J[(9 + k) * 30 + 24 + k] = 1.0; J[(12 + k) * 30 + 27 + k] = 1.0; }
#undef PUT
}

// ---- prior: raw residual [log(q0^-1 q); p-p0; v-v0; bg-bg0; ba-ba0] and Jr^-1 of one frame
__device__ inline void prior_frame_raw(const double *fs, const double *x0, double *r15, double *Jri) {
    double c[4], d[4];
    quat_conj(x0, c);
    quat_mul(c, fs, d);
    logmap(d, r15);                                           // marginalization_error_cost.h:65
    for (int k = 0; k < 12; ++k) r15[3 + k] = fs[4 + k] - x0[4 + k];
    double Jr[9];
    right_jacobian(r15, Jr);
    mat3_inv(Jr, Jri);                                         // :76
}

// ---- plane factor (one plane track), fp64.  J is [6*K] in delta coordinates of the K frames.
__device__ inline void plane_factor(int K, const int32_t *fr, const float *z, const double *frames,
                                    const WinConst &wc, const double *pl, double sic, double *r_out, double *J) {
    double A[2 * kMaxFrames + 1][3], b[2 * kMaxFrames + 1];
    double Rcs[9], qcc[4];
    quat_to_mat(wc.cam_q, Rcs);
    quat_conj(wc.cam_q, qcc);
    double tc[3];
    mat3_tvec(Rcs, wc.cam_p, tc);                              // q_cs^-1 * p_cs
    for (int i = 0; i < K; ++i) {
        const double *fs = frames + fr[i] * kFrameStride;
        double R[9], Rsw[9], Tsw[3];
        quat_to_mat(fs, R);
        // Rsw = (q_cs^-1 q_wc^-1).matrix() = Rcs^T R^T            :68
        for (int a_ = 0; a_ < 3; ++a_)
            for (int b_ = 0; b_ < 3; ++b_) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += Rcs[3 * k + a_] * R[3 * b_ + k];
                Rsw[3 * a_ + b_] = s;
            }
        double pw[3] = {fs[4] - wc.origin[0], fs[5] - wc.origin[1], fs[6] - wc.origin[2]};
        // NOTE: the functor is evaluated in absolute world coordinates; the origin shift would
        // change b because the plane distance is not shifted, so undo it here.
        pw[0] = fs[4]; pw[1] = fs[5]; pw[2] = fs[6];
        mat3_vec(Rsw, pw, Tsw);
        for (int k = 0; k < 3; ++k) Tsw[k] = -Tsw[k] - tc[k];      // :69
        const double u = (double)z[2 * i], v = (double)z[2 * i + 1];
        for (int k = 0; k < 3; ++k) {
            A[2 * i][k] = u * Rsw[6 + k] - Rsw[k];                  // :71-72
            A[2 * i + 1][k] = v * Rsw[6 + k] - Rsw[3 + k];
        }
        b[2 * i] = u * Tsw[2] - Tsw[0];
        b[2 * i + 1] = v * Tsw[2] - Tsw[1];
    }
    const double *nrm = pl;
    const double dist = pl[3];
    for (int k = 0; k < 3; ++k) A[2 * K][k] = nrm[k];             // regularization_weight = 1 :84-85
    b[2 * K] = dist;
    double ATA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ATb[3] = {0, 0, 0};
    for (int i = 0; i < 2 * K + 1; ++i)
        for (int a_ = 0; a_ < 3; ++a_) {
            ATb[a_] += A[i][a_] * b[i];
            for (int b_ = 0; b_ < 3; ++b_) ATA[3 * a_ + b_] += A[i][a_] * A[i][b_];
        }
    double lam[3], V[9], Ainv[9];
    sym3_eig(ATA, lam, V);                                        // :90
    for (int a_ = 0; a_ < 3; ++a_)
        for (int b_ = 0; b_ < 3; ++b_) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += V[3 * a_ + k] * (lam[k] > 1.0e-8 ? 1.0 / lam[k] : 0.0) * V[3 * b_ + k];
            Ainv[3 * a_ + b_] = s;
        }
    double x[3];
    mat3_vec(Ainv, ATb, x);
    for (int k = 0; k < 3; ++k) x[k] = -x[k];                     // :94
    *r_out = (nrm[0] * x[0] + nrm[1] * x[1] + nrm[2] * x[2] - dist) * sic;   // :96,:133
    for (int i = 0; i < K; ++i) {
        const double *fs = frames + fr[i] * kFrameStride;
        const double u = (double)z[2 * i], v = (double)z[2 * i + 1];
        double R[9], Rsw[9];
        quat_to_mat(fs, R);
        for (int a_ = 0; a_ < 3; ++a_)
            for (int b_ = 0; b_ < 3; ++b_) {
                double s = 0.0;
                for (int k = 0; k < 3; ++k) s += Rcs[3 * k + a_] * R[3 * b_ + k];
                Rsw[3 * a_ + b_] = s;
            }
        const double Jb[2][3] = {{-1.0, 0.0, u}, {0.0, -1.0, v}};
        double drdth[3] = {0, 0, 0};
        double M[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};               // dxdAdq + dxdbdq
        for (int rr = 0; rr < 2; ++rr) {
            const double *Ar = A[2 * i + rr];
            const double sres = b[2 * i + rr] + Ar[0] * x[0] + Ar[1] * x[1] + Ar[2] * x[2];
            double AiA[3];                                        // A_row * ATAinv (row vector)
            for (int k = 0; k < 3; ++k) AiA[k] = Ar[0] * Ainv[k] + Ar[1] * Ainv[3 + k] + Ar[2] * Ainv[6 + k];
            double dxdA[9];                                       // sres*ATAinv + (x (A_row ATAinv))^T   :105
            for (int a_ = 0; a_ < 3; ++a_)
                for (int b_ = 0; b_ < 3; ++b_) dxdA[3 * a_ + b_] = sres * Ainv[3 * a_ + b_] + AiA[a_] * x[b_];
            double cj[3], H3[9], dAdq[9], T9[9];
            mat3_vec(Rcs, Jb[rr], cj);                            // q_cs * Jb.row^T
            hat(cj, H3);
            mat3_mul(R, H3, dAdq);                                // :107-108
            mat3_mul(dxdA, dAdq, T9);
            for (int k = 0; k < 9; ++k) M[k] += T9[k];
        }
        {   // dxdbdq = ATAinv A_blk^T Jb Rcs^T hat(qwc^-1 pwc)      :110
            double pw[3] = {fs[4], fs[5], fs[6]}, pb[3], H3[9];
            mat3_tvec(R, pw, pb);
            hat(pb, H3);
            double AtJb[9];                                       // (A_blk^T Jb) 3x3
            for (int a_ = 0; a_ < 3; ++a_)
                for (int b_ = 0; b_ < 3; ++b_) AtJb[3 * a_ + b_] = A[2 * i][a_] * Jb[0][b_] + A[2 * i + 1][a_] * Jb[1][b_];
            double T1[9], T2[9], T3[9];
            mat3_mul(Ainv, AtJb, T1);
            for (int a_ = 0; a_ < 3; ++a_)
                for (int b_ = 0; b_ < 3; ++b_) {
                    double s = 0.0;
                    for (int k = 0; k < 3; ++k) s += T1[3 * a_ + k] * Rcs[3 * b_ + k];   // * Rcs^T
                    T2[3 * a_ + b_] = s;
                }
            mat3_mul(T2, H3, T3);
            for (int k = 0; k < 9; ++k) M[k] += T3[k];
            // drdp = n^T (ATAinv A_blk^T Jb Rsw)                   :117
            mat3_mul(T1, Rsw, T2);
            for (int k = 0; k < 3; ++k)
                J[6 * i + 3 + k] = (nrm[0] * T2[k] + nrm[1] * T2[3 + k] + nrm[2] * T2[6 + k]) * sic;
        }
        for (int k = 0; k < 3; ++k) drdth[k] = nrm[0] * M[k] + nrm[1] * M[3 + k] + nrm[2] * M[6 + k];
        for (int k = 0; k < 3; ++k) J[6 * i + k] = drdth[k] * sic;
    }
}

#ifdef PVIO_SOLVE_STAMPS          // tuning builds of tools/ only: clock64() at the phase boundaries of solve_kernel
__device__ long long g_solve_stamps[16];
#define SOLVE_STAMP(k) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) g_solve_stamps[k] = clock64(); } while (0)
#define CHOL_T0() long long ct__ = clock64()
#define CHOL_ACC(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { const long long n__ = clock64(); g_solve_stamps[k] += n__ - ct__; ct__ = n__; } } while (0)
#else
#define SOLVE_STAMP(k) do { } while (0)
#define CHOL_T0() do { } while (0)
#define CHOL_ACC(k) do { } while (0)
#endif


// Tiled dense Cholesky A = L L^T + solve, all threads of the CTA, fp64, in shared memory.
// nb block rows of 4 (padding rows carry an identity diagonal).  The right-hand side is block row
// nb of the packed tile array (row 0 of its tiles), so the panel / trailing phases carry out the
// forward substitution for free.  Per block column kb:
//   (b) one thread per panel tile row (incl. the rhs row): A_Ik <- A_Ik L_kk^-T;
//   (c) trailing tiles A_IJ -= A_Ik A_Jk^T: one thread per tile (kPerTile: wide CTAs, 64 FMAs from 48 loads) or per
//       tile row (narrow CTAs of the batched visual kernel);
//   (a) LOOK-AHEAD: the owner(s) of the next diagonal tile update it first and thread 0 factors it (4x4 Cholesky +
//       inverse of the factor, a serial chain of ~500 cycles) WHILE the other threads finish the trailing update.
// 2 barriers per block column.  The back substitution L^T x = y runs on warp 0 with warp-level barriers only.
// A must hold (nb+1)(nb+2)/2 tiles of kTP doubles.  Returns false (uniformly) if a pivot is
// not positive / finite (the system must be positive definite, as for ceres' Cholesky-based SPARSE_SCHUR).
// L_kk itself is not needed once its inverse is known (panel and back substitution multiply by inv(L_kk)): the inverse
// overwrites the diagonal tile, no separate array
__device__ __forceinline__ void chol_factor4(double *Akk, int *flag_sm) {
    double L[16], Li[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) { L[e] = Akk[e]; Li[e] = 0.0; }
    bool good = true;
    double idg[4];                       // reciprocals of the diagonal of L (no divisions below)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double d = L[c * 4 + c];
#pragma unroll
        for (int m = 0; m < c; ++m) d -= L[c * 4 + m] * L[c * 4 + m];
        if (!(d > 0.0) || !isfinite(d)) { good = false; d = 1.0; }
        const double id = rsqrt(d);
        idg[c] = id;
        L[c * 4 + c] = d * id;
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            double v = L[r * 4 + c];
#pragma unroll
            for (int m = 0; m < c; ++m) v -= L[r * 4 + m] * L[c * 4 + m];
            L[r * 4 + c] = v * id;
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        Li[c * 4 + c] = idg[c];
#pragma unroll
        for (int r = c + 1; r < 4; ++r) {
            double v = 0.0;
#pragma unroll
            for (int m = c; m < r; ++m) v -= L[r * 4 + m] * Li[m * 4 + c];
            Li[r * 4 + c] = v * idg[r];
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) Akk[e] = Li[e];
    if (!good) *flag_sm = 0;
}

// lower-triangle tile index e -> (ii, jj), jj <= ii
__device__ __forceinline__ void tri_unrank(int e, int &ii, int &jj) {
    ii = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
    while ((ii + 1) * (ii + 2) / 2 <= e) ++ii;
    while (ii * (ii + 1) / 2 > e) --ii;
    jj = e - ii * (ii + 1) / 2;
}

// back substitution x = L^-T y on warp 0 (y: row 0 of the rhs tiles; the diagonal tiles hold the inverse factors)
__device__ __forceinline__ void chol_back_substitute(double *A, double *x, int nb) {
    const int tid = threadIdx.x;
    const double *R = A + tile_off(nb, 0);
    if (tid < 32) {
        for (int i = tid; i < nb * 4; i += 32) x[i] = R[(i >> 2) * kTP + (i & 3)];
        __syncwarp();
        for (int kb = nb - 1; kb >= 0; --kb) {
            const double *Li = A + tile_off(kb, kb);
            double o = 0.0;
            if (tid < 4) {
#pragma unroll
                for (int m = 0; m < 4; ++m) o += (m >= tid) ? Li[m * 4 + tid] * x[kb * 4 + m] : 0.0;   // Linv^T
            }
            __syncwarp();
            if (tid < 4) x[kb * 4 + tid] = o;
            __syncwarp();
            for (int e = tid; e < kb * 4; e += 32) {
                const int J = e >> 2, c = e & 3;
                const double *X = A + tile_off(kb, J);      // tile (kb, J)
                x[e] -= X[c] * x[kb * 4] + X[4 + c] * x[kb * 4 + 1] + X[8 + c] * x[kb * 4 + 2] + X[12 + c] * x[kb * 4 + 3];
            }
            __syncwarp();
        }
    }
}

template <bool kPerTile>
__device__ inline bool chol_solve_tiled(double *A, double *x, int nb, int *flag_sm) {
    const int tid = threadIdx.x, nt = blockDim.x;
    double *R = A + tile_off(nb, 0);                       // rhs block row: tiles (nb, J), row 0 used
    for (int i = tid; i < nb * kTP; i += nt) { const int J = i / kTP, o = i - J * kTP; R[i] = (o < 4) ? x[J * 4 + o] : 0.0; }
    if (tid == 0) *flag_sm = 1;
    __syncthreads();
    if (tid == 0) chol_factor4(A + tile_off(0, 0), flag_sm);
    __syncthreads();
#ifdef PVIO_SOLVE_STAMPS
    if (tid == 0 && blockIdx.x == 0) for (int k = 10; k < 16; ++k) g_solve_stamps[k] = 0;
#endif
    CHOL_T0();
    for (int kb = 0; kb < nb; ++kb) {
        if (*flag_sm == 0) break;                      // uniform
        // (b) panel tiles incl. the rhs block row (I = nb): X <- X * Linv^T, one thread per tile ROW
        for (int e = tid; e < (nb - kb) * 4; e += nt) {
            const int I = kb + 1 + (e >> 2), r = e & 3;
            if (I == nb && r != 0) continue;           // only row 0 of the rhs tiles carries data
            double *X = A + tile_off(I, kb) + r * 4;
            const double *Li = A + tile_off(kb, kb);
            const double x0 = X[0], x1 = X[1], x2 = X[2], x3 = X[3];
            X[0] = x0 * Li[0];
            X[1] = x0 * Li[4] + x1 * Li[5];
            X[2] = x0 * Li[8] + x1 * Li[9] + x2 * Li[10];
            X[3] = x0 * Li[12] + x1 * Li[13] + x2 * Li[14] + x3 * Li[15];
        }
        CHOL_ACC(10);
        __syncthreads();
        CHOL_ACC(11);
        // (c) trailing tiles (I,J), kb < J <= I < nb, plus the rhs row (I = nb, kb < J < nb); tile 0 is the next diagonal tile
        const int n = nb - kb - 1, ntile = n * (n + 1) / 2;
        if (kPerTile) {
            for (int e = tid; e < ntile + n; e += nt) {
                int I, J;
                if (e < ntile) { int ii, jj; tri_unrank(e, ii, jj); I = kb + 1 + ii; J = kb + 1 + jj; }
                else { I = nb; J = kb + 1 + (e - ntile); }
                const double *P = A + tile_off(I, kb), *Q = A + tile_off(J, kb);
                double *C = A + tile_off(I, J);
                double q[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) q[k] = Q[k];
                const int rows = (I == nb) ? 1 : 4;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r >= rows) break;
                    const double p0 = P[r * 4], p1 = P[r * 4 + 1], p2 = P[r * 4 + 2], p3 = P[r * 4 + 3];
#pragma unroll
                    for (int c = 0; c < 4; ++c) C[r * 4 + c] -= p0 * q[c * 4] + p1 * q[c * 4 + 1] + p2 * q[c * 4 + 2] + p3 * q[c * 4 + 3];
                }
                if (e == 0) { CHOL_ACC(12); chol_factor4(C, flag_sm); CHOL_ACC(13); }     // look-ahead: tile (kb+1, kb+1) is complete
            }
        } else {
            for (int e4 = tid; e4 < (ntile + n) * 4; e4 += nt) {
                const int e = e4 >> 2, r = e4 & 3;
                int I, J;
                if (e < ntile) { int ii, jj; tri_unrank(e, ii, jj); I = kb + 1 + ii; J = kb + 1 + jj; }
                else { I = nb; J = kb + 1 + (e - ntile); }
                if (!(I == nb && r != 0)) {
                    const double *P = A + tile_off(I, kb) + r * 4, *Q = A + tile_off(J, kb);
                    double *C = A + tile_off(I, J) + r * 4;
                    const double p0 = P[0], p1 = P[1], p2 = P[2], p3 = P[3];
#pragma unroll
                    for (int c = 0; c < 4; ++c) C[c] -= p0 * Q[c * 4] + p1 * Q[c * 4 + 1] + p2 * Q[c * 4 + 2] + p3 * Q[c * 4 + 3];
                }
                if (e4 < 4 && n > 0) {                 // threads 0..3 hold the four rows of tile (kb+1, kb+1): look-ahead
                    __syncwarp(0xFu);
                    if (e4 == 0) chol_factor4(A + tile_off(kb + 1, kb + 1), flag_sm);
                }
            }
        }
        CHOL_ACC(14);
        __syncthreads();
        CHOL_ACC(15);
    }
    if (*flag_sm == 0) return false;
    // y = L^-1 b now sits in row 0 of the rhs tiles; back substitution x = L^-T y on warp 0
    chol_back_substitute(A, x, nb);
    __syncthreads();
    return true;
}

// diagonal tile C -= L L^T with L = the panel tile (J, k) of its own block row: one tile load, the lower triangle only
// (the strict upper triangle of a diagonal tile is never read: chol_factor4 uses L[r][c], c <= r)
__device__ __forceinline__ void chol_diag_update(double (&c)[16], const double *Lt) {
    const double2 *P2 = reinterpret_cast<const double2 *>(Lt);
    double q[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const double2 v = P2[k]; q[2 * k] = v.x; q[2 * k + 1] = v.y; }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc)
            c[r * 4 + cc] -= q[r * 4] * q[cc * 4] + q[r * 4 + 1] * q[cc * 4 + 1] + q[r * 4 + 2] * q[cc * 4 + 2] + q[r * 4 + 3] * q[cc * 4 + 3];
}

// The same factorisation with the TRAILING MATRIX IN REGISTERS and the critical path on a warp of its own (wide CTA of
// 8 warps, the inertial single window: D = 135 ... 165).  What bounded chol_solve_tiled (profiles/r02j): each block
// column is a serial chain -- update of the next diagonal tile, its 4 x 4 factorisation + inverse by ONE thread (~750
// cycles of dependent fp64 rsqrt / multiply-adds) -- and the warp that held that thread also had its share of trailing
// tiles to do before or after, so the other warps waited ~3.4 K cycles per block column at the barrier.  Here
//   warp 0        owns the DIAGONAL tiles (lane l: tiles l and l + 32), nothing else: per block column it updates the
//                 next diagonal tile, one lane factors it, then the lanes update their remaining diagonal tiles;
//   warps 1..7    own the off-diagonal tiles and the rhs row (tile e = t + s * 224 by descending column, up to kSlots
//                 per thread) IN REGISTERS for the whole factorisation: a block column costs a tile two 128-byte panel
//                 tiles read from shared memory and no write (chol_solve_tiled: 512 B per tile and step).
// Shared memory holds what others need: the inverse factor of each diagonal tile and the finished panel tiles L_Ik
// (also what the back substitution reads).  ONE barrier per block column: trailing update C -= L_Ik L_Jk^T, then the
// owners of the next column's tiles wait for warp 0 to publish the inverse factor (a shared flag, no barrier) and store
// their panel tiles X <- X L^-T.
template <int kSlots>
__device__ inline bool chol_solve_regs(double *A, double *x, int nb, int *flag_sm) {
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int ntask = nb * (nb + 1) / 2, nown = nt - 32;      // off-diagonal tiles + the rhs row
    double *R = A + tile_off(nb, 0);                       // rhs block row: tiles (nb, J), row 0 used
    for (int i = tid; i < nb * kTP; i += nt) { const int J = i / kTP, o = i - J * kTP; R[i] = (o < 4) ? x[J * 4 + o] : 0.0; }
    if (tid == 0) *flag_sm = 1;
    __syncthreads();
    int tI[kSlots], tJ[kSlots];
    double c[kSlots][16];
#pragma unroll
    for (int s = 0; s < kSlots; ++s) {
        tI[s] = -1; tJ[s] = -1;
        if (warp == 0) {
            if (s < 2 && lane + 32 * s < nb) tI[s] = tJ[s] = lane + 32 * s;
        } else {
            // columns in DESCENDING order (column J: its nb - 1 - J off-diagonal tiles, then its rhs tile): the tiles still
            // active at block column kb (J > kb) are the first ones of this order, i.e. they sit in the lowest slots of
            // all threads -- a warp executes as many tile bodies per block column as slots are still alive, not kSlots --
            // and the panel tiles of one column are neighbours (one or two warps, one slot)
            const int e = tid - 32 + s * nown;
            if (e < ntask) { int ii, jj; tri_unrank(e, ii, jj); tJ[s] = nb - 1 - ii; tI[s] = tJ[s] + 1 + jj; }
        }
        if (tI[s] >= 0) {
            const double2 *src = reinterpret_cast<const double2 *>(A + tile_off(tI[s], tJ[s]));
#pragma unroll
            for (int k = 0; k < 8; ++k) { const double2 v = src[k]; c[s][2 * k] = v.x; c[s][2 * k + 1] = v.y; }
        } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) c[s][k] = 0.0;
        }
    }
    if (tid == 0) chol_factor4(A + tile_off(0, 0), flag_sm);      // tile (0, 0) is still intact in shared memory
    __syncthreads();
    // panel tile X <- X L_kk^-T of column `col` for every owned tile of that column (incl. the rhs row), into shared memory
    auto panel = [&](int col) {
        const double *Li = A + tile_off(col, col);
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            if (tJ[s] != col) continue;                // (tI > col for every off-diagonal tile of the column)
            const int rows = (tI[s] == nb) ? 1 : 4;
            double *X = A + tile_off(tI[s], col);
            const double l0 = Li[0], l4 = Li[4], l5 = Li[5], l8 = Li[8], l9 = Li[9], l10 = Li[10], l12 = Li[12], l13 = Li[13],
                         l14 = Li[14], l15 = Li[15];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r >= rows) break;
                const double x0 = c[s][r * 4], x1 = c[s][r * 4 + 1], x2 = c[s][r * 4 + 2], x3 = c[s][r * 4 + 3];
                double2 o0, o1;
                o0.x = x0 * l0;
                o0.y = x0 * l4 + x1 * l5;
                o1.x = x0 * l8 + x1 * l9 + x2 * l10;
                o1.y = x0 * l12 + x1 * l13 + x2 * l14 + x3 * l15;
                reinterpret_cast<double2 *>(X + r * 4)[0] = o0;
                reinterpret_cast<double2 *>(X + r * 4)[1] = o1;
            }
        }
    };
    __shared__ int diag_ready_sm;                      // the last block column whose diagonal tile is factored (inverse in shared memory)
    volatile int *diag_ready = &diag_ready_sm;
    if (tid == 0) *diag_ready = 0;
    if (warp != 0) panel(0);
    __syncthreads();
    // One barrier per block column.  After it the panel tiles of column kb are in shared memory.  Warp 0 updates the next
    // diagonal tile, factors it and PUBLISHES it (diag_ready, no barrier); the other warps update their tiles with column
    // kb -- and the owners of the tiles of column kb + 1 then pick the inverse factor up as soon as it is published and
    // store their finished panel tiles, so that the next block column starts right after the barrier.
    for (int kb = 0; kb < nb; ++kb) {
        if (*flag_sm == 0) break;                      // uniform
        if (warp == 0) {
            const int sd = (kb + 1) >> 5;              // uniform: the slot of the next diagonal tile goes first
#pragma unroll
            for (int pass = 0; pass < 2; ++pass)
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    if ((pass == 0) != (s == sd)) continue;
                    if (tJ[s] > kb) chol_diag_update(c[s], A + tile_off(tJ[s], kb));
                    if (pass == 0 && tJ[s] == kb + 1) {
                        double2 *dst = reinterpret_cast<double2 *>(A + tile_off(kb + 1, kb + 1));
#pragma unroll
                        for (int k = 0; k < 8; ++k) { double2 v; v.x = c[s][2 * k]; v.y = c[s][2 * k + 1]; dst[k] = v; }
                        chol_factor4(A + tile_off(kb + 1, kb + 1), flag_sm);
                        __threadfence_block();
                        *diag_ready = kb + 1;
                    }
                }
        } else {
            // trailing update of the owned off-diagonal / rhs tiles (I, J), kb < J < I; the highest live slot first: it
            // holds the tiles of column kb + 1
            bool next_panel = false;
#pragma unroll
            for (int s = kSlots - 1; s >= 0; --s) {
                if (tJ[s] <= kb) continue;
                next_panel |= tJ[s] == kb + 1;
                const double2 *P2 = reinterpret_cast<const double2 *>(A + tile_off(tI[s], kb));
                const double2 *Q2 = reinterpret_cast<const double2 *>(A + tile_off(tJ[s], kb));
                double q[16];
#pragma unroll
                for (int k = 0; k < 8; ++k) { const double2 v = Q2[k]; q[2 * k] = v.x; q[2 * k + 1] = v.y; }
                const int rows = (tI[s] == nb) ? 1 : 4;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r >= rows) break;
                    const double2 pa = P2[2 * r], pb = P2[2 * r + 1];
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc)
                        c[s][r * 4 + cc] -= pa.x * q[cc * 4] + pa.y * q[cc * 4 + 1] + pb.x * q[cc * 4 + 2] + pb.y * q[cc * 4 + 3];
                }
            }
            if (next_panel) {
                while (*diag_ready < kb + 1) { }
                __threadfence_block();
                panel(kb + 1);
            }
        }
        __syncthreads();
    }
    if (*flag_sm == 0) return false;
    // y = L^-1 b now sits in row 0 of the rhs tiles; back substitution x = L^-T y on warp 0
    chol_back_substitute(A, x, nb);
    __syncthreads();
    return true;
}

constexpr int kSolveImuSlab = 15 * 30 + 16;     // raw J + r per factor
constexpr int kImuRound = 8;                    // IMU factors linearised concurrently by solve_kernel
constexpr int kImuJx = 15 * 32;                 // whitened Jacobian + residual column of one factor, row pitch 32

template <bool kFull>
static __device__ __forceinline__ void solve_body(const SolveArgs &a) {
    const int w = blockIdx.x + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int inertial = H.use_inertial;
    const int stride = inertial ? 15 : 6;
    // Visual-only windows: FF_FIX_POSE frames are constant parameter blocks (bundle_adjustor.cpp:82-87), so the
    // lean kernel leaves them out of the system altogether (cfg2: 48 instead of 60 unknowns, half the
    // factorisation); the full kernel keeps every frame (velocity / bias blocks stay free) and pins the pose rows.
    const unsigned allm = (1u << N) - 1u;
    const bool compact = !kFull && ((~(unsigned)H.fixed_mask) & allm) != 0u;
    const unsigned freem = compact ? (allm & ~(unsigned)H.fixed_mask) : allm;
    const int nfr = __popc(freem);
    const int D = stride * nfr;
    const int mask_c = compact ? 0 : H.fixed_mask;           // pose rows to pin, in the system's own frame numbering
    const int tid = threadIdx.x, nt = blockDim.x;
    const double *frames = a.frames + (size_t)w * a.Ncap * kFrameStride;
    WinCtrl &ctrl = a.ctrl[w];
    if (a.loop && (ctrl.done || ctrl.reuse)) return;         // finished, or the rejected step's linearisation is still valid
    const size_t bsel = a.loop ? (size_t)ctrl.buf : 0;       // buffer set holding the linearisation of the state (LinBufs)
    const int compute_scale = a.loop ? (ctrl.have_scale == 0) : a.compute_scale;
    const double mu = a.mu_override >= 0.0 ? a.mu_override : ctrl.mu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nb = (D + 3) >> 2, Dp = nb * 4;               // block rows of 4; rows >= D are identity padding
    const int nA = (nb + 1) * (nb + 2) / 2 * kTP;           // + the rhs block row
    double *A = reinterpret_cast<double *>(smem_raw);       // packed lower 4x4 tiles
    double *g = A + nA;                                     // [Dp] reduced gradient
    double *gu = g + Dp;                                    // [Dp] unreduced gradient (for |g|_inf)
    double *hcorr = gu + Dp;                                // [Dp] direct - reduced diagonal
    double *xs = hcorr + Dp;                                // [Dp] solution
    double *T = xs + Dp;                                    // [N][36] change of variables
    double *scr = T + N * 36;                               // scratch: staging / IMU slabs / prior vectors / reg copy
    __shared__ double cost_sm[4];
    __shared__ int flag_sm;

    SOLVE_STAMP(0);
    for (int i = tid; i < nA; i += nt) A[i] = 0.0;
    for (int i = tid; i < 4 * Dp; i += nt) g[i] = 0.0;      // g, gu, hcorr, xs contiguous
    if (tid < 4) cost_sm[tid] = 0.0;
    __shared__ double x2_sm[kMaxFrames];                    // |x|^2 of each frame's ambient parameter blocks
    if (tid < N) {
        const double *fs = frames + tid * kFrameStride;
        {   // ceres takes norms of the 4-vector quaternion; constant pose blocks are not parameters
            double x2 = 0.0;
            if (!((H.fixed_mask >> tid) & 1)) for (int k = 0; k < 7; ++k) x2 += fs[k] * fs[k];
            if (inertial) for (int k = 7; k < 16; ++k) x2 += fs[k] * fs[k];
            x2_sm[tid] = x2;
        }
        double R[9];
        quat_to_mat(fs, R);
        const double p[3] = {fs[4] - wc.origin[0], fs[5] - wc.origin[1], fs[6] - wc.origin[2]};
        double Hp[9], B[9];
        hat(p, Hp);
        mat3_mul(Hp, R, B);
        double *Tf = T + tid * 36;
        for (int r_ = 0; r_ < 3; ++r_)
            for (int c_ = 0; c_ < 3; ++c_) {
                Tf[r_ * 6 + c_] = R[3 * r_ + c_];
                Tf[r_ * 6 + 3 + c_] = 0.0;
                Tf[(3 + r_) * 6 + c_] = -B[3 * r_ + c_];
                Tf[(3 + r_) * 6 + 3 + c_] = (r_ == c_) ? -1.0 : 0.0;
            }
    }
    __syncthreads();
    const int npairs = N * (N + 1) / 2;
    const int npairs_cap = a.Ncap * (a.Ncap + 1) / 2;
    if (!kFull) {
        // Lean path (batched visual-only windows): no full staging of the xi blocks -- groups of 6 threads
        // pull one 6x6 block at a time through a small shared buffer, so that ~9 windows fit per SM.
        const double *Hred = a.Hred + bsel * a.bufs.Hred + (size_t)w * npairs_cap * 36;
        const double *Hdd = a.Hdd + bsel * a.bufs.Hdd + (size_t)w * a.Ncap * 36;
        const double *gdir = a.gdir + bsel * a.bufs.g + (size_t)w * a.Ncap * 6;
        const double *gred = a.gred + bsel * a.bufs.g + (size_t)w * a.Ncap * 6;
        double *Ms = scr + Dp;                       // [G][36]
        for (int e = tid; e < nfr * 6; e += nt) {    // direct diagonal diag(T_f^T Xd T_f)
            const int cf = e / 6, i = e - cf * 6;
            const int f = __fns(freem, 0, cf + 1);
            const double *Tf = T + f * 36, *X = Hdd + f * 36;
            double sd = 0.0;
            for (int aa = 0; aa < 6; ++aa) {
                double t = 0.0;
                for (int bb = 0; bb < 6; ++bb) t += X[aa * 6 + bb] * Tf[bb * 6 + i];
                sd += Tf[aa * 6 + i] * t;
            }
            hcorr[cf * stride + i] = sd;
        }
        const int G = nt / 6;
        const int q = tid / 6, i6 = tid - q * 6;
        const int npf = nfr * (nfr + 1) / 2;         // block pairs of the free frames
        for (int p0 = 0; p0 < npf; p0 += G) {
            const int p = p0 + q;
            const bool act = q < G && p < npf;
            int f = 0, gf = 0, cf = 0, cg = 0;
            if (act) {
                while ((cf + 1) * (cf + 2) / 2 <= p) ++cf;
                cg = p - cf * (cf + 1) / 2;
                f = __fns(freem, 0, cf + 1); gf = __fns(freem, 0, cg + 1);
                const double *X = Hred + (f * (f + 1) / 2 + gf) * 36 + i6 * 6, *Tg = T + gf * 36;
                double x[6];
#pragma unroll
                for (int bb = 0; bb < 6; ++bb) x[bb] = X[bb];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    double v = 0.0;
#pragma unroll
                    for (int bb = 0; bb < 6; ++bb) v += x[bb] * Tg[bb * 6 + j];
                    Ms[q * 36 + i6 * 6 + j] = v;       // M = X T_g, row i6
                }
            }
            __syncthreads();
            if (act) {
                const double *Tf = T + f * 36;
                const int j = i6;                        // this thread produces column j of T_f^T M
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    if (cf == cg && j > i) continue;
                    double sv = 0.0;
#pragma unroll
                    for (int aa = 0; aa < 6; ++aa) sv += Tf[aa * 6 + i] * Ms[q * 36 + aa * 6 + j];
                    const int gi = cf * stride + i, gj = cg * stride + j;
                    A[tri(gi, gj)] = sv;
                    if (cf == cg && i == j) hcorr[gi] -= sv;
                }
            }
            __syncthreads();
        }
        for (int e = tid; e < nfr * 6; e += nt) {
            const int cf = e / 6, i = e - cf * 6;
            const int f = __fns(freem, 0, cf + 1);
            const double *Tf = T + f * 36;
            double sr = 0.0, sd = 0.0;
            for (int aa = 0; aa < 6; ++aa) { sr += Tf[aa * 6 + i] * gred[f * 6 + aa]; sd += Tf[aa * 6 + i] * gdir[f * 6 + aa]; }
            g[cf * stride + i] = sr;
            gu[cf * stride + i] = sd;
        }
    } else
    // ---- vision blocks: H_delta[f,gf] = T_f^T X T_g.  Stage the xi-coordinate blocks in shared
    // memory (coalesced), X <- X T_g in place, then T_f^T (X T_g) into the packed system.
    {
        const double *Hred = a.Hred + bsel * a.bufs.Hred + (size_t)w * npairs_cap * 36;
        const double *Hdd = a.Hdd + bsel * a.bufs.Hdd + (size_t)w * a.Ncap * 36;
        const double *gdir = a.gdir + bsel * a.bufs.g + (size_t)w * a.Ncap * 6;
        const double *gred = a.gred + bsel * a.bufs.g + (size_t)w * a.Ncap * 6;
        double *Xs = scr;                        // [npairs][36]
        double *Xd = Xs + npairs * 36;           // [N][36]
        double *gx = Xd + N * 36;                // [2][N][6]
        for (int i = tid; i < npairs * 36; i += nt) Xs[i] = Hred[i];
        for (int i = tid; i < N * 36; i += nt) Xd[i] = Hdd[i];
        for (int i = tid; i < N * 6; i += nt) { gx[i] = gred[i]; gx[N * 6 + i] = gdir[i]; }
        __syncthreads();
        // direct (pre-Schur) diagonal, for the Jacobi scaling / LM diagonal: diag(T_f^T Xd T_f)
        for (int e = tid; e < N * 6; e += nt) {
            const int f = e / 6, i = e - f * 6;
            const double *Tf = T + f * 36, *X = Xd + f * 36;
            double sd = 0.0;
            for (int aa = 0; aa < 6; ++aa) {
                double t = 0.0;
                for (int bb = 0; bb < 6; ++bb) t += X[aa * 6 + bb] * Tf[bb * 6 + i];
                sd += Tf[aa * 6 + i] * t;
            }
            hcorr[f * stride + i] = sd;
        }
        // M = X T_g into a second buffer (one barrier; the in-place form needed two per pass of 252 entries), then
        // T_f^T M into the packed system
        double *Ms = gx + 2 * N * 6;             // [npairs][36]
        for (int e = tid; e < npairs * 36; e += nt) {
            const int p = e / 36, ij = e - p * 36, i = ij / 6, j = ij - i * 6;
            int f = 0;
            while ((f + 1) * (f + 2) / 2 <= p) ++f;
            const int gf = p - f * (f + 1) / 2;
            const double *X = Xs + p * 36 + i * 6, *Tg = T + gf * 36 + j;
            double v = 0.0;
#pragma unroll
            for (int bb = 0; bb < 6; ++bb) v += X[bb] * Tg[bb * 6];
            Ms[e] = v;
        }
        __syncthreads();
        for (int e = tid; e < npairs * 36; e += nt) {
            const int p = e / 36, ij = e - p * 36, i = ij / 6, j = ij - i * 6;
            int f = 0;
            while ((f + 1) * (f + 2) / 2 <= p) ++f;
            const int gf = p - f * (f + 1) / 2;
            if (f == gf && j > i) continue;
            const double *Tf = T + f * 36 + i, *Mx = Ms + p * 36 + j;
            double sv = 0.0;
#pragma unroll
            for (int aa = 0; aa < 6; ++aa) sv += Tf[aa * 6] * Mx[aa * 6];
            const int gi = f * stride + i, gj = gf * stride + j;
            A[tri(gi, gj)] = sv;
            if (f == gf && i == j) hcorr[gi] -= sv;     // direct - reduced diagonal
        }
        for (int e = tid; e < N * 6; e += nt) {
            const int f = e / 6, i = e - f * 6;
            const double *Tf = T + f * 36;
            double sr = 0.0, sd = 0.0;
            for (int aa = 0; aa < 6; ++aa) { sr += Tf[aa * 6 + i] * gx[f * 6 + aa]; sd += Tf[aa * 6 + i] * gx[N * 6 + f * 6 + aa]; }
            g[f * stride + i] = sr;
            gu[f * stride + i] = sd;
        }
    }
    __syncthreads();
    SOLVE_STAMP(1);
    // ---- IMU factors (bundle_adjustor.cpp:220-242): no loss
    if (kFull && inertial && H.n_imu > 0) {
        const int32_t *idx = a.imu_idx + (size_t)w * a.Ncap * 2;
        const double *recs = a.imu_data + (size_t)w * a.Ncap * kImuStride;
        for (int n0 = 0; n0 < H.n_imu; n0 += kImuRound) {  // up to kImuRound factors per round (scratch: raw + whitened J of each)
            const int nb = min(kImuRound, H.n_imu - n0);
            double *Jraw = scr;                              // [R][450], r at [R*450 + n*16]
            double *rraw = scr + kImuRound * 450;
            double *Jx = rraw + kImuRound * 16;              // [R][15][32] whitened [J | r | 0]
            double *Wsm = Jx + kImuRound * kImuJx;           // [R][225] sqrt information matrices of the round
            __shared__ int imu_fr[2 * kImuRound];            // frame pair of each factor of the round (read in every inner loop)
            __shared__ int imu_chain;
            __shared__ double imu_cost[kImuRound];           // r^T r of each factor of the round
            if (tid >= 32 && tid < 32 + 2 * nb) imu_fr[tid - 32] = idx[2 * n0 + tid - 32];
            if (tid < nb) {
                const int n = n0 + tid;
                imu_factor_raw(frames + idx[2 * n] * kFrameStride, frames + idx[2 * n + 1] * kFrameStride,
                               recs + (size_t)n * kImuStride, wc, a.alias_bias, rraw + tid * 16, Jraw + tid * 450);
            } else if (tid >= 32) {
                // the other warps meanwhile stage W (one coalesced pass instead of 15 L2 loads per whitened entry)
                for (int e = tid - 32; e < nb * 225; e += nt - 32) {      // (lanes nb..31 of warp 0 idle with the factor threads)
                    const int n = e / 225;
                    Wsm[e] = __ldg(recs + (size_t)(n0 + n) * kImuStride + 11 + (e - n * 225));
                }
            }
            __syncthreads();
            SOLVE_STAMP(2);
            // whiten: Jx = [W J | W r | 0] (15 x 32 per factor)      :157 and the "sqrt_inv_cov *" lines
            // One thread per COLUMN of a factor: the raw column sits in registers, the rows of W are broadcast loads --
            // one shared-memory load per multiply-add instead of two (the entry-per-thread form was shared-memory bound).
            for (int e = tid; e < nb * 32; e += nt) {
                const int n = e >> 5, col = e & 31;
                double *out = Jx + n * kImuJx + col;
                if (col == 31) { for (int row = 0; row < 15; ++row) out[row * 32] = 0.0; continue; }
                double cv[15];
#pragma unroll
                for (int k = 0; k < 15; ++k) cv[k] = col < 30 ? Jraw[n * 450 + k * 30 + col] : rraw[n * 16 + k];
                const double *Wm = Wsm + n * 225;
                for (int row = 0; row < 15; ++row) {
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < 15; ++k) s += Wm[row * 15 + k] * cv[k];
                    out[row * 32] = s;
                }
            }
            if (tid == 0) {
                bool ch = true;                              // no two factors of equal parity share a frame
                for (int n = 0; n < nb && ch; ++n)
                    for (int m = n + 2; m < nb; m += 2) {
                        const int a0 = imu_fr[2 * n], a1 = imu_fr[2 * n + 1], b0 = imu_fr[2 * m], b1 = imu_fr[2 * m + 1];
                        ch &= (a0 != b0 && a0 != b1 && a1 != b0 && a1 != b1);
                    }
                imu_chain = ch ? 1 : 0;
            }
            __syncthreads();
            // accumulate Jx^T Jx: rows / columns 0..29 are the 30 x 30 block of the system, row 30 the gradient, entry
            // (30, 30) twice the cost.  One thread per 4 x 4 tile of the lower triangle (36 tiles per factor), 15 steps of
            // 2 x 4 loaded values for 16 multiply-adds.  Two factors touch the same entries of the system only if they
            // share a frame; the reference's factors form a chain (frame n, n + 1), so the factors at even positions of
            // the round go first, all at once, then the odd ones.  Any other round: one factor at a time.
            const bool chain = imu_chain != 0;
            const int phases = chain ? 2 : nb;
            for (int ph = 0; ph < phases; ++ph) {
                const int first = ph, step = chain ? 2 : nb, cnt = chain ? (nb - ph + 1) / 2 : 1;
                for (int e = tid; e < cnt * 36; e += nt) {
                    const int q = e / 36, t = e - q * 36;
                    const int n = first + q * step;
                    int tr = 0;
                    while ((tr + 1) * (tr + 2) / 2 <= t) ++tr;
                    const int tc = t - tr * (tr + 1) / 2;    // tile (tr, tc), tc <= tr < 8
                    const double *Jn = Jx + n * kImuJx;
                    double acc[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
#pragma unroll 5
                    for (int k = 0; k < 15; ++k) {
                        const double2 a0 = *reinterpret_cast<const double2 *>(Jn + k * 32 + 4 * tr);
                        const double2 a1 = *reinterpret_cast<const double2 *>(Jn + k * 32 + 4 * tr + 2);
                        const double2 b0 = *reinterpret_cast<const double2 *>(Jn + k * 32 + 4 * tc);
                        const double2 b1 = *reinterpret_cast<const double2 *>(Jn + k * 32 + 4 * tc + 2);
                        const double av[4] = {a0.x, a0.y, a1.x, a1.y}, bv[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * bv[j];
                    }
                    const int fi = imu_fr[2 * n], fj = imu_fr[2 * n + 1];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int ra = 4 * tr + i;
                        if (ra > 30) continue;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int cb_ = 4 * tc + j;
                            if (cb_ > ra || cb_ > 30) continue;          // the lower triangle of the 31 x 31 product, once
                            const double v = acc[i][j];
                            if (ra == 30) {
                                if (cb_ == 30) imu_cost[n] = v;
                                else { const int gb = (cb_ < 15 ? fi * 15 + cb_ : fj * 15 + cb_ - 15); g[gb] += v; gu[gb] += v; }
                            } else {
                                const int ga = (ra < 15 ? fi * 15 + ra : fj * 15 + ra - 15);
                                const int gb = (cb_ < 15 ? fi * 15 + cb_ : fj * 15 + cb_ - 15);
                                A[ga >= gb ? tri(ga, gb) : tri(gb, ga)] += v;
                            }
                        }
                    }
                }
                __syncthreads();
            }
            if (tid == 0) {                                  // in factor order: the cost does not depend on the schedule
                double c = 0.0;
                for (int n = 0; n < nb; ++n) c += imu_cost[n];
                cost_sm[1] += 0.5 * c;
            }
        }
    }

    SOLVE_STAMP(3);
    // ---- marginalisation prior (bundle_adjustor.cpp:126-139): no loss
    if (kFull && inertial && H.n_prior > 0) {
        const int n = H.n_prior, d = 15 * n, dcap = 15 * a.Ncap;
        const int32_t *pf = a.prior_frames + (size_t)w * a.Ncap;
        const double *S = a.prior_S + (size_t)w * dcap * dcap;
        const double *L = a.prior_L + (size_t)w * dcap * dcap;
        const double *ev = a.prior_e + (size_t)w * dcap;
        const double *x0 = a.prior_x0 + (size_t)w * a.Ncap * kFrameStride;
        double *r0 = scr;                 // [d] raw residual
        double *rr = r0 + d;              // [d] r = S r0 + e
        double *vv = rr + d;              // [d] S^T r
        double *Ji = vv + d;              // [n][9] Jr^-1
        double *Bs = Ji + 9 * n;          // [warps][225] block staging of the Hessian pass
        __shared__ int pf_s[kMaxFrames];
        if (tid >= 32 && tid < 32 + n) pf_s[tid - 32] = pf[tid - 32];
        if (tid < n) prior_frame_raw(frames + pf[tid] * kFrameStride, x0 + tid * kFrameStride, r0 + 15 * tid, Ji + 9 * tid);
        __syncthreads();
        // warp per row, kPriorRows rows per pass: the loads of a pass (rows x d / 32 per lane) are all in flight
        // together -- S sits in L2 and a row at a time was one exposed L2 round trip per row
        constexpr int kPriorRows = 5;
        for (int i0 = (tid >> 5) * kPriorRows; i0 < d; i0 += (nt >> 5) * kPriorRows) {
            double sr[kPriorRows];
#pragma unroll
            for (int r = 0; r < kPriorRows; ++r) sr[r] = 0.0;
#pragma unroll 2
            for (int k = tid & 31; k < d; k += 32) {
                double sv[kPriorRows];
#pragma unroll
                for (int r = 0; r < kPriorRows; ++r) sv[r] = i0 + r < d ? __ldg(S + (size_t)(i0 + r) * d + k) : 0.0;
                const double rk = r0[k];
#pragma unroll
                for (int r = 0; r < kPriorRows; ++r) sr[r] += sv[r] * rk;
            }
#pragma unroll
            for (int r = 0; r < kPriorRows; ++r) {
                double s = sr[r];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
                if ((tid & 31) == 0 && i0 + r < d) rr[i0 + r] = ev[i0 + r] + s;     // marginalization_error_cost.h:91
            }
        }
        __syncthreads();
        for (int i = tid; i < d; i += nt) {                    // S^T r: column i, the loads of 8 rows ahead of the sum
            double s = 0.0;
            int k = 0;
            for (; k + 8 <= d; k += 8) {
                double sv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) sv[u] = __ldg(S + (size_t)(k + u) * d + i);
#pragma unroll
                for (int u = 0; u < 8; ++u) s += sv[u] * rr[k + u];
            }
            for (; k < d; ++k) s += __ldg(S + (size_t)k * d + i) * rr[k];
            vv[i] = s;
        }
        if (tid == 0) { double c = 0.0; for (int k = 0; k < d; ++k) c += rr[k] * rr[k]; cost_sm[2] = 0.5 * c; }
        __syncthreads();
        SOLVE_STAMP(4);
        // g += E^T vv ; H += E^T Lambda E  with E = blockdiag(Jr^-1, I, I, I, I) per frame  (:72-88)
        for (int i = tid; i < d; i += nt) {
            const int fi = i / 15, ci = i - fi * 15;
            double s;
            if (ci < 3) { s = 0.0; for (int k = 0; k < 3; ++k) s += Ji[9 * fi + 3 * k + ci] * vv[15 * fi + k]; }
            else s = vv[i];
            const int gi = pf_s[fi] * 15 + ci;
            g[gi] += s; gu[gi] += s;
        }
        // H += E^T Lambda E by 15 x 15 frame blocks, one warp per block (fi, fj) of the lower triangle: the block of
        // Lambda comes in with 8 loads per lane in flight at once (the per-entry loop it replaces was a chain of exposed
        // L2 round trips), the Jr^-1 factors are applied to its first three columns, then rows, in shared memory
        {
            const int lane = tid & 31, wp = tid >> 5, nw = nt >> 5;
            double *B = Bs + wp * 225;
            for (int blk = wp; blk < n * n; blk += nw) {
                const int fi = blk / n, fj = blk - fi * n;
                const int gfi = pf_s[fi], gfj = pf_s[fj];
                if (gfj > gfi) continue;                              // (warp-uniform) the transposed block covers it
                double lv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = lane + 32 * u;
                    const int r = e / 15, c = e - r * 15;
                    lv[u] = e < 225 ? __ldg(L + (size_t)(15 * fi + r) * d + 15 * fj + c) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) if (lane + 32 * u < 225) B[lane + 32 * u] = lv[u];
                __syncwarp();
                double t0 = 0.0, t1 = 0.0;                            // columns 0..2 <- B[:, 0..2] Jr_j^-1
                {
                    const double *Jj = Ji + 9 * fj;
                    const int e0 = lane, e1 = lane + 32;              // 45 outputs (r, cj)
                    if (e0 < 45) { const int r = e0 / 3, cj = e0 - r * 3; for (int m = 0; m < 3; ++m) t0 += B[r * 15 + m] * Jj[3 * m + cj]; }
                    if (e1 < 45) { const int r = e1 / 3, cj = e1 - r * 3; for (int m = 0; m < 3; ++m) t1 += B[r * 15 + m] * Jj[3 * m + cj]; }
                    __syncwarp();
                    if (e0 < 45) B[(e0 / 3) * 15 + e0 % 3] = t0;
                    if (e1 < 45) B[(e1 / 3) * 15 + e1 % 3] = t1;
                    __syncwarp();
                }
                {
                    const double *Jf = Ji + 9 * fi;                   // rows 0..2 <- Jr_i^-T B[0..2, :]
                    const int e0 = lane, e1 = lane + 32;              // 45 outputs (ci, c)
                    t0 = 0.0; t1 = 0.0;
                    if (e0 < 45) { const int ci = e0 / 15, c = e0 - ci * 15; for (int k = 0; k < 3; ++k) t0 += Jf[3 * k + ci] * B[k * 15 + c]; }
                    if (e1 < 45) { const int ci = e1 / 15, c = e1 - ci * 15; for (int k = 0; k < 3; ++k) t1 += Jf[3 * k + ci] * B[k * 15 + c]; }
                    __syncwarp();
                    if (e0 < 45) B[e0] = t0;
                    if (e1 < 45) B[e1] = t1;
                    __syncwarp();
                }
                for (int e = lane; e < 225; e += 32) {                // each (gi, gj) has a unique owner: prior frames are distinct
                    const int r = e / 15, c = e - r * 15;
                    if (fi == fj && c > r) continue;
                    A[tri(gfi * 15 + r, gfj * 15 + c)] += B[e];
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }

    SOLVE_STAMP(5);
    // ---- plane factors (bundle_adjustor.cpp:162-196): CauchyLoss
    if (kFull && H.n_ptracks > 0) {
        const double *pl = a.plane_param + (size_t)w * a.Pcap * 4;
        const int32_t *ptp = a.pt_plane + (size_t)w * a.Tcap;
        const int32_t *ptb = a.pt_begin + (size_t)w * (a.Tcap + 1);
        const int32_t *ptf = a.pt_frame + (size_t)w * a.Ocap;
        const float *ptz = a.pt_z + (size_t)w * a.Ocap * 2;
        const double cb = wc.cauchy_a * wc.cauchy_a;
        // thread per track: the corrected 1 x 6K Jacobian row, spread over the window's frames (zeros where the track has
        // no observation), and the corrected residual go to a scratch row; then thread per ENTRY of the pose block of the
        // system sums over the tracks -- no atomics (fp64 atomicAdd on shared memory is a compare-and-swap loop, and all
        // tracks of a plane hit the same frames: the atomic version took 600 us of the kernel's 820 on cfg4)
        const int PW = 6 * N + 2;                                   // row: [6 N] Jacobian, residual
        // the rows live in the kernel's shared scratch when they fit (inertial windows: the IMU slabs are free again),
        // otherwise in the per-window global scratch
        unsigned dyn_bytes;
        asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_bytes));
        const size_t scr_words = dyn_bytes / sizeof(double) - (size_t)(scr - reinterpret_cast<double *>(smem_raw));
        double *PJ = (size_t)H.n_ptracks * PW <= scr_words ? scr : a.pt_J + (size_t)w * a.Tcap * (6 * a.Ncap + 2);
        for (int t = tid; t < H.n_ptracks; t += nt) {
            const int b0 = ptb[t], K = ptb[t + 1] - b0;
            double r, J[6 * kMaxFrames];
            plane_factor(K, ptf + b0, ptz + 2 * b0, frames, wc, pl + 4 * ptp[t], wc.plane_sic, &r, J);
            const double s = r * r, tt = 1.0 + s / cb;
            const double sc = sqrt(1.0 / tt);
            atomicAdd(&cost_sm[3], 0.5 * cb * log(tt));
            double *row = PJ + (size_t)t * PW;
            for (int i = 0; i < 6 * N; ++i) row[i] = 0.0;
            for (int i = 0; i < 6 * K; ++i) row[ptf[b0 + i / 6] * 6 + (i % 6)] = J[i] * sc;
            row[6 * N] = r * sc;
        }
        __syncthreads();
        const int P6 = 6 * N, T = H.n_ptracks;
        for (int e = tid; e < P6 * (P6 + 1) / 2 + P6; e += nt) {
            if (e >= P6 * (P6 + 1) / 2) {                            // gradient entries
                const int i = e - P6 * (P6 + 1) / 2;
                double sg = 0.0;
                for (int t = 0; t < T; ++t) sg += PJ[(size_t)t * PW + i] * PJ[(size_t)t * PW + P6];
                const int gi = (i / 6) * stride + (i % 6);
                g[gi] += sg; gu[gi] += sg;
                continue;
            }
            int i = (int)((sqrtf(8.0f * e + 1.0f) - 1.0f) * 0.5f);
            while ((i + 1) * (i + 2) / 2 <= e) ++i;
            while (i * (i + 1) / 2 > e) --i;
            const int j = e - i * (i + 1) / 2;                       // j <= i
            double sh = 0.0;
            for (int t = 0; t < T; ++t) sh += PJ[(size_t)t * PW + i] * PJ[(size_t)t * PW + j];
            const int gi = (i / 6) * stride + (i % 6), gj = (j / 6) * stride + (j % 6);
            A[tri(gi, gj)] += sh;                                    // gi >= gj since i >= j
        }
        __syncthreads();
    }

    // ---- optional dump of the assembled reduced system (delta coordinates, no regulariser)
    if (a.Hfull) {
        const int Df = 15 * a.Ncap;
        double *Ho = a.Hfull + (size_t)w * Df * Df, *go = a.gfull + (size_t)w * Df;
        for (int e = tid; e < D * D; e += nt) {
            const int i = e / D, j = e - i * D;
            int fi = i / stride, fj = j / stride;
            const int ci = i - fi * stride, cj = j - fj * stride;
            if (compact) { fi = __fns(freem, 0, fi + 1); fj = __fns(freem, 0, fj + 1); }     // back to the window's frame ids
            Ho[(size_t)(fi * 15 + ci) * Df + fj * 15 + cj] = (i >= j) ? A[tri(i, j)] : A[tri(j, i)];
        }
        for (int i = tid; i < D; i += nt) {
            const int fi = i / stride;
            go[(compact ? __fns(freem, 0, fi + 1) : fi) * 15 + (i % stride)] = g[i];
        }
    }
    SOLVE_STAMP(6);
    // ---- Jacobi scale, LM diagonal, constant blocks
    double *scale = a.pose_scale + (size_t)w * 15 * a.Ncap;
    double my_gdx = 0.0;
    for (int i = tid; i < D; i += nt) {
        const int f = i / stride, c = i - f * stride;
        const double hii = A[tri(i, i)] + hcorr[i];          // diagonal of the UNREDUCED J^T J
        double sc;
        if (compute_scale) { sc = 1.0 / (1.0 + sqrt(fmax(hii, 0.0))); scale[i] = sc; }
        else sc = scale[i];
        const double reg = mu > 0.0 ? lm_reg(hii, sc, mu) : 0.0;
        xs[i] = reg;                                          // keep reg for the scalars below
        A[tri(i, i)] += reg;
        hcorr[i] = hii;                                       // now holds the unreduced diagonal
    }
    __syncthreads();
    const int fixed_mask = mask_c;
    {
        // identity rows / columns for the pinned unknowns: the pose rows of constant frames and the padding rows >= D.
        // Work item (q, j): the q-th pinned index m against column j; a pair of pinned indices is written by the larger one.
        const unsigned fm = (unsigned)fixed_mask & ((1u << nfr) - 1u);
        const int nfix6 = 6 * __popc(fm), npin = nfix6 + (Dp - D);
        for (int e = tid; e < npin * Dp; e += nt) {
            const int q = e / Dp, j = e - q * Dp;
            int m;
            if (q < nfix6) { const int fq = q / 6; m = __fns(fm, 0, fq + 1) * stride + (q - fq * 6); }
            else m = D + (q - nfix6);
            const int fj = j / stride;
            const bool pj = j >= D || (((fm >> fj) & 1u) && (j - fj * stride) < 6);
            if (pj && j > m) continue;
            A[j > m ? tri(j, m) : tri(m, j)] = (j == m) ? 1.0 : 0.0;
        }
    }
    double *reg_keep = scr;       // [D]
    for (int i = tid; i < D; i += nt) {
        const int f = i / stride, c = i - f * stride;
        const bool m = ((fixed_mask >> f) & 1) && c < 6;
        reg_keep[i] = xs[i];
        if (m) { g[i] = 0.0; gu[i] = 0.0; }
        xs[i] = -g[i];
    }
    __syncthreads();
    SOLVE_STAMP(7);
    // wide CTA (8 warps) and a system whose off-diagonal tiles fit into 4 (5) register slots per thread of warps 1..7,
    // its diagonal into the 64 slots of warp 0: the register-resident factorisation
    const int n_off = nb * (nb - 1) / 2 + nb;
    const bool ok = (kFull && nt == 256 && nb <= 64 && n_off <= 4 * 224) ? chol_solve_regs<4>(A, xs, nb, &flag_sm)
                  : (kFull && nt == 256 && nb <= 64 && n_off <= 5 * 224) ? chol_solve_regs<5>(A, xs, nb, &flag_sm)
                                                                         : chol_solve_tiled<kFull>(A, xs, nb, &flag_sm);
    SOLVE_STAMP(8);

    // ---- outputs
    double *dxo = a.dx_pose + (size_t)w * a.Ncap * 15;
    for (int i = tid; i < N * 15; i += nt) {
        const int f = i / 15, c = i - f * 15;
        const bool in_sys = (freem >> f) & 1u;               // dropped (constant) frames do not move
        dxo[i] = (ok && c < stride && in_sys) ? xs[__popc(freem & ((1u << f) - 1u)) * stride + c] : 0.0;
    }
    {
        double gdx = 0.0, rdx = 0.0, gn2 = 0.0, dx2 = 0.0, gmax = 0.0, g2 = 0.0, vrd = 0.0;
        double *vpo = a.v_pose + (size_t)w * a.Ncap * 15;
        for (int i = tid; i < N * 15; i += nt) vpo[i] = 0.0;
        for (int i = tid; i < D; i += nt) {
            const int f = i / stride, c = i - f * stride;
            const bool m = ((fixed_mask >> f) & 1) && c < 6;
            if (m) continue;
            const double dx = ok ? xs[i] : 0.0;
            const double sc = scale[i];
            double d2 = sc * sc * hcorr[i];
            d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
            // NOTE: the reduced gradient g and the unreduced gu differ only through the Schur term;
            // g . dx over the full system is completed by the landmark sweep (acc[1])
            const double v = sc * sc * gu[i] / d2;
            vpo[(compact ? __fns(freem, 0, f + 1) : f) * 15 + c] = v;
            gdx += g[i] * dx;
            rdx += reg_keep[i] * dx * dx;
            gn2 += d2 * (dx / sc) * (dx / sc);
            dx2 += dx * dx;
            g2 += v * gu[i];
            vrd += v * reg_keep[i] * dx;
            gmax = fmax(gmax, fabs(gu[i]));
        }
        __syncthreads();
        double *red = A;                 // the factor is no longer needed
        if (tid < D) { red[tid * 7 + 0] = gdx; red[tid * 7 + 1] = rdx; red[tid * 7 + 2] = gn2; red[tid * 7 + 3] = dx2; red[tid * 7 + 4] = gmax;
                       red[tid * 7 + 5] = g2; red[tid * 7 + 6] = vrd; }
        __syncthreads();
        if (tid < 32) {                  // warp 0: strided partial sums, then a shuffle tree (fixed order: reproducible)
            gdx = rdx = gn2 = dx2 = gmax = g2 = vrd = 0.0;
            const int nred = min(nt, D);
            for (int t = tid; t < nred; t += 32) {
                gdx += red[t * 7]; rdx += red[t * 7 + 1]; gn2 += red[t * 7 + 2]; dx2 += red[t * 7 + 3];
                gmax = fmax(gmax, red[t * 7 + 4]); g2 += red[t * 7 + 5]; vrd += red[t * 7 + 6];
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                gdx += __shfl_xor_sync(0xffffffffu, gdx, off); rdx += __shfl_xor_sync(0xffffffffu, rdx, off);
                gn2 += __shfl_xor_sync(0xffffffffu, gn2, off); dx2 += __shfl_xor_sync(0xffffffffu, dx2, off);
                gmax = fmax(gmax, __shfl_xor_sync(0xffffffffu, gmax, off));
                g2 += __shfl_xor_sync(0xffffffffu, g2, off); vrd += __shfl_xor_sync(0xffffffffu, vrd, off);
            }
        }
        if (tid == 0) {
            ctrl.grad2 = g2;
            ctrl.v_reg_dx = vrd;
            double x2 = 0.0;
            for (int f = 0; f < N; ++f) x2 += x2_sm[f];
            ctrl.g_dot_dx = gdx;
            ctrl.dx_reg_dx = rdx;
            ctrl.gn_norm2 = gn2;
            ctrl.dxnorm2 = dx2;
            ctrl.xnorm2 = x2;
            ctrl.gmax = gmax;
            const double cv = a.cost_vis[bsel * a.bufs.cost + w];
            ctrl.cost_vis = cv;
            ctrl.cost = cv + cost_sm[1] + cost_sm[2] + cost_sm[3];
            ctrl.have_lin = 1;
            ctrl.solve_failed = ok ? 0 : 1;
            if (compute_scale) ctrl.have_scale = 1;
            ctrl.fresh = 1;
        }
    }
    SOLVE_STAMP(9);
}

static __global__ void __launch_bounds__(256) solve_kernel(SolveArgs a) { solve_body<true>(a); }
// visual-only windows (no IMU / prior / plane factors): small register footprint, many CTAs per SM
static __global__ void __launch_bounds__(64, 10) solve_kernel_visual(SolveArgs a) { solve_body<false>(a); }

// Non-vision part of the cost at the candidate state (IMU + prior + plane), one CTA per window.
struct CostArgs {
    const WinHdr *hdr;
    const WinConst *cst;
    const double *frames_cand;
    const double *frames_lin;      // bias linearisation only through the records; unused otherwise
    const int32_t *imu_idx;
    const double *imu_data;
    int alias_bias;                // 1: candidate dbg = cand.bg - current.bg
    const double *frames_cur;
    const int32_t *prior_frames;
    const double *prior_S, *prior_e, *prior_x0;
    const double *plane_param;
    const int32_t *pt_plane, *pt_begin, *pt_frame;
    const float *pt_z;
    int Pcap, Tcap, Ocap, Ncap;
    int w0;
    double *out;                   // [W] non-vision candidate cost
    // step acceptance (tail of the kernel)
    WinCtrl *ctrl;
    const double *acc;             // [W][kAcc] scalars of the sweeps
    double *frames_state;          // [W][Ncap][16] accepted state (overwritten by the candidate on acceptance)
    double *rho_state;             // [W][Mcap]
    const double *rho_cand;
    int Mcap;
    int loop;                      // 1: tr_decide (device-side trust-region loop); 0: plain Gauss-Newton step, `apply` decides
    const double *cost_vis;        // loop: [2][W] reprojection cost of the sweeps; the candidate's is in the window's OTHER buffer set
    size_t cost_stride;
    int apply;
    double beta;                   // loop == 0: the step was beta * dx_gn (model change of the truncated step)
};

// kPart 0: cost of the IMU / prior / plane blocks at the candidate + the step decision (one launch, batches);
// 1: the cost only (-> a.out[w]); 2: the decision only, cost read from a.out[w].  The latency path launches part 1 on a
// parallel branch of the solve graph, beside the candidate's linearisation sweep: both only need the candidate.
template <int kPart>
static __global__ void aux_cost_kernel(CostArgs a) {
    const int w = blockIdx.x + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double *fc = a.frames_cand + (size_t)w * a.Ncap * kFrameStride;
    const double *fcur = a.frames_cur + (size_t)w * a.Ncap * kFrameStride;
    __shared__ double acc;
    __shared__ int accept_sm;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *r0 = reinterpret_cast<double *>(smem_raw);
    WinCtrl &ctrl = a.ctrl[w];
    if (a.loop && (ctrl.done || ctrl.skip)) { if (tid == 0 && kPart != 1) ctrl.fresh = 0; return; }
    if (tid == 0) acc = kPart == 2 ? a.out[w] : 0.0;
    __syncthreads();
    if (kPart != 2 && H.use_inertial) {
        const int32_t *idx = a.imu_idx + (size_t)w * a.Ncap * 2;
        const double *recs = a.imu_data + (size_t)w * a.Ncap * kImuStride;
        __shared__ double imu_c[kMaxFrames];
        const int n_pr = H.n_prior, d = 15 * n_pr, dcap = 15 * a.Ncap;
        const int32_t *pf = a.prior_frames + (size_t)w * a.Ncap;
        const double *x0 = a.prior_x0 + (size_t)w * a.Ncap * kFrameStride;
        double *rc = r0 + 15 * kMaxFrames;                    // [d] squared whitened prior residuals
        for (int n = tid; n < H.n_imu; n += nt) {
            // residual only: the raw evaluator without Jacobian.  Q1 (alias_bias): the bias linearisation point is the
            // CURRENT (accepted) bias of frame i while the state is the candidate
            double r[15];
            const double *rec = recs + (size_t)n * kImuStride;
            imu_factor_raw(fc + idx[2 * n] * kFrameStride, fc + idx[2 * n + 1] * kFrameStride, rec, wc, 0, r, nullptr,
                           a.alias_bias ? fcur + idx[2 * n] * kFrameStride + 10 : nullptr);
            double c = 0.0;
            for (int i = 0; i < 15; ++i) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 15; ++k) s += __ldg(rec + 11 + i * 15 + k) * r[k];
                c += s * s;
            }
            imu_c[n] = c;
        }
        {   // the prior's raw residuals meanwhile on the second warp (a CTA of one warp: after the IMU factors)
            const int pt = nt > 32 ? tid - 32 : tid;
            double Ji[9];
            if (pt >= 0 && pt < n_pr) prior_frame_raw(fc + pf[pt] * kFrameStride, x0 + pt * kFrameStride, r0 + 15 * pt, Ji);
        }
        __syncthreads();
        if (n_pr > 0) {
            const double *S = a.prior_S + (size_t)w * dcap * dcap;
            const double *ev = a.prior_e + (size_t)w * dcap;
            // r = S r0 + e, warp per row, kRows rows per pass with all their loads in flight (S sits in L2)
            constexpr int kRows = 5;
            for (int i0 = (tid >> 5) * kRows; i0 < d; i0 += (nt >> 5) * kRows) {
                double sr[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) sr[r] = 0.0;
#pragma unroll 2
                for (int k = tid & 31; k < d; k += 32) {
                    double sv[kRows];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) sv[r] = i0 + r < d ? __ldg(S + (size_t)(i0 + r) * d + k) : 0.0;
                    const double rk = r0[k];
#pragma unroll
                    for (int r = 0; r < kRows; ++r) sr[r] += sv[r] * rk;
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    double sx = sr[r];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) sx += __shfl_xor_sync(0xffffffffu, sx, off);
                    if ((tid & 31) == 0 && i0 + r < d) { const double v = ev[i0 + r] + sx; rc[i0 + r] = v * v; }
                }
            }
            __syncthreads();
        }
        if (tid == 0) {                                        // fixed summation order: the cost is reproducible
            double c = 0.0;
            for (int n = 0; n < H.n_imu; ++n) c += imu_c[n];
            double cp = 0.0;
            for (int i = 0; i < d; ++i) cp += rc[i];
            acc += 0.5 * c + 0.5 * cp;
        }
        __syncthreads();
    }
    if (kPart != 2 && H.n_ptracks > 0) {
        const double *pl = a.plane_param + (size_t)w * a.Pcap * 4;
        const int32_t *ptp = a.pt_plane + (size_t)w * a.Tcap;
        const int32_t *ptb = a.pt_begin + (size_t)w * (a.Tcap + 1);
        const int32_t *ptf = a.pt_frame + (size_t)w * a.Ocap;
        const float *ptz = a.pt_z + (size_t)w * a.Ocap * 2;
        const double cb = wc.cauchy_a * wc.cauchy_a;
        for (int t = tid; t < H.n_ptracks; t += nt) {
            const int b0 = ptb[t], K = ptb[t + 1] - b0;
            double r, J[6 * kMaxFrames];
            plane_factor(K, ptf + b0, ptz + 2 * b0, fc, wc, pl + 4 * ptp[t], wc.plane_sic, &r, J);
            atomicAdd(&acc, 0.5 * cb * log(1.0 + r * r / cb));
        }
    }
    __syncthreads();
    if (kPart == 1) { if (tid == 0) a.out[w] = acc; return; }
    // ---- step acceptance: what ceres does after evaluating the candidate (trust_region_minimizer.cc)
    if (tid == 0) {
        a.out[w] = acc;
        const double *av = a.acc + (size_t)w * kAcc;
        bool accept;
        if (a.loop) {
            WinCtrl c = ctrl;                   // one round trip of the record instead of a dependent chain of global accesses
            accept = tr_decide(c, av, a.cost_vis[(size_t)(1 - c.buf) * a.cost_stride + w], acc);
            if (accept) c.buf ^= 1;             // the candidate's linearisation becomes the state's
            ctrl = c;
        }
        else {
            ctrl.cand_cost_vis = av[0];
            ctrl.cand_cost = av[0] + acc;
            const double gdx = ctrl.g_dot_dx + av[1];           // g . dx over poses + landmarks (full GN step)
            const double rdx = ctrl.dx_reg_dx + av[2];          // dx^T (mu D) dx
            // model cost change of the step beta * dx_gn:  -(beta g.dx + beta^2/2 dx^T H dx),
            // with dx^T H dx = -g.dx - dx^T (mu D) dx for the regularised Gauss-Newton step
            ctrl.model_change = -a.beta * gdx + 0.5 * a.beta * a.beta * (gdx + rdx);
            accept = a.apply != 0;
        }
        accept_sm = accept ? 1 : 0;
    }
    __syncthreads();
    if (accept_sm) {             // the candidate becomes the state
        const int N = H.N, M = H.M;
        for (int i = tid; i < N * kFrameStride; i += nt) a.frames_state[(size_t)w * a.Ncap * kFrameStride + i] = fc[i];
        for (int i = tid; i < M; i += nt) a.rho_state[(size_t)w * a.Mcap + i] = a.rho_cand[(size_t)w * a.Mcap + i];
    }
}

// |J v|^2 over the IMU / prior / plane blocks (loss-corrected), one CTA per window: the non-vision
// part of the Cauchy-point denominator of the dogleg step.
struct JvAuxArgs {
    CostArgs c;               // same inputs as aux_cost_kernel (frames_cand unused; frames_cur = current state)
    const double *v_pose;     // [W][Ncap][15]
    double *acc;              // [W][kAcc], slot 10
};  // c.loop != 0: runs only for windows whose GN step left the trust region, then picks the dogleg step (tr_after_jv)

static __global__ void jv_aux_kernel(JvAuxArgs ja) {
    const CostArgs &a = ja.c;
    const int w = blockIdx.x + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int tid = threadIdx.x, nt = blockDim.x;
    const double *fr = a.frames_cur + (size_t)w * a.Ncap * kFrameStride;
    const double *v = ja.v_pose + (size_t)w * a.Ncap * 15;
    __shared__ double acc;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *ev = reinterpret_cast<double *>(smem_raw);      // [15 n_prior] E v
    WinCtrl &ctrl = a.ctrl[w];
    if (a.loop && (ctrl.done || ctrl.skip)) return;
    if (a.loop && !ctrl.need_jv) { if (tid == 0) { WinCtrl c = ctrl; tr_after_jv(c, ja.acc + (size_t)w * kAcc); ctrl = c; } return; }
    if (tid == 0) acc = 0.0;
    __syncthreads();
    if (H.use_inertial) {
        const int32_t *idx = a.imu_idx + (size_t)w * a.Ncap * 2;
        const double *recs = a.imu_data + (size_t)w * a.Ncap * kImuStride;
        for (int n = tid; n < H.n_imu; n += nt) {
            double r[15], J[450], u[15];
            const double *rec = recs + (size_t)n * kImuStride;
            const int fi = idx[2 * n], fj = idx[2 * n + 1];
            imu_factor_raw(fr + fi * kFrameStride, fr + fj * kFrameStride, rec, wc, a.alias_bias, r, J);
            for (int k = 0; k < 15; ++k) {
                double s = 0.0;
                for (int c = 0; c < 15; ++c) s += J[k * 30 + c] * v[fi * 15 + c] + J[k * 30 + 15 + c] * v[fj * 15 + c];
                u[k] = s;
            }
            double c2 = 0.0;
            for (int i = 0; i < 15; ++i) {
                double s = 0.0;
                for (int k = 0; k < 15; ++k) s += rec[11 + i * 15 + k] * u[k];
                c2 += s * s;
            }
            atomicAdd(&acc, c2);
        }
        if (H.n_prior > 0) {
            const int n = H.n_prior, d = 15 * n, dcap = 15 * a.Ncap;
            const int32_t *pf = a.prior_frames + (size_t)w * a.Ncap;
            const double *S = a.prior_S + (size_t)w * dcap * dcap;
            const double *x0 = a.prior_x0 + (size_t)w * a.Ncap * kFrameStride;
            if (tid < n) {
                double r15[15], Ji[9];
                prior_frame_raw(fr + pf[tid] * kFrameStride, x0 + tid * kFrameStride, r15, Ji);
                const double *vf = v + pf[tid] * 15;
                for (int k = 0; k < 3; ++k) ev[15 * tid + k] = Ji[3 * k] * vf[0] + Ji[3 * k + 1] * vf[1] + Ji[3 * k + 2] * vf[2];
                for (int k = 3; k < 15; ++k) ev[15 * tid + k] = vf[k];
            }
            __syncthreads();
            double c2 = 0.0;
            for (int i = tid; i < d; i += nt) {
                double s = 0.0;
                for (int k = 0; k < d; ++k) s += S[(size_t)i * d + k] * ev[k];
                c2 += s * s;
            }
            atomicAdd(&acc, c2);
        }
    }
    if (H.n_ptracks > 0) {
        const int stride = 15;
        const double *pl = a.plane_param + (size_t)w * a.Pcap * 4;
        const int32_t *ptp = a.pt_plane + (size_t)w * a.Tcap;
        const int32_t *ptb = a.pt_begin + (size_t)w * (a.Tcap + 1);
        const int32_t *ptf = a.pt_frame + (size_t)w * a.Ocap;
        const float *ptz = a.pt_z + (size_t)w * a.Ocap * 2;
        const double cb = wc.cauchy_a * wc.cauchy_a;
        for (int t = tid; t < H.n_ptracks; t += nt) {
            const int b0 = ptb[t], K = ptb[t + 1] - b0;
            double r, J[6 * kMaxFrames];
            plane_factor(K, ptf + b0, ptz + 2 * b0, fr, wc, pl + 4 * ptp[t], wc.plane_sic, &r, J);
            const double sc2 = 1.0 / (1.0 + r * r / cb);             // rho'
            double s = 0.0;
            for (int i = 0; i < 6 * K; ++i) s += J[i] * v[ptf[b0 + i / 6] * stride + (i % 6)];
            atomicAdd(&acc, sc2 * s * s);
        }
    }
    __syncthreads();
    if (tid == 0) {
        double *av = ja.acc + (size_t)w * kAcc;
        if (acc != 0.0) av[10] += acc;           // one CTA per window owns slot 10
        if (a.loop) { WinCtrl c = ctrl; tr_after_jv(c, av); ctrl = c; }
    }
}

}  // namespace pvio
