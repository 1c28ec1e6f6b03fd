// Back-substitution of the inverse depths, Plus() update and candidate-cost sweep.
//
// After the reduced system has been solved for the pose step, this kernel does what ceres
// does after the Schur solve of an iteration (bundle_adjustor.cpp:249, inside ceres::Solve):
//   d rho_l = -(g_l + h_l . dxi) / (H_ll + reg)                 back-substitution
//   candidate = Plus(x, dx)         (quaternion_parameterization.h:28-31, plain + elsewhere)
//   candidate cost = sum rho(|r|^2)/2 over the reprojection blocks   (step acceptance)
// and accumulates the scalars the trust-region logic needs.  sqrt(w_l) h_l comes from the linearise kernel
// (LinArgs::hs_out: one record of 6 N floats per landmark slot, written there by bulk copies).  Recomputing it from the observation table costs one full linearisation per observation and was
// 60 % of this kernel's instructions (profiles/r01d_update.md); the kernels are issue-bound and HBM is at
// 5 % of its bandwidth, so 2 x 108 KB of extra traffic per window are the cheaper side of the trade.
#pragma once
#include "ba_lin.cuh"

namespace pvio {

struct UpdArgs {
    const WinHdr *hdr;
    const WinConst *cst;
    const ObsRec *obs;
    const LmRec *lms;
    const double *rho;        // [W][Mcap] current
    const double *frames;     // [W][Ncap][16] current
    const WinCtrl *ctrl;
    const double *lm_scale;
    const LmAux *lm_aux;
    const float *hs;          // [W][hs_stride] from the linearise kernel (LinArgs::hs_out)
    size_t hs_stride;
    const double *dx_pose;    // [W][Ncap][15] pose/motion step in delta coordinates
    double *rho_cand;         // [W][Mcap]
    double *frames_cand;      // [W][Ncap][16]
    double *dx_lm;            // [W][Mcap]
    double *acc;              // [W][kAcc]: 0 cand_cost_vis, 1 g.dx(lm), 2 dx.reg.dx(lm), 3 gn_norm2(lm), 4 |step|^2(lm),
                              //   5 |x|^2(lm), 6 ambient |x - x_cand|^2 (frames), 7 |D^-1 S g|^2 (lm), 8 v.reg.dx (lm), 9-10 |J v|^2
    int Ncap, Mcap, Kcap;
    double mu_override;
    int w0;
    // step = step_b * dx_gn - step_a * v, v = S^2 g / clamp(S^2 diag H) the scaled steepest-descent direction:
    // (0, 1) Gauss-Newton, (0, beta) truncated GN, (a, b) dogleg interpolation, (a, 0) Cauchy leg
    double step_a, step_b;
    const double *v_pose;     // [W][Ncap][15] from solve_kernel
};

// Second generation: one LANE per landmark (see ba_lin2.cuh).  Same outputs.
template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 4)
update_tpl_kernel(UpdArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;

    __shared__ FrameSm Fc[kMaxFrames];     // candidate state
    __shared__ double dxi[kMaxFrames][6];  // xi = T delta per frame
    __shared__ double red[8];

    if (tid < N) {
        const double *fs = a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride;
        const double *d = a.dx_pose + ((size_t)w * a.Ncap + tid) * 15;
        double fc[kFrameStride];
        const double *vp = a.v_pose + ((size_t)w * a.Ncap + tid) * 15;
        double de[15];
#pragma unroll
        for (int i = 0; i < 15; ++i) de[i] = a.step_b * d[i] - (a.step_a != 0.0 ? a.step_a * vp[i] : 0.0);
        quat_plus(fs, de, fc);
#pragma unroll
        for (int i = 0; i < 12; ++i) fc[4 + i] = fs[4 + i] + de[3 + i];
        make_frame(fc, wc, Fc[tid]);
        if (blockIdx.x == 0) {
            double *o = a.frames_cand + ((size_t)w * a.Ncap + tid) * kFrameStride;
            double amb = 0.0;
#pragma unroll
            for (int i = 0; i < kFrameStride; ++i) {
                o[i] = fc[i];
                const bool pose = i < 7;
                const bool live = pose ? !((H.fixed_mask >> tid) & 1) : (H.use_inertial != 0);
                if (live) amb += (fc[i] - fs[i]) * (fc[i] - fs[i]);
            }
            atomicAdd(&a.acc[(size_t)w * kAcc + 6], amb);
        }
        double R[9], om[3];
        quat_to_mat(fs, R);
        mat3_vec(R, d, om);
        const double p0 = fs[4] - wc.origin[0], p1 = fs[5] - wc.origin[1], p2 = fs[6] - wc.origin[2];
        dxi[tid][0] = om[0]; dxi[tid][1] = om[1]; dxi[tid][2] = om[2];
        dxi[tid][3] = -(p1 * om[2] - p2 * om[1]) - d[3];
        dxi[tid][4] = -(p2 * om[0] - p0 * om[2]) - d[4];
        dxi[tid][5] = -(p0 * om[1] - p1 * om[0]) - d[5];
    }
    if (tid < 8) red[tid] = 0.0;
    __syncthreads();

    const float W[4] = {(float)wc.sic[0], (float)wc.sic[1], (float)wc.sic[2], (float)wc.sic[3]};
    const float cb = (float)(wc.cauchy_a * wc.cauchy_a);
    const double mu = a.mu_override >= 0.0 ? a.mu_override : a.ctrl[w].mu;
    const ObsRec *obs = a.obs + (size_t)w * a.Kcap;
    const LmRec *lms = a.lms + (size_t)w * a.Mcap;
    const double *rho = a.rho + (size_t)w * a.Mcap;
    const double *lm_scale = a.lm_scale + (size_t)w * a.Mcap;
    const LmAux *aux = a.lm_aux + (size_t)w * a.Mcap;
    double *rho_c = a.rho_cand + (size_t)w * a.Mcap;
    double *dxl = a.dx_lm + (size_t)w * a.Mcap;

    double s_cost = 0.0, s_gdx = 0.0, s_reg = 0.0, s_gn = 0.0, s_dx2 = 0.0, s_x2 = 0.0, s_g2 = 0.0, s_vrd = 0.0;
    // chunk c -> (CTA c mod grid, warp (c / grid) mod 8): a single window spreads over all CTAs
    for (int ci = wv;; ci += 8) {
        const int ch = ci * gridDim.x + blockIdx.x;
        if (ch >= H.n_chunks) break;
        const int lm0 = H.chunk_begin[ch];
        const int cnt = H.chunk_meta[ch] & 0xff;
        const int anchor = H.chunk_meta[ch] >> 8;
        if (lane >= cnt) continue;
        const int l = lm0 + lane;
        const LmRec lr = lms[l];
        const int n_obs = lm_nobs(lr.meta);
        const double rl = rho[l];
        double hdx = 0.0, hll = 0.0;
        double drho = 0.0;
        if (n_obs > 0) {
            const LmAux ax = aux[l];
            hll = ax.hll;
            // h_l . dxi = sqrt(H_ll + reg) * sum_f (sqrt(w) h_lf) . dxi_f over the observing frames and the anchor
            const float2 *hs = reinterpret_cast<const float2 *>(a.hs + (size_t)w * a.hs_stride + (size_t)(ch * 32 + lane) * hs_rec(N));
            for (unsigned fm = lm_mask(lr.meta) | (1u << anchor); fm; fm &= fm - 1) {
                const int f = __ffs(fm) - 1;
                const float2 h01 = hs[f * 3], h23 = hs[f * 3 + 1], h45 = hs[f * 3 + 2];
                hdx += (double)h01.x * dxi[f][0] + (double)h01.y * dxi[f][1] + (double)h23.x * dxi[f][2] +
                       (double)h23.y * dxi[f][3] + (double)h45.x * dxi[f][4] + (double)h45.y * dxi[f][5];
            }
            hdx *= sqrt(ax.hll_reg);
            const double wl = 1.0 / ax.hll_reg;
            drho = isfinite(wl) ? -(ax.gl + hdx) * wl : 0.0;
            const double sc = lm_scale[l];
            const double reg = mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0;
            double d2 = sc * sc * hll;
            d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
            const double vl = sc * sc * ax.gl / d2;          // scaled steepest-descent direction
            s_gdx += hdx * ax.gl * wl + ax.gl * drho;
            s_reg += reg * drho * drho;
            s_gn += d2 * (drho / sc) * (drho / sc);
            s_g2 += vl * ax.gl;
            s_vrd += vl * reg * drho;
            s_x2 += rl * rl;
            drho = a.step_b * drho - a.step_a * vl;
            s_dx2 += drho * drho;
        }
        rho_c[l] = rl + drho;
        dxl[l] = drho;
        if (n_obs > 0) {
            double xc[3];
            float xcf[3], clc[3];
            world_point(Fc[anchor], lr.zrx, lr.zry, rl + drho, xc, xcf, clc);
            for (unsigned j = 0, fm = lm_mask(lr.meta); j < (unsigned)n_obs; ++j, fm &= fm - 1) {
                const ObsRec o = obs[lr.obs_begin + j];
                const int of = __ffs(fm) - 1;
                s_cost += (double)residual_cost<kLoss>(Fc[of], xc, o.zx, o.zy, W, cb);
            }
        }
    }
    double v[8] = {s_cost, s_gdx, s_reg, s_gn, s_dx2, s_x2, s_g2, s_vrd};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0 && v[k] != 0.0) atomicAdd(&red[k], v[k]);
    }
    __syncthreads();
    if (tid < 8) {
        const int slot = tid < 6 ? tid : tid + 1;            // slot 6 is the frames' ambient step norm
        if (red[tid] != 0.0) atomicAdd(a.acc + (size_t)w * kAcc + slot, red[tid]);
    }
}

// |J v|^2 over the reprojection blocks (loss-corrected Jacobians), for the Cauchy point of the dogleg
// step: alpha = |grad|^2 / |J S D^-1 grad|^2 (ceres dogleg_strategy.cc ComputeCauchyPoint).
template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 2)
jv_vision_kernel(UpdArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    __shared__ FrameSm F[kMaxFrames];
    __shared__ double vxi[kMaxFrames][6];
    __shared__ double red;
    if (tid < N) {
        const double *fs = a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride;
        const double *d = a.v_pose + ((size_t)w * a.Ncap + tid) * 15;
        make_frame(fs, wc, F[tid]);
        double R[9], om[3];
        quat_to_mat(fs, R);
        mat3_vec(R, d, om);
        const double p0 = fs[4] - wc.origin[0], p1 = fs[5] - wc.origin[1], p2 = fs[6] - wc.origin[2];
        vxi[tid][0] = om[0]; vxi[tid][1] = om[1]; vxi[tid][2] = om[2];
        vxi[tid][3] = -(p1 * om[2] - p2 * om[1]) - d[3];
        vxi[tid][4] = -(p2 * om[0] - p0 * om[2]) - d[4];
        vxi[tid][5] = -(p0 * om[1] - p1 * om[0]) - d[5];
    }
    if (tid == 0) red = 0.0;
    __syncthreads();
    const float W[4] = {(float)wc.sic[0], (float)wc.sic[1], (float)wc.sic[2], (float)wc.sic[3]};
    const float cb = (float)(wc.cauchy_a * wc.cauchy_a);
    const ObsRec *obs = a.obs + (size_t)w * a.Kcap;
    const LmRec *lms = a.lms + (size_t)w * a.Mcap;
    const double *rho = a.rho + (size_t)w * a.Mcap;
    const double *lm_scale = a.lm_scale + (size_t)w * a.Mcap;
    const LmAux *aux = a.lm_aux + (size_t)w * a.Mcap;
    double acc = 0.0;
    for (int ci = wv;; ci += 8) {
        const int ch = ci * gridDim.x + blockIdx.x;
        if (ch >= H.n_chunks) break;
        const int lm0 = H.chunk_begin[ch], cnt = H.chunk_meta[ch] & 0xff, anchor = H.chunk_meta[ch] >> 8;
        if (lane >= cnt) continue;
        const int l = lm0 + lane;
        const LmRec lr = lms[l];
        const int n_obs = lm_nobs(lr.meta);
        if (n_obs == 0) continue;
        double x[3];
        float xf[3], cl[3];
        world_point(F[anchor], lr.zrx, lr.zry, rho[l], x, xf, cl);
        // first pass: H_ll for the landmark's scaled direction, second pass: the products
        double hll = 0.0;
        for (unsigned j = 0, fm = lm_mask(lr.meta); j < (unsigned)n_obs; ++j, fm &= fm - 1) {
            const ObsRec o = obs[lr.obs_begin + j];
            const int of = __ffs(fm) - 1;
            ObsLin ol;
            linearize_obs<kLoss>(F[of], x, xf, cl, o.zx, o.zy, W, cb, ol);
            hll += (double)(ol.j0 * ol.j0 + ol.j1 * ol.j1);
        }
        const double sc = lm_scale[l];
        double d2 = sc * sc * hll;
        d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
        const double vl = sc * sc * aux[l].gl / d2;
        for (unsigned j = 0, fm = lm_mask(lr.meta); j < (unsigned)n_obs; ++j, fm &= fm - 1) {
            const ObsRec o = obs[lr.obs_begin + j];
            const int of = __ffs(fm) - 1;
            ObsLin ol;
            linearize_obs<kLoss>(F[of], x, xf, cl, o.zx, o.zy, W, cb, ol);
            double r0 = (double)ol.j0 * vl, r1 = (double)ol.j1 * vl;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const double dv = vxi[of][i] - vxi[anchor][i];
                r0 += (double)ol.Y0[i] * dv; r1 += (double)ol.Y1[i] * dv;
            }
            acc += r0 * r0 + r1 * r1;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0 && acc != 0.0) atomicAdd(&red, acc);
    __syncthreads();
    if (tid == 0 && red != 0.0) atomicAdd(a.acc + (size_t)w * kAcc + 9, red);
}

}  // namespace pvio
