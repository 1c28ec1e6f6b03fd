// Back-substitution of the inverse depths, Plus() update and candidate-cost sweep.
//
// After the reduced system has been solved for the pose step, this kernel does what ceres
// does after the Schur solve of an iteration (bundle_adjustor.cpp:249, inside ceres::Solve):
//   d rho_l = -(g_l + h_l . dxi) / (H_ll + reg)                 back-substitution
//   candidate = Plus(x, dx)         (quaternion_parameterization.h:28-31, plain + elsewhere)
//   candidate cost = sum rho(|r|^2)/2 over the reprojection blocks   (step acceptance)
// and accumulates the scalars the trust-region logic needs.  h_l is recomputed from the
// observation table (16 B/observation re-read) instead of being stored (6 floats per
// touched frame per landmark written + read).
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
    const double *dx_pose;    // [W][Ncap][15] pose/motion step in delta coordinates
    double *rho_cand;         // [W][Mcap]
    double *frames_cand;      // [W][Ncap][16]
    double *dx_lm;            // [W][Mcap]
    double *acc;              // [W][8]: cand_cost_vis, g.dx(lm), dx.reg.dx(lm), gn_norm2(lm), dxnorm2(lm), xnorm2(lm)
    int Ncap, Mcap, Kcap;
    double mu_override;
    int w0;
    double beta;              // step scale (1 = full Gauss-Newton step; < 1 when the trust region truncates it)
};

template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 2)
update_cost_kernel(UpdArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x;
    const int lane = tid & (kGroup - 1);
    const int grp = tid / kGroup;

    __shared__ FrameSm F[kMaxFrames];      // current state
    __shared__ FrameSm Fc[kMaxFrames];     // candidate state
    __shared__ double dxi[kMaxFrames][6];  // xi = T delta per frame
    __shared__ double red[8];

    if (tid < N) {
        const double *fs = a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride;
        const double *d = a.dx_pose + ((size_t)w * a.Ncap + tid) * 15;
        make_frame(fs, wc, F[tid]);
        // candidate frame
        double fc[kFrameStride];
        const double db[3] = {a.beta * d[0], a.beta * d[1], a.beta * d[2]};
        quat_plus(fs, db, fc);
#pragma unroll
        for (int i = 0; i < 12; ++i) fc[4 + i] = fs[4 + i] + a.beta * d[3 + i];
        make_frame(fc, wc, Fc[tid]);
        if (blockIdx.x == 0) {
            double *o = a.frames_cand + ((size_t)w * a.Ncap + tid) * kFrameStride;
            double amb = 0.0;   // |x - x_cand|^2 over the ambient coordinates of the free blocks
#pragma unroll
            for (int i = 0; i < kFrameStride; ++i) {
                o[i] = fc[i];
                const bool pose = i < 7;
                const bool live = pose ? !((H.fixed_mask >> tid) & 1) : (H.use_inertial != 0);
                if (live) amb += (fc[i] - fs[i]) * (fc[i] - fs[i]);
            }
            atomicAdd(&a.acc[(size_t)w * 8 + 6], amb);
        }
        // xi = [R dtheta; -[p]x R dtheta - dp]
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

    double s_cost = 0.0, s_gdx = 0.0, s_reg = 0.0, s_gn = 0.0, s_dx2 = 0.0, s_x2 = 0.0;

    for (int ch = blockIdx.x; ch < H.n_chunks; ch += gridDim.x) {
        const int lm0 = H.chunk_begin[ch];
        const int cnt = H.chunk_meta[ch] & 0xff;
        const int anchor = H.chunk_meta[ch] >> 8;
        for (int s = grp; s < kChunk; s += kGroups) {
            const bool lm_ok = s < cnt;
            const int l = lm0 + (lm_ok ? s : 0);
            const LmRec lr = lms[l];
            const int n_obs = lm_ok ? ((lr.meta >> 8) & 0xff) : 0;
            const int n_max = max(n_obs, __shfl_xor_sync(0xffffffffu, n_obs, 16));
            ObsRec o;
            o.frame = -1; o.zx = 0.f; o.zy = 0.f;
            if (lane < n_obs) o = obs[lr.obs_begin + lane];
            int src = -1;
            for (int j = 0; j < n_max; ++j) {
                const int fj = __shfl_sync(0xffffffffu, o.frame, j, kGroup);
                if (fj == lane) src = j;
            }
            const float zx = __shfl_sync(0xffffffffu, o.zx, max(src, 0), kGroup);
            const float zy = __shfl_sync(0xffffffffu, o.zy, max(src, 0), kGroup);
            const bool observed = (src >= 0) && (lane < N);
            const double rl = lm_ok ? rho[l] : 1.0;
            double x[3];
            float xf[3], cl[3];
            world_point(F[anchor], lr.zrx, lr.zry, rl, x, xf, cl);
            double hdx = 0.0, hll = 0.0;
            if (observed) {
                ObsLin ol;
                linearize_obs<kLoss>(F[lane], x, xf, cl, zx, zy, W, cb, ol);
                hll = (double)(ol.j0 * ol.j0 + ol.j1 * ol.j1);
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    hdx += (double)(ol.j0 * ol.Y0[i] + ol.j1 * ol.Y1[i]) * (dxi[lane][i] - dxi[anchor][i]);
            }
#pragma unroll
            for (int off = kGroup / 2; off > 0; off >>= 1) {
                hdx += __shfl_xor_sync(0xffffffffu, hdx, off, kGroup);
                hll += __shfl_xor_sync(0xffffffffu, hll, off, kGroup);
            }
            double drho = 0.0;
            if (n_obs > 0) {
                const LmAux ax = aux[l];
                const double wl = 1.0 / ax.hll_reg;
                drho = isfinite(wl) ? -(ax.gl + hdx) * wl : 0.0;   // full GN back-substitution
                if (lane == 0) {
                    const double sc = lm_scale[l];
                    const double reg = mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0;
                    double d2 = sc * sc * hll;
                    d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
                    s_gdx += hdx * ax.gl * wl + ax.gl * drho;
                    s_reg += reg * drho * drho;
                    s_gn += d2 * (drho / sc) * (drho / sc);
                    s_dx2 += a.beta * a.beta * drho * drho;
                    s_x2 += rl * rl;
                }
            }
            drho *= a.beta;
            if (lm_ok && lane == 0) { rho_c[l] = rl + drho; dxl[l] = drho; }
            // candidate cost of this landmark's observations
            if (observed) {
                double xc[3];
                float xcf[3], clc[3];
                world_point(Fc[anchor], lr.zrx, lr.zry, rl + drho, xc, xcf, clc);
                s_cost += (double)residual_cost<kLoss>(Fc[lane], xc, zx, zy, W, cb);
            }
        }
    }
    // block reduction of the six scalars
    double v[6] = {s_cost, s_gdx, s_reg, s_gn, s_dx2, s_x2};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if ((tid & 31) == 0 && v[k] != 0.0) atomicAdd(&red[k], v[k]);
    }
    __syncthreads();
    if (tid < 6) {
        double *o = a.acc + (size_t)w * 8 + tid;
        if (gridDim.x == 1) *o = red[tid]; else atomicAdd(o, red[tid]);
    }
}

// Second generation: one LANE per landmark (see ba_lin2.cuh).  Same outputs.
template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 2)
update_tpl_kernel(UpdArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;

    __shared__ FrameSm F[kMaxFrames];      // current state
    __shared__ FrameSm Fc[kMaxFrames];     // candidate state
    __shared__ double dxi[kMaxFrames][6];  // xi = T delta per frame
    __shared__ double red[8];

    if (tid < N) {
        const double *fs = a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride;
        const double *d = a.dx_pose + ((size_t)w * a.Ncap + tid) * 15;
        make_frame(fs, wc, F[tid]);
        double fc[kFrameStride];
        const double db[3] = {a.beta * d[0], a.beta * d[1], a.beta * d[2]};
        quat_plus(fs, db, fc);
#pragma unroll
        for (int i = 0; i < 12; ++i) fc[4 + i] = fs[4 + i] + a.beta * d[3 + i];
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
            atomicAdd(&a.acc[(size_t)w * 8 + 6], amb);
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

    double s_cost = 0.0, s_gdx = 0.0, s_reg = 0.0, s_gn = 0.0, s_dx2 = 0.0, s_x2 = 0.0;
    for (int c0 = blockIdx.x * 8; c0 < H.n_chunks; c0 += gridDim.x * 8) {
        const int ch = c0 + wv;
        if (ch >= H.n_chunks) continue;
        const int lm0 = H.chunk_begin[ch];
        const int cnt = H.chunk_meta[ch] & 0xff;
        const int anchor = H.chunk_meta[ch] >> 8;
        if (lane >= cnt) continue;
        const int l = lm0 + lane;
        const LmRec lr = lms[l];
        const int n_obs = (lr.meta >> 8) & 0xff;
        const double rl = rho[l];
        double x[3];
        float xf[3], cl[3];
        world_point(F[anchor], lr.zrx, lr.zry, rl, x, xf, cl);
        double hdx = 0.0, hll = 0.0;
        for (int j = 0; j < n_obs; ++j) {
            const ObsRec o = obs[lr.obs_begin + j];
            ObsLin ol;
            linearize_obs<kLoss>(F[o.frame], x, xf, cl, o.zx, o.zy, W, cb, ol);
            hll += (double)(ol.j0 * ol.j0 + ol.j1 * ol.j1);
#pragma unroll
            for (int i = 0; i < 6; ++i)
                hdx += (double)(ol.j0 * ol.Y0[i] + ol.j1 * ol.Y1[i]) * (dxi[o.frame][i] - dxi[anchor][i]);
        }
        double drho = 0.0;
        if (n_obs > 0) {
            const LmAux ax = aux[l];
            const double wl = 1.0 / ax.hll_reg;
            drho = isfinite(wl) ? -(ax.gl + hdx) * wl : 0.0;
            const double sc = lm_scale[l];
            const double reg = mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0;
            double d2 = sc * sc * hll;
            d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
            s_gdx += hdx * ax.gl * wl + ax.gl * drho;
            s_reg += reg * drho * drho;
            s_gn += d2 * (drho / sc) * (drho / sc);
            s_dx2 += a.beta * a.beta * drho * drho;
            s_x2 += rl * rl;
        }
        drho *= a.beta;
        rho_c[l] = rl + drho;
        dxl[l] = drho;
        if (n_obs > 0) {
            double xc[3];
            float xcf[3], clc[3];
            world_point(Fc[anchor], lr.zrx, lr.zry, rl + drho, xc, xcf, clc);
            for (int j = 0; j < n_obs; ++j) {
                const ObsRec o = obs[lr.obs_begin + j];
                s_cost += (double)residual_cost<kLoss>(Fc[o.frame], xc, o.zx, o.zy, W, cb);
            }
        }
    }
    double v[6] = {s_cost, s_gdx, s_reg, s_gn, s_dx2, s_x2};
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0 && v[k] != 0.0) atomicAdd(&red[k], v[k]);
    }
    __syncthreads();
    if (tid < 6) {
        double *o = a.acc + (size_t)w * 8 + tid;
        if (gridDim.x == 1) *o = red[tid]; else atomicAdd(o, red[tid]);
    }
}

}  // namespace pvio
