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
#include "ba_linearize.cuh"
#include "ba_tr.cuh"

namespace pvio {

struct UpdArgs {
    const WinHdr *hdr;
    const WinConst *cst;
    const ObsRec *obs;
    const LmRec *lms;
    const double *rho;        // [W][Mcap] current
    const double *frames;     // [W][Ncap][16] current
    WinCtrl *ctrl;
    const double *lm_scale;
    const LmAux *lm_aux;
    const void *hs;           // [W][Ncap][Mcap][6] real: unscaled h records of lin_obs_kernel (PipeArgs::hs)
    const FObs *fobs;         // frame-major table (ba_fobs.cuh)
    const int32_t *seg;
    const double *dx_pose;    // [W][Ncap][15] pose/motion step in delta coordinates
    double *rho_cand;         // [W][Mcap]
    double *frames_cand;      // [W][Ncap][16]
    double *dx_lm;            // [W][Mcap]
    double *lm_v;             // [W][Mcap] scaled steepest-descent direction of each inverse depth (kMode 1 -> kMode 2)
    double *acc;              // [W][kAcc]: 0 cand_cost_vis, 1 g.dx(lm), 2 dx.reg.dx(lm), 3 gn_norm2(lm), 4 |step|^2(lm),
                              //   5 |x|^2(lm), 6 ambient |x - x_cand|^2 (frames), 7 |D^-1 S g|^2 (lm), 8 v.reg.dx (lm), 9-10 |J v|^2
    int Ncap, Mcap, Kcap;
    double mu_override;
    int w0;
    // step = step_b * dx_gn - step_a * v, v = S^2 g / clamp(S^2 diag H) the scaled steepest-descent direction:
    // (0, 1) Gauss-Newton, (0, beta) truncated GN, (a, b) dogleg interpolation, (a, 0) Cauchy leg
    double step_a, step_b;
    const double *v_pose;     // [W][Ncap][15] from solve_kernel
    int loop;                 // jv_vision_kernel: 1 = only windows whose GN step left the trust region (WinCtrl::need_jv)
    LinBufs bufs;             // loop: hs / lm_aux of the window's current buffer set (WinCtrl::buf)
};

template <typename real>
__host__ __device__ inline size_t upd_smem_layout(int N, int Mp, int warps, size_t *o_dxi, size_t *o_x, size_t *o_seg, size_t *o_red, size_t *o_ring) {
    size_t off = sizeof(double) * 12 * (size_t)N;                  // candidate camera poses: Rwc[9], c[3]
    *o_dxi = off; off += sizeof(double) * 6 * (size_t)N;           // xi = T delta per frame
    *o_x = off; off += sizeof(double) * 3 * (size_t)Mp;            // candidate world points (SoA)
    *o_red = off; off += sizeof(double) * 8;
    *o_seg = off; off += sizeof(int32_t) * kSegTab;
    off = (off + 15) & ~(size_t)15;
    *o_ring = off; off += sizeof(FObs) * kRing * 32 * (size_t)warps;
    return off;
}
template <typename real>
__host__ __device__ inline size_t upd_smem_bytes(int N, int Mp, int warps) {
    size_t a, b, c, d, e;
    return upd_smem_layout<real>(N, Mp, warps, &a, &b, &c, &d, &e);
}

// Residual-only evaluation of one block (candidate cost): fp64 numerators, the rest in `real`.
template <bool kLoss, typename real>
__device__ __forceinline__ real residual_cost_blk(const double (&Rwc)[9], const double (&c)[3], double x0, double x1, double x2,
                                                  float zx, float zy, const real (&W)[4], bool diag_w, real inv_cauchy_b) {
    const double d0 = x0 - c[0], d1 = x1 - c[1], d2 = x2 - c[2];
    const double y0 = Rwc[0] * d0 + Rwc[3] * d1 + Rwc[6] * d2;
    const double y1 = Rwc[1] * d0 + Rwc[4] * d1 + Rwc[7] * d2;
    const double y2 = Rwc[2] * d0 + Rwc[5] * d1 + Rwc[8] * d2;
    const double nx = y0 - (double)zx * y2, ny = y1 - (double)zy * y2;
    const real iz = rcp_r((real)y2);
    const real u0 = (real)nx * iz, u1 = (real)ny * iz;
    real r0 = W[0] * u0, r1 = W[3] * u1;
    if (!diag_w) { r0 += W[1] * u1; r1 += W[2] * u0; }
    const real s = r0 * r0 + r1 * r1;
    return kLoss ? (real)1 + s * inv_cauchy_b : (real)0.5 * s;      // with the loss: t (the caller accumulates b/2 log t, LogAcc)
}

// Back-substitution (thread per landmark, coalesced reads of the frame-major h records), Plus, and the
// candidate cost over the frame-major table (thread per residual block, target pose in registers).
//   kMode 0  everything in one launch with the step (step_a, step_b) given by the caller: the plain Gauss-Newton
//            iteration of pvio_b200_ba_gn_step / batch_gn_step.  gridDim.x CTAs may share a window (landmark l
//            belongs to CTA (l / 32) mod gridDim.x, the rows are cut as in lin_obs_kernel).
//   kMode 1  back-substitution only (gridDim.x = 1): GN step of the inverse depths and its scalars, then the
//            trust-region test of the step (tr_after_backsub): first half of an iteration of the device-side loop.
//   kMode 2  candidate only: Plus(x, step_b dx_gn - step_a v) with the step chosen by tr_after_jv (its cost comes from
//            the speculative linearise sweep that follows, ba_tr.cuh).
template <bool kLoss, typename real, int kWarps, int kMinBlocks, int kMode>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
update_obs_kernel(UpdArgs a) {
    constexpr int kThreads = kWarps * 32;
    const int w = blockIdx.y + a.w0;
    WinCtrl &ctrl = a.ctrl[w];
    if (kMode == 1 && ctrl.done) return;
    if (kMode == 2 && (ctrl.done || ctrl.skip)) return;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N, M = H.M, Mp = (M + 31) & ~31;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int nsp = N * (N - 1) / 2;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    size_t o_dxi, o_x, o_seg, o_red, o_ring;
    upd_smem_layout<real>(N, Mp, kWarps, &o_dxi, &o_x, &o_seg, &o_red, &o_ring);
    double *Fc = reinterpret_cast<double *>(smem_raw);                  // [N][12] candidate state
    double *dxi = reinterpret_cast<double *>(smem_raw + o_dxi);         // [N][6]
    double *xs = reinterpret_cast<double *>(smem_raw + o_x);            // [3][Mp]
    double *red = reinterpret_cast<double *>(smem_raw + o_red);         // [8]
    int32_t *sg = reinterpret_cast<int32_t *>(smem_raw + o_seg);

    const double step_a = kMode == 2 ? ctrl.step_a : a.step_a;
    const double step_b = kMode == 2 ? ctrl.step_b : a.step_b;
    if (tid < N) {
        const double *fs = a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride;
        const double *d = a.dx_pose + ((size_t)w * a.Ncap + tid) * 15;
        if (kMode != 1) {
            double fc[kFrameStride];
            const double *vp = a.v_pose + ((size_t)w * a.Ncap + tid) * 15;
            double de[15];
#pragma unroll
            for (int i = 0; i < 15; ++i) de[i] = step_b * d[i] - (step_a != 0.0 ? step_a * vp[i] : 0.0);
            quat_plus(fs, de, fc);
#pragma unroll
            for (int i = 0; i < 12; ++i) fc[4 + i] = fs[4 + i] + de[3 + i];
            FrameSm f;
            make_frame(fc, wc, f);
#pragma unroll
            for (int i = 0; i < 9; ++i) Fc[tid * 12 + i] = f.Rwc[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) Fc[tid * 12 + 9 + i] = f.c[i];
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
        }
        if (kMode != 2) {
            double R[9], om[3];
            quat_to_mat(fs, R);
            mat3_vec(R, d, om);
            const double p0 = fs[4] - wc.origin[0], p1 = fs[5] - wc.origin[1], p2 = fs[6] - wc.origin[2];
            dxi[tid * 6 + 0] = om[0]; dxi[tid * 6 + 1] = om[1]; dxi[tid * 6 + 2] = om[2];
            dxi[tid * 6 + 3] = -(p1 * om[2] - p2 * om[1]) - d[3];
            dxi[tid * 6 + 4] = -(p2 * om[0] - p0 * om[2]) - d[4];
            dxi[tid * 6 + 5] = -(p0 * om[1] - p1 * om[0]) - d[5];
        }
    }
    if (tid < 8) red[tid] = 0.0;
    if (kMode == 0) for (int i = tid; i < kSegTab; i += kThreads) sg[i] = a.seg[(size_t)w * kSegTab + i];
    __syncthreads();

    const double mu = a.mu_override >= 0.0 ? a.mu_override : ctrl.mu;
    const LmRec *lms = a.lms + (size_t)w * a.Mcap;
    const double *rho = a.rho + (size_t)w * a.Mcap;
    const double *lm_scale = a.lm_scale + (size_t)w * a.Mcap;
    const size_t bsel = a.loop ? (size_t)ctrl.buf : 0;
    const LmAux *aux = a.lm_aux + bsel * a.bufs.lm_aux + (size_t)w * a.Mcap;
    const real *hs = reinterpret_cast<const real *>(reinterpret_cast<const char *>(a.hs) + bsel * a.bufs.hs) + (size_t)w * a.Ncap * a.Mcap * 6;
    double *rho_c = a.rho_cand + (size_t)w * a.Mcap;
    double *dxl = a.dx_lm + (size_t)w * a.Mcap;
    double *lmv = a.lm_v + (size_t)w * a.Mcap;
    const unsigned fixed = (unsigned)H.fixed_mask & ((1u << N) - 1u);

    // ---- phase 1: back-substitution, candidate inverse depth and world point
    double s_gdx = 0.0, s_reg = 0.0, s_gn = 0.0, s_dx2 = 0.0, s_x2 = 0.0, s_g2 = 0.0, s_vrd = 0.0;
    for (int l = tid; l < Mp; l += kThreads) {
        double x0 = 0.0, x1 = 0.0, x2 = 1.0;
        if (l < M) {
            const LmRec lr = lms[l];
            const int n_obs = lm_nobs(lr.meta), anchor = lm_anchor(lr.meta);
            const double rl = rho[l];
            const bool mine = ((l >> 5) % (int)gridDim.x) == (int)blockIdx.x;
            double drho = 0.0, vl = 0.0;
            if (n_obs > 0 && kMode != 2) {
                const LmAux ax = aux[l];
                // h_l . dxi = sum_t h_lt . (dxi_t - dxi_a): the anchor block of a residual is minus its target block
                double hdx = 0.0;
                const bool a_fixed = (fixed >> anchor) & 1u;
                typedef typename Vec2<real>::type real2;
                unsigned fm = lm_mask(lr.meta);
                if (a_fixed) fm &= ~fixed;                              // record not written; both steps are zero
                const double *da = dxi + anchor * 6;
                // four records in flight at a time (the loads of a group are independent of each other)
                while (fm) {
                    int f4[4];
                    real2 h[4][3];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        f4[q] = fm ? __ffs(fm) - 1 : -1;
                        fm &= fm - 1;
                        if (f4[q] >= 0) {
                            const real2 *hp = reinterpret_cast<const real2 *>(hs + ((size_t)f4[q] * a.Mcap + l) * 6);
                            h[q][0] = hp[0]; h[q][1] = hp[1]; h[q][2] = hp[2];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (f4[q] < 0) continue;
                        const double *df = dxi + f4[q] * 6;
                        hdx += (double)h[q][0].x * (df[0] - da[0]) + (double)h[q][0].y * (df[1] - da[1]) + (double)h[q][1].x * (df[2] - da[2]) +
                               (double)h[q][1].y * (df[3] - da[3]) + (double)h[q][2].x * (df[4] - da[4]) + (double)h[q][2].y * (df[5] - da[5]);
                    }
                }
                const double wl = 1.0 / ax.hll_reg;
                const double hll = ax.hll;
                drho = isfinite(wl) ? -(ax.gl + hdx) * wl : 0.0;
                const double sc = lm_scale[l];
                const double reg = mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0;
                double d2 = sc * sc * hll;
                d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
                vl = sc * sc * ax.gl / d2;                       // scaled steepest-descent direction
                if (mine) {
                    s_gdx += hdx * ax.gl * wl + ax.gl * drho;
                    s_reg += reg * drho * drho;
                    s_gn += d2 * (drho / sc) * (drho / sc);
                    s_g2 += vl * ax.gl;
                    s_vrd += vl * reg * drho;
                    s_x2 += rl * rl;
                }
            }
            if (kMode == 1) { dxl[l] = drho; lmv[l] = vl; continue; }
            if (kMode == 2 && n_obs > 0) { drho = dxl[l]; vl = lmv[l]; }
            drho = step_b * drho - step_a * vl;
            if (mine) { s_dx2 += drho * drho; rho_c[l] = rl + drho; dxl[l] = drho; }
            if (kMode != 2 && n_obs > 0) {
                const double *Fa = Fc + anchor * 12;
                const double ir = 1.0 / (rl + drho);
                const double zx = (double)lr.zrx, zy = (double)lr.zry;
                x0 = (Fa[0] * zx + Fa[1] * zy + Fa[2]) * ir + Fa[9];
                x1 = (Fa[3] * zx + Fa[4] * zy + Fa[5]) * ir + Fa[10];
                x2 = (Fa[6] * zx + Fa[7] * zy + Fa[8]) * ir + Fa[11];
            }
        }
        if (kMode == 0) { xs[l] = x0; xs[Mp + l] = x1; xs[2 * Mp + l] = x2; }
    }
    __syncthreads();

    // ---- phase 2: candidate cost over the rows of the frame-major table
    double s_cost = 0.0;
    if (kMode == 0) {
        const real W[4] = {(real)wc.sic[0], (real)wc.sic[1], (real)wc.sic[2], (real)wc.sic[3]};
        const real cb = (real)(wc.cauchy_a * wc.cauchy_a), inv_cb = (real)(1.0 / (wc.cauchy_a * wc.cauchy_a));
        const bool diag_w = wc.sic[1] == 0.0 && wc.sic[2] == 0.0;
        const FObs *fobs = a.fobs + (size_t)w * a.Kcap;
        const int32_t *sbeg = sg, *srow = sg + kMaxSeg + 1;
        const int rows = srow[nsp];
        const int nwt = gridDim.x * kWarps, wid = blockIdx.x * kWarps + wv;
        const int r_begin = (int)((long long)rows * wid / nwt), r_end = (int)((long long)rows * (wid + 1) / nwt);
        // rows through the per-warp cp.async ring, as in lin_obs_kernel
        FObs *ring = reinterpret_cast<FObs *>(smem_raw + o_ring) + (size_t)wv * kRing * 32;
        RowCursor pf(sbeg, srow, r_begin, rows), cs(sbeg, srow, r_begin, rows);
#pragma unroll
        for (int d = 0; d < kRing; ++d) row_prefetch(pf, r_end, fobs, ring + d * 32, lane);
        int slot = 0;
        while (cs.r < r_end) {
            const int sp = cs.sp, t = cs.t;
            const int seg_end = sbeg[sp + 1];
            const int r_stop = min(r_end, srow[sp + 1]);
            double Rwc[9], c[3];
            {
                const double *Ft = Fc + t * 12;
#pragma unroll
                for (int i = 0; i < 9; ++i) Rwc[i] = Ft[i];
#pragma unroll
                for (int i = 0; i < 3; ++i) c[i] = Ft[9 + i];
            }
            real cacc = 0;
            LogAcc<real> la;
            int nfold = 0;
            for (; cs.r < r_stop; cs.advance()) {
                const int k = cs.k0() + lane;
                asm volatile("cp.async.wait_group %0;" :: "n"(kRing - 1) : "memory");
                const FObs o = ring[slot * 32 + lane];
                row_prefetch(pf, r_end, fobs, ring + slot * 32, lane);
                slot = slot + 1 == kRing ? 0 : slot + 1;
                if (k < seg_end) {
                    const int l = o.lm;
                    const real v = residual_cost_blk<kLoss, real>(Rwc, c, xs[l], xs[Mp + l], xs[2 * Mp + l], o.zx, o.zy, W, diag_w, inv_cb);
                    if (kLoss) la.add(v); else cacc += v;
                }
                if (kLoss && (++nfold & 3) == 0) la.fold();
            }
            if (kLoss) cacc = (real)0.5 * cb * la.total();
            s_cost += (double)cacc;
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    }
    double v[8] = {s_cost, s_gdx, s_reg, s_gn, s_dx2, s_x2, s_g2, s_vrd};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
        if (lane == 0 && v[k] != 0.0) atomicAdd(&red[k], v[k]);
    }
    __syncthreads();
    double *acc = a.acc + (size_t)w * kAcc;
    if (kMode == 1) {
        // one CTA owns the window: plain stores; also clears the slots the later sweeps of this iteration add to
        if (tid < kAcc) {
            double o = 0.0;
            if (tid >= 1 && tid <= 3) o = red[tid];
            else if (tid == 5) o = red[5];
            else if (tid == 7) o = red[6];
            else if (tid == 8) o = red[7];
            acc[tid] = o;
        }
        __syncthreads();
        if (tid == 0) { WinCtrl c = ctrl; tr_after_backsub(c, acc); ctrl = c; }
    } else if (tid < 8) {
        const int slot = tid < 6 ? tid : tid + 1;            // slot 6 is the frames' ambient step norm
        if (kMode == 2 && tid != 4) return;                  // the GN scalars were written by the back-substitution
        if (red[tid] != 0.0) atomicAdd(acc + slot, red[tid]);
    }
}

// |J v|^2 over the reprojection blocks (loss-corrected Jacobians), for the Cauchy point of the dogleg
// step: alpha = |grad|^2 / |J S D^-1 grad|^2 (ceres dogleg_strategy.cc ComputeCauchyPoint).
template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 2)
jv_vision_kernel(UpdArgs a) {
    const int w = blockIdx.y + a.w0;
    if (a.loop && (a.ctrl[w].done || a.ctrl[w].skip || !a.ctrl[w].need_jv)) return;
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
    const LmAux *aux = a.lm_aux + (a.loop ? (size_t)a.ctrl[w].buf * a.bufs.lm_aux : 0) + (size_t)w * a.Mcap;
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
