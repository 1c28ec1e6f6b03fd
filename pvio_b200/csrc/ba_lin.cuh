// Reprojection sweep: per-factor residual + Jacobian, per-landmark Schur elimination and
// accumulation of the reduced pose system -- the hot kernel of the BA path.
//
// Replaces, per Gauss-Newton iteration, what ceres::Solve (bundle_adjustor.cpp:249) does
// with ReprojectionErrorCost::Evaluate (estimation/ceres/reprojection_error_cost.h:40-120)
// over every residual block, the CauchyLoss(1.0) corrector (bundle_adjustor.cpp:58) and the
// SPARSE_SCHUR elimination of the inverse-depth blocks (solver_options.h:27); with
// use_loss = 0 / victim_only = 1 it is the reprojection + landmark-Schur part of
// BundleAdjustor::marginalize_frame (bundle_adjustor.cpp:453-545).
//
// Formulation (DESIGN.md "xi coordinates").  For target frame t, anchor a, landmark l:
//     x_l  = R_wc(a) [z_ref;1]/rho + c_a                 world point (fp64)
//     y    = R_wc(t)^T (x_l - c_t)                        point in the target camera (fp64)
//     r    = W (y.xy/y.z - z_tgt)                         residual; numerator in fp64
//     G    = W dpi(y) R_wc(t)^T                           d r / d x_world  (2x3, fp32)
//     dr   = G X_l (xi_t - xi_a) + G c_l d rho,   X_l = [[x_l]x, I],  c_l = -(x_l - c_a)/rho
// where xi_f = T_f [dtheta_f; dp_f],  T_f = [[R_f, 0], [-[p_f]x R_f, -I]] is a per-frame
// change of variables applied once per iteration in the solve kernel.  This reproduces the
// reference Jacobians :94-113 exactly (J_delta = J_xi T) while making the target and
// anchor blocks of a residual +-the same 2x6 matrix Y = G X_l, so that every direct
// J^T J contribution is one symmetric 6x6 product per observation.
//
// Work decomposition: one CTA per (window, chunk range).  A chunk is <= 32 landmarks with
// the same anchor frame.  Phase A: a 16-lane group per landmark, lane f owns frame f
// (lane = frame keeps the (f,f) accumulators in registers).  Phase B: output-stationary
// threads own 6x3 tiles of the block-lower-triangular Schur sum  sum_l w_l h_l h_l^T  and
// stream the scaled h vectors of the chunk from shared memory.
#pragma once
#include "ba_math.cuh"
#include "ba_types.h"

namespace pvio {

constexpr int kGroup = 16;                  // lanes per landmark group (>= frames)
constexpr int kLinThreads = 256;
constexpr int kGroups = kLinThreads / kGroup;
constexpr int kChunk = 32;                  // landmarks per chunk
constexpr int kStageVals = 33;              // 21 (Y^T Y sym) + 6 (Y^T r) + 6 (w g_l h)

struct FrameSm {
    double Rwc[9];       // body->world times cam->body: camera-to-world rotation
    double c[3];         // camera centre, window-origin relative
    float R32[9];
    float pad[3];
};

struct FrameXf {         // per-frame change of variables xi = T delta and body pose
    double R[9];         // R_f (body to world)
    double p[3];         // p_f - origin
};

// Build the per-frame camera pose from the fp64 frame state.
__device__ __forceinline__ void make_frame(const double *fs, const WinConst &wc, FrameSm &o) {
    double R[9], Rcs[9];
    quat_to_mat(fs, R);
    quat_to_mat(wc.cam_q, Rcs);
    mat3_mul(R, Rcs, o.Rwc);
    double t[3];
    mat3_vec(R, wc.cam_p, t);
#pragma unroll
    for (int i = 0; i < 3; ++i) o.c[i] = fs[4 + i] - wc.origin[i] + t[i];
#pragma unroll
    for (int i = 0; i < 9; ++i) o.R32[i] = (float)o.Rwc[i];
}

// symmetric 6x6 index (i<=j) -> 0..20
__host__ __device__ constexpr int sym6(int i, int j) { return i * 6 - i * (i - 1) / 2 + (j - i); }

// pair index of block (f, g) with g <= f in the block-lower-triangular storage
__host__ __device__ __forceinline__ int pair_idx(int f, int g) { return f * (f + 1) / 2 + g; }

struct ObsLin {            // linearisation of one observation in xi coordinates (loss-corrected)
    float Y0[6], Y1[6];    // rows of Y = sqrt(rho') G X_l
    float r0, r1;          // sqrt(rho') r
    float j0, j1;          // sqrt(rho') G c_l   (d r / d rho)
    float cost;            // rho(s) / 2
};

// Evaluate one observation.  x: world point (fp64), xf: (float)x, cl: d x / d rho (fp32).
template <bool kLoss>
__device__ __forceinline__ void linearize_obs(const FrameSm &F, const double x[3], const float xf[3],
                                              const float cl[3], float zx, float zy,
                                              const float W[4], float cauchy_b, ObsLin &o) {
    const double d0 = x[0] - F.c[0], d1 = x[1] - F.c[1], d2 = x[2] - F.c[2];
    const double y0 = F.Rwc[0] * d0 + F.Rwc[3] * d1 + F.Rwc[6] * d2;
    const double y1 = F.Rwc[1] * d0 + F.Rwc[4] * d1 + F.Rwc[7] * d2;
    const double y2 = F.Rwc[2] * d0 + F.Rwc[5] * d1 + F.Rwc[8] * d2;
    const double nx = y0 - (double)zx * y2;          // fp64: the cancellation happens here
    const double ny = y1 - (double)zy * y2;
    const float iz = 1.0f / (float)y2;
    const float u0 = (float)nx * iz, u1 = (float)ny * iz;
    float r0 = W[0] * u0 + W[1] * u1;
    float r1 = W[2] * u0 + W[3] * u1;
    const float yx = (float)y0 * iz, yy = (float)y1 * iz;
    // A = W * dpi, dpi = [[iz, 0, -yx iz], [0, iz, -yy iz]]
    const float A00 = W[0] * iz, A01 = W[1] * iz, A02 = -(W[0] * yx + W[1] * yy) * iz;
    const float A10 = W[2] * iz, A11 = W[3] * iz, A12 = -(W[2] * yx + W[3] * yy) * iz;
    // G = A * Rwc^T  ->  G[i][k] = sum_m A[i][m] Rwc[k][m]
    float G0[3], G1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        G0[k] = A00 * F.R32[3 * k] + A01 * F.R32[3 * k + 1] + A02 * F.R32[3 * k + 2];
        G1[k] = A10 * F.R32[3 * k] + A11 * F.R32[3 * k + 1] + A12 * F.R32[3 * k + 2];
    }
    float cost;
    if (kLoss) {
        // ceres::CauchyLoss(a), b = a^2: rho = b log(1 + s/b), rho' = 1/(1 + s/b); rho'' < 0 so the
        // Corrector scales r and J by sqrt(rho') (corrector.cc)
        const float s = r0 * r0 + r1 * r1;
        const float t = 1.0f + s / cauchy_b;
        const float sc = rsqrtf(t);
        cost = 0.5f * cauchy_b * logf(t);
        r0 *= sc; r1 *= sc;
#pragma unroll
        for (int k = 0; k < 3; ++k) { G0[k] *= sc; G1[k] *= sc; }
    } else {
        cost = 0.5f * (r0 * r0 + r1 * r1);
    }
    // Y rows: [g x x_l, g]
    o.Y0[0] = G0[1] * xf[2] - G0[2] * xf[1];
    o.Y0[1] = G0[2] * xf[0] - G0[0] * xf[2];
    o.Y0[2] = G0[0] * xf[1] - G0[1] * xf[0];
    o.Y1[0] = G1[1] * xf[2] - G1[2] * xf[1];
    o.Y1[1] = G1[2] * xf[0] - G1[0] * xf[2];
    o.Y1[2] = G1[0] * xf[1] - G1[1] * xf[0];
#pragma unroll
    for (int k = 0; k < 3; ++k) { o.Y0[3 + k] = G0[k]; o.Y1[3 + k] = G1[k]; }
    o.j0 = G0[0] * cl[0] + G0[1] * cl[1] + G0[2] * cl[2];
    o.j1 = G1[0] * cl[0] + G1[1] * cl[1] + G1[2] * cl[2];
    o.r0 = r0; o.r1 = r1; o.cost = cost;
}

// Residual-only evaluation (candidate cost), fp64 numerator.
template <bool kLoss>
__device__ __forceinline__ float residual_cost(const FrameSm &F, const double x[3], float zx, float zy,
                                               const float W[4], float cauchy_b, float *r_out = nullptr) {
    const double d0 = x[0] - F.c[0], d1 = x[1] - F.c[1], d2 = x[2] - F.c[2];
    const double y0 = F.Rwc[0] * d0 + F.Rwc[3] * d1 + F.Rwc[6] * d2;
    const double y1 = F.Rwc[1] * d0 + F.Rwc[4] * d1 + F.Rwc[7] * d2;
    const double y2 = F.Rwc[2] * d0 + F.Rwc[5] * d1 + F.Rwc[8] * d2;
    const double nx = y0 - (double)zx * y2, ny = y1 - (double)zy * y2;
    const float iz = 1.0f / (float)y2;
    const float u0 = (float)nx * iz, u1 = (float)ny * iz;
    const float r0 = W[0] * u0 + W[1] * u1, r1 = W[2] * u0 + W[3] * u1;
    if (r_out) { r_out[0] = r0; r_out[1] = r1; }
    const float s = r0 * r0 + r1 * r1;
    return kLoss ? 0.5f * cauchy_b * logf(1.0f + s / cauchy_b) : 0.5f * s;
}

__device__ __forceinline__ void world_point(const FrameSm &Fa, float zrx, float zry, double rho,
                                            double x[3], float xf[3], float cl[3]) {
    const double ir = 1.0 / rho;
    const double zx = (double)zrx, zy = (double)zry;
    const double v0 = (Fa.Rwc[0] * zx + Fa.Rwc[1] * zy + Fa.Rwc[2]) * ir;
    const double v1 = (Fa.Rwc[3] * zx + Fa.Rwc[4] * zy + Fa.Rwc[5]) * ir;
    const double v2 = (Fa.Rwc[6] * zx + Fa.Rwc[7] * zy + Fa.Rwc[8]) * ir;
    x[0] = v0 + Fa.c[0]; x[1] = v1 + Fa.c[1]; x[2] = v2 + Fa.c[2];
    xf[0] = (float)x[0]; xf[1] = (float)x[1]; xf[2] = (float)x[2];
    cl[0] = (float)(-v0 * ir); cl[1] = (float)(-v1 * ir); cl[2] = (float)(-v2 * ir);
}

// Jacobi scaling / LM regularisation of a diagonal entry, as ceres' dogleg strategy applies
// it (dogleg_strategy.cc): reg = mu * clamp(scale^2 H_ii, 1e-6, 1e32) / scale^2.
__device__ __forceinline__ double lm_reg(double hii, double scale, double mu) {
    const double s2 = scale * scale;
    double d2 = s2 * hii;
    d2 = fmin(fmax(d2, 1.0e-6), 1.0e32);
    return mu * d2 / s2;
}

__host__ __device__ __forceinline__ int hs_rec(int N) { return 6 * N + 2; }

struct LinArgs {
    const WinHdr *hdr;
    const WinConst *cst;
    const ObsRec *obs;       // [W][Kcap]
    const LmRec *lms;        // [W][Mcap]
    const double *rho;       // [W][Mcap]
    const double *frames;    // [W][Ncap][16]
    const WinCtrl *ctrl;     // [W] (mu)
    double *lm_scale;        // [W][Mcap] Jacobi scale of each inverse depth (fixed at iteration 0)
    LmAux *lm_aux;           // [W][Mcap]
    float *hs_out;           // [W][hs_stride]: sqrt(w_l) h_l for the back-substitution of the update kernel, one
                             // record of 6 N floats per landmark SLOT (chunk * 32 + lane): the layout of the
                             // kernels' shared-memory h buffer, so a warp's records leave as one bulk copy.
                             // nullptr: not wanted (marginaliser)
    size_t hs_stride;        // floats per window
    // one record = hs_rec(N) floats: [0, 6N) sqrt(w) h per frame, [6N] sqrt(w) g_l, [6N + 1] the bit pattern of the
    // frame mask (observing frames | anchor; 0 for an empty slot or a non-finite pivot)
    double *Hred;            // [W][npairs_cap][36] block-lower-triangular reduced system (xi coords)
    double *Hdd;             // [W][Ncap][36] direct (pre-Schur) diagonal blocks (xi coords)
    double *gdir;            // [W][Ncap][6] direct gradient (xi coords)
    double *gred;            // [W][Ncap][6] reduced gradient
    double *cost_vis;        // [W]
    int Ncap, Mcap, Kcap;
    int compute_scale;       // 1: this is iteration 0, (re)compute lm_scale
    int victim_only;         // marginaliser: only landmarks flagged in_victim
    double mu_override;      // >= 0: use this mu instead of ctrl->mu
    int w0;                  // first window of this launch (sub-batch pipelining)
};

template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 2)
lin_schur_kernel(LinArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x;
    const int lane = tid & (kGroup - 1);
    const int grp = tid / kGroup;
    const int npairs = N * (N + 1) / 2;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    // carve shared memory
    FrameSm *F = reinterpret_cast<FrameSm *>(smem_raw);                       // [kMaxFrames]
    double *Hs = reinterpret_cast<double *>(F + kMaxFrames);                  // [npairs][36] direct terms
    double *gd = Hs + (kMaxFrames * (kMaxFrames + 1) / 2) * 36;               // [N][6] direct gradient
    double *gs = gd + kMaxFrames * 6;                                         // [N][6] Schur correction
    float *hbuf = reinterpret_cast<float *>(gs + kMaxFrames * 6);             // [kChunk][kMaxFrames][8]
    float *stage = hbuf + kChunk * kMaxFrames * 8;                            // [kGroups][kMaxFrames][33]
    int *msk = reinterpret_cast<int *>(stage + kGroups * kMaxFrames * kStageVals);  // [kChunk]
    double *cost_sm = reinterpret_cast<double *>(msk + kChunk);              // [8]

    // ---- prologue: per-frame camera poses, zero accumulators
    if (tid < N) make_frame(a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride, wc, F[tid]);
    for (int i = tid; i < npairs * 36; i += kLinThreads) Hs[i] = 0.0;
    for (int i = tid; i < kMaxFrames * 12; i += kLinThreads) gd[i] = 0.0;   // gd and gs are contiguous
    if (tid < 8) cost_sm[tid] = 0.0;
    __syncthreads();

    const float W[4] = {(float)wc.sic[0], (float)wc.sic[1], (float)wc.sic[2], (float)wc.sic[3]};
    const float cb = (float)(wc.cauchy_a * wc.cauchy_a);
    const double mu = a.mu_override >= 0.0 ? a.mu_override : a.ctrl[w].mu;

    const ObsRec *obs = a.obs + (size_t)w * a.Kcap;
    const LmRec *lms = a.lms + (size_t)w * a.Mcap;
    const double *rho = a.rho + (size_t)w * a.Mcap;
    double *lm_scale = a.lm_scale + (size_t)w * a.Mcap;
    LmAux *aux = a.lm_aux + (size_t)w * a.Mcap;

    // ---- Phase-B task of this thread: (pair, column half, k-split)
    const int ntask = 2 * npairs;
    const int ksplit = max(1, kLinThreads / ntask);
    const int task = tid % ntask;
    const int kk = tid / ntask;
    const bool b_active = (tid < ntask * ksplit);
    int bf = 0, bg = 0;
    {   // decode pair index -> (bf, bg), bg <= bf
        const int p = task >> 1;
        int f = 0;
        while ((f + 1) * (f + 2) / 2 <= p) ++f;
        bf = f; bg = p - f * (f + 1) / 2;
    }
    const int bhalf = task & 1;
    double acc64[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) acc64[i] = 0.0;

    float cost_acc = 0.f;

    for (int ch = blockIdx.x; ch < H.n_chunks; ch += gridDim.x) {
        const int lm0 = H.chunk_begin[ch];
        const int cnt = H.chunk_meta[ch] & 0xff;
        const int anchor = H.chunk_meta[ch] >> 8;

        // ================= Phase A: linearise, per-landmark Schur scalars =================
        float C[21], bd[6], bs[6];
#pragma unroll
        for (int i = 0; i < 21; ++i) C[i] = 0.f;
#pragma unroll
        for (int i = 0; i < 6; ++i) { bd[i] = 0.f; bs[i] = 0.f; }

        for (int s = grp; s < kChunk; s += kGroups) {
            const bool lm_ok = s < cnt;
            const int l = lm0 + (lm_ok ? s : 0);
            const LmRec lr = lms[l];
            int n_obs = lm_ok ? lm_nobs(lr.meta) : 0;
            if (a.victim_only && !lm_victim(lr.meta)) n_obs = 0;
            // lane = target frame: its record is number popc(frame mask below the lane)
            const unsigned fmask = n_obs > 0 ? lm_mask(lr.meta) : 0u;
            const bool observed = ((fmask >> lane) & 1u) && (lane < N);
            float zx = 0.f, zy = 0.f;
            if (observed) {
                const ObsRec o = obs[lr.obs_begin + __popc(fmask & ((1u << lane) - 1u))];
                zx = o.zx; zy = o.zy;
            }

            double x[3];
            float xf[3], cl[3];
            const double rl = lm_ok ? rho[l] : 1.0;
            world_point(F[anchor], lr.zrx, lr.zry, rl, x, xf, cl);

            ObsLin ol;
            float h[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) h[i] = 0.f;
            float hll = 0.f, gl = 0.f;
            if (observed) {
                linearize_obs<kLoss>(F[lane], x, xf, cl, zx, zy, W, cb, ol);
                hll = ol.j0 * ol.j0 + ol.j1 * ol.j1;
                gl = ol.j0 * ol.r0 + ol.j1 * ol.r1;
#pragma unroll
                for (int i = 0; i < 6; ++i) h[i] = ol.j0 * ol.Y0[i] + ol.j1 * ol.Y1[i];
                cost_acc += ol.cost;
            }
            // group reductions (16 lanes): H_ll, g_l in fp64; h_a = -sum h_t
            double hll_d = (double)hll, gl_d = (double)gl;
            float ha[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) ha[i] = h[i];
#pragma unroll
            for (int off = kGroup / 2; off > 0; off >>= 1) {
                hll_d += __shfl_xor_sync(0xffffffffu, hll_d, off, kGroup);
                gl_d += __shfl_xor_sync(0xffffffffu, gl_d, off, kGroup);
#pragma unroll
                for (int i = 0; i < 6; ++i) ha[i] += __shfl_xor_sync(0xffffffffu, ha[i], off, kGroup);
            }
            int tmask = (int)__ballot_sync(0xffffffffu, observed);
            tmask = (tmask >> (tid & 16)) & 0xffff;
            if (n_obs > 0) {
                // Jacobi scale (fixed at iteration 0) and LM regulariser of the 1x1 landmark block
                double sc;
                if (a.compute_scale) {
                    sc = 1.0 / (1.0 + sqrt(hll_d));
                    if (lane == 0) lm_scale[l] = sc;
                } else {
                    sc = lm_scale[l];
                }
                const double hreg = hll_d + (mu > 0.0 ? lm_reg(hll_d, sc, mu) : 0.0);
                const double wl = 1.0 / hreg;
                const bool finite = isfinite(wl);       // bundle_adjustor.cpp:538 skip
                const float wlf = finite ? (float)wl : 0.f;
                const float sw = sqrtf(wlf);
                const float wg = wlf * (float)gl_d;
                if (lane == 0) { aux[l].hll_reg = hreg; aux[l].gl = gl_d; aux[l].hll = hll_d; }
                float *hb = hbuf + (s * kMaxFrames + lane) * 8;
                if (a.hs_out && (observed || lane == anchor)) {
                    float *hs = a.hs_out + (size_t)w * a.hs_stride + (size_t)(ch * 32 + s) * hs_rec(N);
#pragma unroll
                    for (int i = 0; i < 6; ++i) hs[lane * 6 + i] = (observed ? h[i] : -ha[i]) * sw;
                    if (lane == anchor) { hs[6 * N] = sw * (float)gl_d; hs[6 * N + 1] = __int_as_float(finite ? (tmask | (1 << anchor)) : 0); }
                }
                if (observed) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) hb[i] = h[i] * sw;
                    // register accumulators of frame `lane` as a target
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = i; j < 6; ++j)
                            C[sym6(i, j)] += ol.Y0[i] * ol.Y0[j] + ol.Y1[i] * ol.Y1[j];
#pragma unroll
                    for (int i = 0; i < 6; ++i) {
                        bd[i] += ol.Y0[i] * ol.r0 + ol.Y1[i] * ol.r1;
                        bs[i] += wg * h[i];
                    }
                } else if (lane == anchor) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) hb[i] = -ha[i] * sw;
                }
                if (lane == 0) msk[s] = finite ? (tmask | (1 << anchor)) : 0;
            } else if (lane == 0) {
                msk[s] = 0;
                if (lm_ok && !a.victim_only) { aux[l].hll_reg = 1.0; aux[l].gl = 0.0; aux[l].hll = 0.0; }
            }
        }
        // stage the per-group frame accumulators
        if (lane < N) {
            float *st = stage + (grp * kMaxFrames + lane) * kStageVals;
#pragma unroll
            for (int i = 0; i < 21; ++i) st[i] = C[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) { st[21 + i] = bd[i]; st[27 + i] = bs[i]; }
        }
        __syncthreads();

        // ================= flush direct terms into the CTA accumulators (fp64) =================
        for (int idx = tid; idx < N * kStageVals; idx += kLinThreads) {
            const int f = idx / kStageVals, e = idx - f * kStageVals;
            if (f == anchor) continue;     // the anchor lane never observes (t > a)
            double v = 0.0;
#pragma unroll
            for (int g = 0; g < kGroups; ++g) v += (double)stage[(g * kMaxFrames + f) * kStageVals + e];
            if (v == 0.0) continue;
            if (e < 21) {
                int i = 0, rem = e;
                while (rem >= 6 - i) { rem -= 6 - i; ++i; }
                const int j = i + rem;
                double *dff = Hs + pair_idx(f, f) * 36;
                double *daa = Hs + pair_idx(anchor, anchor) * 36;
                dff[i * 6 + j] += v;
                atomicAdd(&daa[i * 6 + j], v);
                if (i != j) { dff[j * 6 + i] += v; atomicAdd(&daa[j * 6 + i], v); }
                if (f > anchor) {
                    double *dfa = Hs + pair_idx(f, anchor) * 36;
                    dfa[i * 6 + j] -= v;
                    if (i != j) dfa[j * 6 + i] -= v;
                } else {   // target earlier than anchor cannot happen in PVIO (anchor = lowest id) but stay general
                    double *daf = Hs + pair_idx(anchor, f) * 36;
                    daf[i * 6 + j] -= v;
                    if (i != j) daf[j * 6 + i] -= v;
                }
            } else if (e < 27) {
                gd[f * 6 + (e - 21)] += v;
                atomicAdd(&gd[anchor * 6 + (e - 21)], -v);
            } else {
                gs[f * 6 + (e - 27)] += v;
                atomicAdd(&gs[anchor * 6 + (e - 27)], -v);
            }
        }

        // ================= Phase B: Schur sum over the chunk, output stationary =================
        if (b_active) {
            float acc[18];
#pragma unroll
            for (int i = 0; i < 18; ++i) acc[i] = 0.f;
            for (int s = kk; s < cnt; s += ksplit) {
                const int m = msk[s];
                if (((m >> bf) & (m >> bg) & 1) == 0) continue;
                const float4 fa = *reinterpret_cast<const float4 *>(hbuf + (s * kMaxFrames + bf) * 8);
                const float2 fb = *reinterpret_cast<const float2 *>(hbuf + (s * kMaxFrames + bf) * 8 + 4);
                const float *gp = hbuf + (s * kMaxFrames + bg) * 8 + 3 * bhalf;
                const float g0 = gp[0], g1 = gp[1], g2 = gp[2];
                const float hf[6] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y};
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    acc[i * 3 + 0] += hf[i] * g0;
                    acc[i * 3 + 1] += hf[i] * g1;
                    acc[i * 3 + 2] += hf[i] * g2;
                }
            }
#pragma unroll
            for (int i = 0; i < 18; ++i) acc64[i] += (double)acc[i];
        }
        __syncthreads();
    }

    // ---- epilogue: Hred = direct - Schur; write out (plain stores when one CTA owns the window)
    const bool exclusive = (gridDim.x == 1);
    double *Hdd_o = a.Hdd + (size_t)w * a.Ncap * 36;
    for (int i = tid; i < N * 36; i += kLinThreads) {
        const int f = i / 36;
        const double v = Hs[pair_idx(f, f) * 36 + (i - f * 36)];
        if (exclusive) Hdd_o[i] = v; else if (v != 0.0) atomicAdd(&Hdd_o[i], v);
    }
    __syncthreads();
    for (int k = 0; k < ksplit; ++k) {
        if (b_active && kk == k) {
            double *d = Hs + (task >> 1) * 36 + 3 * bhalf;
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) d[i * 6 + j] -= acc64[i * 3 + j];
        }
        __syncthreads();
    }
    const int npairs_cap = a.Ncap * (a.Ncap + 1) / 2;
    double *Hred_o = a.Hred + (size_t)w * npairs_cap * 36;
    for (int i = tid; i < npairs * 36; i += kLinThreads) {
        const double v = Hs[i];
        if (exclusive) Hred_o[i] = v; else if (v != 0.0) atomicAdd(&Hred_o[i], v);
    }
    double *gdir_o = a.gdir + (size_t)w * a.Ncap * 6;
    double *gred_o = a.gred + (size_t)w * a.Ncap * 6;
    for (int i = tid; i < N * 6; i += kLinThreads) {
        const double vd = gd[i], vr = gd[i] - gs[i];
        if (exclusive) { gdir_o[i] = vd; gred_o[i] = vr; }
        else { if (vd != 0.0) atomicAdd(&gdir_o[i], vd); if (vr != 0.0) atomicAdd(&gred_o[i], vr); }
    }
    // cost: warp reduce then shared atomics then one global add
    double cd = (double)cost_acc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cd += __shfl_xor_sync(0xffffffffu, cd, off);
    if ((tid & 31) == 0) atomicAdd(&cost_sm[0], cd);
    __syncthreads();
    if (tid == 0) {
        if (exclusive) a.cost_vis[w] = cost_sm[0]; else atomicAdd(&a.cost_vis[w], cost_sm[0]);
    }
}

constexpr size_t lin_smem_bytes() {
    return sizeof(FrameSm) * kMaxFrames + sizeof(double) * ((kMaxFrames * (kMaxFrames + 1) / 2) * 36 + kMaxFrames * 12) +
           sizeof(float) * (kChunk * kMaxFrames * 8 + kGroups * kMaxFrames * kStageVals) + sizeof(int) * kChunk +
           sizeof(double) * 8;
}

}  // namespace pvio
