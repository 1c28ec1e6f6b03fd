// libpvio_b200: C-ABI entry points (include/pvio_b200.h), host-side packing and launch
// orchestration of the bundle-adjustment kernels.  No torch types, no CPU fallback.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include "api_internal.h"
#include "ba_lin.cuh"
#include "ba_lin2.cuh"
#include "ba_lin3.cuh"
#include "ba_schur.cuh"
#include "ba_schur_tc.cuh"
#include "ba_solve.cuh"
#include "ba_update.cuh"

namespace pvio {

int fail(Handle *h, int code, const char *what, cudaError_t e) {
    if (h) {
        h->err = what ? what : "";
        if (e != cudaSuccess) { h->err += ": "; h->err += cudaGetErrorString(e); }
    }
    return code;
}

template <typename T>
static int alloc(Handle *h, DevBuf<T> &b, size_t n, bool pinned) {
    b.n = n;
    if (n == 0) return 0;
    CK(h, cudaMalloc(&b.d, n * sizeof(T)));
    CK(h, cudaMemset(b.d, 0, n * sizeof(T)));
    if (pinned) {
        CK(h, cudaMallocHost(&b.h, n * sizeof(T)));
        memset(b.h, 0, n * sizeof(T));
    }
    return 0;
}

template <typename T>
static void release(DevBuf<T> &b) {
    if (b.d) cudaFree(b.d);
    if (b.h) cudaFreeHost(b.h);
    b.d = nullptr; b.h = nullptr; b.n = 0;
}

#define TRY(x) do { int rc__ = (x); if (rc__ != 0) return rc__; } while (0)

static int ensure_inertial(Handle *h) {
    if (h->have_inertial) return 0;
    const size_t W = h->W, N = h->Ncap, dcap = 15 * N;
    TRY(alloc(h, h->imu_idx, W * N * 2, true));
    TRY(alloc(h, h->imu_data, W * N * kImuStride, true));
    TRY(alloc(h, h->prior_frames, W * N, true));
    TRY(alloc(h, h->prior_S, W * dcap * dcap, true));
    TRY(alloc(h, h->prior_L, W * dcap * dcap, true));
    TRY(alloc(h, h->prior_e, W * dcap, true));
    TRY(alloc(h, h->prior_x0, W * N * kFrameStride, true));
    h->have_inertial = true;
    return 0;
}

static int ensure_planes(Handle *h, int T, int O) {
    if (h->have_planes && T <= h->Tcap && O <= h->Ocap) return 0;
    release(h->plane_param); release(h->pt_plane); release(h->pt_begin); release(h->pt_frame); release(h->pt_z);
    h->Tcap = std::max(T, 256);
    h->Ocap = std::max(O, 256 * 8);
    const size_t W = h->W;
    TRY(alloc(h, h->plane_param, W * h->Pcap * 4, true));
    TRY(alloc(h, h->pt_plane, W * h->Tcap, true));
    TRY(alloc(h, h->pt_begin, W * (h->Tcap + 1), true));
    TRY(alloc(h, h->pt_frame, W * h->Ocap, true));
    TRY(alloc(h, h->pt_z, W * h->Ocap * 2, true));
    h->have_planes = true;
    return 0;
}

// ------------------------------------------------------------------------ small kernels
__global__ void finalize_kernel(WinCtrl *ctrl, const double *acc, const double *aux_cost, int apply,
                                double *frames, const double *frames_cand, double *rho, const double *rho_cand,
                                const WinHdr *hdr, int Ncap, int Mcap, double beta, int w0) {
    const int w = blockIdx.x + w0;
    WinCtrl &c = ctrl[w];
    const double *a = acc + (size_t)w * kAcc;
    if (threadIdx.x == 0) {
        c.cand_cost_vis = a[0];
        c.cand_cost = a[0] + aux_cost[w];
        const double gdx = c.g_dot_dx + a[1];           // g . dx over poses + landmarks (full GN step)
        const double rdx = c.dx_reg_dx + a[2];          // dx^T (mu D) dx
        // model cost change of the step beta * dx_gn:  -(beta g.dx + beta^2/2 dx^T H dx),
        // with dx^T H dx = -g.dx - dx^T (mu D) dx for the regularised Gauss-Newton step
        c.model_change = -beta * gdx + 0.5 * beta * beta * (gdx + rdx);
    }
    if (apply) {
        const int N = hdr[w].N, M = hdr[w].M;
        for (int i = threadIdx.x; i < N * kFrameStride; i += blockDim.x)
            frames[(size_t)w * Ncap * kFrameStride + i] = frames_cand[(size_t)w * Ncap * kFrameStride + i];
        for (int i = threadIdx.x; i < M; i += blockDim.x) rho[(size_t)w * Mcap + i] = rho_cand[(size_t)w * Mcap + i];
    }
}

// Lambda = S^T S of the marginalisation prior (once per solve; constant across iterations)
__global__ void prior_lambda_kernel(const WinHdr *hdr, const double *S, double *L, int Ncap, int w0) {
    const int w = blockIdx.x + w0;
    const int d = 15 * hdr[w].n_prior, dcap = 15 * Ncap;
    const double *Sw = S + (size_t)w * dcap * dcap;
    double *Lw = L + (size_t)w * dcap * dcap;
    // gridDim.y CTAs share a window (a single window would otherwise leave 147 SMs idle for 0.29 ms)
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < d * d; e += blockDim.x * gridDim.y) {
        const int i = e / d, j = e - i * d;
        double s = 0.0;
        for (int k = 0; k < d; ++k) s += Sw[(size_t)k * d + i] * Sw[(size_t)k * d + j];
        Lw[e] = s;
    }
}

__global__ void init_ctrl_kernel(WinCtrl *ctrl, double mu, double radius) {
    WinCtrl &c = ctrl[blockIdx.x];
    if (threadIdx.x == 0) {
        c.mu = mu; c.radius = radius; c.iteration = 0; c.accepted = 0; c.done = 0;
        c.termination = PVIO_B200_TERM_NO_CONVERGENCE; c.solve_failed = 0; c.have_scale = 0; c.usable = 1;
    }
}

// Landmark post-pass (bundle_adjustor.cpp:277-296): depth test in every observing camera
// (anchor included) and mean pixel reprojection error.  One group of 16 lanes per landmark.
__global__ void postpass_kernel(const WinHdr *hdr, const WinConst *cst, const ObsRec *obs, const LmRec *lms,
                                const double *rho, const double *frames, uint8_t *valid, double *quality,
                                double *err_acc, int Ncap, int Mcap, int Kcap) {
    const int w = blockIdx.y;
    const WinHdr &H = hdr[w];
    const WinConst &wc = cst[w];
    __shared__ FrameSm F[kMaxFrames];
    if (threadIdx.x < H.N) make_frame(frames + ((size_t)w * Ncap + threadIdx.x) * kFrameStride, wc, F[threadIdx.x]);
    __syncthreads();
    const int lane = threadIdx.x & (kGroup - 1);
    const int gid = (blockIdx.x * blockDim.x + threadIdx.x) / kGroup;
    const int ngrp = gridDim.x * blockDim.x / kGroup;
    for (int l0 = 0; l0 < H.M; l0 += ngrp) {
        const int l = l0 + gid;
        const bool ok = l < H.M;
        const LmRec lr = lms[(size_t)w * Mcap + (ok ? l : 0)];
        const int n_obs = ok ? lm_nobs(lr.meta) : 0;
        const int anchor = lm_anchor(lr.meta);
        const unsigned fmask = ok ? lm_mask(lr.meta) : 0u;
        double x[3];
        float xf[3], cl[3];
        world_point(F[anchor], lr.zrx, lr.zry, ok ? rho[(size_t)w * Mcap + l] : 1.0, x, xf, cl);
        // lane = frame: its observation is record popc(mask below the lane); the anchor lane uses z_ref
        int frame = -1;
        float zx = 0.f, zy = 0.f;
        if ((fmask >> lane) & 1u) {
            const ObsRec o = obs[(size_t)w * Kcap + lr.obs_begin + __popc(fmask & ((1u << lane) - 1u))];
            frame = lane; zx = o.zx; zy = o.zy;
        } else if (lane == anchor && ok) {
            frame = anchor; zx = lr.zrx; zy = lr.zry;
        }
        double e = 0.0;
        int bad = 0;
        if (frame >= 0) {
            const FrameSm &Ft = F[frame];
            const double d0 = x[0] - Ft.c[0], d1 = x[1] - Ft.c[1], d2 = x[2] - Ft.c[2];
            const double y0 = Ft.Rwc[0] * d0 + Ft.Rwc[3] * d1 + Ft.Rwc[6] * d2;
            const double y1 = Ft.Rwc[1] * d0 + Ft.Rwc[4] * d1 + Ft.Rwc[7] * d2;
            const double y2 = Ft.Rwc[2] * d0 + Ft.Rwc[5] * d1 + Ft.Rwc[8] * d2;
            if (y2 <= 1.0e-3 || y2 > 50.0) bad = 1;                        // :286
            const double ex = (y0 / y2 - (double)zx) * wc.fx, ey = (y1 / y2 - (double)zy) * wc.fy;
            e = sqrt(ex * ex + ey * ey);                                   // :291
        }
        const unsigned badm = __ballot_sync(0xffffffffu, bad);
        const int grp_bad = (badm >> (threadIdx.x & 16)) & 0xffff;
#pragma unroll
        for (int off = kGroup / 2; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off, kGroup);
        if (ok && lane == 0) {
            const bool v = grp_bad == 0;
            if (valid) valid[(size_t)w * Mcap + l] = v ? 1 : 0;
            if (quality) quality[(size_t)w * Mcap + l] = v ? e / fmax((double)(n_obs + 1), 1.0) : 0.0;
            if (err_acc) { atomicAdd(&err_acc[2 * w], e); atomicAdd(&err_acc[2 * w + 1], (double)(n_obs + 1)); }
        }
    }
}

// ------------------------------------------------------------------------ packing
static int pack_window(Handle *h, int slot, const pvio_b200_window *w, const pvio_b200_state *s) {
    const int N = w->n_frames, M = w->n_landmarks, K = w->n_obs;
    if (slot < 0 || slot >= h->W) return fail(h, PVIO_B200_EINVAL, "slot out of range");
    if (N < 1 || N > h->Ncap || N > kMaxFrames || M > h->Mcap || K > h->Kcap || M < 0 || K < 0)
        return fail(h, PVIO_B200_EINVAL, "window exceeds the handle's capacity");
    WinHdr &H = h->hdr.h[slot];
    WinConst &C = h->cst.h[slot];
    memset(&H, 0, sizeof(H));
    H.N = N; H.M = M; H.K = K; H.use_inertial = w->use_inertial ? 1 : 0;
    for (int f = 0; f < N; ++f) if (w->frame_fixed && w->frame_fixed[f]) H.fixed_mask |= 1 << f;
    memcpy(C.cam_q, w->cam_q_cs, 32); memcpy(C.cam_p, w->cam_p_cs, 24);
    memcpy(C.imu_q, w->imu_q_cs, 32); memcpy(C.imu_p, w->imu_p_cs, 24);
    memcpy(C.sic, w->sqrt_inv_cov, 32);
    C.fx = w->fx; C.fy = w->fy; C.cauchy_a = w->cauchy_a > 0 ? w->cauchy_a : 1.0; C.plane_sic = w->plane_sqrt_inv_cov;
    // frames + origin (mean frame position keeps |x_l| ~ depth in the fp32 Jacobian arithmetic)
    double *fr = h->frames.h + (size_t)slot * h->Ncap * kFrameStride;
    memcpy(fr, s->frames, sizeof(double) * N * kFrameStride);
    double o[3] = {0, 0, 0};
    for (int f = 0; f < N; ++f) for (int k = 0; k < 3; ++k) o[k] += fr[f * kFrameStride + 4 + k];
    for (int k = 0; k < 3; ++k) C.origin[k] = o[k] / N;
    // landmarks sorted by anchor (the reference's first-visit order already is, bundle_adjustor.cpp:92-103)
    std::vector<int32_t> &perm = h->perm[slot];
    perm.resize(M);
    std::iota(perm.begin(), perm.end(), 0);
    bool sorted = true;
    for (int l = 1; l < M; ++l) if (w->lm_anchor[l] < w->lm_anchor[l - 1]) { sorted = false; break; }
    if (!sorted) std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w->lm_anchor[a] < w->lm_anchor[b]; });
    ObsRec *ob = h->obs.h + (size_t)slot * h->Kcap;
    LmRec *lm = h->lms.h + (size_t)slot * h->Mcap;
    double *rh = h->rho.h + (size_t)slot * h->Mcap;
    int k_out = 0, nch = 0;
    for (int lp = 0; lp < M; ++lp) {
        const int l = perm[lp];
        const int a = w->lm_anchor[l];
        const int b0 = w->lm_obs_begin[l], b1 = w->lm_obs_begin[l + 1];
        const int n = b1 - b0;
        if (a < 0 || a >= N || n < 0 || n >= kGroup) return fail(h, PVIO_B200_EINVAL, "landmark with bad anchor / too many observations");
        lm[lp].zrx = (float)w->lm_z_ref[2 * l];
        lm[lp].zry = (float)w->lm_z_ref[2 * l + 1];
        lm[lp].obs_begin = k_out;
        rh[lp] = s->inv_depth[l];
        // records in increasing frame order: the frame index lives in the landmark's mask, not in the record
        unsigned seen = 0;
        for (int k = b0; k < b1; ++k) {
            const int f = w->obs_frame[k];
            if (f <= a || f >= N || (seen >> f) & 1)
                return fail(h, PVIO_B200_EINVAL, "observation frames must be distinct and later than the anchor (Track::first_frame is the lowest id)");
            seen |= 1u << f;
        }
        for (int k = b0; k < b1; ++k) {
            const int f = w->obs_frame[k];
            const int pos = k_out + __builtin_popcount(seen & ((1u << f) - 1u));
            ob[pos].zx = (float)w->obs_z[2 * k];
            ob[pos].zy = (float)w->obs_z[2 * k + 1];
        }
        k_out += n;
        lm[lp].meta = lm_meta(a, (w->lm_in_victim && w->lm_in_victim[l]) ? 1 : 0, n, seen);
        // chunks: <= kChunk landmarks of one anchor
        if (nch == 0 || (H.chunk_meta[nch - 1] >> 8) != a || (H.chunk_meta[nch - 1] & 0xff) == kChunk) {
            if (nch == kMaxChunks || nch == h->Mcap / 32 + h->Ncap + 1) return fail(h, PVIO_B200_EINVAL, "too many landmark chunks");
            H.chunk_begin[nch] = lp;
            H.chunk_meta[nch] = (a << 8);
            ++nch;
        }
        H.chunk_meta[nch - 1] += 1;
    }
    H.n_chunks = nch;
    h->slot_M[slot] = M; h->slot_N[slot] = N; h->slot_K[slot] = K;
    h->max_slot_N = std::max(h->max_slot_N, N);
    h->max_slot_free = std::max(h->max_slot_free, N - __builtin_popcount((unsigned)H.fixed_mask & ((1u << N) - 1u)));
    h->perm_identity[slot] = sorted ? 1 : 0;
    // inertial part
    H.n_imu = w->use_inertial ? w->n_imu : 0;
    H.n_prior = w->use_inertial ? w->n_prior : 0;
    if (H.n_imu > 0 || H.n_prior > 0) {
        TRY(ensure_inertial(h));
        if (H.n_imu > h->Ncap || H.n_prior > h->Ncap) return fail(h, PVIO_B200_EINVAL, "too many IMU / prior frames");
        int32_t *ii = h->imu_idx.h + (size_t)slot * h->Ncap * 2;
        for (int n = 0; n < H.n_imu; ++n) { ii[2 * n] = w->imu_frame_i[n]; ii[2 * n + 1] = w->imu_frame_j[n]; }
        memcpy(h->imu_data.h + (size_t)slot * h->Ncap * kImuStride, w->imu_data, sizeof(double) * H.n_imu * kImuStride);
        const size_t dcap = 15 * (size_t)h->Ncap, d = 15 * (size_t)H.n_prior;
        for (int n = 0; n < H.n_prior; ++n) h->prior_frames.h[(size_t)slot * h->Ncap + n] = w->prior_frames[n];
        if (d > 0) {
            memcpy(h->prior_S.h + (size_t)slot * dcap * dcap, w->prior_S, sizeof(double) * d * d);   // dense d x d, row-major
            memcpy(h->prior_e.h + (size_t)slot * dcap, w->prior_e, sizeof(double) * d);
            memcpy(h->prior_x0.h + (size_t)slot * h->Ncap * kFrameStride, w->prior_state0, sizeof(double) * H.n_prior * kFrameStride);
        }
    }
    // planes
    H.n_planes = w->n_planes; H.n_ptracks = w->n_plane_tracks;
    if (H.n_ptracks > 0) {
        const int O = w->pt_obs_begin[H.n_ptracks];
        TRY(ensure_planes(h, H.n_ptracks, O));
        if (H.n_planes > h->Pcap) return fail(h, PVIO_B200_EINVAL, "too many planes");
        memcpy(h->plane_param.h + (size_t)slot * h->Pcap * 4, w->plane_param, sizeof(double) * 4 * H.n_planes);
        memcpy(h->pt_plane.h + (size_t)slot * h->Tcap, w->pt_plane, sizeof(int32_t) * H.n_ptracks);
        memcpy(h->pt_begin.h + (size_t)slot * (h->Tcap + 1), w->pt_obs_begin, sizeof(int32_t) * (H.n_ptracks + 1));
        memcpy(h->pt_frame.h + (size_t)slot * h->Ocap, w->pt_obs_frame, sizeof(int32_t) * O);
        float *z = h->pt_z.h + (size_t)slot * h->Ocap * 2;
        for (int i = 0; i < 2 * O; ++i) z[i] = (float)w->pt_obs_z[i];
        for (int t = 0; t < H.n_ptracks; ++t)
            if (w->pt_obs_begin[t + 1] - w->pt_obs_begin[t] > kMaxFrames) return fail(h, PVIO_B200_EINVAL, "plane track too long");
    }
    return 0;
}

template <typename T>
static int h2d(Handle *h, DevBuf<T> &b, size_t per, int w0, int n, cudaStream_t st) {
    if (!b.d || !b.h || per == 0) return 0;
    CK(h, cudaMemcpyAsync(b.d + per * w0, b.h + per * w0, per * n * sizeof(T), cudaMemcpyHostToDevice, st));
    return 0;
}

// Host -> device copy of windows [w0, w0 + n) on stream st
static int upload_range(Handle *h, int w0, int n, cudaStream_t st) {
    if (n < 1 || w0 < 0 || w0 + n > h->W) return fail(h, PVIO_B200_EINVAL, "bad window range");
    const size_t N = h->Ncap;
    TRY(h2d(h, h->hdr, 1, w0, n, st));
    TRY(h2d(h, h->cst, 1, w0, n, st));
    TRY(h2d(h, h->obs, h->Kcap, w0, n, st));
    TRY(h2d(h, h->lms, h->Mcap, w0, n, st));
    TRY(h2d(h, h->rho, h->Mcap, w0, n, st));
    TRY(h2d(h, h->frames, N * kFrameStride, w0, n, st));
    bool any_prior = false;
    if (h->have_inertial) {
        const size_t dcap = 15 * N;
        TRY(h2d(h, h->imu_idx, N * 2, w0, n, st));
        TRY(h2d(h, h->imu_data, N * kImuStride, w0, n, st));
        TRY(h2d(h, h->prior_frames, N, w0, n, st));
        TRY(h2d(h, h->prior_S, dcap * dcap, w0, n, st));
        TRY(h2d(h, h->prior_e, dcap, w0, n, st));
        TRY(h2d(h, h->prior_x0, N * kFrameStride, w0, n, st));
        for (int i = w0; i < w0 + n; ++i) any_prior |= h->hdr.h[i].n_prior > 0;
        if (any_prior) {
            prior_lambda_kernel<<<dim3(n, n < 64 ? 32 : 1), 256, 0, st>>>(h->hdr.d, h->prior_S.d, h->prior_L.d, h->Ncap, w0);
            ++h->launches;
        }
    }
    if (h->have_planes) {
        TRY(h2d(h, h->plane_param, (size_t)h->Pcap * 4, w0, n, st));
        TRY(h2d(h, h->pt_plane, h->Tcap, w0, n, st));
        TRY(h2d(h, h->pt_begin, h->Tcap + 1, w0, n, st));
        TRY(h2d(h, h->pt_frame, h->Ocap, w0, n, st));
        TRY(h2d(h, h->pt_z, (size_t)h->Ocap * 2, w0, n, st));
    }
    return 0;
}

static int upload(Handle *h, int n) {
    TRY(upload_range(h, 0, n, h->stream));
    h->n_uploaded = n;
    return 0;
}

int pack_and_upload(Handle *h, const pvio_b200_window *w, const pvio_b200_state *s) {
    TRY(pack_window(h, 0, w, s));
    return upload(h, 1);
}

// ------------------------------------------------------------------------ launches
struct StepCfg {
    double mu = -1.0;          // < 0: per-window ctrl.mu
    double beta = 1.0;             // truncated Gauss-Newton step: step = beta * dx_gn
    double step_a = 0.0;           // dogleg: step = step_b * dx_gn - step_a * v  (step_b defaults to beta)
    double step_b = -1.0;
    int update_grid = 0;           // > 0: CTAs per window of the update sweep
    int apply = 0;
    int compute_scale = 1;
    int alias_bias = 0;
    int dump = 0;
    bool skip_linearize = false;   // reuse the previous linearisation (trust-region retry)
    int w0 = 0;                    // first window (sub-batch pipelining)
    cudaStream_t stream = nullptr; // nullptr: the handle's stream
};

static int lin_grid_x(Handle *h, int n) {
    // one CTA per window once the batch fills the machine; otherwise split a window's chunks
    const int target = 2 * h->sm_count;
    int gx = std::max(1, target / std::max(n, 1));
    return std::min(gx, n * 2 < h->sm_count ? 32 : 16);
}

static size_t solve_smem(Handle *h, int w0, int n, bool lean) {
    // worst case over the windows of the launch
    size_t best = 0;
    for (int i = w0; i < w0 + n; ++i) {
        const WinHdr &H = h->hdr.h[i];
        size_t D = (H.use_inertial ? 15 : 6) * (size_t)H.N;
        if (lean) {                       // the lean kernel drops constant frames from the system
            const int nfree = H.N - __builtin_popcount((unsigned)H.fixed_mask & ((1u << H.N) - 1u));
            if (nfree > 0) D = 6 * (size_t)nfree;
        }
        const size_t nb = (D + 3) / 4, Dp = nb * 4;
        const size_t np_ = (size_t)H.N * (H.N + 1) / 2;
        size_t scr = lean ? Dp + 10 * 36 + 36 : std::max<size_t>(Dp, np_ * 36 + (size_t)H.N * 36 + (size_t)H.N * 12);
        if (H.use_inertial) scr = std::max<size_t>(scr, 2 * (4 * 450 + 64));
        if (H.n_prior > 0) scr = std::max<size_t>(scr, 3 * 15 * (size_t)H.n_prior + 9 * (size_t)H.n_prior);
        const size_t bytes = sizeof(double) * ((nb + 1) * (nb + 2) / 2 * 16 + 4 * Dp + nb * 16 + (size_t)H.N * 36 + scr);
        best = std::max(best, bytes);
    }
    return best;
}

static int run_linearize(Handle *h, int n, const StepCfg &c) {
    const int gx = lin_grid_x(h, n);
    cudaStream_t st = c.stream ? c.stream : h->stream;
    const size_t npc = (size_t)h->Ncap * (h->Ncap + 1) / 2;
    if (gx > 1) CK(h, cudaMemsetAsync(h->Hred.d, 0, sizeof(double) * h->Hred.n, st));
    (void)npc;
    LinArgs a;
    a.hdr = h->hdr.d; a.cst = h->cst.d; a.obs = h->obs.d; a.lms = h->lms.d; a.rho = h->rho.d; a.frames = h->frames.d;
    a.ctrl = h->ctrl.d; a.lm_scale = h->lm_scale.d; a.lm_aux = h->lm_aux.d; a.hs_out = h->hs.d; a.hs_stride = h->hs_stride;
    a.Hred = h->Hred.d; a.Hdd = h->Hdd.d; a.gdir = h->gdir.d; a.gred = h->gred.d; a.cost_vis = h->cost_vis.d;
    a.Ncap = h->Ncap; a.Mcap = h->Mcap; a.Kcap = h->Kcap;
    a.compute_scale = c.compute_scale; a.victim_only = 0; a.mu_override = c.mu; a.w0 = c.w0;
    const int slot = (h->kev_count % 512) * 2;
    if (h->kev.empty()) {
        h->kev.resize(1024);
        for (auto &e : h->kev) CK(h, cudaEventCreate(&e));
    }
    if (!h->capturing) CK(h, cudaEventRecord(h->kev[slot], st));
    // few windows: the group-per-landmark kernel exposes more parallelism per window (latency);
    // many windows: the thread-per-landmark kernel issues ~2x fewer instructions (throughput)
    // (the group kernel owns at most 256 Phase-B tiles: N <= 15)
    // PVIO_B200_TC=1 and windows of <= 10 frames: the Schur SYRK runs on the tensor cores (tcgen05, ba_lin3.cuh).
    // Opt-in: parity-green but 10 % slower than the CUDA-core SYRK on B200 (profiles/r01c_lin_tc.md)
    const bool tc_ok = h->max_slot_N <= kTcMaxFrames && h->use_tc;
    if (n * 2 < h->sm_count && h->Ncap * (h->Ncap + 1) <= kLinThreads) lin_schur_kernel<true><<<dim3(gx, n), kLinThreads, lin_smem_bytes(), st>>>(a);
    else if (tc_ok && h->tc_gs == 4) lin_tc_kernel<true, 4><<<dim3(gx, n), kLinThreads, lin3_smem_bytes(std::min(h->Ncap, kTcMaxFrames)), st>>>(a);
    else if (tc_ok && h->tc_gs == 2) lin_tc_kernel<true, 2><<<dim3(gx, n), kLinThreads, lin3_smem_bytes(std::min(h->Ncap, kTcMaxFrames)), st>>>(a);
    else if (tc_ok) lin_tc_kernel<true, 1><<<dim3(gx, n), kLinThreads, lin3_smem_bytes(std::min(h->Ncap, kTcMaxFrames)), st>>>(a);
    else if (h->split_schur && gx == 1) {
        // split stage (default for whole-window CTAs): Phase A (records + direct blocks, ba_lin4.cuh), then the Schur
        // sum as its own kernel at ~3x the occupancy (ba_schur.cuh): 0.83 ms instead of 0.90 ms per 4096 cfg2 windows
        switch (h->split_shape) {
            case 1: lin_a_kernel<true, 6, 3><<<dim3(1, n), 192, lin4_smem_bytes<6>(h->Ncap), st>>>(a); break;
            case 2: lin_a_kernel<true, 4, 4><<<dim3(1, n), 128, lin4_smem_bytes<4>(h->Ncap), st>>>(a); break;
            case 3: lin_a_kernel<true, 4, 5><<<dim3(1, n), 128, lin4_smem_bytes<4>(h->Ncap), st>>>(a); break;
            default: lin_a_kernel<true, 8, 2><<<dim3(1, n), 256, lin4_smem_bytes<8>(h->Ncap), st>>>(a); break;
        }
        // CTA = the tiles of the free-frame pairs x 4 k-split lanes (x 2 / x 1 when that exceeds 256 threads)
        const int nfree = std::max(1, h->max_slot_free), ntask = nfree * (nfree + 1) / 2;
        const int ks = ntask * 4 <= kSchurThreads ? 4 : (ntask * 2 <= kSchurThreads ? 2 : 1);
        const int nthr = std::min(kSchurThreads, ((ntask * ks + 31) / 32) * 32);
        if (h->tc_mode == 2 && h->max_slot_N <= kTcMaxFrames) schur_tc_kernel<<<n, 128, schur_tc_smem_bytes(std::min(h->Ncap, kTcMaxFrames)), st>>>(a);
        else if (nthr <= 160) schur_kernel<160, 4><<<n, nthr, schur_smem_bytes(h->Ncap), st>>>(a);
        else schur_kernel<256, 2><<<n, nthr, schur_smem_bytes(h->Ncap), st>>>(a);
        ++h->launches;
    }
    else lin_tpl_kernel<true><<<dim3(gx, n), kLinThreads, lin2_smem_bytes(h->Ncap), st>>>(a);
    if (!h->capturing) { CK(h, cudaEventRecord(h->kev[slot + 1], st)); ++h->kev_count; }
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

static int run_solve(Handle *h, int n, const StepCfg &c) {
    SolveArgs a;
    memset(&a, 0, sizeof(a));
    a.hdr = h->hdr.d; a.cst = h->cst.d; a.frames = h->frames.d; a.ctrl = h->ctrl.d;
    a.Hred = h->Hred.d; a.Hdd = h->Hdd.d; a.gdir = h->gdir.d; a.gred = h->gred.d; a.cost_vis = h->cost_vis.d;
    a.imu_idx = h->imu_idx.d; a.imu_data = h->imu_data.d; a.alias_bias = c.alias_bias;
    a.prior_frames = h->prior_frames.d; a.prior_S = h->prior_S.d; a.prior_L = h->prior_L.d; a.prior_e = h->prior_e.d;
    a.prior_x0 = h->prior_x0.d;
    a.plane_param = h->plane_param.d; a.pt_plane = h->pt_plane.d; a.pt_begin = h->pt_begin.d; a.pt_frame = h->pt_frame.d;
    a.pt_z = h->pt_z.d; a.Pcap = h->Pcap; a.Tcap = h->Tcap; a.Ocap = h->Ocap;
    a.pose_scale = h->pose_scale.d; a.dx_pose = h->dx_pose.d; a.v_pose = h->v_pose.d;
    a.Hfull = c.dump ? h->Hfull.d : nullptr; a.gfull = c.dump ? h->gfull.d : nullptr;
    a.Ncap = h->Ncap; a.compute_scale = c.compute_scale; a.mu_override = c.mu; a.w0 = c.w0;
    a.dbg = nullptr;
    if (!h->capturing && getenv("PVIO_B200_SOLVE_STAMPS")) {       // profiling aid: clock64 stamps of the solve kernel's phases
        static long long *dbg = nullptr;
        long long hst[16];
        if (!dbg) { cudaMalloc(&dbg, 16 * sizeof(long long)); cudaMemset(dbg, 0, 16 * sizeof(long long)); }
        else {
            cudaStreamSynchronize(h->stream);
            cudaMemcpy(hst, dbg, sizeof(hst), cudaMemcpyDeviceToHost);
            fprintf(stderr, "solve stamps (clk):");
            for (int i = 1; i < 6; ++i) fprintf(stderr, " %lld", hst[i] - hst[i - 1]);
            fprintf(stderr, "\n");
        }
        a.dbg = dbg;
    }
    cudaStream_t st = c.stream ? c.stream : h->stream;
    // visual-only batches: the lean kernel (no IMU / prior / plane code, no full staging), a narrow CTA
    // per window so that many windows are resident per SM; otherwise the full kernel with a wide CTA
    bool visual = n >= 64;
    for (int i = c.w0; i < c.w0 + n && visual; ++i) {
        const WinHdr &H = h->hdr.h[i];
        if (H.use_inertial || H.n_ptracks > 0) visual = false;
    }
    size_t smem = solve_smem(h, c.w0, n, visual);
    if (visual && smem > 48 * 1024) { visual = false; smem = solve_smem(h, c.w0, n, false); }
    if (smem > 220 * 1024) return fail(h, PVIO_B200_EINVAL, "reduced system too large for shared memory");
    if (visual) solve_kernel_visual<<<n, 64, smem, st>>>(a);
    else solve_kernel<<<n, 256, smem, st>>>(a);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

static CostArgs make_cost_args(Handle *h, const StepCfg &c) {
    CostArgs k;
    memset(&k, 0, sizeof(k));
    k.hdr = h->hdr.d; k.cst = h->cst.d; k.frames_cand = h->frames_cand.d; k.frames_cur = h->frames.d;
    k.imu_idx = h->imu_idx.d; k.imu_data = h->imu_data.d; k.alias_bias = c.alias_bias;
    k.prior_frames = h->prior_frames.d; k.prior_S = h->prior_S.d; k.prior_e = h->prior_e.d; k.prior_x0 = h->prior_x0.d;
    k.plane_param = h->plane_param.d; k.pt_plane = h->pt_plane.d; k.pt_begin = h->pt_begin.d; k.pt_frame = h->pt_frame.d;
    k.pt_z = h->pt_z.d; k.Pcap = h->Pcap; k.Tcap = h->Tcap; k.Ocap = h->Ocap; k.Ncap = h->Ncap; k.out = h->aux_cost.d;
    k.w0 = c.w0;
    return k;
}

// |J v|^2 for the Cauchy point of the dogleg step (acc slots 9 and 10); single window path
static int run_jv(Handle *h, const StepCfg &c) {
    cudaStream_t st = c.stream ? c.stream : h->stream;
    CK(h, cudaMemsetAsync(h->acc.d + 9, 0, sizeof(double) * 2, st));
    UpdArgs u;
    memset(&u, 0, sizeof(u));
    u.hdr = h->hdr.d; u.cst = h->cst.d; u.obs = h->obs.d; u.lms = h->lms.d; u.rho = h->rho.d; u.frames = h->frames.d;
    u.ctrl = h->ctrl.d; u.lm_scale = h->lm_scale.d; u.lm_aux = h->lm_aux.d; u.hs = h->hs.d; u.hs_stride = h->hs_stride; u.dx_pose = h->dx_pose.d; u.acc = h->acc.d;
    u.Ncap = h->Ncap; u.Mcap = h->Mcap; u.Kcap = h->Kcap; u.mu_override = c.mu; u.w0 = 0; u.v_pose = h->v_pose.d;
    jv_vision_kernel<true><<<dim3(16, 1), kLinThreads, 0, st>>>(u);
    JvAuxArgs ja;
    ja.c = make_cost_args(h, c);
    ja.v_pose = h->v_pose.d; ja.acc = h->acc.d;
    jv_aux_kernel<<<1, 64, sizeof(double) * 15 * kMaxFrames, st>>>(ja);
    h->launches += 2;
    CK(h, cudaGetLastError());
    return 0;
}

static int run_update(Handle *h, int n, const StepCfg &c) {
    cudaStream_t st = c.stream ? c.stream : h->stream;
    CK(h, cudaMemsetAsync(h->acc.d + (size_t)kAcc * c.w0, 0, sizeof(double) * kAcc * n, st));
    UpdArgs u;
    u.hdr = h->hdr.d; u.cst = h->cst.d; u.obs = h->obs.d; u.lms = h->lms.d; u.rho = h->rho.d; u.frames = h->frames.d;
    u.ctrl = h->ctrl.d; u.lm_scale = h->lm_scale.d; u.lm_aux = h->lm_aux.d; u.hs = h->hs.d; u.hs_stride = h->hs_stride; u.dx_pose = h->dx_pose.d;
    u.rho_cand = h->rho_cand.d; u.frames_cand = h->frames_cand.d; u.dx_lm = h->dx_lm.d; u.acc = h->acc.d;
    u.Ncap = h->Ncap; u.Mcap = h->Mcap; u.Kcap = h->Kcap; u.mu_override = c.mu; u.w0 = c.w0;
    u.step_a = c.step_a; u.step_b = c.step_b >= 0.0 ? c.step_b : c.beta; u.v_pose = h->v_pose.d;
    const int gx = lin_grid_x(h, n);
    // few windows: one chunk per CTA-warp so that a single window spreads over 16 CTAs
    const int ugx = n * 2 < h->sm_count ? 16 : 1;
    (void)gx;
    update_tpl_kernel<true><<<dim3(ugx, n), kLinThreads, 0, st>>>(u);
    ++h->launches;
    CostArgs k = make_cost_args(h, c);
    aux_cost_kernel<<<n, 64, sizeof(double) * 15 * kMaxFrames, st>>>(k);
    ++h->launches;
    finalize_kernel<<<n, 128, 0, st>>>(h->ctrl.d, h->acc.d, h->aux_cost.d, c.apply, h->frames.d, h->frames_cand.d,
                                      h->rho.d, h->rho_cand.d, h->hdr.d, h->Ncap, h->Mcap, c.beta, c.w0);
    ++h->launches;
    CK(h, cudaGetLastError());
    return 0;
}

static int run_step_raw(Handle *h, int n, const StepCfg &c, int kind) {
    if (!c.skip_linearize) {
        TRY(run_linearize(h, n, c));
        TRY(run_solve(h, n, c));
    }
    if (kind == 0) TRY(run_update(h, n, c));
    return 0;
}

// kind 0: linearise + solve (unless skipped) + update sweep; kind 1: linearise + solve only.
// Small launches (latency path) are captured once into a CUDA graph and replayed: 6-8 nodes cost one
// launch instead of 6-8.
static int run_step(Handle *h, int n, const StepCfg &c, int kind = 0) {
    const bool graphable = h->use_graphs && n * 2 < h->sm_count && c.w0 == 0 && c.stream == nullptr && !c.dump &&
                           !getenv("PVIO_B200_SOLVE_STAMPS");
    if (!graphable) return run_step_raw(h, n, c, kind);
    const WinHdr &H = h->hdr.h[0];
    int shape = 0;
    for (int i = 0; i < n; ++i) shape = shape * 31 + h->hdr.h[i].N * 4 + h->hdr.h[i].use_inertial * 2 + (h->hdr.h[i].n_ptracks > 0);
    Handle::GraphKey key(n, shape, H.N, kind, c.compute_scale, c.skip_linearize ? 1 : 0, c.apply, c.alias_bias,
                         (int)(solve_smem(h, 0, n, false) >> 4), c.mu, c.step_a, c.step_b >= 0.0 ? c.step_b : c.beta);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        const int64_t l0 = h->launches;
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        CK(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->capturing = true;
        const int rc = run_step_raw(h, n, c, kind);
        h->capturing = false;
        const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
        if (rc != 0) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) return fail(h, PVIO_B200_ECUDA, "cudaStreamEndCapture", e);
        CK(h, cudaGraphInstantiate(&ge, g, 0));
        cudaGraphDestroy(g);
        const int nl = (int)(h->launches - l0);
        h->launches = l0;
        it = h->graphs.emplace(key, std::make_pair(ge, nl)).first;
    }
    CK(h, cudaGraphLaunch(it->second.first, h->stream));
    h->launches += it->second.second;
    return 0;
}

// pinned staging -> caller arrays (landmark order un-permuted), windows [w0, w0 + n)
static int scatter_dx(Handle *h, int w0, int n, double *dx, int64_t dx_stride, double *costs) {
    for (int i = w0; i < w0 + n; ++i) {
        const int N = h->slot_N[i], M = h->slot_M[i];
        if (dx) {
            double *o = dx + (size_t)i * dx_stride;
            memcpy(o, h->dx_pose.h + (size_t)i * h->Ncap * 15, sizeof(double) * N * 15);
            const double *dl = h->dx_lm.h + (size_t)i * h->Mcap;
            if (h->perm_identity[i]) memcpy(o + N * 15, dl, sizeof(double) * M);
            else { const std::vector<int32_t> &perm = h->perm[i]; for (int lp = 0; lp < M; ++lp) o[N * 15 + perm[lp]] = dl[lp]; }
        }
        if (costs) { costs[2 * i] = h->ctrl.h[i].cost; costs[2 * i + 1] = h->ctrl.h[i].cand_cost; }
        if (h->ctrl.h[i].solve_failed) return fail(h, PVIO_B200_ENUMERIC, "reduced system not positive definite");
    }
    return 0;
}

static int download_dx(Handle *h, int n, double *dx, int64_t dx_stride, double *costs) {
    CK(h, cudaMemcpyAsync(h->dx_pose.h, h->dx_pose.d, sizeof(double) * h->Ncap * 15 * n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->dx_lm.h, h->dx_lm.d, sizeof(double) * h->Mcap * n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->ctrl.h, h->ctrl.d, sizeof(WinCtrl) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    return scatter_dx(h, 0, n, dx, dx_stride, costs);
}

}  // namespace pvio

using namespace pvio;

// ======================================================================== C ABI
extern "C" {

const char *pvio_b200_version(void) { return "pvio_b200 0.1 (sm_100a)"; }

int pvio_b200_create(int device, int max_windows, int max_frames, int max_landmarks, int max_obs,
                     pvio_b200_handle *out) {
    if (!out) return PVIO_B200_EINVAL;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device >= ndev) return PVIO_B200_ENODEV;
    if (max_windows < 1 || max_frames < 1 || max_frames > kMaxFrames || max_landmarks < 1 || max_obs < 1) return PVIO_B200_EINVAL;
    Handle *h = new Handle();
    h->device = device; h->W = max_windows; h->Ncap = max_frames; h->Mcap = max_landmarks; h->Kcap = max_obs;
    if (cudaSetDevice(device) != cudaSuccess) { delete h; return PVIO_B200_ENODEV; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    h->sm_count = prop.multiProcessorCount;
    *out = reinterpret_cast<pvio_b200_handle>(h);
    CK(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CK(h, cudaEventCreate(&h->ev0)); CK(h, cudaEventCreate(&h->ev1));
    CK(h, cudaEventCreate(&h->evk0)); CK(h, cudaEventCreate(&h->evk1));
    const size_t W = h->W, N = h->Ncap, M = h->Mcap, K = h->Kcap, npc = N * (N + 1) / 2;
    TRY(alloc(h, h->hdr, W, true)); TRY(alloc(h, h->cst, W, true));
    TRY(alloc(h, h->obs, W * K, true)); TRY(alloc(h, h->lms, W * M, true));
    TRY(alloc(h, h->rho, W * M, true)); TRY(alloc(h, h->frames, W * N * kFrameStride, true));
    TRY(alloc(h, h->ctrl, W, true));
    TRY(alloc(h, h->rho_cand, W * M, false)); TRY(alloc(h, h->frames_cand, W * N * kFrameStride, false));
    TRY(alloc(h, h->lm_scale, W * M, false)); TRY(alloc(h, h->lm_aux, W * M, false));
    h->hs_stride = (size_t)(M / 32 + N + 1) * 32 * hs_rec(N);
    TRY(alloc(h, h->hs, W * h->hs_stride, false));
    TRY(alloc(h, h->dx_lm, W * M, true)); TRY(alloc(h, h->dx_pose, W * N * 15, true));
    TRY(alloc(h, h->pose_scale, W * N * 15, false)); TRY(alloc(h, h->v_pose, W * N * 15, false));
    {   // the reduced-system outputs of the linearise kernel live in ONE allocation so that the
        // multi-CTA-per-window mode (atomic accumulation) needs a single memset per launch
        const size_t n_sys = W * (npc * 36 + N * 36 + N * 6 + N * 6 + 1);
        TRY(alloc(h, h->Hred, n_sys, false));
        h->Hdd.d = h->Hred.d + W * npc * 36; h->Hdd.n = 0;
        h->gdir.d = h->Hdd.d + W * N * 36; h->gdir.n = 0;
        h->gred.d = h->gdir.d + W * N * 6; h->gred.n = 0;
        h->cost_vis.d = h->gred.d + W * N * 6; h->cost_vis.n = 0;
    } TRY(alloc(h, h->acc, W * kAcc, true)); TRY(alloc(h, h->aux_cost, W, false));
    TRY(alloc(h, h->Hfull, (15 * N) * (15 * N), true)); TRY(alloc(h, h->gfull, 15 * N, true));
    // unallocated optional buffers still need valid (dummy) device pointers? kernels never touch them
    h->perm.resize(W); h->perm_identity.assign(W, 1); h->slot_M.assign(W, 0); h->slot_N.assign(W, 0); h->slot_K.assign(W, 0);
    CK(h, cudaFuncSetAttribute(lin_schur_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_smem_bytes()));
    CK(h, cudaFuncSetAttribute(lin_tpl_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin2_smem_bytes(kMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_tc_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin3_smem_bytes(kTcMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_tc_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin3_smem_bytes(kTcMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_tc_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin3_smem_bytes(kTcMaxFrames)));
    { const char *e = getenv("PVIO_B200_TC_GS"); h->tc_gs = e ? atoi(e) : 2; }
    { const char *e = getenv("PVIO_B200_TC"); h->use_tc = e && e[0] == '1'; h->tc_mode = e ? atoi(e) : 0; }
    CK(h, cudaFuncSetAttribute(schur_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)schur_tc_smem_bytes(kTcMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_a_kernel<true, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin4_smem_bytes<8>(kMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_a_kernel<true, 6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin4_smem_bytes<6>(kMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_a_kernel<true, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin4_smem_bytes<4>(kMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_a_kernel<true, 4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin4_smem_bytes<4>(kMaxFrames)));
    { const char *e = getenv("PVIO_B200_SPLIT_SHAPE"); h->split_shape = e ? atoi(e) : 3; }    // 4 warps x 5 CTAs per SM (96 registers) measured best
    CK(h, cudaFuncSetAttribute(schur_kernel<160, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)schur_smem_bytes(kMaxFrames)));
    CK(h, cudaFuncSetAttribute(schur_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)schur_smem_bytes(kMaxFrames)));
    { const char *e = getenv("PVIO_B200_SPLIT"); h->split_schur = !(e && e[0] == '0'); }     // default on; 0: the fused kernel
    CK(h, cudaFuncSetAttribute(lin_tpl_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin2_smem_bytes(kMaxFrames)));
    CK(h, cudaFuncSetAttribute(lin_schur_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lin_smem_bytes()));
    CK(h, cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    CK(h, cudaFuncSetAttribute(solve_kernel_visual, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
    init_ctrl_kernel<<<(int)W, 32, 0, h->stream>>>(h->ctrl.d, 1e-8, 1e4);
    ++h->launches;
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

void pvio_b200_destroy(pvio_b200_handle hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    if (h->pnp_dev) cudaFree(h->pnp_dev);
    if (h->pnp_host) cudaFreeHost(h->pnp_host);
    klt_free(h);
    release(h->hdr); release(h->cst); release(h->obs); release(h->lms); release(h->rho); release(h->frames);
    release(h->ctrl); release(h->rho_cand); release(h->frames_cand); release(h->lm_scale); release(h->lm_aux); release(h->hs);
    release(h->dx_lm); release(h->dx_pose); release(h->pose_scale); release(h->v_pose); release(h->Hred);
    release(h->acc); release(h->aux_cost);
    release(h->Hfull); release(h->gfull);
    release(h->imu_idx); release(h->imu_data); release(h->prior_frames); release(h->prior_S); release(h->prior_L);
    release(h->prior_e); release(h->prior_x0);
    release(h->plane_param); release(h->pt_plane); release(h->pt_begin); release(h->pt_frame); release(h->pt_z);
    for (auto &kv : h->graphs) cudaGraphExecDestroy(kv.second.first);
    for (auto &e : h->kev) cudaEventDestroy(e);
    cudaEventDestroy(h->ev0); cudaEventDestroy(h->ev1); cudaEventDestroy(h->evk0); cudaEventDestroy(h->evk1);
    for (auto &e : h->ev_up) cudaEventDestroy(e);
    for (auto &e : h->ev_done) cudaEventDestroy(e);
    for (auto &e : h->ev_down) cudaEventDestroy(e);
    if (h->stream_up) cudaStreamDestroy(h->stream_up);
    if (h->stream_down) cudaStreamDestroy(h->stream_down);
    cudaStreamDestroy(h->stream);
    delete h;
}

const char *pvio_b200_last_error(pvio_b200_handle hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    return h ? h->err.c_str() : "null handle";
}

int64_t pvio_b200_kernel_launches(pvio_b200_handle hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    return h ? h->launches : 0;
}

int pvio_b200_sync(pvio_b200_handle hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int pvio_b200_timer_start(pvio_b200_handle hh) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    CK(h, cudaEventRecord(h->ev0, h->stream));
    return 0;
}

int pvio_b200_timer_stop(pvio_b200_handle hh, float *ms) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    CK(h, cudaEventRecord(h->ev1, h->stream));
    CK(h, cudaEventSynchronize(h->ev1));
    CK(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return 0;
}

// which = 0: duration of the most recent linearise+Schur launch; which = 1: MEAN duration over the
// launches since the last reset (at most the latest 512); which = -1: reset the accumulation.
int pvio_b200_last_kernel_ms(pvio_b200_handle hh, int which, float *ms) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (which < 0) { h->kev_count = 0; if (ms) *ms = 0.f; return 0; }
    if (h->kev_count == 0) { *ms = 0.f; return 0; }     // graph-replayed launches are not individually timed
    CK(h, cudaStreamSynchronize(h->stream));
    const int n = which == 0 ? 1 : std::min(h->kev_count, 512);
    double tot = 0.0;
    for (int i = 0; i < n; ++i) {
        const int slot = ((h->kev_count - 1 - i) % 512) * 2;
        float t = 0.f;
        CK(h, cudaEventElapsedTime(&t, h->kev[slot], h->kev[slot + 1]));
        tot += t;
    }
    *ms = (float)(tot / n);
    return 0;
}

int pvio_b200_batch_set_window(pvio_b200_handle hh, int slot, const pvio_b200_window *w, const pvio_b200_state *s) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !w || !s) return PVIO_B200_EINVAL;
    return pack_window(h, slot, w, s);
}

int pvio_b200_batch_replicate(pvio_b200_handle hh, int n) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (n < 1 || n > h->W) return fail(h, PVIO_B200_EINVAL, "bad window count");
    const size_t N = h->Ncap;
    for (int i = 1; i < n; ++i) {
        h->hdr.h[i] = h->hdr.h[0]; h->cst.h[i] = h->cst.h[0];
        memcpy(h->obs.h + (size_t)i * h->Kcap, h->obs.h, sizeof(ObsRec) * h->slot_K[0]);
        memcpy(h->lms.h + (size_t)i * h->Mcap, h->lms.h, sizeof(LmRec) * h->slot_M[0]);
        memcpy(h->rho.h + (size_t)i * h->Mcap, h->rho.h, sizeof(double) * h->slot_M[0]);
        memcpy(h->frames.h + (size_t)i * N * kFrameStride, h->frames.h, sizeof(double) * N * kFrameStride);
        h->perm[i] = h->perm[0]; h->perm_identity[i] = h->perm_identity[0]; h->slot_M[i] = h->slot_M[0]; h->slot_N[i] = h->slot_N[0]; h->slot_K[i] = h->slot_K[0];
        if (h->have_inertial) {
            const size_t dcap = 15 * N;
            memcpy(h->imu_idx.h + (size_t)i * N * 2, h->imu_idx.h, sizeof(int32_t) * N * 2);
            memcpy(h->imu_data.h + (size_t)i * N * kImuStride, h->imu_data.h, sizeof(double) * N * kImuStride);
            memcpy(h->prior_frames.h + (size_t)i * N, h->prior_frames.h, sizeof(int32_t) * N);
            memcpy(h->prior_S.h + (size_t)i * dcap * dcap, h->prior_S.h, sizeof(double) * dcap * dcap);
            memcpy(h->prior_e.h + (size_t)i * dcap, h->prior_e.h, sizeof(double) * dcap);
            memcpy(h->prior_x0.h + (size_t)i * N * kFrameStride, h->prior_x0.h, sizeof(double) * N * kFrameStride);
        }
        if (h->have_planes) {
            memcpy(h->plane_param.h + (size_t)i * h->Pcap * 4, h->plane_param.h, sizeof(double) * h->Pcap * 4);
            memcpy(h->pt_plane.h + (size_t)i * h->Tcap, h->pt_plane.h, sizeof(int32_t) * h->Tcap);
            memcpy(h->pt_begin.h + (size_t)i * (h->Tcap + 1), h->pt_begin.h, sizeof(int32_t) * (h->Tcap + 1));
            memcpy(h->pt_frame.h + (size_t)i * h->Ocap, h->pt_frame.h, sizeof(int32_t) * h->Ocap);
            memcpy(h->pt_z.h + (size_t)i * h->Ocap * 2, h->pt_z.h, sizeof(float) * h->Ocap * 2);
        }
    }
    return 0;
}

int pvio_b200_batch_upload(pvio_b200_handle hh, int n) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    return upload(h, n);
}

int pvio_b200_batch_gn_step(pvio_b200_handle hh, int n, double mu, int apply) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (n < 1 || n > h->n_uploaded) return fail(h, PVIO_B200_EINVAL, "windows not uploaded");
    StepCfg c;
    c.mu = mu; c.apply = apply; c.compute_scale = 1;
    return run_step(h, n, c);
}

int pvio_b200_batch_download(pvio_b200_handle hh, int n, double *dx, int64_t dx_stride, double *costs) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    return download_dx(h, n, dx, dx_stride, costs);
}

// End-to-end step with HOST buffers.  Large batches are cut into sub-batches and pipelined over
// three streams: host->device copy of sub-batch i+1 overlaps the kernels of sub-batch i and the
// device->host copy of sub-batch i-1 (PCIe is full duplex), so the call costs
// max(copy, compute) instead of their sum.
int pvio_b200_batch_gn_step_host(pvio_b200_handle hh, int n, double mu, double *dx, int64_t dx_stride, double *costs) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (n < 1 || n > h->W) return fail(h, PVIO_B200_EINVAL, "bad window count");
    // sub-batch schedule: small batches first and last (the first upload and the last kernels + download are the
    // only parts of the pipeline that nothing overlaps), 512-window batches in between
    static const int sub_env = getenv("PVIO_B200_SUB") ? atoi(getenv("PVIO_B200_SUB")) : 0;     // uniform size (experiments)
    std::vector<int> sizes;
    if (n < 1024) sizes.push_back(n);
    else if (sub_env > 0) { for (int r = n; r > 0; r -= sub_env) sizes.push_back(std::min(sub_env, r)); }
    else {
        int mid = n - 2 * (128 + 256);
        sizes.push_back(128); sizes.push_back(256);
        while (mid > 0) { const int m = std::min(512, mid); sizes.push_back(m); mid -= m; }
        sizes.push_back(256); sizes.push_back(128);
    }
    const int nsub = (int)sizes.size();
    if (nsub == 1) {
        TRY(upload(h, n));
        StepCfg c;
        c.mu = mu; c.apply = 0; c.compute_scale = 1;
        TRY(run_step(h, n, c));
        return download_dx(h, n, dx, dx_stride, costs);
    }
    std::vector<int> starts(nsub, 0);
    for (int i = 1; i < nsub; ++i) starts[i] = starts[i - 1] + sizes[i - 1];
    if (!h->stream_up) {
        CK(h, cudaStreamCreateWithFlags(&h->stream_up, cudaStreamNonBlocking));
        CK(h, cudaStreamCreateWithFlags(&h->stream_down, cudaStreamNonBlocking));
    }
    while ((int)h->ev_up.size() < nsub) {
        cudaEvent_t a, b;
        CK(h, cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        h->ev_up.push_back(a); h->ev_done.push_back(b);
        cudaEvent_t c_;
        CK(h, cudaEventCreateWithFlags(&c_, cudaEventDisableTiming));
        h->ev_down.push_back(c_);
    }
    for (int i = 0; i < nsub; ++i) {
        const int w0 = starts[i], m = sizes[i];
        TRY(upload_range(h, w0, m, h->stream_up));
        CK(h, cudaEventRecord(h->ev_up[i], h->stream_up));
        CK(h, cudaStreamWaitEvent(h->stream, h->ev_up[i], 0));
        StepCfg c;
        c.mu = mu; c.apply = 0; c.compute_scale = 1; c.w0 = w0;
        TRY(run_step(h, m, c));
        CK(h, cudaEventRecord(h->ev_done[i], h->stream));
        CK(h, cudaStreamWaitEvent(h->stream_down, h->ev_done[i], 0));
        CK(h, cudaMemcpyAsync(h->dx_pose.h + (size_t)w0 * h->Ncap * 15, h->dx_pose.d + (size_t)w0 * h->Ncap * 15,
                              sizeof(double) * h->Ncap * 15 * m, cudaMemcpyDeviceToHost, h->stream_down));
        CK(h, cudaMemcpyAsync(h->dx_lm.h + (size_t)w0 * h->Mcap, h->dx_lm.d + (size_t)w0 * h->Mcap,
                              sizeof(double) * h->Mcap * m, cudaMemcpyDeviceToHost, h->stream_down));
        CK(h, cudaMemcpyAsync(h->ctrl.h + w0, h->ctrl.d + w0, sizeof(WinCtrl) * m, cudaMemcpyDeviceToHost, h->stream_down));
        CK(h, cudaEventRecord(h->ev_down[i], h->stream_down));
    }
    h->n_uploaded = n;
    // scatter each sub-batch as soon as it has landed, while the GPU works on the later ones
    for (int i = 0; i < nsub; ++i) {
        const int w0 = starts[i], m = sizes[i];
        CK(h, cudaEventSynchronize(h->ev_down[i]));
        TRY(scatter_dx(h, w0, m, dx, dx_stride, costs));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    return 0;
}

int pvio_b200_ba_gn_step(pvio_b200_handle hh, const pvio_b200_window *w, const pvio_b200_state *s, double mu,
                         double *dx, double *cost, double *new_cost, double *Hred, double *gred) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !w || !s) return PVIO_B200_EINVAL;
    TRY(pack_window(h, 0, w, s));
    TRY(upload(h, 1));
    StepCfg c;
    c.mu = mu; c.apply = 0; c.compute_scale = 1; c.dump = (Hred || gred) ? 1 : 0;
    if (c.dump) {
        CK(h, cudaMemsetAsync(h->Hfull.d, 0, sizeof(double) * h->Hfull.n, h->stream));
        CK(h, cudaMemsetAsync(h->gfull.d, 0, sizeof(double) * h->gfull.n, h->stream));
    }
    TRY(run_step(h, 1, c));
    double costs[2];
    const int N = w->n_frames, M = w->n_landmarks;
    TRY(download_dx(h, 1, dx, (int64_t)N * 15 + M, costs));
    if (cost) *cost = costs[0];
    if (new_cost) *new_cost = costs[1];
    if (c.dump) {
        const size_t Df = 15 * (size_t)h->Ncap, Dn = 15 * (size_t)N;
        CK(h, cudaMemcpyAsync(h->Hfull.h, h->Hfull.d, sizeof(double) * Df * Df, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaMemcpyAsync(h->gfull.h, h->gfull.d, sizeof(double) * Df, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        if (Hred) for (size_t i = 0; i < Dn; ++i) memcpy(Hred + i * Dn, h->Hfull.h + i * Df, sizeof(double) * Dn);
        if (gred) memcpy(gred, h->gfull.h, sizeof(double) * Dn);
    }
    return 0;
}

// Trust-region loop: the minimiser logic of ceres::Solve as PVIO configures it
// (solver_options.h:26-33; TrustRegionMinimizer + dogleg defaults of Ceres 1.14), driven from
// the host with one small device->host read per iteration.  When the Gauss-Newton step leaves the
// trust region the Cauchy point is computed with one extra J.v sweep and the traditional dogleg
// interpolation is applied, as dogleg_strategy.cc does.
int pvio_b200_ba_solve(pvio_b200_handle hh, const pvio_b200_window *w, pvio_b200_state *s,
                       const pvio_b200_options *opt, pvio_b200_summary *summary, uint8_t *valid, double *quality) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !w || !s) return PVIO_B200_EINVAL;
    const int max_iter = opt ? opt->max_iterations : 10;
    const int alias = opt ? opt->alias_bias : 1;
    const double max_time = (opt && opt->max_time > 0) ? opt->max_time : 1e6;
    const auto t_begin = std::chrono::steady_clock::now();
    TRY(pack_window(h, 0, w, s));
    TRY(upload(h, 1));
    CK(h, cudaEventRecord(h->ev0, h->stream));
    double radius = (opt && opt->initial_trust_region_radius > 0) ? opt->initial_trust_region_radius : 1e4, mu = 1e-8;
    bool reuse = false;
    StepCfg c;
    c.apply = 0; c.alias_bias = alias; c.compute_scale = 1;
    pvio_b200_summary sm;
    memset(&sm, 0, sizeof(sm));
    sm.termination = PVIO_B200_TERM_NO_CONVERGENCE; sm.usable = 1;
    auto read_ctrl = [&](WinCtrl &o, double *acc) -> int {
        CK(h, cudaMemcpyAsync(h->ctrl.h, h->ctrl.d, sizeof(WinCtrl), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaMemcpyAsync(h->acc.h, h->acc.d, sizeof(double) * kAcc, cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        o = h->ctrl.h[0];
        memcpy(acc, h->acc.h, sizeof(double) * kAcc);
        return 0;
    };
    WinCtrl ct;
    double acc[kAcc];
    int it = 0;
    bool first = true;
    double cost = 0.0;
    while (it < max_iter) {
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() > max_time) break;
        ++it;
        c.mu = mu; c.beta = 1.0; c.skip_linearize = reuse;
        TRY(run_step(h, 1, c));
        TRY(read_ctrl(ct, acc));
        if (first) {
            sm.initial_cost = ct.cost;
            first = false;
            if (ct.gmax <= 1e-10) { sm.termination = PVIO_B200_TERM_CONVERGENCE; --it; cost = ct.cost; break; }
        }
        c.compute_scale = 0;
        cost = ct.cost;
        if (ct.solve_failed) {                 // dogleg_strategy.cc: retry the GN solve with a larger mu
            mu *= 10.0;
            reuse = false;
            if (mu > 1.0) { sm.termination = PVIO_B200_TERM_FAILURE; sm.usable = 0; break; }
            --it;
            continue;
        }
        // ---- DoglegStrategy::ComputeStep (TRADITIONAL_DOGLEG), all norms in the diag-scaled space
        const double gn_norm = std::sqrt(ct.gn_norm2 + acc[3]);
        const double gdx = ct.g_dot_dx + acc[1];                 // g . dx_gn  (= grad . gn in scaled space)
        const double rdx = ct.dx_reg_dx + acc[2];                // dx_gn^T (mu D) dx_gn
        double sa = 0.0, sb = 1.0, step_norm = gn_norm;
        double grad2 = 0.0, v_rd = 0.0, jv2 = 0.0;
        if (gn_norm > radius) {
            // Cauchy point: alpha = |grad|^2 / |J S D^-1 grad|^2 needs one J.v sweep
            TRY(run_jv(h, c));
            double acc2[kAcc];
            WinCtrl ct2;
            TRY(read_ctrl(ct2, acc2));
            grad2 = ct.grad2 + acc[7];
            v_rd = ct.v_reg_dx + acc[8];
            jv2 = acc2[9] + acc2[10];
            const double g_norm = std::sqrt(grad2);
            const double alpha = grad2 / jv2;
            if (g_norm * alpha >= radius) {                       // scaled steepest descent to the boundary
                sa = radius / g_norm; sb = 0.0;
            } else {                                              // dogleg interpolation (dogleg_strategy.cc)
                const double b_dot_a = -alpha * gdx;
                const double a2 = (alpha * g_norm) * (alpha * g_norm);
                const double bma2 = a2 - 2.0 * b_dot_a + gn_norm * gn_norm;
                const double cc = b_dot_a - a2;
                const double dd = std::sqrt(cc * cc + bma2 * (radius * radius - a2));
                const double beta = cc <= 0 ? (dd - cc) / bma2 : (radius * radius - a2) / (dd + cc);
                sa = alpha * (1.0 - beta); sb = beta;
            }
            step_norm = radius;
            c.step_a = sa; c.step_b = sb; c.skip_linearize = true;
            TRY(run_step(h, 1, c));
            TRY(read_ctrl(ct, acc));
            c.step_a = 0.0; c.step_b = -1.0;
        }
        // model cost change -(g.s + s^T H s / 2) of s = sb dx_gn - sa v, using H dx_gn = -g - (mu D) dx_gn
        const double dHd = -gdx - rdx, vHd = -grad2 - v_rd;
        const double sHs = sb * sb * dHd - 2.0 * sa * sb * vHd + sa * sa * jv2;
        const double model_change = -(sb * gdx - sa * grad2) - 0.5 * sHs;
        const double beta = sb;
        if (!(model_change > 0.0)) { radius *= 0.5; reuse = true; continue; }       // invalid step
        const double x_norm = std::sqrt(ct.xnorm2 + acc[5]);
        const double step_amb = std::sqrt(acc[6] + acc[4]);
        if (step_amb <= 1e-8 * (x_norm + 1e-8)) { sm.termination = PVIO_B200_TERM_CONVERGENCE; break; }
        if (std::fabs(cost - ct.cand_cost) <= 1e-6 * cost) { sm.termination = PVIO_B200_TERM_CONVERGENCE; break; }
        const double rel = (cost - ct.cand_cost) / model_change;
        if (rel > 1e-3) {
            // accept: candidate becomes the state
            finalize_kernel<<<1, 128, 0, h->stream>>>(h->ctrl.d, h->acc.d, h->aux_cost.d, 1, h->frames.d, h->frames_cand.d,
                                                     h->rho.d, h->rho_cand.d, h->hdr.d, h->Ncap, h->Mcap, beta, 0);
            ++h->launches;
            ++sm.accepted_steps;
            cost = ct.cand_cost;
            if (rel < 0.25) radius *= 0.5;
            if (rel > 0.75) radius = std::max(radius, 3.0 * step_norm);
            mu = std::max(1e-8, 2.0 * mu / 10.0);
            reuse = false;
            // gradient tolerance at the new point is checked after the next linearisation
            if (it < max_iter) {
                StepCfg g = c;
                g.mu = mu; g.beta = 1.0; g.skip_linearize = false;
                // peek: linearise + solve only, to read |g|_inf and the re-evaluated cost (ceres re-evaluates
                // the cost at the accepted point together with the Jacobian)
                TRY(run_step(h, 1, g, 1));
                TRY(read_ctrl(ct, acc));
                cost = ct.cost;
                reuse = true;                // the next iteration reuses this linearisation
                if (ct.gmax <= 1e-10) { sm.termination = PVIO_B200_TERM_CONVERGENCE; break; }
            }
        } else {
            radius *= 0.5;
            reuse = true;
        }
        if (radius <= 1e-32) { sm.termination = PVIO_B200_TERM_CONVERGENCE; break; }
    }
    CK(h, cudaEventRecord(h->ev1, h->stream));
    sm.iterations = it;
    sm.final_cost = cost;
    sm.final_radius = radius; sm.final_mu = mu;
    // read back the state
    const int N = w->n_frames, M = w->n_landmarks;
    CK(h, cudaMemcpyAsync(h->frames.h, h->frames.d, sizeof(double) * N * kFrameStride, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->rho.h, h->rho.d, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
    uint8_t *d_valid = nullptr;
    double *d_quality = nullptr;
    if ((!opt || opt->run_postpass) && (valid || quality)) {
        CK(h, cudaMalloc(&d_valid, M > 0 ? M : 1));
        CK(h, cudaMalloc(&d_quality, sizeof(double) * (M > 0 ? M : 1)));
        postpass_kernel<<<dim3(8, 1), 256, 0, h->stream>>>(h->hdr.d, h->cst.d, h->obs.d, h->lms.d, h->rho.d, h->frames.d,
                                                          d_valid, d_quality, nullptr, h->Ncap, h->Mcap, h->Kcap);
        ++h->launches;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    sm.solve_seconds = ms * 1e-3;
    memcpy(s->frames, h->frames.h, sizeof(double) * N * kFrameStride);
    const std::vector<int32_t> &perm = h->perm[0];
    for (int lp = 0; lp < M; ++lp) s->inv_depth[perm[lp]] = h->rho.h[lp];
    if (d_valid) {
        std::vector<uint8_t> hv(M);
        std::vector<double> hq(M);
        CK(h, cudaMemcpy(hv.data(), d_valid, M, cudaMemcpyDeviceToHost));
        CK(h, cudaMemcpy(hq.data(), d_quality, sizeof(double) * M, cudaMemcpyDeviceToHost));
        for (int lp = 0; lp < M; ++lp) {
            if (valid) valid[perm[lp]] = hv[lp];
            if (quality) quality[perm[lp]] = hq[lp];
        }
        cudaFree(d_valid); cudaFree(d_quality);
    }
    if (summary) *summary = sm;
    return 0;
}

int pvio_b200_reprojection_error(pvio_b200_handle hh, const pvio_b200_window *w, const pvio_b200_state *s, double *error) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !w || !s || !error) return PVIO_B200_EINVAL;
    TRY(pack_window(h, 0, w, s));
    TRY(upload(h, 1));
    CK(h, cudaMemsetAsync(h->acc.d, 0, sizeof(double) * 8, h->stream));
    postpass_kernel<<<dim3(8, 1), 256, 0, h->stream>>>(h->hdr.d, h->cst.d, h->obs.d, h->lms.d, h->rho.d, h->frames.d,
                                                      nullptr, nullptr, h->acc.d, h->Ncap, h->Mcap, h->Kcap);
    ++h->launches;
    CK(h, cudaMemcpyAsync(h->acc.h, h->acc.d, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    *error = h->acc.h[0] / std::max(h->acc.h[1], 1.0);
    return 0;
}

int pvio_b200_ba_marginalize(pvio_b200_handle hh, const pvio_b200_window *w, const pvio_b200_state *s, int index,
                             double *S_out, double *e_out, double *H_out, double *b_out) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !w || !s) return PVIO_B200_EINVAL;
    return marginalize_impl(h, w, s, index, S_out, e_out, H_out, b_out);
}

int pvio_b200_pnp_solve(pvio_b200_handle hh, const pvio_b200_pnp_problem *problem, double *frame,
                        const pvio_b200_options *opt, pvio_b200_summary *summary) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !problem || !frame) return PVIO_B200_EINVAL;
    return pnp_solve_impl(h, problem, frame, opt, summary);
}

int pvio_b200_klt_track(pvio_b200_handle hh, const uint8_t *prev, const uint8_t *next, int width, int height, int stride,
                        const float *prev_pts, float *next_pts, uint8_t *status, float *err, int n_points,
                        int max_level, int max_iter, double eps) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !prev || !next || !prev_pts || !next_pts || !status) return PVIO_B200_EINVAL;
    return klt_track_impl(h, prev, next, width, height, stride, prev_pts, next_pts, status, err, n_points, max_level,
                          max_iter, eps);
}

int pvio_b200_klt_track_raw(pvio_b200_handle hh, const uint8_t *prev, const uint8_t *next, int width, int height, int stride,
                            const float *prev_pts, float *next_pts, uint8_t *status, float *err, int n_points,
                            int max_level, int max_iter, double eps, double clahe_clip, int tiles_x, int tiles_y,
                            uint8_t *prev_eq, uint8_t *next_eq) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !prev || !next || !prev_pts || !next_pts || !status) return PVIO_B200_EINVAL;
    if (!(clahe_clip > 0.0)) return fail(h, PVIO_B200_EINVAL, "klt_track_raw: clahe_clip must be positive");
    return klt_track_impl(h, prev, next, width, height, stride, prev_pts, next_pts, status, err, n_points, max_level,
                          max_iter, eps, clahe_clip, tiles_x, tiles_y, prev_eq, next_eq);
}

int pvio_b200_clahe(pvio_b200_handle hh, const uint8_t *src, int width, int height, int stride, double clip_limit,
                    int tiles_x, int tiles_y, uint8_t *dst) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !src || !dst) return PVIO_B200_EINVAL;
    return clahe_impl(h, src, width, height, stride, clip_limit, tiles_x, tiles_y, dst);
}

}  // extern "C"
