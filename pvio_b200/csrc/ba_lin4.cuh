// Linearise stage, first half of the split pipeline (second half: schur_kernel, ba_schur.cuh).
// Phase A of lin_tpl_kernel (ba_lin2.cuh) as its own kernel: thread per landmark, per-frame transpose-reduction
// of the direct terms, records (sqrt(w) h per frame, sqrt(w) g_l, frame mask: LinArgs::hs_out) staged in shared
// memory and handed over with one bulk copy per warp.  Without the Schur tiles and their fp64 accumulator the
// CTA shape is free: kWarps x 32 threads, kMinBlocks CTAs per SM (template parameters, chosen by measurement).
#pragma once
#include "ba_lin2.cuh"

namespace pvio {

constexpr int kDta = 28;                        // doubles per (target, anchor) block: 21 sym + 6 grad + 1 pad

template <int kWarps>
__host__ __device__ inline size_t lin4_smem_bytes(int N) {
    const size_t nsp = (size_t)N * (N - 1) / 2;
    return sizeof(FrameSm) * N + sizeof(double) * ((nsp + 1) * kDta + 8) + 16 + sizeof(float) * (size_t)kWarps * 32 * (6 * N + 2);
}

template <bool kLoss, int kWarps, int kMinBlocks>
__global__ void __launch_bounds__(kWarps * 32, kMinBlocks)
lin_a_kernel(LinArgs a) {
    constexpr int kThreads = kWarps * 32;
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int nsp = N * (N - 1) / 2;
    const int R = hs_rec(N);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameSm *F = reinterpret_cast<FrameSm *>(smem_raw);
    double *Dta = reinterpret_cast<double *>(F + N);                // [nsp + 1][kDta] direct blocks (sized by the window: 5 CTAs per SM)
    double *cost_sm = Dta + (nsp + 1) * kDta;                       // [8]
    float *hbuf = reinterpret_cast<float *>(smem_raw + ((sizeof(FrameSm) * N + sizeof(double) * ((nsp + 1) * kDta + 8) + 15) & ~(size_t)15));

    if (tid < N) make_frame(a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride, wc, F[tid]);
    for (int i = tid; i < (nsp + 1) * kDta + 8; i += kThreads) Dta[i] = 0.0;
    __syncthreads();

    const float W[4] = {(float)wc.sic[0], (float)wc.sic[1], (float)wc.sic[2], (float)wc.sic[3]};
    const float cb = (float)(wc.cauchy_a * wc.cauchy_a);
    const double mu = a.mu_override >= 0.0 ? a.mu_override : a.ctrl[w].mu;
    const ObsRec *obs = a.obs + (size_t)w * a.Kcap;
    const LmRec *lms = a.lms + (size_t)w * a.Mcap;
    const double *rho = a.rho + (size_t)w * a.Mcap;
    double *lm_scale = a.lm_scale + (size_t)w * a.Mcap;
    LmAux *aux = a.lm_aux + (size_t)w * a.Mcap;
    const unsigned fixed = (unsigned)H.fixed_mask & ((1u << N) - 1u);
    float *hb = hbuf + (size_t)tid * R;                             // this thread's record

    float cost_acc = 0.f;
    for (int ch = blockIdx.x * kWarps + wv; ch < H.n_chunks; ch += gridDim.x * kWarps) {      // warps are independent: no CTA barrier
        const int lm0 = H.chunk_begin[ch];
        const int cnt = H.chunk_meta[ch] & 0xff;
        const int anchor = H.chunk_meta[ch] >> 8;
        const bool lm_ok = lane < cnt;
        const int l = lm0 + (lm_ok ? lane : 0);
        const LmRec lr = lms[l];
        const int n_obs = lm_ok ? lm_nobs(lr.meta) : 0;
        unsigned fm = lm_mask(lr.meta);
        const int n_max = __reduce_max_sync(0xffffffffu, n_obs);
        const double rl = lm_ok ? rho[l] : 1.0;
        double x[3];
        float xf[3], cl[3];
        world_point(F[anchor], lr.zrx, lr.zry, rl, x, xf, cl);
        double hll = 0.0, gl = 0.0;
        float ha[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int tmask = 0;
        // the previous chunk's bulk copy of this warp's records must have read the buffer before it is rewritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        for (int j = 0; j < n_max; ++j) {
            const bool act = j < n_obs;
            ObsRec o;
            o.zx = 0.f; o.zy = 0.f;
            if (act) o = obs[lr.obs_begin + j];
            const int t = act ? __ffs(fm) - 1 : 0;                  // j-th set bit of the frame mask
            fm &= fm - 1;
            float q[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) q[i] = 0.f;
            if (act) {
                ObsLin ol;
                linearize_obs<kLoss>(F[t], x, xf, cl, o.zx, o.zy, W, cb, ol);
                hll += (double)(ol.j0 * ol.j0 + ol.j1 * ol.j1);
                gl += (double)(ol.j0 * ol.r0 + ol.j1 * ol.r1);
                cost_acc += ol.cost;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const float h = ol.j0 * ol.Y0[i] + ol.j1 * ol.Y1[i];
                    ha[i] -= h;
                    hb[t * 6 + i] = h;                               // unscaled; scaled by sqrt(w) below
                }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int jj = i; jj < 6; ++jj) q[sym6(i, jj)] = ol.Y0[i] * ol.Y0[jj] + ol.Y1[i] * ol.Y1[jj];
#pragma unroll
                for (int i = 0; i < 6; ++i) q[21 + i] = ol.Y0[i] * ol.r0 + ol.Y1[i] * ol.r1;
                tmask |= 1 << t;
            }
            // reduce the direct terms of this iteration over the landmarks of the warp, per target frame
            unsigned todo = __ballot_sync(0xffffffffu, act);
            while (todo) {
                const int leader = __ffs(todo) - 1;
                const int tf = __shfl_sync(0xffffffffu, t, leader);
                const unsigned peers = __ballot_sync(0xffffffffu, act && t == tf);
                todo &= ~peers;
                if ((fixed >> tf) & (fixed >> anchor) & 1u) continue;          // both blocks constant: nothing to assemble
                const bool mine = act && t == tf;                              // q is zero for inactive lanes
                const float tot = (peers == (todo | peers)) ? transpose_reduce32(q, true, lane) : transpose_reduce32(q, mine, lane);
                if (lane < kDirVals && tf != anchor) {
                    const int sp = tf > anchor ? spair(tf, anchor) : spair(anchor, tf);
                    atomicAdd(&Dta[sp * kDta + lane], (double)tot);              // +sum Y^T Y; the epilogue applies the signs
                }
            }
        }
        // per-landmark Schur scalars; scale the record
        if (n_obs > 0) {
            double sc;
            if (a.compute_scale) { sc = 1.0 / (1.0 + sqrt(hll)); lm_scale[l] = sc; }
            else sc = lm_scale[l];
            const double hreg = hll + (mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0);
            const double wl = 1.0 / hreg;
            const bool finite = isfinite(wl);                       // bundle_adjustor.cpp:538 skip
            const float sw = finite ? sqrtf((float)wl) : 0.f;
            aux[l].hll_reg = hreg; aux[l].gl = gl; aux[l].hll = hll;
            int m = tmask;
            while (m) {
                const int t = __ffs(m) - 1;
                m &= m - 1;
#pragma unroll
                for (int i = 0; i < 6; ++i) hb[t * 6 + i] *= sw;
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) hb[anchor * 6 + i] = ha[i] * sw;
            hb[6 * N] = sw * (float)gl;
            hb[6 * N + 1] = __int_as_float(finite ? (tmask | (1 << anchor)) : 0);
        } else {
            hb[6 * N] = 0.f;
            hb[6 * N + 1] = __int_as_float(0);
            if (lm_ok) { aux[l].hll_reg = 1.0; aux[l].gl = 0.0; aux[l].hll = 0.0; }
        }
        // the warp's 32 records leave as ONE bulk copy (TMA engine): shared -> global
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            float *dst = a.hs_out + (size_t)w * a.hs_stride + (size_t)ch * 32 * R;
            const float *src = hbuf + (size_t)wv * 32 * R;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(dst), "r"((uint32_t)__cvta_generic_to_shared(src)), "r"(32 * R * 4) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // before the buffer is freed

    // ---- direct part of the reduced system (schur_kernel subtracts the Schur sum from Hred / gred)
    __syncthreads();
    const bool exclusive = (gridDim.x == 1);
    const int npairs_cap = a.Ncap * (a.Ncap + 1) / 2;
    double *Hred_o = a.Hred + (size_t)w * npairs_cap * 36;
    double *Hdd_o = a.Hdd + (size_t)w * a.Ncap * 36;
    double *gdir_o = a.gdir + (size_t)w * a.Ncap * 6;
    double *gred_o = a.gred + (size_t)w * a.Ncap * 6;
    for (int e = tid; e < N * 36; e += kThreads) {                  // diagonal blocks: sum of the pair blocks touching f
        const int f = e / 36, ij = e - f * 36, i = ij / 6, j = ij - i * 6;
        const int se = i <= j ? sym6(i, j) : sym6(j, i);
        double d = 0.0;
        for (int g = 0; g < N; ++g) {
            if (g == f) continue;
            d += Dta[(f > g ? spair(f, g) : spair(g, f)) * kDta + se];
        }
        if (exclusive) { Hdd_o[e] = d; Hred_o[pair_idx(f, f) * 36 + ij] = d; }
        else if (d != 0.0) { atomicAdd(&Hdd_o[e], d); atomicAdd(&Hred_o[pair_idx(f, f) * 36 + ij], d); }
    }
    for (int e = tid; e < nsp * 36; e += kThreads) {                // off-diagonal blocks (f > g): -D(f, g)
        const int sp = e / 36, ij = e - sp * 36, i = ij / 6, j = ij - i * 6;
        int f = 1;
        while ((f + 1) * f / 2 <= sp) ++f;
        const int g = sp - f * (f - 1) / 2;
        const int se = i <= j ? sym6(i, j) : sym6(j, i);
        const double v = -Dta[sp * kDta + se];
        if (exclusive) Hred_o[pair_idx(f, g) * 36 + ij] = v; else if (v != 0.0) atomicAdd(&Hred_o[pair_idx(f, g) * 36 + ij], v);
    }
    for (int e = tid; e < N * 6; e += kThreads) {                   // gradients (f is the target when f > g)
        const int f = e / 6, i = e - f * 6;
        double d = 0.0;
        for (int g = 0; g < N; ++g) {
            if (g == f) continue;
            const double v = Dta[(f > g ? spair(f, g) : spair(g, f)) * kDta + 21 + i];
            d += (f > g) ? v : -v;
        }
        if (exclusive) { gdir_o[e] = d; gred_o[e] = d; }
        else if (d != 0.0) { atomicAdd(&gdir_o[e], d); atomicAdd(&gred_o[e], d); }
    }
    double cd = (double)cost_acc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cd += __shfl_xor_sync(0xffffffffu, cd, off);
    if (lane == 0) atomicAdd(&cost_sm[0], cd);
    __syncthreads();
    if (tid == 0) {
        if (exclusive) a.cost_vis[w] = cost_sm[0]; else atomicAdd(&a.cost_vis[w], cost_sm[0]);
    }
}

}  // namespace pvio
