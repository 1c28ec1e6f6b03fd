// Reprojection sweep, second generation: THREAD PER LANDMARK.
//
// Same mathematics and outputs as ba_lin.cuh (xi-coordinate formulation, see there), different
// work decomposition, chosen after profiling the first version (profiles/r01_lin_schur_n592.md:
// 318 K warp-instructions per window, 18 of 32 lanes useful in the per-landmark groups):
//
//   Phase A  one WARP per chunk (<= 32 landmarks with a common anchor), one LANE per landmark.
//            The lane walks its landmark's observations; H_ll, g_l and h_anchor are plain
//            thread-local sums (no shuffles).  All lanes of the warp visit the same target
//            frame in the same iteration (tracks are contiguous runs of frames after the
//            anchor), so the per-frame direct terms  sum_l Y^T Y, Y^T r  are reduced across the
//            32 landmarks with ONE 31-shuffle transpose-reduction per iteration; lanes that see a
//            different frame fall back to a per-frame loop (correct for any visibility pattern).
//   Phase B  output-stationary threads own 6x3 tiles of  sum_l w_l h_l h_l^T  over the up to 256
//            landmarks the 8 warps just produced (scaled h vectors in shared memory), fp32
//            accumulation flushed to fp64 every 16 landmarks.
// Atomics: only fp64 shared-memory adds of the (target, anchor) direct blocks (27 per warp
// iteration) and, when a window is split over several CTAs, the final global accumulation.
#pragma once
#include "ba_lin.cuh"

namespace pvio {

constexpr int kSuper = 256;                 // landmarks per super-chunk (8 warps x 32 lanes)
constexpr int kDirVals = 27;                // 21 (Y^T Y sym) + 6 (Y^T r)

// Sum over the 32 lanes of (mine ? q[i] : 0) for i = 0..31; lane i returns the total of entry i.
// Transpose-reduction: 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.  q is not modified.
__device__ __forceinline__ float transpose_reduce32(const float (&q)[32], bool mine, int lane) {
    float v[16];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float lo = mine ? q[k] : 0.f, hi = mine ? q[k + 16] : 0.f;
            const float send = up ? lo : hi, keep = up ? hi : lo;
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const bool up = lane & 8;
        const float send = up ? v[k] : v[k + 8], keep = up ? v[k + 8] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool up = lane & 4;
        const float send = up ? v[k] : v[k + 4], keep = up ? v[k + 4] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const bool up = lane & 2;
        const float send = up ? v[k] : v[k + 2], keep = up ? v[k + 2] : v[k];
        v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    return v[0];
}

// strictly-lower pair index (t > a)
__device__ __forceinline__ int spair(int t, int a) { return t * (t - 1) / 2 + a; }

__host__ __device__ inline size_t lin2_smem_bytes(int N) {
    const size_t npairs = (size_t)N * (N + 1) / 2, nsp = (size_t)N * (N - 1) / 2;
    return sizeof(FrameSm) * kMaxFrames                 // camera poses
           + sizeof(double) * (npairs * 36)              // Schur sum (block lower triangular)
           + sizeof(double) * (nsp + 1) * 33             // direct (target, anchor) blocks + gradients
           + sizeof(double) * (size_t)N * 6              // Schur gradient correction
           + sizeof(float) * ((size_t)kSuper * (N * 6 + 2))   // records: scaled h vectors, sqrt(w) g_l, frame mask
           + sizeof(double) * 8 + 16;                    // + alignment slack of the h buffer
}

// Epilogue shared by the thread-per-landmark kernels: assemble the direct blocks, subtract the Schur sum.
// Hdir[t,t] += D(t,a), Hdir[a,a] += D(t,a), Hdir[t,a] = -D(t,a);  g[t] += d(t,a), g[a] -= d(t,a)
// (target/anchor Jacobians are +-Y).  The pair storage keeps D for the (max,min) frame pair, with
// the gradient sign relative to the TARGET; a target earlier than its anchor cannot occur in PVIO
// (the anchor is the lowest-id frame) and is rejected by the packer.
__device__ __forceinline__ void lin_epilogue(const LinArgs &a, int w, int N, int tid, const double *Ss, const double *Dta, const double *gsc) {
    const int nsp = N * (N - 1) / 2;
    const bool exclusive = (gridDim.x == 1);
    const int npairs_cap = a.Ncap * (a.Ncap + 1) / 2;
    double *Hred_o = a.Hred + (size_t)w * npairs_cap * 36;
    double *Hdd_o = a.Hdd + (size_t)w * a.Ncap * 36;
    double *gdir_o = a.gdir + (size_t)w * a.Ncap * 6;
    double *gred_o = a.gred + (size_t)w * a.Ncap * 6;
    // diagonal blocks (direct): thread per (f, i, j)
    for (int e = tid; e < N * 36; e += kLinThreads) {
        const int f = e / 36, ij = e - f * 36, i = ij / 6, j = ij - i * 6;
        const int se = i <= j ? sym6(i, j) : sym6(j, i);
        double d = 0.0;
        for (int g = 0; g < N; ++g) {
            if (g == f) continue;
            d += Dta[(f > g ? spair(f, g) : spair(g, f)) * 33 + se];
        }
        const double red = d - Ss[pair_idx(f, f) * 36 + ij];
        if (exclusive) { Hdd_o[e] = d; Hred_o[pair_idx(f, f) * 36 + ij] = red; }
        else { if (d != 0.0) atomicAdd(&Hdd_o[e], d); if (red != 0.0) atomicAdd(&Hred_o[pair_idx(f, f) * 36 + ij], red); }
    }
    // off-diagonal blocks (f > g): -D(f,g) - S(f,g)
    for (int e = tid; e < nsp * 36; e += kLinThreads) {
        const int sp = e / 36, ij = e - sp * 36, i = ij / 6, j = ij - i * 6;
        int f = 1;
        while ((f + 1) * f / 2 <= sp) ++f;
        const int g = sp - f * (f - 1) / 2;
        const int se = i <= j ? sym6(i, j) : sym6(j, i);
        const double v = -Dta[sp * 33 + se] - Ss[pair_idx(f, g) * 36 + ij];
        if (exclusive) Hred_o[pair_idx(f, g) * 36 + ij] = v; else if (v != 0.0) atomicAdd(&Hred_o[pair_idx(f, g) * 36 + ij], v);
    }
    // gradients: the landmarks are sorted by anchor, and a pair (t,a) only ever has t as the target
    for (int e = tid; e < N * 6; e += kLinThreads) {
        const int f = e / 6, i = e - f * 6;
        double d = 0.0;
        for (int g = 0; g < N; ++g) {
            if (g == f) continue;
            const double v = Dta[(f > g ? spair(f, g) : spair(g, f)) * 33 + 21 + i];
            d += (f > g) ? v : -v;          // f is the target when f > g, the anchor otherwise
        }
        const double r = d - gsc[e];
        if (exclusive) { gdir_o[e] = d; gred_o[e] = r; }
        else { if (d != 0.0) atomicAdd(&gdir_o[e], d); if (r != 0.0) atomicAdd(&gred_o[e], r); }
    }
}

template <bool kLoss>
__global__ void __launch_bounds__(kLinThreads, 2)
lin_tpl_kernel(LinArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int npairs = N * (N + 1) / 2, nsp = N * (N - 1) / 2;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameSm *F = reinterpret_cast<FrameSm *>(smem_raw);
    double *Ss = reinterpret_cast<double *>(F + kMaxFrames);        // [npairs][36] Schur sum
    double *Dta = Ss + npairs * 36;                                 // [nsp][33] direct blocks: 21 sym + 6 grad (+6 pad)
    double *gsc = Dta + (nsp + 1) * 33;                             // [N][6] sum_l w g_l h_f
    // [kSuper][N][6], 16-byte aligned (offset arithmetic keeps the shared address space): bulk-copy source
    float *hbuf = reinterpret_cast<float *>(smem_raw + ((sizeof(FrameSm) * kMaxFrames + sizeof(double) * (npairs * 36 + (nsp + 1) * 33 + N * 6) + 15) & ~(size_t)15));
    const int R = hs_rec(N);                                        // record stride (LinArgs::hs_out layout)
    double *cost_sm = reinterpret_cast<double *>(hbuf + kSuper * R);   // [8]  (kSuper * R floats is a multiple of 8 bytes)

    if (tid < N) make_frame(a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride, wc, F[tid]);
    for (int i = tid; i < npairs * 36 + (nsp + 1) * 33 + N * 6; i += kLinThreads) Ss[i] = 0.0;   // Ss, Dta, gsc contiguous
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

    // ---- Phase-B task of this thread: one 6x6 block pair x k-split (the k-split lanes are adjacent).
    // Pairs are enumerated over the FREE frames only: FF_FIX_POSE frames are constant parameter blocks
    // (bundle_adjustor.cpp:82-87), their rows of the reduced system are never read by the solve.
    const unsigned fixed = a.victim_only ? 0u : ((unsigned)H.fixed_mask & ((1u << N) - 1u));
    const unsigned freem = ~fixed & ((1u << N) - 1u);
    const int nfree = __popc(freem);
    const int ntask = nfree * (nfree + 1) / 2;
    const int ksplit = (ntask * 4 <= kLinThreads) ? 4 : ((ntask * 2 <= kLinThreads) ? 2 : 1);
    const int task = tid / ksplit, kk = tid % ksplit;
    const bool b_active = task < ntask;
    int bf = 0, bg = 0;
    if (ntask > 0) {
        const int p = min(task, ntask - 1);
        int f = 0;
        while ((f + 1) * (f + 2) / 2 <= p) ++f;
        const int g = p - f * (f + 1) / 2;
        bf = __fns(freem, 0, f + 1); bg = __fns(freem, 0, g + 1);        // f-th / g-th free frame
    }
    const bool b_diag = (bf == bg);

    float cost_acc = 0.f;

    for (int c0 = blockIdx.x * 8; c0 < H.n_chunks; c0 += gridDim.x * 8) {
        // ========================= Phase A: one warp per chunk, one lane per landmark =========================
        const int ch = c0 + wv;
        const bool ch_ok = ch < H.n_chunks;
        const int lm0 = ch_ok ? H.chunk_begin[ch] : 0;
        const int cnt = ch_ok ? (H.chunk_meta[ch] & 0xff) : 0;
        const int anchor = ch_ok ? (H.chunk_meta[ch] >> 8) : 0;
        const int slot = wv * 32 + lane;
        {
            const bool lm_ok = lane < cnt;
            const int l = lm0 + (lm_ok ? lane : 0);
            const LmRec lr = lms[l];
            int n_obs = lm_ok ? lm_nobs(lr.meta) : 0;
            if (a.victim_only && !lm_victim(lr.meta)) n_obs = 0;
            unsigned fm = lm_mask(lr.meta);
            const int n_max = __reduce_max_sync(0xffffffffu, n_obs);
            const double rl = lm_ok ? rho[l] : 1.0;
            double x[3];
            float xf[3], cl[3];
            world_point(F[anchor], lr.zrx, lr.zry, rl, x, xf, cl);
            double hll = 0.0, gl = 0.0;
            float ha[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            int tmask = 0;
            float *hb = hbuf + (size_t)slot * R;
            // the previous pass's bulk copy of this warp's records must have read the buffer before it is rewritten
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
            for (int j = 0; j < n_max; ++j) {
                const bool act = j < n_obs;
                ObsRec o;
                o.zx = 0.f; o.zy = 0.f;
                if (act) o = obs[lr.obs_begin + j];
                const int t = act ? __ffs(fm) - 1 : 0;          // j-th set bit of the frame mask
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
                        hb[t * 6 + i] = h;                       // unscaled; scaled by sqrt(w) below
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
                    if ((fixed >> tf) & (fixed >> anchor) & 1u) continue;      // both blocks constant: nothing to assemble
                    const bool mine = act && t == tf;      // q is zero for inactive lanes
                    const float tot = (peers == (todo | peers)) ? transpose_reduce32(q, true, lane) : transpose_reduce32(q, mine, lane);
                    if (lane < kDirVals && tf != anchor) {
                        const int sp = tf > anchor ? spair(tf, anchor) : spair(anchor, tf);
                        // sign convention: the block stores +sum Y^T Y; the epilogue applies the signs
                        atomicAdd(&Dta[sp * 33 + lane], (double)tot);
                    }
                }
            }
            // per-landmark Schur scalars (thread local)
            if (n_obs > 0) {
                double sc;
                if (a.compute_scale) { sc = 1.0 / (1.0 + sqrt(hll)); lm_scale[l] = sc; }
                else sc = lm_scale[l];
                const double hreg = hll + (mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0);
                const double wl = 1.0 / hreg;
                const bool finite = isfinite(wl);                 // bundle_adjustor.cpp:538 skip
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
                if (lm_ok && !a.victim_only) { aux[l].hll_reg = 1.0; aux[l].gl = 0.0; aux[l].hll = 0.0; }
            }
            // the warp's 32 records of 6 N floats go to the update kernel as ONE bulk copy (TMA engine, no LSU
            // instructions): shared -> global, read completion awaited before the next pass reuses the buffer
            if (a.hs_out && ch_ok) {
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
        }
        __syncthreads();

        // ========================= Phase B: Schur sum over the super-chunk =========================
        // fp32 accumulation over the <= kSuper / ksplit landmarks of this thread (<= 64 terms: relative
        // error of the partial ~ 3e-7, averaged down over the partials), then ONE fp64 flush.
        {
            float acc[36], accg[6];
#pragma unroll
            for (int i = 0; i < 36; ++i) acc[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 6; ++i) accg[i] = 0.f;
            if (b_active) {
                for (int s = kk; s < kSuper; s += ksplit) {
                    const int m = __float_as_int(hbuf[(size_t)s * R + 6 * N + 1]);
                    if (((m >> bf) & (m >> bg) & 1) == 0) continue;
                    const float *hf = hbuf + (size_t)s * R + bf * 6;
                    const float *hg = hbuf + (size_t)s * R + bg * 6;
                    const float2 f01 = *reinterpret_cast<const float2 *>(hf);
                    const float2 f23 = *reinterpret_cast<const float2 *>(hf + 2);
                    const float2 f45 = *reinterpret_cast<const float2 *>(hf + 4);
                    const float2 g01 = *reinterpret_cast<const float2 *>(hg);
                    const float2 g23 = *reinterpret_cast<const float2 *>(hg + 2);
                    const float2 g45 = *reinterpret_cast<const float2 *>(hg + 4);
                    const float hfv[6] = {f01.x, f01.y, f23.x, f23.y, f45.x, f45.y};
                    const float hgv[6] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y};
#pragma unroll
                    for (int i = 0; i < 6; ++i)
#pragma unroll
                        for (int j = 0; j < 6; ++j) acc[i * 6 + j] += hfv[i] * hgv[j];
                    if (b_diag) {
                        const float sg = hbuf[(size_t)s * R + 6 * N];
#pragma unroll
                        for (int j = 0; j < 6; ++j) accg[j] += hgv[j] * sg;
                    }
                }
            }
            // flush: sum over the k-split lanes (butterfly; ALL lanes run it, idle ones carry zeros, so the shuffles are
            // plain full-mask SHFL), lane kk adds rows kk, kk + ksplit, ... in fp64
#pragma unroll
            for (int i = 0; i < 36; ++i) {
                float v = acc[i];
                if (ksplit >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (ksplit >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
                acc[i] = v;
            }
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                float v = accg[j];
                if (ksplit >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (ksplit >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
                accg[j] = v;
            }
            if (b_active) {
                double *dst = Ss + pair_idx(bf, bg) * 36;
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    if ((i % ksplit) == kk) {
#pragma unroll
                        for (int j = 0; j < 6; ++j) dst[i * 6 + j] += (double)acc[i * 6 + j];
                    }
                if (b_diag && kk == 0) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) gsc[bf * 6 + j] += (double)accg[j];
                }
            }
        }
        __syncthreads();
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");        // before the h buffer is freed

    lin_epilogue(a, w, N, tid, Ss, Dta, gsc);
    double cd = (double)cost_acc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cd += __shfl_xor_sync(0xffffffffu, cd, off);
    if (lane == 0) atomicAdd(&cost_sm[0], cd);
    __syncthreads();
    if (tid == 0) {
        if (gridDim.x == 1) a.cost_vis[w] = cost_sm[0]; else atomicAdd(&a.cost_vis[w], cost_sm[0]);
    }
}

// ------------------------------------------------------------------------------------------------
// Back-substitution + Plus + candidate cost, thread per landmark (same role as update_cost_kernel).
struct UpdArgs;   // ba_update.cuh

}  // namespace pvio
