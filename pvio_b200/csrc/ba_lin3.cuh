// Reprojection sweep, third generation: thread-pair per landmark + tcgen05 Schur SYRK.
//
// Same mathematics and outputs as ba_lin2.cuh.  The profile of the second generation
// (profiles/r01b_lin_tpl_n592.md) put 42 % of its instructions into Phase B, the rank-K update
//     S += sum_l w_l h_l h_l^T,   h_l in R^(6N)
// which is a plain SYRK with K = landmarks: GEMM-shaped, so it belongs on the tensor cores.  Here
//   Phase A  a PAIR of lanes owns a landmark (even / odd observation records), a warp owns 16
//            landmarks of one anchor-homogeneous chunk, the CTA 128 landmarks per pass.  The direct
//            terms sum Y^T Y, Y^T r are transpose-reduced over the 16 lanes of equal parity (they see
//            the same target frame when tracks are contiguous runs of frames) and added to the
//            (target, anchor) blocks in shared memory with fp64 atomics, as before.
//   Phase B  the lane pair scales h_l by sqrt(w_l), splits it into TF32 hi + lo and stores column l of
//            the 64 x 128 operand matrix  A = [sqrt(w) h ; sqrt(w) g_l]  (K-major UMMA layout,
//            tc_syrk.cuh); ONE thread issues 16 x 3 tcgen05.mma (3xTF32) computing A A^T into a ring of
//            four 64 x 64 fp32 TMEM accumulators (short partial sums: the tensor core's accumulate
//            truncates, tc_syrk.cuh); all warps read them back (tcgen05.ld), add them in fp32 registers
//            and then to the fp64 Schur sum in shared memory.  The 61st row of A makes the Schur
//            gradient correction sum_l w g_l h_l the 61st column of the same product.
// Status (profiles/r01c_lin_tc.md): parity-green, 138 K warp-instructions per cfg2 window instead of
// 151 K, but 0.98-1.03 ms per 4096 windows against 0.89 ms for lin_tpl_kernel: four passes of 128
// landmarks with CTA-wide barriers around the issue -> commit -> tcgen05.ld round trip serialise what
// the CUDA-core version overlaps, and two CTAs per SM cannot hide it.  Opt-in (PVIO_B200_TC=1) until the
// operands are double-buffered behind a dedicated MMA / read-back warp.
// Limits: 6 N + 1 <= 64 rows, i.e. N <= 10 frames (the reference's windows are 8 + 1, euroc.yaml:47);
// larger windows and the marginaliser stay on lin_tpl_kernel.
#pragma once
#include "ba_lin2.cuh"
#include "tc_syrk.cuh"

namespace pvio {

constexpr int kTcMaxFrames = 10;

// Sums of q[i] over the 16 lanes of the caller's parity: lane 2p + h returns entries 2p and 2p + 1 of its
// parity class (stages xor 16, 8, 4, 2 of the transpose reduction; xor 1 would mix the classes).
__device__ __forceinline__ void transpose_reduce_2x16(const float (&q)[32], int lane, float &e0, float &e1) {
    float v[16];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float send = up ? q[k] : q[k + 16], keep = up ? q[k + 16] : q[k];
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
    e0 = v[0]; e1 = v[1];
}

// Add 32 consecutive columns (first column kC0, compile time) of one accumulator row to the fp64 Schur sum:
// column n < R belongs to block (f, n / 6) and is stored when that block is in the lower triangle; column R is
// the Schur gradient correction of this row.
template <int kC0>
__device__ __forceinline__ void schur_flush_row(const float (&facc)[32], double *row, double *gsc_m, int f, int R) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int n = kC0 + i;
        const int g = n / 6, j6 = n - g * 6;                 // compile time
        if (n < R) { if (g <= f) row[g * 36 + j6] += (double)facc[i]; }
        else if (n == R) *gsc_m += (double)facc[i];
    }
}

__host__ __device__ inline size_t lin3_smem_bytes(int N) {
    const size_t npairs = (size_t)N * (N + 1) / 2, nsp = (size_t)N * (N - 1) / 2;
    return sizeof(FrameSm) * kMaxFrames                 // camera poses
           + sizeof(double) * (npairs * 36)              // Schur sum (block lower triangular)
           + sizeof(double) * (nsp + 1) * 33             // direct (target, anchor) blocks + gradients
           + sizeof(double) * (size_t)N * 6              // Schur gradient correction
           + sizeof(double) * 8                          // cost
           + 48                                          // mbarriers + TMEM slot
           + 16                                          // alignment slack of the operand buffers
           + sizeof(float) * 2 * tc::kBufFloats;         // A_hi, A_lo
}

template <bool kLoss, int kGS>
__global__ void __launch_bounds__(kLinThreads, 2)
lin_tc_kernel(LinArgs a) {
    const int w = blockIdx.y + a.w0;
    const WinHdr &H = a.hdr[w];
    const WinConst &wc = a.cst[w];
    const int N = H.N;
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int npairs = N * (N + 1) / 2, nsp = N * (N - 1) / 2;
    const int R = 6 * N;                                            // row of sqrt(w) g_l

    extern __shared__ __align__(16) unsigned char smem_raw[];
    FrameSm *F = reinterpret_cast<FrameSm *>(smem_raw);
    double *Ss = reinterpret_cast<double *>(F + kMaxFrames);        // [npairs][36] Schur sum
    double *Dta = Ss + npairs * 36;                                 // [nsp + 1][33] direct blocks: 21 sym + 6 grad (+6 pad)
    double *gsc = Dta + (nsp + 1) * 33;                             // [N][6] sum_l w g_l h_f
    double *cost_sm = gsc + N * 6;                                  // [8]
    uint64_t *bars = reinterpret_cast<uint64_t *>(cost_sm + 8);     // [kAccs]
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + tc::kAccs);
    // operand buffers at a 16-byte aligned OFFSET of the (16-byte aligned) dynamic shared memory, so that the
    // compiler keeps the shared address space (a pointer rounded through uintptr_t becomes generic: LD.E / ST.E)
    const unsigned a_byte_off = (unsigned)((sizeof(FrameSm) * kMaxFrames + sizeof(double) * (npairs * 36 + (nsp + 1) * 33 + N * 6 + 8) +
                                            sizeof(uint64_t) * tc::kAccs + 8 + 15) & ~(size_t)15);
    float *a_hi = reinterpret_cast<float *>(smem_raw + a_byte_off);
    float *a_lo = a_hi + tc::kBufFloats;

    if (wv == 0) tc::tmem_alloc(tslot);
    if (tid == 0) {
        for (int j = 0; j < tc::kAccs; ++j) tc::mbar_init(&bars[j], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < N) make_frame(a.frames + ((size_t)w * a.Ncap + tid) * kFrameStride, wc, F[tid]);
    for (int i = tid; i < npairs * 36 + (nsp + 1) * 33 + N * 6 + 8; i += kLinThreads) Ss[i] = 0.0;   // Ss, Dta, gsc, cost contiguous
    for (int i = tid; i < 2 * tc::kBufFloats; i += kLinThreads) a_hi[i] = 0.f;     // rows > 6 N stay zero for the whole kernel
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t taddr = *tslot;
    uint32_t phase = 0;

    const float W[4] = {(float)wc.sic[0], (float)wc.sic[1], (float)wc.sic[2], (float)wc.sic[3]};
    const float cb = (float)(wc.cauchy_a * wc.cauchy_a);
    const double mu = a.mu_override >= 0.0 ? a.mu_override : a.ctrl[w].mu;
    const ObsRec *obs = a.obs + (size_t)w * a.Kcap;
    const LmRec *lms = a.lms + (size_t)w * a.Mcap;
    const double *rho = a.rho + (size_t)w * a.Mcap;
    double *lm_scale = a.lm_scale + (size_t)w * a.Mcap;
    LmAux *aux = a.lm_aux + (size_t)w * a.Mcap;
    const unsigned fixed = a.victim_only ? 0u : ((unsigned)H.fixed_mask & ((1u << N) - 1u));
    const int h2 = lane & 1, pl = lane >> 1;
    const int n_units = H.n_chunks * 2;                             // half chunks of <= 16 landmarks

    float cost_acc = 0.f;

    for (int u0 = blockIdx.x * 8; u0 < n_units; u0 += gridDim.x * 8) {
        // ========================= Phase A: one warp per half chunk, one lane PAIR per landmark =========================
        const int u = u0 + wv;
        const bool u_ok = u < n_units;
        const int ch = u_ok ? (u >> 1) : 0;
        const int lm0 = H.chunk_begin[ch];
        const int cnt = u_ok ? (H.chunk_meta[ch] & 0xff) : 0;
        const int anchor = H.chunk_meta[ch] >> 8;
        const int li = (u & 1) * 16 + pl;
        const bool lm_ok = li < cnt;
        const int l = lm0 + (lm_ok ? li : 0);
        const int koff = tc::k_off(wv * 16 + pl);                   // column of A owned by this lane pair
        const LmRec lr = lms[l];
        int n_obs = lm_ok ? lm_nobs(lr.meta) : 0;
        if (a.victim_only && !lm_victim(lr.meta)) n_obs = 0;
        const unsigned seen = n_obs > 0 ? lm_mask(lr.meta) : 0u;
        unsigned fm = seen;
        if (h2) fm &= fm - 1;                                       // the odd lane starts at the second record
        const int n_mine = (n_obs + 1 - h2) >> 1;
        const int n_max = __reduce_max_sync(0xffffffffu, n_mine);
        const double rl = lm_ok ? rho[l] : 1.0;
        double x[3];
        float xf[3], cl[3];
        world_point(F[anchor], lr.zrx, lr.zry, rl, x, xf, cl);
        double hll = 0.0, gl = 0.0;
        float ha[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int it = 0; it < n_max; ++it) {
            const bool act = it < n_mine;
            ObsRec o;
            o.zx = 0.f; o.zy = 0.f;
            if (act) o = obs[lr.obs_begin + 2 * it + h2];
            const int t = act ? __ffs(fm) - 1 : 0;                  // (2 it + h2)-th set bit of the frame mask
            fm &= fm - 1;
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
                const int rb = koff + tc::m_off(6 * t), cross = 8 - ((6 * t) & 7);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    const float h = ol.j0 * ol.Y0[i] + ol.j1 * ol.Y1[i];
                    ha[i] -= h;
                    a_hi[tc::frame_row_off(rb, cross, i)] = h;   // unscaled; scaled and split below
                }
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int jj = i; jj < 6; ++jj) q[sym6(i, jj)] = ol.Y0[i] * ol.Y0[jj] + ol.Y1[i] * ol.Y1[jj];
#pragma unroll
                for (int i = 0; i < 6; ++i) q[21 + i] = ol.Y0[i] * ol.r0 + ol.Y1[i] * ol.r1;
            }
            // reduce the direct terms of this iteration over the landmarks of the warp, per target frame
            const unsigned actm = __ballot_sync(0xffffffffu, act);
            const unsigned evn = actm & 0x55555555u, odd = actm & 0xaaaaaaaau;
            const int te = __shfl_sync(0xffffffffu, t, evn ? __ffs(evn) - 1 : 0);
            const int to = __shfl_sync(0xffffffffu, t, odd ? __ffs(odd) - 1 : 0);
            const int tmine = h2 ? to : te;
            if (__all_sync(0xffffffffu, !act || t == tmine)) {
                // the common case: each parity class sees one frame
                float e0, e1;
                transpose_reduce_2x16(q, lane, e0, e1);
                const unsigned grp = h2 ? odd : evn;
                if (grp && tmine != anchor && !((fixed >> tmine) & (fixed >> anchor) & 1u)) {
                    const int sp = tmine > anchor ? spair(tmine, anchor) : spair(anchor, tmine);
                    const int i0 = lane & ~1;
                    // sign convention: the block stores +sum Y^T Y; the epilogue applies the signs
                    if (i0 < kDirVals) atomicAdd(&Dta[sp * 33 + i0], (double)e0);
                    if (i0 + 1 < kDirVals) atomicAdd(&Dta[sp * 33 + i0 + 1], (double)e1);
                }
            } else {
                unsigned todo = actm;
                while (todo) {
                    const int leader = __ffs(todo) - 1;
                    const int tf = __shfl_sync(0xffffffffu, t, leader);
                    const unsigned peers = __ballot_sync(0xffffffffu, act && t == tf);
                    todo &= ~peers;
                    if ((fixed >> tf) & (fixed >> anchor) & 1u) continue;
                    const float tot = transpose_reduce32(q, act && t == tf, lane);
                    if (lane < kDirVals && tf != anchor) {
                        const int sp = tf > anchor ? spair(tf, anchor) : spair(anchor, tf);
                        atomicAdd(&Dta[sp * 33 + lane], (double)tot);
                    }
                }
            }
        }
        // ---- per-landmark Schur scalars: combine the pair
        hll += __shfl_xor_sync(0xffffffffu, hll, 1);
        gl += __shfl_xor_sync(0xffffffffu, gl, 1);
#pragma unroll
        for (int i = 0; i < 6; ++i) ha[i] += __shfl_xor_sync(0xffffffffu, ha[i], 1);
        float sw = 0.f;
        if (n_obs > 0) {
            double sc;
            if (a.compute_scale) { sc = 1.0 / (1.0 + sqrt(hll)); if (!h2) lm_scale[l] = sc; }
            else sc = lm_scale[l];
            const double hreg = hll + (mu > 0.0 ? lm_reg(hll, sc, mu) : 0.0);
            const double wl = 1.0 / hreg;
            const bool finite = isfinite(wl);                       // bundle_adjustor.cpp:538 skip
            sw = finite ? sqrtf((float)wl) : 0.f;
            if (!h2) { aux[l].hll_reg = hreg; aux[l].gl = gl; aux[l].hll = hll; }
        } else if (lm_ok && !a.victim_only && !h2) {
            aux[l].hll_reg = 1.0; aux[l].gl = 0.0; aux[l].hll = 0.0;
        }
        float *hs = (a.hs_out && lm_ok) ? a.hs_out + (size_t)w * a.hs_stride + (size_t)(ch * 32 + li) * hs_rec(N) : nullptr;
        if (hs && !h2) { hs[6 * N] = sw * (float)gl; hs[6 * N + 1] = __int_as_float(sw != 0.f ? (int)(seen | (1u << anchor)) : 0); }
        __syncwarp();                                               // the partner's unscaled h values
        // ---- column `slot` of A: every row below 6 N + 1 is (re)written, zeros where the landmark is not seen
        for (int f = h2; f < N; f += 2) {
            const bool is_anchor = (f == anchor);
            const bool is_seen = (seen >> f) & 1u;
            const int rb = koff + tc::m_off(6 * f), cross = 8 - ((6 * f) & 7);
            const bool live = (is_anchor || is_seen) && sw != 0.f;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const int off = tc::frame_row_off(rb, cross, i);
                float hi = 0.f, lo = 0.f;
                if (live) {
                    const float v = (is_anchor ? ha[i] : a_hi[off]) * sw;
                    tc::split_tf32(v, hi, lo);
                    if (hs) hs[f * 6 + i] = v;
                } else if (hs && (is_anchor || is_seen)) {
                    hs[f * 6 + i] = 0.f;
                }
                a_hi[off] = hi;
                a_lo[off] = lo;
            }
        }
        if (!h2) {
            const int off = koff + tc::m_off(R);
            float hi, lo;
            tc::split_tf32(sw != 0.f ? sw * (float)gl : 0.f, hi, lo);
            a_hi[off] = hi;
            a_lo[off] = lo;
        }
        tc::fence_async_smem();
        __syncthreads();

        // ========================= Phase B: A A^T on the tensor cores =========================
        {
            float facc[32];
            tc::syrk_pass<kGS>(taddr, a_hi, a_lo, bars, phase, tid, facc);
            const int m = (wv & 3) * 16 + lane;                     // accumulator row held by this lane (lanes < 16)
            if (lane < 16 && m < R) {
                const int f = m / 6, i6 = m - f * 6;
                double *row = Ss + pair_idx(f, 0) * 36 + i6 * 6;    // block (f, g) of this row: + 36 g
                if ((wv >> 2) == 0) schur_flush_row<0>(facc, row, gsc + m, f, R);
                else schur_flush_row<32>(facc, row, gsc + m, f, R);
            }
        }
        __syncthreads();
    }

    lin_epilogue(a, w, N, tid, Ss, Dta, gsc);
    double cd = (double)cost_acc;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) cd += __shfl_xor_sync(0xffffffffu, cd, off);
    if (lane == 0) atomicAdd(&cost_sm[0], cd);
    __syncthreads();
    if (tid == 0) {
        if (gridDim.x == 1) a.cost_vis[w] = cost_sm[0]; else atomicAdd(&a.cost_vis[w], cost_sm[0]);
    }
    if (wv == 0) tc::tmem_free(taddr);
}

}  // namespace pvio
