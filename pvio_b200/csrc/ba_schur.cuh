// Schur sum as its own kernel (second half of the split linearise stage).
//
// lin_tpl_kernel<kLoss, false> leaves one record per landmark slot in global memory (LinArgs::hs_out: sqrt(w) h
// per frame, sqrt(w) g_l, frame mask) and the DIRECT part of the reduced system in Hred / gred.  This kernel
// streams a window's records through shared memory -- bulk copies (TMA engine) into a two-slab ring signalled
// by mbarriers, issued by one thread -- accumulates
//     S(f, g) = sum_l (sqrt(w) h_lf)(sqrt(w) h_lg)^T   over the FREE frame pairs,   c(f) = sum_l sqrt(w) g_l sqrt(w) h_lf
// in the same output-stationary 6 x 6 register tiles as the fused kernel (fp32 partial sums of <= 64 terms,
// fp64 from there), and subtracts them from Hred / gred.
// Why split: the fused kernel is latency-bound at 16 warps per SM (119 registers, 94 KB of shared memory); this
// half needs ~60 registers and 48 KB, so it runs at 3x the occupancy, and the first half sheds the tile code.
#pragma once
#include "ba_lin4.cuh"

namespace pvio {

constexpr int kSlab = 64;                       // records per staged slab
constexpr int kSchurThreads = 256;

__host__ __device__ inline size_t schur_smem_bytes(int N) {
    const size_t npairs = (size_t)N * (N + 1) / 2;
    return sizeof(double) * (npairs * 36 + (size_t)N * 6) + 32 + 16 + sizeof(float) * 2 * kSlab * (6 * N + 2);
}

// kMaxThreads / kMinBlocks: <160, 4> covers up to 36 tiles x 4 lanes (8 free frames, cfg2) at 4 CTAs per SM,
// <256, 2> everything else
template <int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks)
schur_kernel(LinArgs a) {
    const int w = blockIdx.x + a.w0;
    const WinHdr &H = a.hdr[w];
    const int N = H.N, R = hs_rec(N);
    const int tid = threadIdx.x, nt = blockDim.x;                   // nt: the host sizes the CTA to the tile count
    const int npairs = N * (N + 1) / 2;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Ss = reinterpret_cast<double *>(smem_raw);              // [npairs][36]
    double *gsc = Ss + npairs * 36;                                 // [N][6]
    uint64_t *bars = reinterpret_cast<uint64_t *>(gsc + N * 6);     // [2]
    float *slab = reinterpret_cast<float *>(smem_raw + ((sizeof(double) * (npairs * 36 + N * 6) + 16 + 15) & ~(size_t)15));   // [2][kSlab][R]

    const unsigned fixed = (unsigned)H.fixed_mask & ((1u << N) - 1u);
    const unsigned freem = ~fixed & ((1u << N) - 1u);
    const int nfree = __popc(freem);
    const int ntask = nfree * (nfree + 1) / 2;
    const int ksplit = (ntask * 4 <= nt) ? 4 : ((ntask * 2 <= nt) ? 2 : 1);
    const int task = tid / ksplit, kk = tid % ksplit;
    const bool b_active = task < ntask;
    int bf = 0, bg = 0;
    if (ntask > 0) {
        const int p = min(task, ntask - 1);
        int f = 0;
        while ((f + 1) * (f + 2) / 2 <= p) ++f;
        const int g = p - f * (f + 1) / 2;
        bf = __fns(freem, 0, f + 1); bg = __fns(freem, 0, g + 1);
    }
    const bool b_diag = (bf == bg);

    const int n_slots = H.n_chunks * 32;
    const int n_slab = (n_slots + kSlab - 1) / kSlab;
    const float *src = a.hs_out + (size_t)w * a.hs_stride;
    auto issue = [&](int i) {                                       // one thread: bulk copy of slab i into ring slot i & 1
        const int cnt = min(kSlab, n_slots - i * kSlab);
        const uint32_t bytes = (uint32_t)(cnt * R * 4);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[i & 1]);
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slab + (size_t)(i & 1) * kSlab * R);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst), "l"(src + (size_t)i * kSlab * R), "r"(bytes), "r"(bar) : "memory");
    };
    if (tid == 0) {
        for (int j = 0; j < 2; ++j)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"((uint32_t)__cvta_generic_to_shared(&bars[j])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < npairs * 36 + N * 6; i += nt) Ss[i] = 0.0;                // Ss, gsc contiguous
    __syncthreads();
    if (tid == 0) { if (n_slab > 0) issue(0); if (n_slab > 1) issue(1); }

    float acc[36], accg[6];
#pragma unroll
    for (int i = 0; i < 36; ++i) acc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) accg[i] = 0.f;

    for (int i = 0; i < n_slab; ++i) {
        {   // wait for slab i
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&bars[i & 1]);
            const uint32_t parity = (uint32_t)((i >> 1) & 1);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        const int cnt = min(kSlab, n_slots - i * kSlab);
        const float *buf = slab + (size_t)(i & 1) * kSlab * R;
        if (b_active) {
            for (int s = kk; s < cnt; s += ksplit) {
                const float *rec = buf + (size_t)s * R;
                const int m = __float_as_int(rec[6 * N + 1]);
                if (((m >> bf) & (m >> bg) & 1) == 0) continue;
                const float *hf = rec + bf * 6, *hg = rec + bg * 6;
                const float2 f01 = *reinterpret_cast<const float2 *>(hf), f23 = *reinterpret_cast<const float2 *>(hf + 2),
                             f45 = *reinterpret_cast<const float2 *>(hf + 4);
                const float2 g01 = *reinterpret_cast<const float2 *>(hg), g23 = *reinterpret_cast<const float2 *>(hg + 2),
                             g45 = *reinterpret_cast<const float2 *>(hg + 4);
                const float hfv[6] = {f01.x, f01.y, f23.x, f23.y, f45.x, f45.y};
                const float hgv[6] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y};
#pragma unroll
                for (int ii = 0; ii < 6; ++ii)
#pragma unroll
                    for (int j = 0; j < 6; ++j) acc[ii * 6 + j] += hfv[ii] * hgv[j];
                if (b_diag) {
                    const float sg = rec[6 * N];
#pragma unroll
                    for (int j = 0; j < 6; ++j) accg[j] += hgv[j] * sg;
                }
            }
        }
        // every 4 slabs (256 records: <= 64 fp32 terms per lane, like the fused kernel) and at the end: fp64 flush.
        // ALL lanes run the shuffles (idle lanes carry zeros) so that they compile to plain full-mask SHFL.
        if ((i & 3) == 3 || i == n_slab - 1) {
#pragma unroll
            for (int e = 0; e < 36; ++e) {
                float v = acc[e];
                if (ksplit >= 2) v += __shfl_xor_sync(0xffffffffu, v, 1);
                if (ksplit >= 4) v += __shfl_xor_sync(0xffffffffu, v, 2);
                acc[e] = v;
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
                for (int ii = 0; ii < 6; ++ii)
                    if ((ii % ksplit) == kk) {
#pragma unroll
                        for (int j = 0; j < 6; ++j) dst[ii * 6 + j] += (double)acc[ii * 6 + j];
                    }
                if (b_diag && kk == 0) {
#pragma unroll
                    for (int j = 0; j < 6; ++j) gsc[bf * 6 + j] += (double)accg[j];
                }
            }
#pragma unroll
            for (int e = 0; e < 36; ++e) acc[e] = 0.f;
#pragma unroll
            for (int j = 0; j < 6; ++j) accg[j] = 0.f;
        }
        __syncthreads();                                            // slab i consumed by everybody
        if (tid == 0 && i + 2 < n_slab) issue(i + 2);
    }
    // ---- reduced system: direct part (already there) minus the Schur sum
    const int npairs_cap = a.Ncap * (a.Ncap + 1) / 2;
    double *Hred_o = a.Hred + (size_t)w * npairs_cap * 36;
    double *gred_o = a.gred + (size_t)w * a.Ncap * 6;
    for (int e = tid; e < npairs * 36; e += nt) { const double v = Ss[e]; if (v != 0.0) Hred_o[e] -= v; }
    for (int e = tid; e < N * 6; e += nt) { const double v = gsc[e]; if (v != 0.0) gred_o[e] -= v; }
}

}  // namespace pvio
