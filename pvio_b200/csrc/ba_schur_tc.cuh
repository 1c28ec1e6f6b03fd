// Schur sum on the tensor cores (opt-in: PVIO_B200_TC=2), second half of the split linearise stage.
// Same inputs / outputs as schur_kernel (ba_schur.cuh).  Per 32-record slab:
//   records (bulk copy ring, 2 slots)  ->  transpose + 3xTF32 split into the K-major UMMA operand buffers
//   (tc_syrk.cuh layout; entries of frames a landmark does not see are written as zeros)  ->  thread 0 issues
//   2 groups x 2 k-steps x 3 tcgen05.mma into 2 TMEM accumulators (16-landmark partial sums: the tensor core's
//   accumulate truncates, tc_syrk.cuh)  ->  64 lanes read them back (tcgen05.ld) and add them in fp32 registers
//   -> fp64 Schur sum in shared memory every 256 records  ->  Hred / gred -= at the end.
// 51 KB of shared memory and 128 TMEM columns per CTA: 4 CTAs per SM.
#pragma once
#include "ba_schur.cuh"
#include "tc_syrk.cuh"

namespace pvio {

constexpr int kTcSlab = 32;                     // records per slab = 4 k-steps
constexpr int kTcOpFloats = (kTcSlab / 8) * tc::kStepFloats;   // one operand buffer (hi or lo)

__host__ __device__ inline size_t schur_tc_smem_bytes(int N) {
    const size_t npairs = (size_t)N * (N + 1) / 2;
    return sizeof(double) * (npairs * 36 + (size_t)N * 6) + 64 + 16 + sizeof(float) * (2 * kTcSlab * (6 * N + 2) + 2 * kTcOpFloats) + 16;
}

__global__ void __launch_bounds__(128, 4)
schur_tc_kernel(LinArgs a) {
    const int w = blockIdx.x + a.w0;
    const WinHdr &H = a.hdr[w];
    const int N = H.N, R = hs_rec(N), RG = 6 * N;                   // RG: operand row of sqrt(w) g_l
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int npairs = N * (N + 1) / 2;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *Ss = reinterpret_cast<double *>(smem_raw);              // [npairs][36]
    double *gsc = Ss + npairs * 36;                                 // [N][6]
    uint64_t *bars = reinterpret_cast<uint64_t *>(gsc + N * 6);     // [0,1]: record ring, [2,3]: MMA groups
    uint32_t *tslot = reinterpret_cast<uint32_t *>(bars + 4);
    const size_t off0 = (sizeof(double) * (npairs * 36 + N * 6) + 64 + 15) & ~(size_t)15;
    float *slab = reinterpret_cast<float *>(smem_raw + off0);       // [2][kTcSlab][R]
    float *a_hi = reinterpret_cast<float *>(smem_raw + ((off0 + sizeof(float) * 2 * kTcSlab * R + 15) & ~(size_t)15));
    float *a_lo = a_hi + kTcOpFloats;

    const int n_slots = H.n_chunks * 32;
    const int n_slab = n_slots / kTcSlab;                            // chunks are 32 slots: whole slabs
    const float *src = a.hs_out + (size_t)w * a.hs_stride;
    auto issue_load = [&](int i) {
        const uint32_t bytes = (uint32_t)(kTcSlab * R * 4);
        const uint32_t bar = tc::smem_u32(&bars[i & 1]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(tc::smem_u32(slab + (size_t)(i & 1) * kTcSlab * R)), "l"(src + (size_t)i * kTcSlab * R), "r"(bytes), "r"(bar) : "memory");
    };
    if (wv == 0) {                                                   // 128 TMEM columns: two 64 x 64 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc::smem_u32(tslot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int j = 0; j < 4; ++j) tc::mbar_init(&bars[j], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < npairs * 36 + N * 6; i += 128) Ss[i] = 0.0;
    for (int i = tid; i < 2 * kTcOpFloats; i += 128) a_hi[i] = 0.f;  // rows >= 6 N + 1 stay zero
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t taddr = *tslot;
    if (tid == 0) { if (n_slab > 0) issue_load(0); if (n_slab > 1) issue_load(1); }

    // read-back ownership: lanes < 16 of warp wv hold accumulator rows wv * 16 + lane, all 64 columns
    const int m_row = wv * 16 + lane;
    const bool row_ok = lane < 16 && m_row <= RG;
    float facc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) facc[i] = 0.f;
    const uint64_t dh = tc::make_desc(tc::smem_u32(a_hi)), dl = tc::make_desc(tc::smem_u32(a_lo));

    for (int i = 0; i < n_slab; ++i) {
        tc::mbar_wait(&bars[i & 1], (uint32_t)((i >> 1) & 1));
        const float *buf = slab + (size_t)(i & 1) * kTcSlab * R;
        // ---- transpose + split: thread -> (slot k = tid & 31, rows m = (tid >> 5), + 4, ...): a warp writes one row
        // of 32 landmarks, which the padded K-major layout spreads over 32 banks
        {
            const int k = tid & 31;
            const float *rec = buf + (size_t)k * R;
            const int mask = __float_as_int(rec[RG + 1]);
            const int ko = tc::k_off(k);
            for (int m = wv; m <= RG; m += 4) {
                const int f = m / 6;
                float v = 0.f;
                if (m == RG) v = mask ? rec[RG] : 0.f;
                else if ((mask >> f) & 1) v = rec[m];
                float hi, lo;
                tc::split_tf32(v, hi, lo);
                const int o = ko + tc::m_off(m);
                a_hi[o] = hi; a_lo[o] = lo;
            }
        }
        tc::fence_async_smem();
        tc::fence_before();
        __syncthreads();                                             // operands complete; previous accumulators read; slab free
        if (tid == 0) {
            tc::fence_after();
            if (i + 2 < n_slab) issue_load(i + 2);
            for (int g = 0; g < 2; ++g) {
                const uint32_t td = taddr + (uint32_t)(g * 64);
                for (int q = 0; q < 2; ++q) {
                    const uint64_t step = (uint64_t)((g * 2 + q) * ((tc::kStepFloats * 4) >> 4));
                    tc::mma_tf32(td, dh + step, dh + step, q > 0 ? 1u : 0u);
                    tc::mma_tf32(td, dh + step, dl + step, 1u);
                    tc::mma_tf32(td, dl + step, dh + step, 1u);
                }
                tc::mma_commit(&bars[2 + g]);
            }
        }
        for (int g = 0; g < 2; ++g) {
            tc::mbar_wait(&bars[2 + g], (uint32_t)(i & 1));
            tc::fence_after();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[16];
                tc::tmem_ld16(taddr + ((uint32_t)(wv * 32) << 16) + (uint32_t)(g * 64 + c * 16), v);
#pragma unroll
                for (int e = 0; e < 16; ++e) facc[c * 16 + e] += v[e];
            }
        }
        // ---- fp64 flush every 8 slabs (256 records) and at the end
        if ((i & 7) == 7 || i == n_slab - 1) {
            if (row_ok) {
                if (m_row < RG) {
                    const int f = m_row / 6, i6 = m_row - f * 6;
                    double *row = Ss + pair_idx(f, 0) * 36 + i6 * 6;
#pragma unroll
                    for (int n = 0; n < 64; ++n) {
                        const int g = n / 6, j6 = n - g * 6;          // compile time
                        if (n < RG) { if (g <= f) row[g * 36 + j6] += (double)facc[n]; }
                        else if (n == RG) gsc[m_row] += (double)facc[n];
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 64; ++e) facc[e] = 0.f;
        }
    }
    tc::fence_before();
    __syncthreads();
    const int npairs_cap = a.Ncap * (a.Ncap + 1) / 2;
    double *Hred_o = a.Hred + (size_t)w * npairs_cap * 36;
    double *gred_o = a.gred + (size_t)w * a.Ncap * 6;
    for (int e = tid; e < npairs * 36; e += 128) { const double v = Ss[e]; if (v != 0.0) Hred_o[e] -= v; }
    for (int e = tid; e < N * 6; e += 128) { const double v = gsc[e]; if (v != 0.0) gred_o[e] -= v; }
    if (wv == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(128) : "memory");
}

}  // namespace pvio
