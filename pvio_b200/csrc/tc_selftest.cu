// Self-test of the tcgen05 3xTF32 SYRK building block (tc_syrk.cuh) on caller-provided data:
// D = sum_k a_k a_k^T for K vectors of 64 floats.  Diagnostic entry point (tests/test_gpu_tc.py);
// the product path uses the same device functions inside lin_tc_kernel.
#include "api_internal.h"
#include "tc_syrk.cuh"

namespace pvio {

template <int kGS>
__global__ void __launch_bounds__(256, 1) syrk_selftest_kernel(const float *A, int K, double *D) {
    extern __shared__ __align__(128) unsigned char raw[];
    float *a_hi = reinterpret_cast<float *>(raw);
    float *a_lo = a_hi + tc::kBufFloats;
    double *Dsm = reinterpret_cast<double *>(raw + 2 * tc::kBufFloats * 4 + 128);
    __shared__ __align__(8) uint64_t bars[tc::kAccs];
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tslot);
    if (tid == 0) {
        for (int j = 0; j < tc::kAccs; ++j) tc::mbar_init(&bars[j], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < 64 * 64; i += 256) Dsm[i] = 0.0;
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t taddr = tslot;
    uint32_t phase = 0;
    for (int k0 = 0; k0 < K; k0 += tc::kPassK) {
        for (int idx = tid; idx < 64 * tc::kPassK; idx += 256) {
            const int k = idx >> 6, m = idx & 63;
            const float v = (k0 + k < K) ? A[(size_t)(k0 + k) * 64 + m] : 0.f;
            float hi, lo;
            tc::split_tf32(v, hi, lo);
            a_hi[tc::a_off(m, k)] = hi;
            a_lo[tc::a_off(m, k)] = lo;
        }
        tc::fence_async_smem();
        __syncthreads();
        float facc[32];
        tc::syrk_pass<kGS>(taddr, a_hi, a_lo, bars, phase, tid, facc);
        if (lane < 16) {
            double *row = Dsm + ((warp & 3) * 16 + lane) * 64 + (warp >> 2) * 32;
#pragma unroll
            for (int i = 0; i < 32; ++i) row[i] += (double)facc[i];
        }
        __syncthreads();
    }
    for (int i = tid; i < 64 * 64; i += 256) D[i] = Dsm[i];
    __syncthreads();
    if (warp == 0) tc::tmem_free(taddr);
}

// ---- debug variant: one pass, raw dump of the 128 x 64 TMEM block.  mode 0: product layout, 3xTF32;
// 1: unpadded K-major strides, one TF32 MMA; 2: tcgen05.st pattern only (no MMA); 3: product layout, one TF32 MMA
__global__ void __launch_bounds__(256, 1) syrk_debug_kernel(const float *A, int K, float *out, int mode) {
    extern __shared__ __align__(128) unsigned char raw[];
    float *a_hi = reinterpret_cast<float *>(raw);
    float *a_lo = a_hi + tc::kBufFloats;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tslot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (warp == 0) tc::tmem_alloc(&tslot);
    if (tid == 0) { tc::mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t taddr = tslot;
    if (warp < 4) {      // clear / pattern
        for (int c = 0; c < 64; ++c) {
            const uint32_t v = mode == 2 ? __float_as_uint((float)((warp * 32 + lane) * 100 + c)) : 0u;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(taddr + ((uint32_t)(warp * 32) << 16) + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (mode != 2) {
        for (int idx = tid; idx < 64 * tc::kPassK; idx += 256) {
            const int k = idx >> 6, m = idx & 63;
            const float v = (k < K) ? A[(size_t)k * 64 + m] : 0.f;
            float hi, lo;
            tc::split_tf32(v, hi, lo);
            int off;
            if (mode == 1) off = (k >> 3) * 512 + (m >> 3) * 64 + ((k & 7) >> 2) * 32 + (m & 7) * 4 + (k & 3);
            else off = tc::a_off(m, k);
            a_hi[off] = hi; a_lo[off] = lo;
        }
        tc::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc::fence_after();
            if (mode == 1) {
                // K-major, no swizzle: SBO = 256 B between 8-row groups, LBO = 128 B between 4-k groups; 2048 B per k-step
                uint64_t d = 0;
                d |= (uint64_t)((tc::smem_u32(a_hi) >> 4) & 0x3fffu);
                d |= (uint64_t)(128u >> 4) << 16;
                d |= (uint64_t)(256u >> 4) << 32;
                d |= (uint64_t)1 << 46;
                const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (8u << 17) | (4u << 24);
                for (int s = 0; s < tc::kPassK / 8; ++s) {
                    const uint64_t ds = d + (uint64_t)(s * (2048 >> 4));   // unpadded strides
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                                 :: "r"(taddr), "l"(ds), "l"(ds), "r"(idesc), "r"(s > 0 ? 1u : 0u) : "memory");
                }
                tc::mma_commit(&bar);
            } else if (mode == 3) {
                const uint64_t dh = tc::make_desc(tc::smem_u32(a_hi));
                for (int s = 0; s < tc::kPassK / 8; ++s) {
                    const uint64_t st = (uint64_t)(s * ((tc::kStepFloats * 4) >> 4));
                    tc::mma_tf32(taddr, dh + st, dh + st, s > 0 ? 1u : 0u);
                }
                tc::mma_commit(&bar);
            } else {
                const uint64_t dh = tc::make_desc(tc::smem_u32(a_hi)), dl = tc::make_desc(tc::smem_u32(a_lo));
                for (int s = 0; s < tc::kPassK / 8; ++s) {
                    const uint64_t st = (uint64_t)(s * ((tc::kStepFloats * 4) >> 4));
                    tc::mma_tf32(taddr, dh + st, dh + st, s > 0 ? 1u : 0u);
                    tc::mma_tf32(taddr, dh + st, dl + st, 1u);
                    tc::mma_tf32(taddr, dl + st, dh + st, 1u);
                }
                tc::mma_commit(&bar);
            }
        }
        tc::mbar_wait(&bar, 0);
        tc::fence_after();
    }
    if (warp < 4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float v[16];
            tc::tmem_ld16(taddr + ((uint32_t)(warp * 32) << 16) + c * 16, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 64 + c * 16 + i] = v[i];
        }
    }
    if (tid == 0) out[128 * 64] = __uint_as_float(taddr);
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_free(taddr);
}

}  // namespace pvio

using namespace pvio;

extern "C" int pvio_b200_selftest_syrk(pvio_b200_handle hh, const float *A, int K, double *D) {
    const char *env = getenv("PVIO_B200_TC_GS");
    const int gs = env ? atoi(env) : 1;
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !A || !D || K < 1) return PVIO_B200_EINVAL;
    cudaSetDevice(h->device);
    float *dA = nullptr;
    double *dD = nullptr;
    if (cudaMalloc(&dA, sizeof(float) * 64 * (size_t)K) != cudaSuccess) return PVIO_B200_ECUDA;
    if (cudaMalloc(&dD, sizeof(double) * 64 * 64) != cudaSuccess) { cudaFree(dA); return PVIO_B200_ECUDA; }
    cudaMemcpy(dA, A, sizeof(float) * 64 * (size_t)K, cudaMemcpyHostToDevice);
    const size_t smem = 2 * tc::kBufFloats * 4 + 128 + 64 * 64 * 8;
    cudaFuncSetAttribute(syrk_selftest_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(syrk_selftest_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (gs == 2) syrk_selftest_kernel<2><<<1, 256, smem, h->stream>>>(dA, K, dD);
    else syrk_selftest_kernel<1><<<1, 256, smem, h->stream>>>(dA, K, dD);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(D, dD, sizeof(double) * 64 * 64, cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dD);
    if (e != cudaSuccess) { h->err = std::string("selftest_syrk: ") + cudaGetErrorString(e); return PVIO_B200_ECUDA; }
    return PVIO_B200_OK;
}

extern "C" int pvio_b200_selftest_syrk_raw(pvio_b200_handle hh, const float *A, int K, float *out, int mode) {
    Handle *h = reinterpret_cast<Handle *>(hh);
    if (!h || !A || !out || K < 1 || K > tc::kPassK) return PVIO_B200_EINVAL;
    cudaSetDevice(h->device);
    float *dA = nullptr, *dO = nullptr;
    cudaMalloc(&dA, sizeof(float) * 64 * (size_t)K);
    cudaMalloc(&dO, sizeof(float) * (128 * 64 + 4));
    cudaMemset(dO, 0xff, sizeof(float) * (128 * 64 + 4));
    cudaMemcpy(dA, A, sizeof(float) * 64 * (size_t)K, cudaMemcpyHostToDevice);
    const size_t smem = 2 * tc::kBufFloats * 4 + 128;
    cudaFuncSetAttribute(syrk_debug_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    syrk_debug_kernel<<<1, 256, smem, h->stream>>>(dA, K, dO, mode);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(out, dO, sizeof(float) * (128 * 64 + 4), cudaMemcpyDeviceToHost);
    cudaFree(dA); cudaFree(dO);
    if (e != cudaSuccess) { h->err = std::string("selftest_syrk_raw: ") + cudaGetErrorString(e); return PVIO_B200_ECUDA; }
    return PVIO_B200_OK;
}
