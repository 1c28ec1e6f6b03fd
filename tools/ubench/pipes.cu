// Micro-benchmarks of the sm_100a issue / pipe rates the BA kernels are designed around (run under gpurun):
//   FFMA vs FFMA2 (fma.rn.f32x2) vs DFMA throughput per SM, shared-memory fp32 / fp64 atomics, LDS.64 / LDS.128.
// Prints warp-instructions per clock per SM for each.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) pipe_kernel(float *out, int iters, float seed) {
    float2 a[8];
    double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); d[i] = seed + i; }
    const float2 b = make_float2(1.0001f, 0.9999f);
    const float bs = 1.0001f;
    const double bd = 1.0000001;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, bs, 0.5f); a[i].y = fmaf(a[i].y, bs, 0.25f); }          // 2 FFMA
                else if (MODE == 1) a[i] = __ffma2_rn(a[i], b, b);                                         // 1 FFMA2
                else if (MODE == 2) a[i] = __ffma2_rn(make_float2(bs, bs), a[i], b);                        // FFMA2 scalar-broadcast form
                else if (MODE == 3) d[i] = fma(d[i], bd, 0.5);                                             // DFMA
                else if (MODE == 4) { a[i].x = fmaf(a[i].x, bs, 0.5f); d[i] = fma(d[i], bd, 0.5); }         // FFMA + DFMA mix
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + (float)d[i];
    if (s == 12345.678f) out[0] = s;
}

template <int MODE>
__global__ void __launch_bounds__(256) smem_kernel(float *out, int iters) {
    __shared__ float sf[2048];
    __shared__ double sd[1024];
    __shared__ __align__(16) float4 sv[1024];
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += 256) sf[i] = 0.f;
    for (int i = tid; i < 1024; i += 256) { sd[i] = 0.0; sv[i] = make_float4(1, 2, 3, 4); }
    __syncthreads();
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int idx = (tid + r * 37 + it) & 1023;
            if (MODE == 0) atomicAdd(&sf[idx], 1.0f);                 // native fp32 shared atomic, conflict-free
            else if (MODE == 1) atomicAdd(&sd[idx], 1.0);             // fp64 shared atomic (CAS loop)
            else if (MODE == 2) { const float4 v = sv[idx]; acc += v.x + v.w; }      // LDS.128
            else if (MODE == 3) { const float2 v = reinterpret_cast<float2 *>(sv)[idx * 2]; acc += v.x + v.y; }   // LDS.64
            else if (MODE == 4) { acc += sf[idx]; }                   // LDS.32
        }
    }
    __syncthreads();
    if (acc == 12345.678f || sf[tid] == -1.f || sd[tid] == -1.0) out[0] = acc;
}


// Dependent-chain latencies seen by ONE warp (clock64 around an unrolled chain): DFMA, DMUL->DFMA, rsqrt(double), FFMA,
// an LDS round trip, a CTA barrier of 8 warps.
template <int MODE>
__global__ void lat_kernel(long long *out, double seed, int n) {
    __shared__ double sm[64];
    double x = seed, y = 1.0000001;
    float xf = (float)seed;
    sm[threadIdx.x & 63] = seed;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (MODE == 0) x = fma(x, y, 0.5);
        else if (MODE == 1) x = rsqrt(x + 2.0);
        else if (MODE == 2) xf = fmaf(xf, 1.0001f, 0.5f);
        else if (MODE == 3) { x = sm[((int)x) & 63] + 1.0; }          // LDS -> convert -> address
        else if (MODE == 4) { __syncthreads(); }
        else if (MODE == 5) x = 1.0 / (x + 2.0);
        else if (MODE == 6) x = sqrt(x + 2.0);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[MODE] = (t1 - t0);
    if (x == 12345.678 || xf == 1234.5f) out[15] = 1;
}

template <typename F>
static float time_it(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, %d kHz\n", p.name, p.multiProcessorCount, clk_khz);
    float *out;
    CK(cudaMalloc(&out, 16));
    const int iters = 2000, grid = p.multiProcessorCount * 8;
    const double clocks = 1.0;
    (void)clocks;
    auto rep = [&](const char *name, float ms, double warp_instr_per_thread_iter, int its) {
        // warp instructions issued per SM per clock (at the reported max clock)
        const double wi = (double)grid * 8 /*warps*/ * its * warp_instr_per_thread_iter;
        const double cyc = ms * 1e-3 * clk_khz * 1e3;
        printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SM\n", name, ms, wi / cyc / p.multiProcessorCount);
    };
    rep("FFMA  (2 per pair)", time_it([&] { pipe_kernel<0><<<grid, 256>>>(out, iters, 1.f); }), 128, iters);
    rep("FFMA2 (vector x vector)", time_it([&] { pipe_kernel<1><<<grid, 256>>>(out, iters, 1.f); }), 64, iters);
    rep("FFMA2 (scalar broadcast)", time_it([&] { pipe_kernel<2><<<grid, 256>>>(out, iters, 1.f); }), 64, iters);
    rep("DFMA", time_it([&] { pipe_kernel<3><<<grid, 256>>>(out, iters, 1.f); }), 64, iters);
    rep("FFMA + DFMA (1:1)", time_it([&] { pipe_kernel<4><<<grid, 256>>>(out, iters, 1.f); }), 128, iters);
    rep("ATOMS.ADD.F32", time_it([&] { smem_kernel<0><<<grid, 256>>>(out, iters); }), 8, iters);
    rep("atomicAdd(double) shared (CAS)", time_it([&] { smem_kernel<1><<<grid, 256>>>(out, iters); }), 8, iters);
    rep("LDS.128", time_it([&] { smem_kernel<2><<<grid, 256>>>(out, iters); }), 8, iters);
    rep("LDS.64", time_it([&] { smem_kernel<3><<<grid, 256>>>(out, iters); }), 8, iters);
    rep("LDS.32", time_it([&] { smem_kernel<4><<<grid, 256>>>(out, iters); }), 8, iters);
    {
        long long *d_lat, h_lat[16] = {0};
        CK(cudaMalloc(&d_lat, sizeof(h_lat)));
        CK(cudaMemset(d_lat, 0, sizeof(h_lat)));
        const int n = 4096;
        lat_kernel<0><<<1, 32>>>(d_lat, 1.5, n); lat_kernel<1><<<1, 32>>>(d_lat, 1.5, n); lat_kernel<2><<<1, 32>>>(d_lat, 1.5, n);
        lat_kernel<3><<<1, 32>>>(d_lat, 1.5, n); lat_kernel<4><<<1, 256>>>(d_lat, 1.5, n); lat_kernel<5><<<1, 32>>>(d_lat, 1.5, n);
        lat_kernel<6><<<1, 32>>>(d_lat, 1.5, n);
        CK(cudaMemcpy(h_lat, d_lat, sizeof(h_lat), cudaMemcpyDeviceToHost));
        const char *nm[7] = {"DFMA dependent chain", "rsqrt(double) chain (+ DADD)", "FFMA dependent chain", "LDS.64 -> F2I -> address chain (+ DADD)",
                             "__syncthreads, 8 warps", "1.0 / x (double) chain (+ DADD)", "sqrt(double) chain (+ DADD)"};
        for (int i = 0; i < 7; ++i) printf("%-44s %7.1f cycles per step\n", nm[i], (double)h_lat[i] / n);
    }
    return 0;
}
