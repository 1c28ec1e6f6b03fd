// tcgen05 (5th-generation tensor core) helpers for the Schur SYRK of the linearise kernel:
//     S  +=  sum_k  a_k a_k^T         a_k = [ sqrt(w_l) h_l (6N rows) ; sqrt(w_l) g_l ]   (<= 64 rows)
// over the landmarks k of one pass.  This is the only GEMM-shaped contraction of the window
// (DESIGN.md 4.1): M = N = 64 (dof rows), K = landmarks.
//
// Precision: the operands are fp32 Jacobian products; one TF32 pass (10-bit mantissa) is far outside
// the 1e-5 bound on dx, so each operand is split  a = hi + lo  (both TF32-representable) and three
// MMAs  hi hi^T + hi lo^T + lo hi^T  recover ~22 bits ("3xTF32"); TMEM partial sums are kept short (see
// syrk_pass) and the fp64 shared-memory Schur sum takes over after every pass (<= 128 terms).
//
// Shared-memory operand layout (both A and B operands read the SAME buffer: the product is A A^T):
// "K-major, no swizzle" canonical layout of the UMMA shared-memory descriptor.  A core matrix is 8 rows (m)
// x 4 landmarks (k) = 8 x 16 B, rows 16 B apart.  One MMA (K = 8 for TF32) reads, per group of 8 rows, two
// core matrices LBO apart; row groups are SBO apart; consecutive MMAs (k-steps) are kStepBytes apart.
// The strides are padded (LBO 144 B, SBO 288 B, step 2336 B instead of 128 / 256 / 2048) so that the 32
// lanes of a warp, which own 32 different landmarks k and store the same row m, hit 32 different banks:
//   float offset of (m, k) = (k / 8) * 584 + (m / 8) * 72 + ((k % 8) / 4) * 36 + (m % 8) * 4 + (k % 4)
// (MN-major operands would give 16-byte vector stores, but the MN-major / no-swizzle descriptor produced
// all-zero products on B200 in our tests; K-major is the layout verified by tests/test_gpu_tc.py.)
// The accumulator D (64 x 64 fp32) lives in TMEM: row m = 16 * j + i  ->  lane 32 * j + i, column n.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pvio {
namespace tc {

constexpr int kPassK = 128;                     // landmarks per pass
constexpr int kRows = 64;                       // UMMA M = N
constexpr int kLboFloats = 36, kSboFloats = 72, kStepFloats = 584;
constexpr int kBufFloats = (kPassK / 8) * kStepFloats;   // one operand buffer (hi or lo): 37 376 B
constexpr int kAccs = 4;                        // ring of TMEM accumulators
constexpr int kTmemCols = 64 * kAccs;

__device__ __forceinline__ int a_off(int m, int k) {
    return (k >> 3) * kStepFloats + (m >> 3) * kSboFloats + ((k & 7) >> 2) * kLboFloats + (m & 7) * 4 + (k & 3);
}

// a_off(m, k) = m_off(m) + k_off(k); the 6 rows of frame f (m0 = 6 f) are m_off(m0) + 4 i, plus 40 once the
// rows cross into the next group of 8 (i >= 8 - (m0 & 7)): lets callers walk a frame without re-deriving it.
__device__ __forceinline__ int m_off(int m) { return (m >> 3) * kSboFloats + (m & 7) * 4; }
__device__ __forceinline__ int k_off(int k) { return (k >> 3) * kStepFloats + ((k & 7) >> 2) * kLboFloats + (k & 3); }
__device__ __forceinline__ int frame_row_off(int base, int cross, int i) { return base + 4 * i + (i >= cross ? 40 : 0); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// round-to-nearest TF32 (result is an fp32 bit pattern with the low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// a = hi + lo with hi on the TF32 grid (round half away from zero by integer arithmetic: cvt.rna.tf32 is
// emulated with ~4 instructions on sm_100a) and lo = a - hi exact in fp32; the tensor core ignores the low
// 13 mantissa bits of lo, an error of 2^-22 |a| with the sign of lo, i.e. unbiased with respect to a.
__device__ __forceinline__ void split_tf32(float a, float &hi, float &lo) {
    hi = __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
    lo = a - hi;
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (bit layout: start >> 4 [0,14), LBO >> 4 [16,30),
// SBO >> 4 [32,46), version = 1 [46,48), layout type [61,64) = 0)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((uint32_t)(kLboFloats * 4) >> 4) << 16;   // LBO: next core matrix along k (4 landmarks)
    d |= (uint64_t)((uint32_t)(kSboFloats * 4) >> 4) << 32;   // SBO: next group of 8 rows
    d |= (uint64_t)1 << 46;
    return d;
}

// instruction descriptor: D fp32 [4,6), A/B TF32 [7,10) [10,13), both K-major (bits 15, 16 clear),
// N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) |
                            ((uint32_t)(kRows >> 3) << 17) | ((uint32_t)(kRows >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    }
}

// one full warp allocates / frees the accumulator columns
__device__ __forceinline__ void tmem_alloc(uint32_t *slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(kTmemCols) : "memory");
}

// this thread's TMEM lane (32 * (warp % 4) + lane), 16 consecutive columns starting at col
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// A A^T of one pass (kPassK landmarks = 16 k-steps) in 3xTF32, called by ALL 256 threads of the CTA after
// the operand buffers were fenced (fence_async_smem) and the CTA synchronised.
//
// The tensor core truncates (rounds toward zero) when it aligns products to the accumulator, which biases
// same-sign sums: measured relative bias of a diagonal entry -1.3e-7 after 8 landmarks, -1.2e-6 after 128
// (tools/tc_debug.py).  So TMEM only ever holds the sum of kGS k-steps: thread 0 issues groups of kGS x 3
// MMAs round-robin into kAccs accumulators, every group is read back (tcgen05.ld) as soon as its commit
// arrives and added in round-to-nearest fp32 to the caller's registers; the caller adds them to the fp64
// Schur sum once per pass.
// Thread (warp wv, lane < 16) owns accumulator row (wv % 4) * 16 + lane, columns (wv / 4) * 32 .. + 31.
template <int kGS>
__device__ __forceinline__ void syrk_pass(uint32_t taddr, const float *a_hi, const float *a_lo, uint64_t *bars,
                                          uint32_t &phase, int tid, float (&facc)[32]) {
    constexpr int kGroups = (kPassK / 8) / kGS;
    static_assert(kGroups % kAccs == 0, "groups must fill whole rounds");
    const int wv = tid >> 5;
    const uint64_t dh = make_desc(smem_u32(a_hi)), dl = make_desc(smem_u32(a_lo));
    const uint32_t tld = taddr + ((uint32_t)((wv & 3) * 32) << 16) + (uint32_t)((wv >> 2) * 32);
#pragma unroll
    for (int i = 0; i < 32; ++i) facc[i] = 0.f;
    for (int r = 0; r < kGroups / kAccs; ++r) {
        if (tid == 0) {
            fence_after();
            for (int j = 0; j < kAccs; ++j) {
                const uint32_t td = taddr + (uint32_t)(j * 64);
                for (int q = 0; q < kGS; ++q) {
                    const int ks = (r * kAccs + j) * kGS + q;
                    const uint64_t step = (uint64_t)(ks * ((kStepFloats * 4) >> 4));   // start-address field, 16-byte units
                    mma_tf32(td, dh + step, dh + step, q > 0 ? 1u : 0u);   // hi hi^T
                    mma_tf32(td, dh + step, dl + step, 1u);                // hi lo^T
                    mma_tf32(td, dl + step, dh + step, 1u);                // lo hi^T
                }
                mma_commit(&bars[j]);
            }
        }
        for (int j = 0; j < kAccs; ++j) {
            if (wv == 0) mbar_wait(&bars[j], phase);    // one warp polls; the others block on the CTA barrier
            __syncthreads();
            fence_after();
            float v[16];
            tmem_ld16(tld + (uint32_t)(j * 64), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) facc[i] += v[i];
            tmem_ld16(tld + (uint32_t)(j * 64 + 16), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) facc[16 + i] += v[i];
        }
        phase ^= 1u;
        fence_before();
        __syncthreads();            // every accumulator has been read before the next round overwrites it
    }
}

}  // namespace tc
}  // namespace pvio
