// Device-side (packed) layout of a batch of sliding windows.  See DESIGN.md "Data layout".
// Observation / landmark records are the 16-byte fp32 records SURVEY.md 8(d) counts as
// algorithmic bytes; all STATE (poses, motion, inverse depths) stays fp64 because the
// residual must be evaluated in fp64 to meet the 1e-5 relative tolerance on dx.
#pragma once
#include <stdint.h>

namespace pvio {

constexpr int kMaxFrames = 16;      // PVIO_B200_MAX_FRAMES
constexpr int kFrameStride = 16;    // doubles per frame state
constexpr int kImuStride = 288;     // doubles per IMU factor record
constexpr int kMaxChunks = 96;      // anchor-homogeneous chunks of <= 32 landmarks per window
constexpr int kAcc = 16;            // doubles per window accumulated by the update / J.v sweeps
constexpr int kMaxSeg = kMaxFrames * (kMaxFrames - 1) / 2;   // (target, anchor) segments of the frame-major table
constexpr int kSegTab = 2 * (kMaxSeg + 1);                   // ints per window: seg_begin[kMaxSeg + 1], seg_row[kMaxSeg + 1]

struct __align__(8) ObsRec {        // one reprojection residual block (non-anchor observation), 8 B
    float zx, zy;                   // normalised keypoint in the target frame
};                                  // SURVEY's 16-B record also carried the landmark and frame indices: the
                                    // landmark is implied by the CSR offset in LmRec, the frame by the
                                    // landmark's frame mask (observations are stored in increasing frame
                                    // order, so record j belongs to the j-th set bit): half the bytes to move

struct __align__(16) LmRec {        // one inverse-depth landmark
    float zrx, zry;                 // normalised keypoint in the anchor frame
    int32_t meta;                   // anchor | in_victim << 4 | n_obs << 8 | target-frame mask << 16
    int32_t obs_begin;              // first ObsRec of this landmark
};
#if defined(__CUDACC__)
#define PVIO_HD __host__ __device__ __forceinline__
#else
#define PVIO_HD inline
#endif
PVIO_HD int lm_anchor(int32_t meta) { return meta & 0xf; }
PVIO_HD int lm_victim(int32_t meta) { return (meta >> 4) & 1; }
PVIO_HD int lm_nobs(int32_t meta) { return (meta >> 8) & 0x1f; }
PVIO_HD unsigned lm_mask(int32_t meta) { return (unsigned)meta >> 16; }
PVIO_HD int32_t lm_meta(int anchor, int victim, int n_obs, unsigned mask) {
    return (int32_t)((unsigned)anchor | ((unsigned)victim << 4) | ((unsigned)n_obs << 8) | (mask << 16));
}

// Frame-major observation table (built on the device from the landmark-major one, ba_fobs.cuh): the residual
// blocks sorted by (target frame t, anchor frame a, landmark), segment sp = t (t - 1) / 2 + a.  A warp of the
// linearise / update sweeps works on 32 consecutive entries of ONE segment ("row"), so the target frame is
// warp-uniform (held in registers) and the per-(t, a) direct blocks are plain register sums.
struct __align__(16) FObs { float zx, zy; int32_t lm, pad; };     // keypoint in the target frame, packed landmark index

// per-landmark sums of the linearise sweep (H_ll, g_l), formed by the Schur kernel
struct LmSum { double hll, gl; };

struct WinHdr {                     // per-window integers
    int32_t N, M, K, use_inertial;
    int32_t n_imu, n_prior, n_planes, n_ptracks;
    int32_t fixed_mask;             // bit f: FF_FIX_POSE
    int32_t n_chunks;
    int32_t pad0, pad1;
    int32_t chunk_begin[kMaxChunks];   // first landmark of chunk c
    int32_t chunk_meta[kMaxChunks];    // count | anchor << 8
};

struct WinConst {                   // per-window doubles
    double cam_q[4], cam_p[3];
    double imu_q[4], imu_p[3];
    double sic[4];                  // sqrt_inv_cov 2x2
    double fx, fy, cauchy_a, plane_sic;
    double origin[3];               // window origin subtracted from all positions (mean frame position)
    double pad;
};

struct __align__(16) WinCtrl {      // per-window solver state (device resident)
    double mu, radius;
    double cost, cand_cost;         // 0.5 * sum rho(s): current / candidate
    double cost_vis, cand_cost_vis; // reprojection part accumulated by the sweeps
    double g_dot_dx, dx_reg_dx, gn_norm2, gmax, xnorm2, dxnorm2;
    double model_change;
    double grad2, v_reg_dx;         // |D^-1 S g|^2 and v^T (mu D) dx over the pose block (dogleg)
    // device-side trust-region loop (ba_tr.cuh)
    double step_a, step_b;          // step = step_b * dx_gn - step_a * v of the current iteration
    double step_norm;               // its norm in the scaled space
    double initial_cost;
    double max_time_ns, t_start_ns; // solver_options.h:30; %globaltimer at the start of the solve
    int32_t iteration, accepted, done, termination;
    int32_t solve_failed, have_scale, usable, reuse;   // reuse: the next iteration keeps linearisation and GN solve
    int32_t need_jv, skip, max_iter, fresh;            // skip: the rest of this iteration's kernels do nothing
    int32_t buf, have_lin, pad0, pad1;                 // buf: buffer set (LinBufs) holding the linearisation of the STATE;
                                                       // have_lin: it is valid for the current state and mu
};

// The linearisation of a window (records, pivots, direct blocks, reduced system, reprojection cost) lives in one of two
// buffer sets.  An iteration of the trust-region loop linearises its CANDIDATE into the other set: the cost of that
// sweep decides the step, and when the step is accepted the Jacobians of the new state are already there (WinCtrl::buf
// flips) -- the residual-only evaluation ceres does for the decision and the relinearisation that follows it are ONE
// sweep here.  Offsets of set 1 from set 0: bytes for the type-erased arrays, elements for the others.
struct LinBufs {
    size_t hs, jr, lm_w;            // bytes
    size_t lm_msk, lm_aux, Hred, Hdd, g, cost;   // elements
};

// per-landmark Schur scalars written by the linearise kernel, read by the update kernel
struct __align__(16) LmAux {
    double hll_reg;                 // H_ll + mu * clamp(.)  (the pivot the Schur complement divides by)
    double gl;                      // g_l
    double hll;                     // H_ll
    double pad;
};

}  // namespace pvio
