// Internal handle of libpvio_b200 (host side).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/pvio_b200.h"
#include "ba_types.h"

namespace pvio {

template <typename T>
struct DevBuf {
    T *d = nullptr;      // device
    T *h = nullptr;      // pinned host staging (optional)
    size_t n = 0;
};

struct KltState;         // klt.cu

struct Handle {
    int device = 0;
    int W = 1, Ncap = 0, Mcap = 0, Kcap = 0;
    int Pcap = 4, Tcap = 0, Ocap = 0;            // planes / plane tracks / plane observations
    cudaStream_t stream = nullptr;
    cudaStream_t stream_up = nullptr, stream_down = nullptr;   // copy streams of the pipelined host path
    std::vector<cudaEvent_t> ev_up, ev_done, ev_down;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;

    // vision (always allocated)
    DevBuf<WinHdr> hdr;
    DevBuf<WinConst> cst;
    DevBuf<ObsRec> obs;
    DevBuf<LmRec> lms;
    DevBuf<double> rho, frames;
    DevBuf<WinCtrl> ctrl;
    DevBuf<double> rho_cand, frames_cand, lm_scale, dx_lm, dx_pose, pose_scale, v_pose;
    DevBuf<LmAux> lm_aux;
    DevBuf<float> hs;                            // [W][hs_stride] sqrt(w_l) h_l records, linearise -> update
    size_t hs_stride = 0;                        // (Mcap / 32 + Ncap + 1) chunks x 32 slots x 6 Ncap floats
    DevBuf<double> Hred, Hdd, gdir, gred, cost_vis, acc, aux_cost;
    DevBuf<double> Hfull, gfull;                 // debug dump (single window only)
    // inertial / prior / planes (allocated on first use)
    bool have_inertial = false;
    DevBuf<int32_t> imu_idx, prior_frames;
    DevBuf<double> imu_data, prior_S, prior_L, prior_e, prior_x0;
    bool have_planes = false;
    DevBuf<double> plane_param;
    DevBuf<int32_t> pt_plane, pt_begin, pt_frame;
    DevBuf<float> pt_z;
    // host bookkeeping per slot
    std::vector<std::vector<int32_t>> perm;      // packed landmark -> caller landmark
    std::vector<int> slot_M, slot_N, slot_K;
    std::vector<uint8_t> perm_identity;
    int n_uploaded = 0;
    int max_slot_free = 0;        // most free (non FF_FIX_POSE) frames of any window packed so far
    int max_slot_N = 0;           // largest window packed so far (selects the tensor-core linearise kernel)
    int tc_gs = 1;                // k-steps per TMEM partial sum (PVIO_B200_TC_GS, experiments)
    int split_shape = 3;          // CTA shape of lin_a_kernel (experiments)
    bool split_schur = true;      // PVIO_B200_SPLIT=1: linearise stage as two kernels (Phase A, then schur_kernel)
    int tc_mode = 0;              // PVIO_B200_TC: 1 = fused tcgen05 linearise kernel, 2 = split stage with the tcgen05 Schur kernel
    bool use_tc = false;          // PVIO_B200_TC=1: tcgen05 Schur SYRK (lin_tc_kernel) instead of the CUDA-core one
    float last_lin_ms = 0.f;
    // ring of event pairs around every linearise+Schur launch (roofline timing without host syncs)
    std::vector<cudaEvent_t> kev;
    int kev_count = 0;
    // CUDA graphs of the single-window launch sequences (latency path), keyed by everything that
    // shapes the launches; replayed on later calls (device buffers are persistent)
    typedef std::tuple<int, int, int, int, int, int, int, int, int, double, double, double> GraphKey;
    std::map<GraphKey, std::pair<cudaGraphExec_t, int>> graphs;
    bool capturing = false;
    bool use_graphs = true;
    KltState *klt = nullptr;
    double *pnp_dev = nullptr, *pnp_host = nullptr;   // pnp.cu: device buffer + pinned staging, grown on demand
    size_t pnp_words = 0;
};

int fail(Handle *h, int code, const char *what, cudaError_t e = cudaSuccess);

#define CK(h, call)                                                        \
    do {                                                                   \
        cudaError_t e__ = (call);                                          \
        if (e__ != cudaSuccess) return fail((h), PVIO_B200_ECUDA, #call, e__); \
    } while (0)

// api.cu: pack window into slot 0 and copy it to the device
int pack_and_upload(Handle *h, const pvio_b200_window *w, const pvio_b200_state *s);
// klt.cu
int klt_track_impl(Handle *h, const uint8_t *prev, const uint8_t *next, int width, int height, int stride,
                   const float *prev_pts, float *next_pts, uint8_t *status, float *err, int n_points,
                   int max_level, int max_iter, double eps, double clahe_clip = 0.0, int tiles_x = 0, int tiles_y = 0,
                   uint8_t *prev_eq = nullptr, uint8_t *next_eq = nullptr);
int clahe_impl(Handle *h, const uint8_t *src, int width, int height, int stride, double clip, int tiles_x, int tiles_y, uint8_t *dst);
void klt_free(Handle *h);
// pnp.cu
int pnp_solve_impl(Handle *h, const pvio_b200_pnp_problem *pb, double *frame, const pvio_b200_options *opt,
                   pvio_b200_summary *summary);
// ba_marg.cu
int marginalize_impl(Handle *h, const pvio_b200_window *w, const pvio_b200_state *s, int index,
                     double *S_out, double *e_out, double *H_out, double *b_out);

}  // namespace pvio
