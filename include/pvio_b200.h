/* pvio_b200 -- C ABI of the B200-native sliding-window bundle-adjustment + KLT backend.
 *
 * Drop-in boundary for ONE hot path of zju3dv/PVIO (reference paths relative to
 * /root/reference):
 *   - pvio_b200_ba_solve        replaces the ceres::Solve call in
 *                               pvio/src/pvio/estimation/bundle_adjustor.cpp:244-249
 *                               (BundleAdjustorSolver::solve :63-299; post-pass :277-296)
 *   - pvio_b200_ba_gn_step      one linearise -> Schur -> solve -> Plus -> cost iteration
 *                               (test / benchmark unit; what Ceres does once per iteration)
 *   - pvio_b200_ba_marginalize  replaces BundleAdjustor::marginalize_frame,
 *                               bundle_adjustor.cpp:348-599
 *   - pvio_b200_klt_track       replaces cv::calcOpticalFlowPyrLK in
 *                               pvio-extra/src/pvio/extra/opencv_image.cpp:103
 *                               (OpenCvImage::track_keypoints :88-136)
 *   - pvio_b200_reprojection_error  BundleAdjustor::compute_reprojection_error :321-336
 *
 * Conventions: plain pointers and sizes, caller-owned arrays, the callee
 * never frees them; doubles at the boundary (the reference's state is fp64); no
 * exceptions; integer return codes (0 = ok, < 0 = error, see PVIO_B200_E*); one
 * in-flight call per handle (the reference calls solve from one thread,
 * core/frontend_worker.cpp:58-77); handles are independent across GPUs.
 * There is no CPU fallback: every entry point fails with PVIO_B200_ENODEV without a GPU.
 */
#ifndef PVIO_B200_H
#define PVIO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVIO_B200_OK 0
#define PVIO_B200_ENODEV (-1)   /* no CUDA device / driver error            */
#define PVIO_B200_EINVAL (-2)   /* bad argument or window exceeds capacity   */
#define PVIO_B200_ENOMEM (-3)
#define PVIO_B200_ECUDA (-4)    /* kernel / runtime failure (see last_error) */
#define PVIO_B200_ENUMERIC (-5) /* reduced system not positive definite      */

#define PVIO_B200_MAX_FRAMES 16      /* N: window frames (reference: 9..11)  */
#define PVIO_B200_FRAME_STRIDE 16    /* doubles per frame: q(x,y,z,w) p v bg ba, estimation/state.h:43-69 */
#define PVIO_B200_IMU_STRIDE 288     /* doubles per IMU factor, layout below */

/* IMU factor record = the PreIntegrator outputs the cost functor reads
 * (estimation/preintegrator.h:29-44, ceres/preintegration_error_cost.h:53-76):
 *   [0] dt | [1..4] dq (x,y,z,w) | [5..7] dp | [8..10] dv |
 *   [11..235] sqrt_inv_cov 15x15 row-major |
 *   [236..244] dq_dbg | [245..253] dp_dbg | [254..262] dp_dba | [263..271] dv_dbg |
 *   [272..280] dv_dba (3x3 row-major) | [281..283] bg0 | [284..286] ba0 | [287] pad
 * bg0/ba0: the bias linearisation point (frame_i->motion at integrate() time). */
#define PVIO_B200_IMU_DT 0
#define PVIO_B200_IMU_DQ 1
#define PVIO_B200_IMU_DP 5
#define PVIO_B200_IMU_DV 8
#define PVIO_B200_IMU_SQRT_INV_COV 11
#define PVIO_B200_IMU_DQ_DBG 236
#define PVIO_B200_IMU_DP_DBG 245
#define PVIO_B200_IMU_DP_DBA 254
#define PVIO_B200_IMU_DV_DBG 263
#define PVIO_B200_IMU_DV_DBA 272
#define PVIO_B200_IMU_BG0 281
#define PVIO_B200_IMU_BA0 284

/* One sliding window, gathered by the shim from Map/Frame/Track exactly in the order
 * bundle_adjustor.cpp:75-242 walks them.  Keypoints are normalised image coordinates. */
typedef struct pvio_b200_window {
    int32_t n_frames;            /* map->frame_num()                              :75  */
    int32_t n_landmarks;         /* VALID non-PLANE tracks with a parameter block :91-103 */
    int32_t n_obs;               /* reprojection residual blocks (anchor excluded) :149 */
    int32_t use_inertial;        /* solve(map, config, use_inertial)              :63  */
    const uint8_t *frame_fixed;  /* [N] Frame::flag(FF_FIX_POSE)                  :79  */
    double cam_q_cs[4], cam_p_cs[3]; /* Frame::camera (estimation/state.h:38-41)       */
    double imu_q_cs[4], imu_p_cs[3]; /* Frame::imu                                     */
    double sqrt_inv_cov[4];      /* Frame::sqrt_inv_cov 2x2 row-major (core/core.cpp:112-116) */
    double fx, fy;               /* Frame::K(0,0), K(1,1): post-pass pixel error  :291 */
    double cauchy_a;             /* ceres::CauchyLoss(1.0)                        :58  */
    const int32_t *lm_anchor;    /* [M] index of Track::first_frame()                  */
    const double *lm_z_ref;      /* [M][2] first_keypoint()                            */
    const int32_t *lm_obs_begin; /* [M+1] CSR offsets into obs_*                       */
    const uint8_t *lm_in_victim; /* [M] track observed by the frame to marginalise :454 (may be NULL) */
    const int32_t *obs_frame;    /* [K] sorted by landmark, then frame index           */
    const double *obs_z;         /* [K][2]                                             */
    int32_t n_imu;               /* pre-integration factors                      :220-242 */
    const int32_t *imu_frame_i;  /* [n_imu]                                            */
    const int32_t *imu_frame_j;  /* [n_imu]                                            */
    const double *imu_data;      /* [n_imu][PVIO_B200_IMU_STRIDE]                      */
    int32_t n_prior;             /* frames related to the marginalisation prior  :126-139 */
    const int32_t *prior_frames; /* [n_prior] window index of each related frame       */
    const double *prior_S;       /* [(15 n)^2] row-major sqrt_inv_cov (marginalization_error_cost.h:103) */
    const double *prior_e;       /* [15 n] infovec                                     */
    const double *prior_state0;  /* [n_prior][16] pose_0 / motion_0 snapshots :44-45   */
    int32_t n_planes;            /* planes with >= 20 tracks                     :165  */
    const double *plane_param;   /* [n_planes][4] normal(3), distance                  */
    double plane_sqrt_inv_cov;   /* sqrt(1 / config->plane_distance_cov())       :181  */
    int32_t n_plane_tracks;
    const int32_t *pt_plane;     /* [T] plane index of each plane-track          :183  */
    const int32_t *pt_obs_begin; /* [T+1] CSR offsets                                  */
    const int32_t *pt_obs_frame; /* [...] every observation of the track         :186  */
    const double *pt_obs_z;      /* [...][2]                                           */
} pvio_b200_window;

/* In/out state: Frame::pose, Frame::motion, Track::landmark.inv_depth -- overwritten in
 * place as ceres does (bundle_adjustor.cpp:77-101). */
typedef struct pvio_b200_state {
    double *frames;    /* [N][PVIO_B200_FRAME_STRIDE] */
    double *inv_depth; /* [M] */
} pvio_b200_state;

typedef struct pvio_b200_options {
    int32_t max_iterations; /* config->solver_iteration_limit() (10)   solver_options.h:29 */
    double max_time;        /* config->solver_time_limit() seconds     solver_options.h:30 */
    int32_t alias_bias;     /* 1: bias linearisation point follows the accepted state (quirk Q1,
                               the reference's behaviour); 0: frozen at solve entry          */
    int32_t run_postpass;   /* 1: run the landmark depth / pixel-error pass :277-296         */
    double initial_trust_region_radius; /* 0: ceres default 1e4 (PVIO does not override it);
                               other values exist to exercise the dogleg legs in tests         */
} pvio_b200_options;

#define PVIO_B200_TERM_CONVERGENCE 0
#define PVIO_B200_TERM_NO_CONVERGENCE 1
#define PVIO_B200_TERM_FAILURE 2

typedef struct pvio_b200_summary {
    int32_t iterations;      /* trust-region iterations executed                          */
    int32_t accepted_steps;
    int32_t termination;     /* PVIO_B200_TERM_*                                          */
    int32_t usable;          /* ceres Summary::IsSolutionUsable()                   :298  */
    double initial_cost, final_cost;
    double final_radius, final_mu;
    double solve_seconds;    /* device time of the solve (CUDA events)                    */
} pvio_b200_summary;

typedef struct pvio_b200_handle_s *pvio_b200_handle;

/* ---- lifetime ---------------------------------------------------------------------- */
/* Owns device buffers, one stream, CUDA graphs.  max_windows > 1 enables the batched
 * entry points (BASELINE config 5 and the roofline launch). */
int pvio_b200_create(int device, int max_windows, int max_frames, int max_landmarks,
                     int max_obs, pvio_b200_handle *out);
void pvio_b200_destroy(pvio_b200_handle h);
const char *pvio_b200_last_error(pvio_b200_handle h);
/* number of kernels this library launched since creation (bench.py "gpu_launches") */
int64_t pvio_b200_kernel_launches(pvio_b200_handle h);
const char *pvio_b200_version(void);

/* ---- single window: the reference-facing calls ------------------------------------- */
/* Full solve; state updated in place.  valid/quality ([M], may be NULL) receive the
 * post-pass result (TF_VALID after the depth test, landmark.quality). */
int pvio_b200_ba_solve(pvio_b200_handle h, const pvio_b200_window *w, pvio_b200_state *s,
                       const pvio_b200_options *opt, pvio_b200_summary *summary,
                       uint8_t *valid, double *quality);

/* One regularised Gauss-Newton iteration at `s` (not modified): dx = -(H + mu D)^-1 g in
 * the local coordinates [N][15] then [M] (theta, p, v, bg, ba per frame; d inv_depth per
 * landmark; masked coordinates are 0), the cost at s, and the cost after Plus(s, dx).
 * Hred/gred (may be NULL) receive the reduced system in the same [N][15] layout
 * (row-major D x D, D = 15 N). */
int pvio_b200_ba_gn_step(pvio_b200_handle h, const pvio_b200_window *w, const pvio_b200_state *s,
                         double mu, double *dx, double *cost, double *new_cost,
                         double *Hred, double *gred);

/* New prior over the remaining frames after marginalising frame `index`:
 * S_out [(15(N-1))^2] row-major, e_out [15(N-1)]; H_out/b_out (may be NULL) the
 * information matrix / vector before the eigen factorisation. */
int pvio_b200_ba_marginalize(pvio_b200_handle h, const pvio_b200_window *w, const pvio_b200_state *s,
                             int index, double *S_out, double *e_out, double *H_out, double *b_out);

/* ---- resident sliding window ----------------------------------- */
/* The window persists in the handle between keyframes instead of being rebuilt from Map / Frame / Track at every
 * solve (bundle_adjustor.cpp:75-242): the shim reports what CHANGED.  The marginalisation prior stays ON THE DEVICE:
 * pvio_b200_window_drop_victim leaves S, e and the linearisation point where the solve kernels read them.
 * One resident window per handle.  Frame arguments are window indices (0 = oldest). */
/* Start an empty window.  Of `constants` the scalar members are kept (extrinsics, sqrt_inv_cov, fx, fy, cauchy_a,
 * use_inertial); its arrays are ignored. */
int pvio_b200_window_reset(pvio_b200_handle h, const pvio_b200_window *constants);
/* sliding_window_tracker.cpp:113-118.  state [PVIO_B200_FRAME_STRIDE]; imu_record [PVIO_B200_IMU_STRIDE]: the
 * pre-integration factor (previous frame, this frame), NULL for the first frame / a visual-only window. */
int pvio_b200_window_append_frame(pvio_b200_handle h, const double *state, int fixed, const double *imu_record);
/* New tracks, each with its first observation z [n][2] in frame[i] and an initial inverse depth (the shim's
 * triangulation, track.cpp:83-106).  ids_out [n] (may be NULL): handles for the calls below. */
int pvio_b200_window_add_tracks(pvio_b200_handle h, int n, const int32_t *frame, const double *z,
                                const double *inv_depth, int32_t *ids_out);
/* Track::add_keypoint (track.cpp:32-37): observations arrive in increasing frame order per track. */
int pvio_b200_window_add_observations(pvio_b200_handle h, int n, const int32_t *track, const int32_t *frame, const double *z);
int pvio_b200_window_remove_track(pvio_b200_handle h, int32_t track);
/* A prior given by the caller over frames 0 .. n_prior-1 (the 1e15 gauge prior of the first window,
 * sliding_window_tracker.cpp:100-112); later priors come from pvio_b200_window_drop_victim. */
int pvio_b200_window_set_prior(pvio_b200_handle h, int n_prior, const double *S, const double *e, const double *state0);
/* BundleAdjustor::solve on the resident window (tracks with >= 2 observations take part); states and inverse
 * depths are updated in the handle. */
int pvio_b200_window_solve(pvio_b200_handle h, const pvio_b200_options *opt, pvio_b200_summary *summary);
/* Map::marginalize_frame(0) (map.cpp:76-88): marginaliser -> new prior (device resident), the victim's observations
 * leave their tracks, tracks anchored in it are re-anchored from the current estimate (track.cpp:42-49). */
int pvio_b200_window_drop_victim(pvio_b200_handle h);
/* Read back: frames [n_frames][PVIO_B200_FRAME_STRIDE]; per queried track its inverse depth, anchor frame (-1: gone)
 * and observation count.  Any output pointer may be NULL. */
int pvio_b200_window_get(pvio_b200_handle h, int32_t *n_frames, double *frames, int n_tracks, const int32_t *tracks,
                         double *inv_depth, int32_t *anchor_frame, int32_t *n_observations);

/* Mean pixel reprojection error over all observations of the window's landmarks. */
int pvio_b200_reprojection_error(pvio_b200_handle h, const pvio_b200_window *w,
                                 const pvio_b200_state *s, double *error);

/* ---- batched windows (independent problems, one CTA-group each) --------------------- */
/* Pack window `slot` (0 <= slot < max_windows) into the pinned staging area in device
 * layout (fp32 observation table, fp64 state).  Nothing is copied to the GPU yet. */
int pvio_b200_batch_set_window(pvio_b200_handle h, int slot, const pvio_b200_window *w,
                               const pvio_b200_state *s);
/* Replicate slot 0 into slots 1..n-1 (synthetic batches of identical shape). */
int pvio_b200_batch_replicate(pvio_b200_handle h, int n);
/* Host -> device copy of the first n packed windows (async on the handle's stream). */
int pvio_b200_batch_upload(pvio_b200_handle h, int n);
/* One GN iteration for the first n windows, device-resident; apply != 0 keeps the
 * updated state on the device (as an accepted step). */
int pvio_b200_batch_gn_step(pvio_b200_handle h, int n, double mu, int apply);
/* Device -> host: dx [n][15 N + M] (stride dx_stride doubles), costs [n][2]. */
int pvio_b200_batch_download(pvio_b200_handle h, int n, double *dx, int64_t dx_stride, double *costs);
/* upload + gn_step + download in one call: the end-to-end path bench.py times. */
int pvio_b200_batch_gn_step_host(pvio_b200_handle h, int n, double mu, double *dx,
                                 int64_t dx_stride, double *costs);
/* Full trust-region solve (ceres::Solve semantics per window: dogleg, per-window termination, at most
 * opt->max_iterations iterations) of the first n uploaded windows, entirely on the device: no host round trip per
 * iteration.  The device-resident states are overwritten with the solutions (BASELINE config 5: independent windows). */
int pvio_b200_batch_solve(pvio_b200_handle h, int n, const pvio_b200_options *opt);
/* Device -> host: solved states frames [n][frames_stride] (16 doubles per frame), inv_depth [n][inv_depth_stride] in
 * the caller's landmark order, one summary per window (solve_seconds is not filled).  Any pointer may be NULL. */
int pvio_b200_batch_download_state(pvio_b200_handle h, int n, double *frames, int64_t frames_stride, double *inv_depth,
                                   int64_t inv_depth_stride, pvio_b200_summary *summaries);
/* upload + batch_solve + download_state in one call with HOST buffers, pipelined over sub-batches: one host -> device
 * copy buys up to max_iterations iterations per window. */
int pvio_b200_batch_solve_host(pvio_b200_handle h, int n, const pvio_b200_options *opt, double *frames, int64_t frames_stride,
                               double *inv_depth, int64_t inv_depth_stride, pvio_b200_summary *summaries);
int pvio_b200_sync(pvio_b200_handle h);
/* CUDA-event timing on the handle's stream (bench.py): start / stop return ms. */
int pvio_b200_timer_start(pvio_b200_handle h);
int pvio_b200_timer_stop(pvio_b200_handle h, float *ms);
/* device-side time (ms) of the linearise + Schur stage from CUDA events around its launches: which = 0 the most recent
 * stage, 1 the mean stage since the last reset, 2 / 3 the mean of the linearise / Schur kernel alone, -1 reset */
int pvio_b200_last_kernel_ms(pvio_b200_handle h, int which, float *ms);

/* ---- visual-inertial PnP ------------------------------------------------ */
/* Replaces the ceres::Solve of visual_inertial_pnp, pvio/src/pvio/estimation/pnp.cpp:32-100: the new
 * frame's pose (and v, bg, ba when use_inertial) against PreIntegrationPriorCost
 * (ceres/preintegration_error_cost.h:167-206; last_frame is constant) and pose-only reprojection
 * blocks on constant world points (ceres/reprojection_error_cost.h:128-203, CauchyLoss(1.0)):
 * Track::get_landmark_point() for ordinary tracks, the plane-cast point for TF_PLANE tracks. */
typedef struct pvio_b200_pnp_problem {
    int32_t n_points;
    int32_t use_inertial;
    const double *points;        /* [n][3] world points                                        */
    const double *z;             /* [n][2] normalised keypoints in the new frame               */
    double cam_q_cs[4], cam_p_cs[3], imu_q_cs[4], imu_p_cs[3];
    double sqrt_inv_cov[4];
    double cauchy_a;
    const double *last_frame;    /* [16] map->last_frame() state (constant)          pnp.cpp:44 */
    const double *imu_data;      /* [PVIO_B200_IMU_STRIDE] frame->preintegration (bg0/ba0 ignored:
                                    they alias last_frame's biases)                             */
} pvio_b200_pnp_problem;

/* frame: [16] in/out (q xyzw, p, v, bg, ba).  One kernel launch runs the whole trust-region solve. */
int pvio_b200_pnp_solve(pvio_b200_handle h, const pvio_b200_pnp_problem *problem, double *frame,
                        const pvio_b200_options *opt, pvio_b200_summary *summary);

/* Same, from the RAW 8-bit frames: CLAHE (cv::createCLAHE(clahe_clip, Size(tiles_x, tiles_y))->apply,
 * pvio-extra/src/pvio/extra/opencv_image.cpp:138-143: clip 6.0, 8 x 8 tiles) runs on the device between the upload and
 * the pyramid build, bit-exact with OpenCV.  width / height must be multiples of tiles_x / tiles_y (752x480 and
 * 512x512 are).  prev_eq / next_eq (optional, [height][width]): the equalised frames, e.g. to hand to GFTT. */
int pvio_b200_klt_track_raw(pvio_b200_handle h, const uint8_t *prev, const uint8_t *next,
                            int width, int height, int stride, const float *prev_pts,
                            float *next_pts, uint8_t *status, float *err, int n_points,
                            int max_level, int max_iter, double eps, double clahe_clip, int tiles_x, int tiles_y,
                            uint8_t *prev_eq, uint8_t *next_eq);

/* The call OpenCvImage::track_keypoints maps to in a running tracker (opencv_image.cpp:88-136, caller
 * core/feature_tracker.cpp:92: prev is always the previous call's next): frames carry ids (!= 0) and the handle keeps the
 * finished pyramids (CLAHE output + 4 levels) of the last two frames on the device, so a steady-state call uploads and
 * builds ONE pyramid; an image whose id is cached may be passed as NULL.  clahe_clip > 0: raw frames, CLAHE on the
 * device; 0: frames are already equalised.  border > 0: the 20-pixel border rejection of :106-108 runs on the device. */
int pvio_b200_klt_track_cached(pvio_b200_handle h, uint64_t prev_id, const uint8_t *prev, uint64_t next_id, const uint8_t *next,
                               int width, int height, int stride, const float *prev_pts, float *next_pts, uint8_t *status,
                               float *err, int n_points, int max_level, int max_iter, double eps, double clahe_clip,
                               int tiles_x, int tiles_y, int border);

/* OpenCvImage::detect_keypoints (pvio-extra/src/pvio/extra/opencv_image.cpp:54-86): GFTT with the Harris response
 * (GFTTDetector::create(1000, 1e-3, 20, 3, true), :183) on the preprocessed frame, PVIO's Poisson-disk filter
 * (utility/poisson_disk_filter.h) against the `existing` keypoints [n_existing][2] with radius keypoint_distance, 20-pixel
 * border.  The frame is taken from the pyramid cache of pvio_b200_klt_track_cached when frame_id is cached there
 * (image may then be NULL); otherwise `image` is uploaded (clahe_clip > 0: raw frame, equalised on the device first).
 * keypoints_out [max_out][2] pixel coordinates in the reference's order (decreasing response), *n_out their number.
 * gftt_out [1000][2] / n_gftt (may be NULL): the corners before the Poisson filter, for parity checks against
 * cv::goodFeaturesToTrack. */
int pvio_b200_detect_keypoints(pvio_b200_handle h, uint64_t frame_id, const uint8_t *image, int width, int height, int stride,
                               double clahe_clip, int tiles_x, int tiles_y, const double *existing, int n_existing,
                               double keypoint_distance, int max_out, double *keypoints_out, int *n_out,
                               float *gftt_out, int *n_gftt);

/* The F-matrix outlier rejection of OpenCvImage::track_keypoints (pvio-extra/src/pvio/extra/opencv_image.cpp:121-129):
 * cv::findFundamentalMat(p, q, cv::FM_RANSAC, threshold, confidence, mask) with OpenCV's semantics -- n >= 15: RANSAC on
 * 7-point models, at most max_iters (OpenCV: 1000) iterations with the adaptive iteration count; 8 <= n < 15: OpenCV's
 * LMedS branch; n == 7: mask all ones; n < 7: mask zero.  p, q: [n][2] pixel coordinates (the cv::Point2f values).
 * schedule == NULL: the samples are the ones OpenCV itself draws (cv::RNG((uint64)-1) and getSubset restated), so the
 * mask is cv2's; otherwise int32 [n_schedule][7] injected sample indices (a row starting with -1 ends the run).
 * Every iteration of the schedule is evaluated on the device at once; the host replays OpenCV's acceptance rule over
 * the inlier counts (see csrc/fmat.cu).  mask: [n] 0/1.  F (may be NULL): the winning model, row-major 3x3.
 * info (may be NULL): int32[4] = iterations the serial loop runs, winning iteration, winning model, method (1 RANSAC,
 * 2 LMedS, 3 seven points). */
int pvio_b200_find_fundamental_mask(pvio_b200_handle h, int n, const float *p, const float *q, double threshold, double confidence,
                                    int max_iters, const int32_t *schedule, int n_schedule, uint8_t *mask, double *F, int32_t *info);

/* Host only (no device, no handle): the sample schedule the call above uses when schedule == NULL, i.e. the subsets
 * cv::findFundamentalMat draws for these n >= 8 matches with cv::RNG((uint64)-1): int32 [iters][7].  Returns the number of
 * iterations that have a subset (< iters only if getSubset gives up: degenerate, collinear input), or a negative error. */
int pvio_b200_fm_sample_schedule(int n, const float *p, const float *q, int iters, int32_t *schedule);

/* OpenCvImage::track_keypoints as a whole (opencv_image.cpp:88-136): pvio_b200_klt_track_cached (LK on the cached pyramids,
 * 20-pixel border on the device) followed by the F-matrix rejection above over the surviving matches when there are at
 * least 8 (:121).  next_pts holds the initial guess on entry (:91-96: the caller's next_keypoints, or a copy of prev_pts);
 * on return status[i] != 0 marks the matches the reference keeps and next_pts[i] their positions. */
int pvio_b200_track_keypoints(pvio_b200_handle h, uint64_t prev_id, const uint8_t *prev, uint64_t next_id, const uint8_t *next,
                              int width, int height, int stride, const float *prev_pts, float *next_pts, uint8_t *status,
                              int n_points, int max_level, int max_iter, double eps, double clahe_clip, int tiles_x, int tiles_y,
                              int border, double ransac_threshold, double ransac_confidence);

/* CLAHE of one frame (same kernels); dst is [height][width]. */
int pvio_b200_clahe(pvio_b200_handle h, const uint8_t *src, int width, int height, int stride, double clip_limit,
                    int tiles_x, int tiles_y, uint8_t *dst);

/* ---- IMU pre-integration ---------------------------------------------------------------- */
/* PreIntegrator::integrate(t, bg, ba, true, true) (pvio/src/pvio/estimation/preintegrator.cpp:85-98) for
 * n_factors independent factors, one warp each.  samples: rows (t, w xyz, a xyz) of every factor back to
 * back, begin[n_factors + 1] the CSR offsets; the last sample of a factor is integrated up to t_end[i].
 * bias: [n_factors][6] (bg, ba) -- the biases of the factor's FIRST frame (bundle_adjustor.cpp:228).
 * noise_cov: [4][9] row-major 3x3 cov_w, cov_a, cov_bg, cov_ba (preintegrator.h:63-66).
 * records: [n_factors][PVIO_B200_IMU_STRIDE], the layout pvio_b200_window::imu_data takes. */
int pvio_b200_preintegrate(pvio_b200_handle h, int n_factors, const int32_t *begin, const double *samples,
                           const double *t_end, const double *bias, const double *noise_cov, double *records);

/* ---- triangulation of new tracks ---------------------------------------------------------- */
/* Track::triangulate (pvio/src/pvio/map/track.cpp:83-106): multi-view DLT (stereo.h:67-75) and the parallax /
 * depth checks of triangulate_point_scored (stereo.h:104-128), one thread per track.
 * P: [n_frames][12] row-major 3x4 [R | T] of the CAMERA in each frame (R = q_wc^-1, T = -R p_wc, track.cpp:90-95).
 * begin / obs_frame / obs_z: CSR list of ALL keypoints of each track (>= 2 views).
 * points: [n_tracks][3] (q.hnormalized() when valid; otherwise the unit direction, whose sign is that of the
 * singular vector and as unspecified as in the reference, stereo.h:122-126);
 * valid: has_parallax; score: mean squared reprojection error in normalised coordinates. */
int pvio_b200_triangulate(pvio_b200_handle h, int n_frames, const double *P, int n_tracks, const int32_t *begin,
                          const int32_t *obs_frame, const double *obs_z, double *points, uint8_t *valid, double *score);

/* ---- diagnostics ---------------------------------------------------------------------- */
/* Device self-test of the Lie-group helpers (geometry/lie_algebra.h:25-42, lie_algebra.cpp:22-59,
 * quaternion_parameterization.h:28-31): w [n][3] rotation vectors; out [n][32] = expmap(w) (4, xyzw),
 * logmap(expmap(w)) (3), right_jacobian(w) (9, row-major), its inverse (9), Plus(expmap(w), w) (4), 3 spare.
 * Not part of the reference interface. */
int pvio_b200_selftest_lie(pvio_b200_handle h, int n, const double *w, double *out);

/* ---- KLT ----------------------------------------------------------------------------- */
/* Pyramidal Lucas-Kanade with OpenCV's semantics: winSize 21x21, maxLevel levels above
 * level 0, criteria COUNT+EPS (max_iter, eps), OPTFLOW_USE_INITIAL_FLOW (next_pts holds
 * the guess on entry), minEigThreshold 1e-4.  prev/next: 8-bit single-channel images
 * (post-CLAHE).  status[i] = 1 if found; err[i] as cv (mean abs patch difference). */
int pvio_b200_klt_track(pvio_b200_handle h, const uint8_t *prev, const uint8_t *next,
                        int width, int height, int stride, const float *prev_pts,
                        float *next_pts, uint8_t *status, float *err, int n_points,
                        int max_level, int max_iter, double eps);

#ifdef __cplusplus
}
#endif
#endif /* PVIO_B200_H */
