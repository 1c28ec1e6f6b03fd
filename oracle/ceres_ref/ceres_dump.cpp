// ceres_dump: runs the REFERENCE's own BundleAdjustor::solve (bundle_adjustor.cpp:308-319 -> ceres::Solve :249, with the
// reference's cost functors, parameterisation, loss and solver options) on windows exported by
// tests/golden/export_windows.py, and writes the state after k = 1 .. K iterations.  Needs Eigen3, Ceres and the PVIO
// library: it is NOT built in this image (neither is installed) -- see README.md.  Nothing here restates reference
// arithmetic: the program only moves numbers between flat files and the reference's Map / Frame / Track objects.
//
// Window file (little-endian): int32 header {N, M, K, use_inertial, n_imu, n_prior, iters}; then doubles in the order of
// export_windows.py: per frame 16 (q xyzw, p, v, bg, ba) + fixed flag; extrinsics cam q,p / imu q,p; K (fx fy cx cy);
// sqrt_inv_cov 2x2; per landmark anchor, z_ref, inverse depth, n_obs then (frame, z) pairs; per IMU factor the noise
// covariances, n samples, samples (t, w, a), t_end; prior: frames, S, e.
#include <cstdio>
#include <fstream>
#include <memory>
#include <string>
#include <vector>
#include <pvio/pvio.h>
#include <pvio/estimation/bundle_adjustor.h>
#include <pvio/estimation/factor.h>
#include <pvio/map/frame.h>
#include <pvio/map/map.h>
#include <pvio/map/track.h>

using namespace pvio;

struct Reader {
    std::ifstream f;
    explicit Reader(const std::string &p) : f(p, std::ios::binary) {}
    int32_t i32() { int32_t v; f.read(reinterpret_cast<char *>(&v), 4); return v; }
    double f64() { double v; f.read(reinterpret_cast<char *>(&v), 8); return v; }
    template <typename V> void vec(V &v, int n) { for (int i = 0; i < n; ++i) v(i) = f64(); }
};

class DumpConfig : public Config {          // only the solver limits are read by BundleAdjustor::solve
  public:
    size_t iters = 10;
    size_t solver_iteration_limit() const override { return iters; }
    double solver_time_limit() const override { return 1.0e6; }
    // the remaining pure virtuals of pvio::Config are not touched by solve(); give them the shipped EuRoC values
    vector<2> camera_resolution() const override { return {752, 480}; }
    matrix<3> camera_intrinsic() const override { return matrix<3>::Identity(); }
    quaternion camera_to_center_rotation() const override { return quaternion::Identity(); }
    vector<3> camera_to_center_translation() const override { return vector<3>::Zero(); }
    quaternion imu_to_center_rotation() const override { return quaternion::Identity(); }
    vector<3> imu_to_center_translation() const override { return vector<3>::Zero(); }
    matrix<2> keypoint_pixel_error_cov() const override { return matrix<2>::Identity(); }
    matrix<3> imu_gyro_white_noise() const override { return matrix<3>::Identity(); }
    matrix<3> imu_accel_white_noise() const override { return matrix<3>::Identity(); }
    matrix<3> imu_gyro_random_walk() const override { return matrix<3>::Identity(); }
    matrix<3> imu_accel_random_walk() const override { return matrix<3>::Identity(); }
};

static void build(Reader &r, Map &map, int32_t *hdr) {
    for (int i = 0; i < 7; ++i) hdr[i] = r.i32();
    const int N = hdr[0], M = hdr[1], use_inertial = hdr[3], n_imu = hdr[4], n_prior = hdr[5];
    std::vector<std::unique_ptr<Frame>> frames;
    for (int i = 0; i < N; ++i) {
        auto fr = std::make_unique<Frame>();
        vector<4> q; r.vec(q, 4);
        fr->pose.q = quaternion(q(3), q(0), q(1), q(2));
        r.vec(fr->pose.p, 3); r.vec(fr->motion.v, 3); r.vec(fr->motion.bg, 3); r.vec(fr->motion.ba, 3);
        if (r.f64() != 0.0) fr->flag(FrameFlag::FF_FIX_POSE) = true;
        frames.push_back(std::move(fr));
    }
    vector<4> q; vector<3> p;
    ExtrinsicParams cam, imu;
    r.vec(q, 4); r.vec(p, 3); cam.q_cs = quaternion(q(3), q(0), q(1), q(2)); cam.p_cs = p;
    r.vec(q, 4); r.vec(p, 3); imu.q_cs = quaternion(q(3), q(0), q(1), q(2)); imu.p_cs = p;
    vector<4> Kv; r.vec(Kv, 4);
    matrix<2> sic; sic(0, 0) = r.f64(); sic(0, 1) = r.f64(); sic(1, 0) = r.f64(); sic(1, 1) = r.f64();
    for (auto &fr : frames) {
        fr->camera = cam; fr->imu = imu; fr->sqrt_inv_cov = sic;
        fr->K = matrix<3>::Identity(); fr->K(0, 0) = Kv(0); fr->K(1, 1) = Kv(1); fr->K(0, 2) = Kv(2); fr->K(1, 2) = Kv(3);
    }
    std::vector<Frame *> fp;
    for (auto &fr : frames) { fp.push_back(fr.get()); map.put_frame(std::move(fr)); }
    for (int l = 0; l < M; ++l) {                 // tracks in first-visit order; keypoints are normalised coordinates
        const int anchor = r.i32();
        vector<2> z; r.vec(z, 2);
        const double rho = r.f64();
        const int n_obs = r.i32();
        fp[anchor]->append_keypoint(z);
        Track *t = fp[anchor]->get_track(fp[anchor]->keypoint_num() - 1, create_if_empty);
        for (int k = 0; k < n_obs; ++k) {
            const int f = r.i32();
            r.vec(z, 2);
            fp[f]->append_keypoint(z);
            t->add_keypoint(fp[f], fp[f]->keypoint_num() - 1);
        }
        t->landmark.inv_depth = rho;
        t->flag(TrackFlag::TF_VALID) = true;
    }
    for (int n = 0; n < n_imu; ++n) {             // raw samples: solve() re-integrates with frame_i's biases (:224)
        const int j = r.i32();
        PreIntegrator &pre = fp[j]->preintegration;
        for (matrix<3> *c : {&pre.cov_w, &pre.cov_a, &pre.cov_bg, &pre.cov_ba}) for (int a = 0; a < 9; ++a) (*c)(a / 3, a % 3) = r.f64();
        const int ns = r.i32();
        for (int s = 0; s < ns; ++s) { ImuData d; d.t = r.f64(); r.vec(d.w, 3); r.vec(d.a, 3); pre.data.push_back(d); }
        const double t_end = r.f64();
        pre.integrate(t_end, fp[j - 1]->motion.bg, fp[j - 1]->motion.ba, true, true);
    }
    if (use_inertial && n_prior > 0) {
        std::vector<Frame *> rel;
        for (int i = 0; i < n_prior; ++i) rel.push_back(fp[r.i32()]);
        const int d = 15 * n_prior;
        matrix<> S(d, d); vector<> e(d);
        for (int i = 0; i < d; ++i) for (int k = 0; k < d; ++k) S(i, k) = r.f64();
        for (int i = 0; i < d; ++i) e(i) = r.f64();
        map.set_marginalization_factor(Factor::create_marginalization_error(S, e, std::move(rel)));   // snapshots pose_0 / motion_0 now
    }
}

int main(int argc, char **argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: ceres_dump <window.bin> <out.bin>\n"); return 2; }
    int32_t hdr[7];
    { Reader probe(argv[1]); for (int i = 0; i < 7; ++i) hdr[i] = probe.i32(); }
    std::ofstream out(argv[2], std::ios::binary);
    const int K = hdr[6];
    for (int k = 1; k <= K; ++k) {                // ceres is deterministic: k iterations from the same start = state after iteration k
        Map map;
        Reader r(argv[1]);
        build(r, map, hdr);
        DumpConfig cfg; cfg.iters = (size_t)k;
        BundleAdjustor ba;
        const bool usable = ba.solve(&map, &cfg, hdr[3] != 0);
        const double u = usable ? 1.0 : 0.0;
        out.write(reinterpret_cast<const char *>(&u), 8);
        for (size_t i = 0; i < map.frame_num(); ++i) {
            const Frame *f = map.get_frame(i);
            const double s[16] = {f->pose.q.x(), f->pose.q.y(), f->pose.q.z(), f->pose.q.w(), f->pose.p(0), f->pose.p(1), f->pose.p(2),
                                  f->motion.v(0), f->motion.v(1), f->motion.v(2), f->motion.bg(0), f->motion.bg(1), f->motion.bg(2),
                                  f->motion.ba(0), f->motion.ba(1), f->motion.ba(2)};
            out.write(reinterpret_cast<const char *>(s), sizeof(s));
        }
        for (size_t t = 0; t < map.track_num(); ++t) { const double rho = map.get_track(t)->landmark.inv_depth; out.write(reinterpret_cast<const char *>(&rho), 8); }
    }
    return 0;
}
