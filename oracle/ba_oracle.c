/* TEST / BASELINE INFRASTRUCTURE ONLY -- dependency-free fp64 C restatement of ONE Gauss-Newton
 * iteration of PVIO's sliding-window BA on a reprojection-only window (BASELINE config 2), the way
 * the reference + Ceres perform it on the CPU:
 *   per residual block  ReprojectionErrorCost::Evaluate   estimation/ceres/reprojection_error_cost.h:40-120
 *   CauchyLoss(1.0) corrector                             estimation/bundle_adjustor.cpp:58 (ceres corrector.cc)
 *   local parameterisation (tangent Jacobians as is)      estimation/ceres/quaternion_parameterization.h:33-36
 *   Schur elimination of the inverse depths, dense Cholesky of the reduced camera system,
 *   back-substitution (what SPARSE_SCHUR does, solver_options.h:27), Jacobi scaling and the
 *   mu*diag regulariser of the dogleg Gauss-Newton step (Ceres 1.14 dogleg_strategy.cc),
 *   Plus (quaternion_parameterization.h:28-31) and the candidate cost.
 * Ceres / Eigen are not installed, so this is the timed "reference CPU path" of bench.py
 * (cpu_baseline.kind = "port").  It is checked against oracle/ba_oracle.py in tests/.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library.  Single-threaded per window like the reference (num_threads = 1,
 * solver_options.h:31); the batch entry point spreads independent windows over OpenMP threads.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

static void q2m(const double *q, double *R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}
static void mv(const double *A, const double *v, double *o) { for (int i = 0; i < 3; ++i) o[i] = A[3*i]*v[0] + A[3*i+1]*v[1] + A[3*i+2]*v[2]; }
static void mtv(const double *A, const double *v, double *o) { for (int i = 0; i < 3; ++i) o[i] = A[i]*v[0] + A[3+i]*v[1] + A[6+i]*v[2]; }
/* C(2x3) = A(2x3) * B(3x3) or B^T */
static void m23(const double *A, const double *B, int transB, double *C) {
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) {
        double s = 0; for (int k = 0; k < 3; ++k) s += A[3*i+k] * (transB ? B[3*j+k] : B[3*k+j]); C[3*i+j] = s; }
}
static void m23hat(const double *A, const double *v, double sgn, double *C) { /* C = sgn * A * hat(v) */
    for (int i = 0; i < 2; ++i) {
        const double a0 = A[3*i], a1 = A[3*i+1], a2 = A[3*i+2];
        C[3*i+0] = sgn * (a1 * v[2] - a2 * v[1]);
        C[3*i+1] = sgn * (a2 * v[0] - a0 * v[2]);
        C[3*i+2] = sgn * (a0 * v[1] - a1 * v[0]);
    }
}

typedef struct { double R[9], p[3]; } Pose;

/* reprojection_error_cost.h:40-120.  J: 2x13 = [q_tgt(3) p_tgt(3) q_ref(3) p_ref(3) rho] */
static void reproj(const Pose *tg, const Pose *rf, double rho, const double *zt, const double *zr,
                   const double *Rcs, const double *pcs, const double *W, double *r, double *J) {
    double yref[3] = {zr[0] / rho, zr[1] / rho, 1.0 / rho}, yrc[3], x[3], d[3], ytc[3], yt[3], t[3];
    mv(Rcs, yref, yrc); for (int i = 0; i < 3; ++i) yrc[i] += pcs[i];
    mv(rf->R, yrc, x); for (int i = 0; i < 3; ++i) { x[i] += rf->p[i]; d[i] = x[i] - tg->p[i]; }
    mtv(tg->R, d, ytc);
    for (int i = 0; i < 3; ++i) t[i] = ytc[i] - pcs[i];
    mtv(Rcs, t, yt);
    const double u0 = yt[0] / yt[2] - zt[0], u1 = yt[1] / yt[2] - zt[1];
    r[0] = W[0] * u0 + W[1] * u1; r[1] = W[2] * u0 + W[3] * u1;
    if (!J) return;
    const double iz = 1.0 / yt[2];
    double dp[6] = {iz, 0, -yt[0] * iz * iz, 0, iz, -yt[1] * iz * iz}, A[6], Dtc[6], Dx[6], Drc[6], T[6];
    for (int j = 0; j < 3; ++j) { A[j] = W[0] * dp[j] + W[1] * dp[3 + j]; A[3 + j] = W[2] * dp[j] + W[3] * dp[3 + j]; }
    m23(A, Rcs, 1, Dtc);            /* dr_dy_tgt_center = A * Rcs^T          :75  */
    m23(Dtc, tg->R, 1, Dx);         /* dr_dx = . * R_tgt^T                   :79  */
    m23(Dx, rf->R, 0, Drc);         /* dr_dy_ref_center = dr_dx * R_ref      :86  */
    m23hat(Dtc, ytc, 1.0, T);       /* :95 */
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) {
        J[13*i + j] = T[3*i + j];
        J[13*i + 3 + j] = -Dx[3*i + j];     /* :100 */
        J[13*i + 9 + j] = Dx[3*i + j];      /* :109 */
    }
    m23hat(Drc, yrc, -1.0, T);      /* :104 */
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 3; ++j) J[13*i + 6 + j] = T[3*i + j];
    double c[3]; mv(Rcs, yref, c);
    for (int i = 0; i < 2; ++i) J[13*i + 12] = -(Drc[3*i] * c[0] + Drc[3*i+1] * c[1] + Drc[3*i+2] * c[2]) / rho;  /* :113 */
}

static void quat_plus(const double *q, const double *w, double *o) {
    const double a = sqrt(w[0]*w[0] + w[1]*w[1] + w[2]*w[2]);
    double e[4] = {0, 0, 0, 1};
    if (a > 0) { const double s = sin(0.5 * a) / a; e[0] = s*w[0]; e[1] = s*w[1]; e[2] = s*w[2]; e[3] = cos(0.5 * a); }
    double t[4] = { q[3]*e[0] + q[0]*e[3] + q[1]*e[2] - q[2]*e[1], q[3]*e[1] + q[1]*e[3] + q[2]*e[0] - q[0]*e[2],
                    q[3]*e[2] + q[2]*e[3] + q[0]*e[1] - q[1]*e[0], q[3]*e[3] - q[0]*e[0] - q[1]*e[1] - q[2]*e[2] };
    const double n = 1.0 / sqrt(t[0]*t[0] + t[1]*t[1] + t[2]*t[2] + t[3]*t[3]);
    for (int i = 0; i < 4; ++i) o[i] = t[i] * n;
}

static double lm_reg(double hii, double mu) {
    const double s = 1.0 / (1.0 + sqrt(hii > 0 ? hii : 0)), s2 = s * s;
    double d2 = s2 * hii; if (d2 < 1e-6) d2 = 1e-6; if (d2 > 1e32) d2 = 1e32;
    return mu * d2 / s2;
}

static double window_cost(int N, int M, const Pose *P, const double *rho, const int32_t *anchor, const double *zref,
                          const int32_t *ob, const int32_t *of, const double *oz, const double *Rcs, const double *pcs,
                          const double *W, double b) {
    double cost = 0;
    for (int l = 0; l < M; ++l) for (int k = ob[l]; k < ob[l + 1]; ++k) {
        double r[2];
        reproj(&P[of[k]], &P[anchor[l]], rho[l], oz + 2 * k, zref + 2 * l, Rcs, pcs, W, r, 0);
        cost += 0.5 * b * log(1.0 + (r[0]*r[0] + r[1]*r[1]) / b);
    }
    return cost;
}

/* One GN iteration.  frames [N][16] (q xyzw, p, v, bg, ba), dx out [15N + M]. returns 0 or -1 (not SPD) */
int ba_oracle_gn_step(int N, int M, const uint8_t *fixed, const double *cam_q, const double *cam_p, const double *W,
                      double cauchy_a, const int32_t *anchor, const double *zref, const int32_t *ob, const int32_t *of,
                      const double *oz, const double *frames, const double *rho, double mu, double *dx, double *cost_out,
                      double *new_cost_out) {
    const size_t P6 = 6 * (size_t)N;
    const double b = cauchy_a * cauchy_a;
    double Rcs[9]; q2m(cam_q, Rcs);
    Pose *P = (Pose *)malloc(sizeof(Pose) * N);
    for (int f = 0; f < N; ++f) { q2m(frames + 16 * f, P[f].R); memcpy(P[f].p, frames + 16 * f + 4, 24); }
    double *H = (double *)calloc((size_t)P6 * P6, 8), *g = (double *)calloc(P6, 8), *hd = (double *)calloc(P6, 8);
    double *Hpl = (double *)calloc((size_t)M * P6, 8), *Hll = (double *)calloc(M, 8), *gl = (double *)calloc(M, 8);
    double cost = 0;
    for (int l = 0; l < M; ++l) {
        const int a = anchor[l];
        double *h = Hpl + (size_t)l * P6;
        for (int k = ob[l]; k < ob[l + 1]; ++k) {
            const int t = of[k];
            double r[2], J[26];
            reproj(&P[t], &P[a], rho[l], oz + 2 * k, zref + 2 * l, Rcs, cam_p, W, r, J);
            const double s = r[0]*r[0] + r[1]*r[1], tt = 1.0 + s / b, sc = sqrt(1.0 / tt);   /* corrector */
            cost += 0.5 * b * log(tt);
            r[0] *= sc; r[1] *= sc; for (int i = 0; i < 26; ++i) J[i] *= sc;
            const int col[2] = {6 * t, 6 * a};
            for (int bi = 0; bi < 2; ++bi) for (int i = 0; i < 6; ++i) {
                const double j0 = J[6 * bi + i], j1 = J[13 + 6 * bi + i];
                g[col[bi] + i] += j0 * r[0] + j1 * r[1];
                h[col[bi] + i] += j0 * J[12] + j1 * J[25];
                for (int bj = 0; bj < 2; ++bj) for (int j = 0; j < 6; ++j)
                    H[(size_t)(col[bi] + i) * P6 + col[bj] + j] += j0 * J[6 * bj + j] + j1 * J[13 + 6 * bj + j];
            }
            Hll[l] += J[12] * J[12] + J[25] * J[25];
            gl[l] += J[12] * r[0] + J[25] * r[1];
        }
    }
    for (size_t i = 0; i < P6; ++i) hd[i] = H[(size_t)i * P6 + i];
    /* Schur complement with the regularised 1x1 landmark blocks */
    for (int l = 0; l < M; ++l) {
        if (ob[l + 1] == ob[l]) { Hll[l] = 1.0; continue; }
        Hll[l] += lm_reg(Hll[l], mu);
        const double w = 1.0 / Hll[l];
        const double *h = Hpl + (size_t)l * P6;
        for (size_t i = 0; i < P6; ++i) {
            if (h[i] == 0.0) continue;
            const double hw = h[i] * w;
            g[i] -= hw * gl[l];
            for (size_t j = 0; j < P6; ++j) H[(size_t)i * P6 + j] -= hw * h[j];
        }
    }
    for (size_t i = 0; i < P6; ++i) H[(size_t)i * P6 + i] += lm_reg(hd[i], mu);
    for (size_t i = 0; i < P6; ++i) if (fixed[i / 6]) {
        for (size_t j = 0; j < P6; ++j) { H[(size_t)i * P6 + j] = 0; H[(size_t)j * P6 + i] = 0; }
        H[(size_t)i * P6 + i] = 1; g[i] = 0;
    }
    /* dense Cholesky + solve */
    int rc = 0;
    for (size_t k = 0; k < P6 && rc == 0; ++k) {
        double d = H[(size_t)k * P6 + k];
        for (size_t m = 0; m < k; ++m) d -= H[(size_t)k * P6 + m] * H[(size_t)k * P6 + m];
        if (!(d > 0)) { rc = -1; break; }
        d = sqrt(d); H[(size_t)k * P6 + k] = d;
        for (size_t i = k + 1; i < P6; ++i) {
            double s = H[(size_t)i * P6 + k];
            for (size_t m = 0; m < k; ++m) s -= H[(size_t)i * P6 + m] * H[(size_t)k * P6 + m];
            H[(size_t)i * P6 + k] = s / d;
        }
    }
    double *xp = (double *)calloc(P6, 8);
    if (rc == 0) {
        for (size_t i = 0; i < P6; ++i) { double s = -g[i]; for (size_t m = 0; m < i; ++m) s -= H[(size_t)i * P6 + m] * xp[m]; xp[i] = s / H[(size_t)i * P6 + i]; }
        for (int i = (int)P6 - 1; i >= 0; --i) { double s = xp[i]; for (size_t m = i + 1; m < P6; ++m) s -= H[(size_t)m * P6 + i] * xp[m]; xp[i] = s / H[(size_t)i * P6 + i]; }
    }
    memset(dx, 0, sizeof(double) * (15 * N + M));
    for (int f = 0; f < N; ++f) for (int i = 0; i < 6; ++i) dx[15 * f + i] = xp[6 * f + i];
    for (int l = 0; l < M; ++l) {
        if (ob[l + 1] == ob[l]) continue;
        const double *h = Hpl + (size_t)l * P6;
        double s = gl[l];
        for (size_t i = 0; i < P6; ++i) s += h[i] * xp[i];
        dx[15 * N + l] = -s / Hll[l];
    }
    /* Plus + candidate cost */
    Pose *Pc = (Pose *)malloc(sizeof(Pose) * N);
    double *rc_ = (double *)malloc(sizeof(double) * M);
    for (int f = 0; f < N; ++f) {
        double q[4]; quat_plus(frames + 16 * f, dx + 15 * f, q); q2m(q, Pc[f].R);
        for (int i = 0; i < 3; ++i) Pc[f].p[i] = frames[16 * f + 4 + i] + dx[15 * f + 3 + i];
    }
    for (int l = 0; l < M; ++l) rc_[l] = rho[l] + dx[15 * N + l];
    if (cost_out) *cost_out = cost;
    if (new_cost_out) *new_cost_out = window_cost(N, M, Pc, rc_, anchor, zref, ob, of, oz, Rcs, cam_p, W, b);
    free(P); free(Pc); free(rc_); free(H); free(g); free(hd); free(Hpl); free(Hll); free(gl); free(xp);
    return rc;
}

/* n_windows independent copies of the same window over a pool of POSIX threads (the image
 * has no libgomp); returns the number of threads used */
typedef struct {
    int n_windows, N, M; const uint8_t *fixed; const double *cam_q, *cam_p, *W; double cauchy_a;
    const int32_t *anchor; const double *zref; const int32_t *ob, *of; const double *oz, *frames, *rho; double mu;
    double *dx_all, *costs; volatile int next;
} BatchJob;

static void *batch_worker(void *arg) {
    BatchJob *j = (BatchJob *)arg;
    for (;;) {
        const int w = __sync_fetch_and_add(&j->next, 1);
        if (w >= j->n_windows) break;
        ba_oracle_gn_step(j->N, j->M, j->fixed, j->cam_q, j->cam_p, j->W, j->cauchy_a, j->anchor, j->zref, j->ob, j->of,
                          j->oz, j->frames, j->rho, j->mu, j->dx_all + (size_t)w * (15 * j->N + j->M), j->costs + 2 * w,
                          j->costs + 2 * w + 1);
    }
    return 0;
}

int ba_oracle_gn_step_batch(int n_windows, int n_threads, int N, int M, const uint8_t *fixed, const double *cam_q,
                            const double *cam_p, const double *W, double cauchy_a, const int32_t *anchor,
                            const double *zref, const int32_t *ob, const int32_t *of, const double *oz,
                            const double *frames, const double *rho, double mu, double *dx_all, double *costs) {
    if (n_threads <= 0) n_threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (n_threads > n_windows) n_threads = n_windows;
    if (n_threads < 1) n_threads = 1;
    BatchJob j = {n_windows, N, M, fixed, cam_q, cam_p, W, cauchy_a, anchor, zref, ob, of, oz, frames, rho, mu, dx_all, costs, 0};
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
    for (int t = 1; t < n_threads; ++t) pthread_create(&th[t], 0, batch_worker, &j);
    batch_worker(&j);
    for (int t = 1; t < n_threads; ++t) pthread_join(th[t], 0);
    free(th);
    return n_threads;
}
