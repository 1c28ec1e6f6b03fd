#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on its config, on N GPUs of one node.

Metric: Gauss-Newton iterations/sec on the 10-keyframe x 500-landmark window (BASELINE config 2:
reprojection-only, frames 0 and 1 fixed, K_res = 4500).  One "step" = one full GN iteration
(linearise all factors + Cauchy weights -> Schur-eliminate inverse depths -> solve the reduced
system -> back-substitute -> Plus -> candidate cost) over a batch of W independent cfg2 windows
per GPU; value = windows * iterations / second over all GPUs.  W = 4096 by default so the inputs
(377 MB) exceed the 126 MB L2.  Single-window latency (the reference's one-solve-at-a-time use),
the solve() call and KLT tracks/s are reported as extra keys of the same line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--windows W]
Multi-GPU: python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N (one rank per GPU;
windows are independent, so there is no data-path collective: scaling is weak, NCCL carries only
the start barrier and the max-over-ranks of the device time).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GN iters/sec on 10-KF x 500-landmark window (window-iterations/s)"
UNIT = "window-iterations/s"


def bytes_alg(N, M, K, D):
    """compulsory bytes of one GN iteration of one window in the fp32 device layout."""
    return 16 * K + 16 * M + 64 * N + 4 * (D * D + D) + 4 * M


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc = [], None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def cpu_reference(win, st, n_sample, threads=0):
    """The reference's CPU implementation of the path: oracle/ba_oracle.c (Ceres/Eigen are absent, so
    the fp64 C restatement stands in, kind = "port"), independent windows over all host threads."""
    from oracle import c_oracle
    c_oracle.gn_step_batch(win, st, 8, threads)          # warm-up / page-in
    t = time.perf_counter()
    _, _, used = c_oracle.gn_step_batch(win, st, n_sample, threads)
    dt = time.perf_counter() - t
    return n_sample / dt, used, dt


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    from synthetic import synth
    win, st, _ = synth.make_cfg2()
    ncpu = os.cpu_count() or 1
    n_sample = max(64, min(args.windows, 128 * ncpu))     # ~10-30 s of CPU work in total over the run
    times = []
    for i in range(args.warmup + args.steps):
        v, used, dt = cpu_reference(win, st, n_sample)
        if i >= args.warmup:
            times.append(dt)
    tot = sum(times)
    value = n_sample * args.steps / tot
    one, _, _ = cpu_reference(win, st, 32, threads=1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cfg2 (10 KF x 500 landmarks, K_res=4500, reprojection-only GN) x {n_sample} windows per step",
                   "windows_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": "port",
                         "sample": f"{n_sample} independent cfg2 window-iterations per step over all host threads; "
                                   "oracle/ba_oracle.c (fp64 restatement; Ceres/Eigen are not installed)",
                         "single_thread_value": one},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_gpu(args):
    rank, local_rank, world = dist_env()
    import torch
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from pvio_b200 import klt, pnp, imu
    from synthetic import synth
    from pvio_b200.bundle_adjustor import BundleAdjustor

    W = args.windows
    win, st, _ = synth.make_cfg2()
    N, M, K = win.N, win.M, win.K
    D = 6 * int(np.sum(win.frame_fixed == 0))
    stride = 15 * N + M
    ba = BundleAdjustor(device=local_rank, max_windows=W, max_frames=N, max_landmarks=512, max_obs=4608)
    ba.batch_set(0, win, st)
    ba.batch_replicate(W)
    ba.batch_upload(W)
    ba.sync()

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    launches0 = ba.kernel_launches
    for _ in range(max(args.warmup, 3)):
        ba.batch_gn_step(W, 1e-8, apply=False)
    ba.sync()
    ba.last_kernel_ms(-1)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    l0 = ba.kernel_launches
    ba.timer_start()
    for _ in range(args.steps):
        ba.batch_gn_step(W, 1e-8, apply=False)
    ms = ba.timer_stop()
    barrier()
    launches = ba.kernel_launches - l0
    lin_ms = ba.last_kernel_ms(1)
    clocks = sampler.stop() if sampler else None
    if use_dist:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        lt = torch.tensor([launches], device="cuda", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    else:
        ms_max = ms
    value = world * W * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C-ABI with HOST buffers: H2D of the packed batch + step + D2H of dx
    dx = np.zeros((W, stride)); costs = np.zeros((W, 2))
    ba.batch_gn_step_host(W, stride, 1e-8, dx, costs)
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ba.batch_gn_step_host(W, stride, 1e-8, dx, costs)
    e2e_s = time.perf_counter() - t0
    barrier()
    if use_dist:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * W * e2e_steps / e2e_s
    h2d = W * (8 * 4608 + 16 * 512 + 8 * 512 + 8 * 16 * N + 832 + 224)
    d2h = W * (8 * 15 * N + 8 * 512 + 168)

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (linearise + Schur), CUDA events around every launch
    balg = bytes_alg(N, M, K, D)
    peak, peak_src = measured_peak()
    achieved = balg * W / (lin_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "lin_schur_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp))["dram_bytes_per_window"] * W
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "lin_a_kernel + schur_kernel (linearise + Schur stage, timed together)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "bytes_alg_per_window": balg, "windows_per_launch": W, "kernel_ms": lin_ms,
                "kernel_share_of_step": lin_ms / (ms / args.steps),
                "note": "algorithmic bytes / CUDA-event time of the stage (two launches: linearisation 0.49 ms, Schur sum 0.34 ms; SURVEY's byte figure covers both); DRAM traffic of the stage = 264 KB per window: 55 KB of inputs (8-byte observation records: below the 92 KB algorithmic figure), 65 KB of sqrt(w) h records written, 139 KB read back by the Schur kernel, 5 KB of outputs (a deliberate compute-for-bandwidth trade: the records also save the update kernel one linearisation per observation, and HBM is at 7 % while the SMs are latency-bound); "
                        "arithmetic intensity ~33 flop/B is above the fp32 ridge (11.5 flop/B), so on CUDA cores the kernel is "
                        "FP32-issue bound, see DESIGN.md 4.1"}

    # ---- single window: latency of one GN iteration and of a whole solve() through the C-ABI
    ba1 = BundleAdjustor(device=local_rank, max_windows=1, max_frames=N, max_landmarks=512, max_obs=4608)
    ba1.batch_set(0, win, st); ba1.batch_upload(1)
    for _ in range(5):
        ba1.batch_gn_step(1, 1e-8, apply=False)
    ba1.sync(); ba1.timer_start()
    for _ in range(50):
        ba1.batch_gn_step(1, 1e-8, apply=False)
    us_iter = ba1.timer_stop() * 1e3 / 50
    ba1.solve(win, st, max_iterations=10)
    t0 = time.perf_counter()
    for _ in range(5):
        _, summ = ba1.solve(win, st, max_iterations=10)
    solve_ms = (time.perf_counter() - t0) * 1e3 / 5
    single = {"gn_iters_per_s": 1e6 / us_iter, "us_per_iteration": us_iter, "solve_call_ms": solve_ms,
              "solve_iterations": int(summ["iterations"]), "solve_iters_per_s_e2e": summ["iterations"] / (solve_ms * 1e-3)}
    # the window a running VIO solves at every keyframe (cfg3: 9 frames, 8 IMU factors, 120-dim prior), one at a time
    try:
        w3, s3, _ = synth.make_cfg3()
        ba3 = BundleAdjustor(device=local_rank, max_windows=1, max_frames=w3.N, max_landmarks=320, max_obs=2048)
        ba3.solve(w3, s3, max_iterations=10)
        t0 = time.perf_counter()
        for _ in range(5):
            _, summ3 = ba3.solve(w3, s3, max_iterations=10)
        single["inertial_window_solve_call_ms"] = (time.perf_counter() - t0) * 1e3 / 5
        single["inertial_window_iterations"] = int(summ3["iterations"])
        single["inertial_window_device_ms"] = summ3["solve_seconds"] * 1e3
        ba3.close()
    except Exception as e:          # an extra, never allowed to take the headline line down
        single["inertial_window_error"] = str(e)

    # ---- KLT tracks/s (752x480, 500 points, 21x21, 4 levels) through the C-ABI with host images
    prev, nxt, pts, _ = synth.make_klt_pair()
    klt.track_keypoints(ba1, prev, nxt, pts)
    t0 = time.perf_counter()
    for _ in range(20):
        klt.track_keypoints(ba1, prev, nxt, pts)
    klt_s = (time.perf_counter() - t0) / 20
    klt_info = {"tracks_per_s_e2e": len(pts) / klt_s, "ms_per_frame_pair": klt_s * 1e3, "points": int(len(pts))}
    # from the RAW frames: CLAHE(6, 8x8) on the device + pyramids + LK (OpenCvImage::preprocess + track_keypoints)
    klt.track_keypoints(ba1, prev, nxt, pts, clahe_clip=6.0)
    t0 = time.perf_counter()
    for _ in range(20):
        klt.track_keypoints(ba1, prev, nxt, pts, clahe_clip=6.0)
    raw_s = (time.perf_counter() - t0) / 20
    klt_info["raw_frames_tracks_per_s_e2e"] = len(pts) / raw_s
    klt_info["raw_frames_ms_per_frame_pair"] = raw_s * 1e3
    try:
        import cv2
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
        p0 = pts.reshape(-1, 1, 2)
        cv2.calcOpticalFlowPyrLK(prev, nxt, p0.copy(), p0.copy(), winSize=(21, 21), maxLevel=3, criteria=crit,
                                 flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        t0 = time.perf_counter()
        for _ in range(10):
            cv2.calcOpticalFlowPyrLK(prev, nxt, p0.copy(), p0.copy(), winSize=(21, 21), maxLevel=3, criteria=crit,
                                     flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        cv_s = (time.perf_counter() - t0) / 10
        klt_info["cv2_tracks_per_s"] = len(pts) / cv_s
        klt_info["cv2_threads"] = cv2.getNumThreads()
        cl = cv2.createCLAHE(6.0, (8, 8))
        t0 = time.perf_counter()
        for _ in range(10):
            a_, b_ = cl.apply(prev), cl.apply(nxt)
            cv2.calcOpticalFlowPyrLK(a_, b_, p0.copy(), p0.copy(), winSize=(21, 21), maxLevel=3, criteria=crit,
                                     flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        klt_info["cv2_raw_frames_tracks_per_s"] = len(pts) / ((time.perf_counter() - t0) / 10)
    except Exception as e:      # cv2 is the reference's KLT; report if it is unavailable
        klt_info["cv2_tracks_per_s"] = None
        klt_info["cv2_error"] = str(e)

    # ---- visual_inertial_pnp (150 points + IMU prior): one kernel launch per solve, host buffers in/out
    d = synth.make_pnp()
    pargs = (d['frame'], d['last'], d['imu'], d['pts'], d['zs'], d['cam_q'], d['cam_p'], d['imu_q'], d['imu_p'], d['W'], True)
    _, psum = pnp.visual_inertial_pnp(ba1, *pargs)
    t0 = time.perf_counter()
    for _ in range(20):
        _, psum = pnp.visual_inertial_pnp(ba1, *pargs)
    pnp_info = {"ms_per_solve_e2e": (time.perf_counter() - t0) * 1e3 / 20, "kernel_ms": psum["solve_seconds"] * 1e3,
                "iterations": int(psum["iterations"]), "points": int(len(d['pts']))}

    # ---- IMU pre-integration (PreIntegrator::integrate): 8 factors x 4096 windows, 40 samples each, host buffers in/out
    rng = np.random.default_rng(648)
    nf, ks = 32768, 40
    tt = np.arange(ks) / 200.0
    base = np.c_[tt, rng.normal(0, 0.3, (ks, 3)), rng.normal(0, 1.0, (ks, 3)) + np.array([0.0, 0.0, 9.81])]
    factors = [(base, tt[-1] + 0.004, np.zeros(3), np.zeros(3))] * nf
    noise = (np.eye(3) * 2.8791e-8, np.eye(3) * 4.0e-6, np.eye(3) * 3.7608e-10, np.eye(3) * 9.0e-6)
    begin = np.arange(nf + 1, dtype=np.int32) * ks
    samples = np.ascontiguousarray(np.tile(base, (nf, 1)))
    t_end = np.full(nf, tt[-1] + 0.004); bias = np.zeros((nf, 6)); rec = np.zeros((nf, 288))
    noise_a = np.ascontiguousarray(np.array([c.reshape(9) for c in noise]))
    import ctypes as C
    from pvio_b200 import _lib as L
    fn = ba1.lib.pvio_b200_preintegrate
    fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_double),
                   C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    call = lambda: fn(ba1.h, nf, L._ptr(begin, C.c_int32), L._ptr(samples, C.c_double), L._ptr(t_end, C.c_double),
                      L._ptr(bias, C.c_double), L._ptr(noise_a, C.c_double), L._ptr(rec, C.c_double))
    call()
    t0 = time.perf_counter()
    for _ in range(3):
        call()
    imu_s = (time.perf_counter() - t0) / 3
    from oracle import imu_oracle
    t0 = time.perf_counter()
    for _ in range(20):
        imu_oracle.integrate(base, tt[-1] + 0.004, np.zeros(3), np.zeros(3), *noise)
    imu_cpu = 20 / (time.perf_counter() - t0)
    imu_info = {"factors_per_s_e2e": nf / imu_s, "factors": nf, "samples_per_factor": ks,
                "numpy_oracle_factors_per_s_1core": imu_cpu}

    # ---- CPU baseline beside it (bounded sample, all host threads; plus one thread like num_threads=1)
    ncpu = os.cpu_count() or 1
    n_sample = max(256, 128 * ncpu)
    cpu_v, cpu_used, cpu_dt = cpu_reference(win, st, n_sample)
    cpu_one, _, _ = cpu_reference(win, st, 64, threads=1)
    cpu_baseline = {"value": cpu_v, "unit": UNIT, "cores": cpu_used, "kind": "port",
                    "sample": f"{n_sample} independent cfg2 window-iterations over all host threads "
                              f"({cpu_dt:.2f} s wall); oracle/ba_oracle.c, fp64, Ceres/Eigen unavailable",
                    "single_thread_value": cpu_one}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 Jacobians + f64 residuals/accumulation/solve", "data": "synthetic",
        "config": {"workload": f"cfg2 (10 KF x 500 landmarks, K_res=4500, reprojection-only GN, frames 0-1 fixed) x {W} "
                               "independent windows per GPU per step",
                   "windows_per_gpu": W, "l2_policy": f"device-resident inputs {(8 * K + 24 * M + 128 * N + 1056) * W / 1e6:.0f} MB + {(M // 32 + 1) * 32 * (6 * N + 2) * 4 * W / 1e6:.0f} MB of intermediate records per step exceed the 126 MB L2",
                   "parallelism": f"independent windows, {world} GPU(s), no data-path collective"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "note": "pvio_b200_batch_gn_step_host: pinned host buffers -> device, one GN iteration, dx back"},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "single_window": single, "klt": klt_info, "pnp": pnp_info, "imu_preintegration": imu_info,
    }
    print(json.dumps(line))
    ba.close(); ba1.close()
    if use_dist:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--windows", type=int, default=4096, help="independent cfg2 windows per GPU per step")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
