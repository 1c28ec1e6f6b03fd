"""Sliding-window SEQUENCE on top of one long synthetic run (test harness: plays the role of PVIO's
SlidingWindowTracker / Map bookkeeping, core/sliding_window_tracker.cpp:75-125, map/map.cpp:76-88, map/track.cpp:42-49).

A run of F keyframes (synth.make_cfg3 with N = F: trajectory, landmarks with their track frames, one IMU factor per
consecutive pair) is cut into windows of N frames.  Per keyframe the caller (a) solves the window, (b) marginalises its
oldest frame into a prior over the remaining N - 1 frames, (c) shifts: drops the oldest frame, appends the next one.
`Chain` carries one estimator's state along the run (frame states, inverse depths with their anchors, the prior), so a
GPU chain and an oracle chain can be advanced side by side with their own numbers."""
import dataclasses

import numpy as np

from pvio_b200.window import Window, State
from . import so3, synth


class Run:
    def __init__(self, F=26, N=6, M=260, seed=700):
        self.F, self.N = F, N
        self.w, self.guess, self.truth = synth.make_cfg3(seed=seed, N=F, M=M, prior='none')
        w = self.w
        self.tracks = []                      # per landmark: sorted frame list and keypoints
        for l in range(w.M):
            fr = [int(w.lm_anchor[l])] + [int(f) for f in w.obs_frame[w.lm_obs_begin[l]:w.lm_obs_begin[l + 1]]]
            zs = [w.lm_z_ref[l]] + [w.obs_z[k] for k in range(w.lm_obs_begin[l], w.lm_obs_begin[l + 1])]
            self.tracks.append((fr, zs))

    def cam_pose(self, q, p):
        return so3.qmul(q, self.w.cam_q_cs), p + so3.qrot(q, self.w.cam_p_cs)


class Chain:
    """One estimator's view of the run: absolute frame states [F], per-landmark (anchor frame, inverse depth), prior."""

    def __init__(self, run):
        self.run = run
        g = run.guess
        self.q, self.p, self.v, self.bg, self.ba = g.q.copy(), g.p.copy(), g.v.copy(), g.bg.copy(), g.ba.copy()
        self.anchor = np.array([t[0][0] for t in run.tracks])
        self.rho = g.rho.copy()
        self.prior = None                     # dict(frames (absolute), S, e, x0 [n][16])

    def _reanchor(self, l, new_anchor):
        """Track::first_frame changes when the old anchor leaves the window: the inverse depth is re-expressed in the next
        observing frame from the CURRENT estimate (map/track.cpp:42-49)."""
        run = self.run
        fr, zs = run.tracks[l]
        a = int(self.anchor[l])
        qc, pc = run.cam_pose(self.q[a], self.p[a])
        z = zs[fr.index(a)]
        x = so3.qrot(qc, np.array([z[0], z[1], 1.0]) / self.rho[l]) + pc
        qn, pn = run.cam_pose(self.q[new_anchor], self.p[new_anchor])
        y = so3.qrot(so3.qconj(qn), x - pn)
        self.anchor[l] = new_anchor
        self.rho[l] = 1.0 / y[2]

    def window(self, k):
        """Window over absolute frames [k, k + N): (Window, State, landmark ids)."""
        run, N = self.run, self.run.N
        src = run.w
        w = dataclasses.replace(src)
        w.N = N
        w.frame_fixed = np.zeros(N, dtype=np.uint8)
        lm_ids, anchors, zref, begins, of, oz, victim, rho = [], [], [], [0], [], [], [], []
        for l, (fr, zs) in enumerate(run.tracks):
            inw = [(f, z) for f, z in zip(fr, zs) if k <= f < k + N]
            if len(inw) < 2:
                continue
            if self.anchor[l] != inw[0][0]:
                if self.anchor[l] < inw[0][0]:
                    self._reanchor(l, inw[0][0])
                else:
                    continue
            lm_ids.append(l)
            anchors.append(inw[0][0] - k)
            zref.append(inw[0][1])
            for f, z in inw[1:]:
                of.append(f - k)
                oz.append(z)
            begins.append(len(of))
            victim.append(1 if inw[0][0] == k else 0)
            rho.append(self.rho[l])
        w.M, w.K = len(lm_ids), len(of)
        w.lm_anchor = np.array(anchors, dtype=np.int32)
        w.lm_z_ref = np.array(zref).reshape(-1, 2)
        w.lm_obs_begin = np.array(begins, dtype=np.int32)
        w.obs_frame = np.array(of, dtype=np.int32)
        w.obs_z = np.array(oz).reshape(-1, 2)
        w.lm_in_victim = np.array(victim, dtype=np.uint8)
        sl = slice(k, k + N - 1)
        w.n_imu = N - 1
        w.imu_frame_i = np.arange(0, N - 1, dtype=np.int32)
        w.imu_frame_j = np.arange(1, N, dtype=np.int32)
        for name in ('imu_dt', 'imu_dq', 'imu_dp', 'imu_dv', 'imu_sqrt_inv_cov', 'imu_dq_dbg', 'imu_dp_dbg', 'imu_dp_dba',
                     'imu_dv_dbg', 'imu_dv_dba', 'imu_bg0', 'imu_ba0'):
            setattr(w, name, getattr(src, name)[sl].copy())
        st = State(self.q[k:k + N].copy(), self.p[k:k + N].copy(), self.v[k:k + N].copy(), self.bg[k:k + N].copy(),
                   self.ba[k:k + N].copy(), np.array(rho))
        if self.prior is None:
            synth.gauge_prior(w, st)          # first window: 1e15 on the pose of frame 0 (sliding_window_tracker.cpp:100-112)
        else:
            pr = self.prior
            w.n_prior = len(pr['frames'])
            w.prior_frames = (np.array(pr['frames']) - k).astype(np.int32)
            w.prior_S, w.prior_e = pr['S'], pr['e']
            x0 = pr['x0']
            w.prior_q0, w.prior_p0, w.prior_v0 = x0[:, 0:4].copy(), x0[:, 4:7].copy(), x0[:, 7:10].copy()
            w.prior_bg0, w.prior_ba0 = x0[:, 10:13].copy(), x0[:, 13:16].copy()
        w.n_planes, w.n_ptracks = 0, 0
        w.validate()
        return w, st, lm_ids

    def store(self, k, st, lm_ids):
        N = self.run.N
        self.q[k:k + N], self.p[k:k + N], self.v[k:k + N] = st.q, st.p, st.v
        self.bg[k:k + N], self.ba[k:k + N] = st.bg, st.ba
        for i, l in enumerate(lm_ids):
            self.rho[l] = st.rho[i]

    def set_prior(self, k, S, e, st):
        """The prior produced by marginalising frame k: over absolute frames k + 1 .. k + N - 1, linearised at st."""
        N = self.run.N
        x0 = np.concatenate([st.q[1:], st.p[1:], st.v[1:], st.bg[1:], st.ba[1:]], axis=1)
        self.prior = dict(frames=list(range(k + 1, k + N)), S=S, e=e, x0=x0)


class ResidentPlayer:
    """Feeds a run into the RESIDENT window of a handle (pvio_b200.resident.ResidentWindow): what the shim does with the
    resident API -- per keyframe: solve, drop the victim, append the next frame with its observations."""

    def __init__(self, ba, run):
        from pvio_b200 import _lib
        from pvio_b200.resident import ResidentWindow
        self.run, self.rw, self.ids = run, ResidentWindow(ba, run.w), {}
        g = run.guess
        self.rec = _lib.PackedArgs(run.w, g).keep["imu"].reshape(-1, 288)     # factor f couples absolute frames (f, f + 1)
        for f in range(run.N):
            self.rw.append_frame(self.state16(f), False, self.rec[f - 1] if f > 0 else None)
            self.feed(f, f)
        w0, _, _ = Chain(run).window(0)                                       # first window: the 1e15 gauge prior
        x0 = np.concatenate([w0.prior_q0, w0.prior_p0, w0.prior_v0, w0.prior_bg0, w0.prior_ba0], axis=1)
        self.rw.set_prior(w0.prior_S, w0.prior_e, x0)
        self.seconds = 0.0                                                    # wall time spent inside the C-ABI calls

    def state16(self, f):
        g = self.run.guess
        return np.concatenate([g.q[f], g.p[f], g.v[f], g.bg[f], g.ba[f]])

    def feed(self, f_abs, f_win):
        new, seen = [], []
        for l, (fr, zs) in enumerate(self.run.tracks):
            if f_abs in fr:
                (seen if l in self.ids else new).append((l, zs[fr.index(f_abs)]))
        if seen:
            self.rw.add_observations([self.ids[l] for l, _ in seen], [f_win] * len(seen), [z for _, z in seen])
        if new:
            got = self.rw.add_tracks([f_win] * len(new), [z for _, z in new], [self.run.guess.rho[l] for l, _ in new])
            for (l, _), i in zip(new, got):
                self.ids[l] = int(i)

    def solve(self, max_iterations):
        import time
        t = time.perf_counter()
        s = self.rw.solve(max_iterations=max_iterations)
        self.seconds += time.perf_counter() - t
        return s

    def shift(self, k):
        """drop frame k (the oldest), append absolute frame k + N"""
        import time
        f_abs = k + self.run.N
        t = time.perf_counter()
        self.rw.drop_victim()
        self.rw.append_frame(self.state16(f_abs), False, self.rec[f_abs - 1])
        self.seconds += time.perf_counter() - t
        self.feed(f_abs, self.run.N - 1)
