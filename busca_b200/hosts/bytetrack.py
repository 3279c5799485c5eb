"""A compact ByteTrack-with-BUSCA host tracker, written from scratch after the control flow of the reference adapter
``BYTETracker.update`` (adapters/ByteTrack/yolox/tracker/byte_tracker.py:226-456) so that the plug-in can be driven over
whole sequences - with the adapter's exact call pattern into BUSCA - on a box where the reference is not installed:

  per frame   3 x get_image_crops (second-round, first-round, all considered detections; byte_tracker.py:278-282)
              rounds 1 / 2: IoU cost (+ score fusion), linear assignment                     (:312-362)
              Step 3b:  T x get_image_crops with ONE box (the Kalman proposal of every unmatched track, :468-479),
                        center_distance(tracks, all considered detections) (:489),
                        associate_embeddings(...) (:491-502), decision reliable & p[t, D+t] > busca_thresh (:504-526),
                        winners that are still Tracked are updated at their Kalman box with update_mems=False (:385-391)
              unconfirmed tracks, new tracks, lost / removed bookkeeping, duplicate removal  (:399-444, 685-698)

Everything numeric that is not BUSCA's is injected (``iou_fn``, ``center_distance_fn``, ``reliable_fn``, ``rounds``) so the same
driver runs on the GPU library or, in the CPU tests, on the oracle.  With ``rounds=DeviceRounds(engine)`` the association rounds
themselves run on the device: batched Kalman predict / update, IoU cost + score fusion + assignment in one
call per round, duplicate removal.  The detection-coverage gate (8f row 2, ``reliable_thresh``) needs ``reliable_fn``;
camera-motion compensation (8f row 3) needs ``camera_motion_fn``: configs that enable them without it raise.

Pinned: tests/test_host_bytetrack.py replays tests/golden/adapter_seq.npz, produced by the UNMODIFIED reference adapter,
and requires identical track ids frame by frame.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import scipy.linalg
from scipy.optimize import linear_sum_assignment

NEW, TRACKED, LOST, REMOVED = 0, 1, 2, 3


class KalmanXYAH:
    """Constant-velocity filter on (cx, cy, aspect, height) and their velocities
    (adapters/.../mot_online/kalman_filter.py:22-225): noise proportional to the box height."""
    W_POS, W_VEL = 1.0 / 20, 1.0 / 160

    def __init__(self):
        self.F = np.eye(8)
        for i in range(4):
            self.F[i, 4 + i] = 1.0
        self.H = np.eye(4, 8)

    def initiate(self, z):
        mean = np.r_[z, np.zeros_like(z)]
        h = z[3]
        std = [2 * self.W_POS * h, 2 * self.W_POS * h, 1e-2, 2 * self.W_POS * h,
               10 * self.W_VEL * h, 10 * self.W_VEL * h, 1e-5, 10 * self.W_VEL * h]
        return mean, np.diag(np.square(std))

    def predict_many(self, mean, cov):
        h = mean[:, 3]
        std = np.r_[[self.W_POS * h, self.W_POS * h, 1e-2 * np.ones_like(h), self.W_POS * h],
                    [self.W_VEL * h, self.W_VEL * h, 1e-5 * np.ones_like(h), self.W_VEL * h]]
        q = np.square(std).T
        Q = np.asarray([np.diag(q[i]) for i in range(len(mean))])
        mean = np.dot(mean, self.F.T)
        left = np.dot(self.F, cov).transpose((1, 0, 2))
        return mean, np.dot(left, self.F.T) + Q

    def update(self, mean, cov, z):
        h = mean[3]
        std = [self.W_POS * h, self.W_POS * h, 1e-1, self.W_POS * h]
        pm = np.dot(self.H, mean)
        pc = np.linalg.multi_dot((self.H, cov, self.H.T)) + np.diag(np.square(std))
        chol, lower = scipy.linalg.cho_factor(pc, lower=True, check_finite=False)
        gain = scipy.linalg.cho_solve((chol, lower), np.dot(cov, self.H.T).T, check_finite=False).T
        return mean + np.dot(z - pm, gain.T), cov - np.linalg.multi_dot((gain, pc, gain.T))


def tlwh_to_xyah(tlwh):
    r = np.asarray(tlwh).copy()
    r[:2] += r[2:] / 2
    r[2] /= r[3]
    return r


class Track:
    """What BUSCA reads from a track or a detection: images_mem, tlwh_mem, scale, tlwh, tlbr (byte_tracker.py:23-161)."""

    def __init__(self, tlwh, score, image=None, scale=None):
        self._tlwh = np.asarray(tlwh, dtype=np.float64)
        self.mean = self.cov = None
        self.state = NEW
        self.is_activated = False
        self.score = score
        self.scale = scale
        self.tracklet_len = 0
        self.track_id = 0
        self.frame_id = self.start_frame = 0
        self.tlwh_mem = [self._tlwh.copy()]
        self.images_mem = [] if image is None else [image]

    @property
    def tlwh(self):
        if self.mean is None:
            return self._tlwh.copy()
        r = self.mean[:4].copy()
        r[2] *= r[3]
        r[:2] -= r[2:] / 2
        return r

    @property
    def tlbr(self):
        r = self.tlwh.copy()
        r[2:] += r[:2]
        return r

    def _absorb(self, other, update_mems):
        self.score, self.scale = other.score, other.scale
        if update_mems:
            self.tlwh_mem.extend(other.tlwh_mem)
            self.images_mem.extend(other.images_mem)


def apply_camera_motion(t: "Track", warp: np.ndarray):
    """STrack.apply_camera_motion + BYTETracker.warp_pos (byte_tracker.py:123-138, 660-664): the track centre (frame coordinates) goes
    through the 2x3 warp in FLOAT32 (the reference builds torch.Tensor([x, y, 1]) and multiplies with the float32 warp matrix) and is
    written back into the float64 state."""
    pos = (t._tlwh[:2] if t.mean is None else t.mean[:2]) * t.scale
    w = np.asarray(warp, np.float32)
    x, y = np.float32(pos[0]), np.float32(pos[1])
    # torch.mm's float32 summation order for a [2,3] x [3,1] product as measured on the reference run (torch 2.11 CPU; 5000 of 5000
    # random cases): (w1*y + w2*1) + w0*x, every operation rounded to float32
    new = ((w[:, 1] * y + w[:, 2]) + w[:, 0] * x).astype(np.float32) / np.float32(t.scale)
    if t.mean is None:
        t._tlwh[:2] = new
    else:
        t.mean[:2] = new


def assign(cost, thresh):
    """lap.lapjv(cost, extend_cost=True, cost_limit=thresh) as matching.linear_assignment uses it (matching.py:39-50):
    the optimum of the cost matrix extended by thresh/2 'unassigned' blocks.  Returns (matches[k,2], unmatched rows, unmatched cols)."""
    n, m = cost.shape
    if cost.size == 0:
        return np.empty((0, 2), dtype=int), list(range(n)), list(range(m))
    ext = np.full((n + m, n + m), thresh / 2.0)
    ext[n:, m:] = 0.0
    ext[:n, :m] = cost
    rows, cols = linear_sum_assignment(ext)
    x = np.full(n, -1, dtype=int)
    for r, c in zip(rows, cols):
        if r < n and c < m:
            x[r] = c
    matches = np.asarray([[i, j] for i, j in enumerate(x) if j >= 0], dtype=int).reshape(-1, 2)
    taken = set(matches[:, 1].tolist())
    return matches, [i for i in range(n) if x[i] < 0], [j for j in range(m) if j not in taken]


class DeviceRounds:
    """The association rounds on libbusca_b200 (csrc/rounds.cu) behind the four operations the driver needs."""

    def __init__(self, engine):
        self.e = engine

    def predict(self, mean, cov, tracked):
        return self.e.kalman_predict(mean, cov, tracked)

    def update(self, mean, cov, xyah):
        return self.e.kalman_update(mean, cov, xyah)

    def match(self, a_tlbr, b_tlbr, scores, thresh):
        x, _y, _ = self.e.match_round(a_tlbr, b_tlbr, scores, thresh)
        return x

    def duplicates(self, a_tlbr, a_age, b_tlbr, b_age):
        return self.e.duplicate_tracks(a_tlbr, a_age, b_tlbr, b_age, 0.15)


def _merge(a: List[Track], b: List[Track]) -> List[Track]:
    seen = {t.track_id for t in a}
    out = list(a)
    for t in b:
        if t.track_id not in seen:
            seen.add(t.track_id)
            out.append(t)
    return out


def _minus(a: List[Track], b: List[Track]) -> List[Track]:
    drop = {t.track_id for t in b}
    return [t for t in a if t.track_id not in drop]


class ByteTrackHost:
    def __init__(self, busca, args, iou_fn: Callable, center_distance_fn: Callable, frame_rate: int = 30,
                 reliable_fn: Optional[Callable] = None, camera_motion_fn: Optional[Callable] = None, rounds=None):
        """``busca``: object with get_image_crops / associate_embeddings (busca_b200.network.BUSCA).  ``args``: the
        tracker namespace of option.load_args_from_config plus track_thresh, track_buffer, match_thresh, mot20.
        ``iou_fn(a_tlbr[N,4], b_tlbr[M,4]) -> [N,M]`` (+1 convention), ``center_distance_fn(tracks, dets) -> [T,D]``,
        ``reliable_fn(frame_shape, tracks, p) -> bool`` (BYTETracker.is_reliable), ``camera_motion_fn(previous_frame, frame) ->
        2x3 float32 warp`` (cv2.findTransformECC, byte_tracker.py:626-657), ``rounds``: None (numpy / scipy on the host) or a DeviceRounds."""
        if hasattr(args, "reliable_thresh") and reliable_fn is None:
            raise NotImplementedError("the detection-coverage gate (reliable_thresh) needs reliable_fn")
        if getattr(args, "use_camera_motion_compensation", False) and camera_motion_fn is None:
            raise NotImplementedError("camera-motion compensation needs camera_motion_fn")
        self.busca, self.args = busca, args
        self.iou_fn, self.cdist_fn, self.reliable_fn, self.camera_motion_fn, self.rounds = iou_fn, center_distance_fn, reliable_fn, camera_motion_fn, rounds
        self.last_image = None
        self.det_thresh = args.track_thresh + 0.1
        self.max_time_lost = int(frame_rate / 30.0 * args.track_buffer)
        self.kf = KalmanXYAH()
        self.tracked: List[Track] = []
        self.lost: List[Track] = []
        self.removed: List[Track] = []
        self.frame_id = 0
        self._next_id = 0
        self.last_busca = None            # (matches, u_track, probs of the Kalman slots, reliable) of the last Step 3b

    # ---- helpers -------------------------------------------------------------------------------------------
    def _iou_cost(self, a: Sequence[Track], b: Sequence[Track]):
        if len(a) == 0 or len(b) == 0:
            return np.zeros((len(a), len(b)))
        return 1.0 - self.iou_fn(np.ascontiguousarray([t.tlbr for t in a], dtype=np.float64), np.ascontiguousarray([t.tlbr for t in b], dtype=np.float64))

    def _fuse(self, cost, dets):
        if cost.size == 0 or self.args.mot20:
            return cost
        scores = np.array([d.score for d in dets])
        return 1.0 - (1.0 - cost) * scores[None, :].repeat(cost.shape[0], axis=0)

    def _match(self, a: Sequence[Track], b: Sequence[Track], fuse: bool, thresh: float):
        """One round: IoU cost (+ score fusion) and the assignment with a cost limit -> (matches, unmatched a, unmatched b)."""
        if self.rounds is None or len(a) == 0 or len(b) == 0:
            cost = self._iou_cost(a, b)
            return assign(self._fuse(cost, b) if fuse else cost, thresh)
        scores = np.array([d.score for d in b], dtype=np.float64) if (fuse and not self.args.mot20) else None
        x = self.rounds.match(np.ascontiguousarray([t.tlbr for t in a], dtype=np.float64), np.ascontiguousarray([t.tlbr for t in b], dtype=np.float64),
                              scores, thresh)
        matches = np.asarray([[i, j] for i, j in enumerate(x) if j >= 0], dtype=int).reshape(-1, 2)
        taken = set(matches[:, 1].tolist())
        return matches, [i for i in range(len(a)) if x[i] < 0], [j for j in range(len(b)) if j not in taken]

    def _update_many(self, jobs):
        """jobs: (track, detection, update_mems, reactivate) of ONE round - every track at most once, so the Kalman updates are
        independent and run as one batch on the device."""
        if self.rounds is None or not jobs:
            for t, d, um, re in jobs:
                self._update(t, d, update_mems=um, reactivate=re)
            return
        mean, cov = self.rounds.update(np.asarray([t.mean for t, *_ in jobs]), np.asarray([t.cov for t, *_ in jobs]),
                                       np.asarray([tlwh_to_xyah(d.tlwh) for _, d, *_ in jobs]))
        for k, (t, d, um, re) in enumerate(jobs):
            t.mean, t.cov = mean[k], cov[k]
            t.tracklet_len = 0 if re else t.tracklet_len + 1
            t.state, t.is_activated, t.frame_id = TRACKED, True, self.frame_id
            t._absorb(d, um)

    def _activate(self, t: Track):
        self._next_id += 1
        t.track_id = self._next_id
        t.mean, t.cov = self.kf.initiate(tlwh_to_xyah(t._tlwh))
        t.tracklet_len = 0
        t.state = TRACKED
        t.is_activated = self.frame_id == 1
        t.frame_id = t.start_frame = self.frame_id

    def _update(self, t: Track, det: Track, update_mems: bool, reactivate: bool = False):
        t.mean, t.cov = self.kf.update(t.mean, t.cov, tlwh_to_xyah(det.tlwh))
        t.tracklet_len = 0 if reactivate else t.tracklet_len + 1
        t.state, t.is_activated, t.frame_id = TRACKED, True, self.frame_id
        t._absorb(det, update_mems)

    def _predict(self, pool: List[Track]):
        if not pool:
            return
        mean = np.asarray([t.mean.copy() for t in pool])
        cov = np.asarray([t.cov for t in pool])
        for i, t in enumerate(pool):
            if t.state != TRACKED:
                mean[i][7] = 0
        if self.rounds is not None:
            mean = np.asarray([t.mean for t in pool])            # the device zeroes the height velocity of non-Tracked tracks itself
            mean, cov = self.rounds.predict(mean, cov, np.array([t.state == TRACKED for t in pool], np.uint8))
        else:
            mean, cov = self.kf.predict_many(mean, cov)
        for t, m, c in zip(pool, mean, cov):
            t.mean, t.cov = m, c

    # ---- Step 3b -------------------------------------------------------------------------------------------
    def _kalman_candidates(self, pool: List[Track], frame):
        out = []
        for t in pool:                                       # one single-box crop call per track, as byte_tracker.py:468-479
            img = self.busca.get_image_crops(image=frame, bboxes=[t.tlbr * t.scale], normalize=False)[0]
            out.append(Track(t.tlwh, np.float32(0.10000001), image=img, scale=t.scale))
        return out

    def _third_round(self, pool: List[Track], considered: List[Track], kalman: List[Track]):
        a = self.args
        dists = self.cdist_fn(pool, considered)
        probs, reliable = self.busca.associate_embeddings(
            tracks_embeddings=pool, dets_embeddings=considered, dists_matrix=dists, seq_len=a.seq_len, num_candidates=a.num_candidates,
            use_broader_memory=a.use_broader_memory, extra_kalman_candidates=kalman, select_highest_candidate=a.select_highest_candidate,
            highest_candidate_minimum_thresh=getattr(a, "highest_candidate_minimum_thresh", None), plot_results=False, normalize_ims=True)
        if probs is None:
            self.last_busca = ([], list(range(len(pool))), np.zeros(0), np.zeros(0, bool))
            return [], list(range(len(pool)))
        D = len(considered)
        pk = np.array([probs[t, D + t] for t in range(len(pool))])
        matches = [[i, pk[i]] for i in range(len(pool)) if reliable[i] and pk[i] > a.busca_thresh]
        hit = {m[0] for m in matches}
        u = [i for i in range(len(pool)) if i not in hit]
        self.last_busca = (matches, u, pk, np.asarray(reliable).copy())
        return matches, u

    # ---- one frame -----------------------------------------------------------------------------------------
    def update(self, results: np.ndarray, img_info, img_size, current_frame=None) -> List[Track]:
        a = self.args
        self.frame_id += 1
        self.last_busca = None
        activated, refind, lost, removed = [], [], [], []
        results = np.asarray(results)
        scores, boxes = results[:, 4], results[:, :4].copy()
        scale = min(img_size[0] / float(img_info[0]), img_size[1] / float(img_info[1]))
        boxes /= scale
        first = scores > a.track_thresh
        second = np.logical_and(scores > 0.1, scores < a.track_thresh)
        considered = np.logical_or(first, second)
        use_busca = getattr(a, "use_busca", False) and getattr(a, "busca_thresh", 0) > 0
        if use_busca:
            im2 = self.busca.get_image_crops(image=current_frame, bboxes=boxes[second] * scale, normalize=False)
            im1 = self.busca.get_image_crops(image=current_frame, bboxes=boxes[first] * scale, normalize=False)
            imc = self.busca.get_image_crops(image=current_frame, bboxes=boxes[considered] * scale, normalize=False)
        else:
            im2, im1, imc = [None] * int(second.sum()), [None] * int(first.sum()), [None] * int(considered.sum())

        def make(mask, images):
            return [Track(np.r_[b[:2], b[2:] - b[:2]], s, image=im, scale=scale) for b, s, im in zip(boxes[mask], scores[mask], images)]

        dets1 = make(first, im1)
        dets_all = make(considered, imc) if len(boxes) else []
        unconfirmed = [t for t in self.tracked if not t.is_activated]
        confirmed = [t for t in self.tracked if t.is_activated]

        # round 1: confirmed + lost tracks vs high-score detections
        pool = _merge(confirmed, self.lost)
        self._predict(pool)
        matches, u_trk, u_det = self._match(pool, dets1, True, a.match_thresh)
        jobs = []
        for it, idet in matches:
            t, d = pool[it], dets1[idet]
            jobs.append((t, d, d.score >= self.det_thresh, t.state != TRACKED))
            (activated if t.state == TRACKED else refind).append(t)
        self._update_many(jobs)

        # round 2: still-tracked leftovers vs low-score detections
        dets2 = make(second, im2)
        r_tracked = [pool[i] for i in u_trk if pool[i].state == TRACKED]
        r_lost = [pool[i] for i in u_trk if pool[i].state != TRACKED]
        matches, u_trk2, _ = self._match(r_tracked, dets2, False, 0.5)
        mems2 = not getattr(a, "transformer_update_mems_only_first_round", False)
        self._update_many([(r_tracked[it], dets2[idet], mems2, False) for it, idet in matches])
        activated.extend(r_tracked[it] for it, _ in matches)
        unassigned = _merge([r_tracked[i] for i in u_trk2], r_lost)
        u_final = list(range(len(unassigned)))

        # Step 3b: BUSCA
        if use_busca and not (hasattr(a, "reliable_thresh") and not self.reliable_fn(current_frame.shape, self.tracked, a.reliable_thresh)):
            if getattr(a, "use_camera_motion_compensation", False) and self.frame_id > 1:
                warp = np.asarray(self.camera_motion_fn(self.last_image, current_frame), np.float32)
                for t in unassigned:
                    apply_camera_motion(t, warp)
            kalman = self._kalman_candidates(unassigned, current_frame)
            m3, u_final = self._third_round(unassigned, dets_all, kalman)
            winners = [it for it, _p in m3 if unassigned[it].state == TRACKED]   # Lost winners are dropped (byte_tracker.py:389)
            self._update_many([(unassigned[it], kalman[it], False, False) for it in winners])
            activated.extend(unassigned[it] for it in winners)
        for it in u_final:
            t = unassigned[it]
            if t.state != LOST:
                t.state = LOST
                lost.append(t)

        # unconfirmed tracks vs what round 1 left over
        rest = [dets1[i] for i in u_det]
        matches, u_unc, u_rest = self._match(unconfirmed, rest, True, 0.7)
        self._update_many([(unconfirmed[it], rest[idet], True, False) for it, idet in matches])
        activated.extend(unconfirmed[it] for it, _ in matches)
        for it in u_unc:
            unconfirmed[it].state = REMOVED
            removed.append(unconfirmed[it])
        for i in u_rest:
            if rest[i].score >= self.det_thresh:
                self._activate(rest[i])
                activated.append(rest[i])
        for t in self.lost:
            if self.frame_id - t.frame_id > self.max_time_lost:
                t.state = REMOVED
                removed.append(t)

        self.tracked = _merge(_merge([t for t in self.tracked if t.state == TRACKED], activated), refind)
        self.lost = _minus(self.lost, self.tracked)
        self.lost.extend(lost)
        self.lost = _minus(self.lost, self.removed)
        self.removed.extend(removed)
        self.removed = [t for t in self.removed if self.frame_id - t.frame_id < 10 * self.max_time_lost]
        self._dedupe()
        if getattr(a, "use_camera_motion_compensation", False):
            self.last_image = None if current_frame is None else np.copy(current_frame)
        return [t for t in self.tracked if t.is_activated]

    def _dedupe(self):
        if self.rounds is not None and self.tracked and self.lost:
            da, db = self.rounds.duplicates(np.ascontiguousarray([t.tlbr for t in self.tracked], dtype=np.float64),
                                            [t.frame_id - t.start_frame for t in self.tracked],
                                            np.ascontiguousarray([t.tlbr for t in self.lost], dtype=np.float64),
                                            [t.frame_id - t.start_frame for t in self.lost])
            self.tracked = [t for t, d in zip(self.tracked, da) if not d]
            self.lost = [t for t, d in zip(self.lost, db) if not d]
            return
        cost = self._iou_cost(self.tracked, self.lost)
        drop_a, drop_b = set(), set()
        for p, q in zip(*np.where(cost < 0.15)):
            ta, tb = self.tracked[p], self.lost[q]
            if ta.frame_id - ta.start_frame > tb.frame_id - tb.start_frame:
                drop_b.add(q)
            else:
                drop_a.add(p)
        self.tracked = [t for i, t in enumerate(self.tracked) if i not in drop_a]
        self.lost = [t for i, t in enumerate(self.lost) if i not in drop_b]
