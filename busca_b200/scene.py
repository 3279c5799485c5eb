"""Synthetic per-frame workloads of the hot path, shared by bench.py, the parity tests and the golden generator.

``Scene`` is ONE frame's worth of tracker state at a BASELINE.json scale: T unmatched tracks with L-deep histories, D
current-frame detections, one Kalman state per track.  It can be fed to
  * the reference / the oracle / the plug-in API as duck-typed tracks (``objects(crop_fn)``),
  * the device-resident entry point ``busca_frame_step_dev`` (``setup_resident`` / ``step_resident``),
and both must give the same answer: every box that reaches both paths is built so that the two derivations agree bit
for bit (detector boxes are float32-representable, as a detector emits them; ltwh = exact difference of the corners).
"""
from __future__ import annotations

from typing import Callable, List

import numpy as np

from . import synth

FILLER_LTWH = np.array([250.0, 250.0, 500.0, 500.0])     # incomplete history (network.py:304-308)


def predict_boxes(mean: np.ndarray, tracked: np.ndarray):
    """STrack.multi_predict mean + STrack.tlwh / tlbr (byte_tracker.py:50-61, 140-161) in numpy fp64, operation by
    operation (the device kernel uses the same IEEE operations, so the results are bit-equal)."""
    m = np.array(mean, dtype=np.float64, copy=True)
    m[~np.asarray(tracked, bool), 7] = 0.0
    m[:, :4] = m[:, :4] + m[:, 4:]
    tlwh = m[:, :4].copy()
    tlwh[:, 2] = tlwh[:, 2] * tlwh[:, 3]
    tlwh[:, 0] = tlwh[:, 0] - tlwh[:, 2] / 2
    tlwh[:, 1] = tlwh[:, 1] - tlwh[:, 3] / 2
    tlbr = tlwh.copy()
    tlbr[:, 2:] = tlbr[:, 2:] + tlbr[:, :2]
    return m, tlwh, tlbr


def scene_frames(seed: int, n_frames: int = 3):
    frames = [synth.make_frame(seed * 10 + 1)]
    for i in range(1, n_frames):
        frames.append(synth.next_frame(frames[-1], seed * 10 + 1 + i))
    return frames


class Scene:
    def __init__(self, T: int, D: int, L: int, C: int, seed: int, n_frames: int = 3, short_frac: float = 0.05, frames=None):
        """``frames``: optional list of frames to use instead of synthesising ``make_frame(seed*10+1)``, ... (benchmarks with
        many sequences share a few frame sets; boxes, histories and detections still follow ``seed``)."""
        rng = np.random.default_rng(seed)
        self.T, self.D, self.L, self.C, self.seed = T, D, L, C, seed
        self.frames = list(frames) if frames is not None else scene_frames(seed, n_frames)
        H, W = self.frames[0].shape[:2]
        box = synth.random_boxes(rng, T, H, W)                     # ltwh
        vel = rng.normal(0, 3, (T, 2))
        # Kalman state (cx, cy, a, h, vx, vy, va, vh) one step before the current frame
        self.mean = np.concatenate([box[:, :2] + box[:, 2:] / 2, (box[:, 2] / box[:, 3])[:, None], box[:, 3:4], vel,
                                    rng.normal(0, 1e-3, (T, 1)), rng.normal(0, 0.5, (T, 1))], axis=1)
        self.tracked = (rng.uniform(size=T) < 0.8)
        # history: L observations per track along its motion; a few tracks have a short history -> unreliable
        self.reliable = rng.uniform(size=T) >= short_frac
        hist = np.empty((T, L, 4))
        for i in range(L):
            b = box.copy()
            b[:, :2] -= vel * (L - i)
            b[:, 2:] *= 1 + 0.01 * rng.standard_normal((T, 2))
            hist[:, i] = b
        self.hist_ltwh = hist
        self.hist_tlbr = hist.copy()
        self.hist_tlbr[..., 2:] += self.hist_tlbr[..., :2]
        self.mem_ltwh = hist.copy()
        self.mem_ltwh[~self.reliable] = FILLER_LTWH
        # detections: half near the tracks, half elsewhere; corners float32-representable (a detector's output dtype)
        det = synth.random_boxes(rng, D, H, W)
        n_near = min(D, T) // 2
        det[:n_near] = box[:n_near] + np.concatenate([vel[:n_near] + rng.normal(0, 6, (n_near, 2)), np.zeros((n_near, 2))], 1)
        tlbr = det.copy()
        tlbr[:, 2:] += tlbr[:, :2]
        self.det_tlbr = tlbr.astype(np.float32).astype(np.float64)
        self.det_ltwh = self.det_tlbr.copy()
        self.det_ltwh[:, 2:] = self.det_tlbr[:, 2:] - self.det_tlbr[:, :2]     # exact
        self.pred_mean, self.pred_tlwh, self.pred_tlbr = predict_boxes(self.mean, self.tracked)
        self.model = None

    # ---- duck-typed objects (reference / oracle / plug-in API) -------------------------------------------------
    def objects(self, crop_fn: Callable, frame_index: int = 0):
        """(tracks, dets, kalman) as an adapter would hand them to BUSCA in Step 3b; ``crop_fn(frame, boxes) -> uint8 [N,384,128,3]``."""
        T, D, L = self.T, self.D, self.L
        frame = self.frames[frame_index]
        crops = crop_fn(self.frames[0], self.hist_tlbr.reshape(-1, 4)).reshape(T, L, 384, 128, 3)
        tracks = []
        for t in range(T):
            tr = synth.SynthTrack(self.pred_tlwh[t], scale=1.0)
            n = L if self.reliable[t] else L - 3
            tr.images_mem = [crops[t, i] for i in range(L - n, L)]
            tr.tlwh_mem = [self.hist_ltwh[t, i].copy() for i in range(L - n, L)]
            tracks.append(tr)
        dcrops = crop_fn(frame, self.det_tlbr.astype(np.float32)) if D else []
        dets = []
        for j in range(D):
            d = synth.SynthTrack(self.det_ltwh[j], scale=1.0)
            d.tlwh_mem = [d._tlwh.copy()]
            d.images_mem = [dcrops[j]]
            dets.append(d)
        kcrops = crop_fn(frame, self.pred_tlbr)
        kal = []
        for t in range(T):
            k = synth.SynthTrack(self.pred_tlwh[t], scale=1.0, score=0.10000001)
            k.images_mem = [kcrops[t]]
            kal.append(k)
        self._keepalive = (crops, dcrops, kcrops)
        return tracks, dets, kal

    # ---- device-resident path ----------------------------------------------------------------------------------
    def setup_resident(self, model, busca_thresh: float = 0.3, select_highest: bool = False, own_frame: bool = False):
        """``own_frame``: keep this scene's frame in its own HBM buffer (several scenes sharing one context)."""
        from ._lib import StepArgs
        self.model = model
        eng = model.engine
        T, D, L, C = self.T, self.D, self.L, self.C
        eng.upload_frame(self.frames[0])
        mem_slots = eng.alloc_slots(T * L).reshape(T, L)
        eng.crop(self.hist_tlbr.reshape(-1, 4), mem_slots.reshape(-1), to_host=False)
        mem_slots = mem_slots.copy()
        mem_slots[~self.reliable] = -1
        self.det_slots = eng.alloc_slots(D)
        self.kal_slots = eng.alloc_slots(T)
        a = StepArgs(T=T, D=D, L=L, C=C)
        a.track_mean_dev = eng.to_dev(self.mean)
        a.tracked_dev = eng.to_dev(self.tracked.astype(np.uint8))
        a.det_tlbr_dev = eng.to_dev(self.det_tlbr)
        a.mem_slots_dev = eng.to_dev(mem_slots.astype(np.int32))
        a.mem_ltwh_dev = eng.to_dev(self.mem_ltwh)
        a.det_slots_dev = eng.to_dev(self.det_slots)
        a.kal_slots_dev = eng.to_dev(self.kal_slots)
        a.busca_thresh = busca_thresh
        a.select_highest = int(select_highest)
        a.reliable_dev = eng.to_dev(self.reliable.astype(np.uint8))
        self.probs_dev = eng.dev_alloc(T * (C + 2) * 4)
        self.keep_dev = eng.dev_alloc(max(T, 16))
        self.cand_dev = eng.dev_alloc(T * C * 4)
        a.probs_dev = self.probs_dev
        a.keep_dev = self.keep_dev
        a.cand_dev = self.cand_dev
        if own_frame:
            f = np.ascontiguousarray(self.frames[0])
            a.frame_dev = eng.to_dev(f)
            a.frame_H, a.frame_W = f.shape[0], f.shape[1]
        self.step_args = a

    def step_resident(self):
        self.model.engine.frame_step_dev(self.step_args)

    def read_resident(self):
        eng = self.model.engine
        T, C = self.T, self.C
        return dict(probs=eng.from_dev(self.probs_dev, (T, C + 2), np.float32), keep=eng.from_dev(self.keep_dev, (T,), np.uint8).astype(bool),
                    cand=eng.from_dev(self.cand_dev, (T, C), np.int32))

    # ---- plug-in API path (host buffers) -----------------------------------------------------------------------
    def setup_e2e(self, model):
        self.model = model
        T, L = self.T, self.L
        crops = model.get_image_crops(self.frames[0], self.hist_tlbr.reshape(-1, 4), normalize=False).reshape(T, L, 384, 128, 3)
        self._hist_crops = crops
        self.tracks = []
        for t in range(T):
            tr = synth.SynthTrack(self.pred_tlwh[t], scale=1.0)
            n = L if self.reliable[t] else L - 3
            tr.images_mem = [crops[t, i] for i in range(L - n, L)]
            tr.tlwh_mem = [self.hist_ltwh[t, i].copy() for i in range(L - n, L)]
            self.tracks.append(tr)
        self.h2d = self.d2h = 0

    def step_e2e(self, i: int, busca_thresh: float = 0.3, details: bool = False):
        """One frame through the reference-facing API with host buffers: motion proposals, the D detection crops and the T
        Kalman crops (batched calls), center_distance, associate_embeddings, the decision (byte_tracker.py:504-526)."""
        from . import tracking
        m = self.model
        frame = self.frames[i % len(self.frames)]
        T, D, L, C = self.T, self.D, self.L, self.C
        _mo, tlwh, tlbr = m.engine.motion_proposals(self.mean, self.tracked)
        det_crops = m.get_image_crops(frame, self.det_tlbr.astype(np.float32), normalize=False)
        dets = []
        for j in range(D):
            d = synth.SynthTrack(self.det_ltwh[j], scale=1.0)
            d.tlwh_mem = [d._tlwh]
            d.images_mem = [det_crops[j]]
            dets.append(d)
        kal_crops = m.get_image_crops(frame, tlbr, normalize=False)
        kal = []
        for t in range(T):
            k = synth.SynthTrack(tlwh[t], scale=1.0)
            k.images_mem = [kal_crops[t]]
            kal.append(k)
            self.tracks[t]._tlwh = tlwh[t]
        dists = tracking.center_distance(tlbr, self.det_tlbr, engine=m.engine)
        pm, reliable = m.associate_embeddings(self.tracks, dets, dists, L, C, use_broader_memory=True, select_highest_candidate=False,
                                              extra_kalman_candidates=kal, normalize_ims=True)
        keep = reliable & (pm[np.arange(T), D + np.arange(T)] > busca_thresh)
        self.h2d = frame.nbytes + self.mean.nbytes + T + 2 * (D + T) * 32 + (T + D) * 32 + dists.nbytes + (T * L + D + T) * 4 + T * L * 32 + (D + T) * 32
        self.d2h = det_crops.nbytes + kal_crops.nbytes + T * 64 + dists.nbytes + T * (C + 2) * 4 + T * C * 4 + T * (C + 2) * 512 * 4 + T * 512 * 4
        if details:
            return dict(keep=keep, probs_matrix=pm, reliable=reliable, dists=dists, tlwh=tlwh, tlbr=tlbr)
        return keep
