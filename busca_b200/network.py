"""Host mirror of the reference's ``busca.network.BUSCA`` (busca/network.py:11-507): the object every adapter
constructs and calls.  Same constructor argument, same public methods, same return types; everything
numeric runs in libbusca_b200.so on the B200 (no CPU fallback).

    tracker = BUSCA(args.transformer).to(device); tracker.load_pretrained(ckpt, ignore_reid_fc=True); tracker.eval()
    crops  = tracker.get_image_crops(image=frame, bboxes=boxes, normalize=False)        # uint8 [N,384,128,3]
    probs, reliable = tracker.associate_embeddings(tracks, dets, dists, seq_len, num_candidates, ...)

Patches never have to leave the GPU: ``get_image_crops`` returns a real numpy array (adapters store its rows in
``track.images_mem``) AND remembers which patch-bank slot holds each row; ``associate_embeddings`` resolves every
``images_mem`` entry back to its slot by host address and only uploads arrays it has never seen.

Contract inherited from the adapters (byte_tracker.py:40-42, 117-121: crops are appended, never edited): a crop row handed
out by ``get_image_crops`` is IMMUTABLE - the registry maps its address to the device patch, so writing into the row
afterwards would not reach the GPU.  Copy the row (a copy is an unknown array and is uploaded) to change pixels.
"""
from __future__ import annotations

import os

import weakref
from typing import Dict, List, Optional

import numpy as np

from . import custom_layers, tracking
from .engine import PATCH_BYTES, PATCH_SHAPE, Engine


def _device_index(device) -> int:
    if device is None:
        return 0
    if not isinstance(device, (str, int)):
        idx = getattr(device, "index", None)              # torch.device
        if isinstance(idx, int):
            return idx
    if isinstance(device, int):
        return device
    s = str(device)
    if s.startswith("cpu"):
        raise RuntimeError("busca_b200 runs on a B200 only (args.device is '%s'); there is no CPU path" % s)
    return int(s.split(":")[1]) if ":" in s else 0


class _PatchRegistry:
    """host address of a crop row  ->  patch-bank slot.  Slots are released when the numpy array that
    ``get_image_crops`` returned (the base of every row view) is garbage collected."""

    def __init__(self, engine: Engine):
        self.engine = engine
        self.slot_of: Dict[int, int] = {}
        self._by_id: Dict[int, tuple] = {}

    def register(self, arr: np.ndarray, slots: np.ndarray, owner=None):
        """``owner``: the object that outlives every view of ``arr`` (the pinned block for pooled arrays; ``arr`` itself
        when it owns its data - numpy points a row view's ``.base`` at the owner, not at ``arr``)."""
        base = arr.ctypes.data
        sl = slots.tolist()
        keys = [base + i * PATCH_BYTES for i in range(len(sl))]
        self.slot_of.update(zip(keys, sl))
        weakref.finalize(arr if owner is None else owner, _PatchRegistry._release, weakref.ref(self), keys, sl)

    @staticmethod
    def _release(self_ref, keys, slots):
        self = self_ref()
        if self is None or self.engine.h is None:
            return
        for k in keys:
            self.slot_of.pop(k, None)
        self.engine.free_slots(slots)
        # (the per-object cache entries died with their arrays: the owner outlives every view)

    def lookup(self, patch: np.ndarray) -> Optional[int]:
        if patch.dtype != np.uint8 or patch.shape != PATCH_SHAPE or not patch.flags["C_CONTIGUOUS"]:
            return None
        return self.slot_of.get(patch.ctypes.data)

    def lookup_cached(self, patch) -> Optional[int]:
        """``lookup`` for the per-frame loops: adapters keep the SAME array objects in ``track.images_mem`` from frame
        to frame, so the slot is remembered per object (id + weak reference, so a recycled id can never alias)."""
        hit = self._by_id.get(id(patch))
        if hit is not None and hit[0]() is patch:
            return hit[1]
        if not isinstance(patch, np.ndarray):
            return None
        s = self.lookup(patch)
        if s is not None:
            key = id(patch)
            by_id = self._by_id
            self._by_id[key] = (weakref.ref(patch, lambda _r, key=key, by_id=by_id: by_id.pop(key, None)), s)
        return s


class BUSCA:
    def __init__(self, args):
        self.args = args
        self.dim_embedding = args.dim_embedding
        self.dim_model = args.trans_dim
        if args.input_flavour != "MEM-SEP-CAN-BAD" or args.output_flavour != "CAN" or not args.encode_separator_as_reference \
                or args.encode_special_tokens:
            # every shipped YAML uses this one flavour; CLS-* flavours are broken in the reference (encodings.py:161)
            raise NotImplementedError("busca_b200 implements input_flavour MEM-SEP-CAN-BAD / output_flavour CAN only")
        self.activation = custom_layers.effective_activation(args.activation, getattr(args, "follow_reference_activation", True))
        self.precision = getattr(args, "precision", "fp32")
        self.legacy_float64_sentinel = bool(getattr(args, "legacy_float64_sentinel", True))
        self.engine = Engine(device=_device_index(getattr(args, "device", None)), d_model=self.dim_model, nhead=args.nhead,
                             ff_size=args.ff_size, num_layers=args.num_layer, activation=self.activation,
                             precision=self.precision, sentinel_fp64=self.legacy_float64_sentinel,
                             bank_slots=int(getattr(args, "bank_slots", 2048)))
        # opt-in (args.defer_crop_copies / BUSCA_DEFER_CROP_COPIES=1): get_image_crops returns before the device->host copy of the crops
        # has landed; the returned arrays hold valid bytes after the next BUSCA call that waits for the device (associate_embeddings,
        # center_distance, sync()).  For adapters that only STORE the crops in between - every shipped one does (byte_tracker.py:278-282,
        # 468-479) - it takes the per-call wait out of the T single-box crop calls.  Off by default: strict reference semantics.
        self.verify_patches = bool(getattr(args, "verify_patches", False)) or os.environ.get("BUSCA_VERIFY_PATCHES") == "1"
        if bool(getattr(args, "defer_crop_copies", False)) or os.environ.get("BUSCA_DEFER_CROP_COPIES") == "1":
            self.engine.set_option("defer_crop_copies", 1)
        self.expected_image_size = (384, 128)           # ReID_Encoder.PRETRAINED_SIZE (network.py:512)
        self._registry = _PatchRegistry(self.engine)
        self.attentions = None
        self.logits = None
        self.mem_logits = None
        self.store_logits = True
        self.training = False

    # nn.Module-ish surface the adapters touch (byte_tracker.py:217-221)
    def to(self, device):
        return self

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    # ---- weights: load_pretrained (network.py:432-467) -------------------------------------------------
    def load_pretrained(self, path, ignore_reid=False, ignore_reid_fc=False):
        import torch
        state_dict = torch.load(path, map_location="cpu")
        sd = state_dict["model_state_dict"] if "model_state_dict" in state_dict else state_dict
        self.load_state_dict(sd, ignore_reid=ignore_reid, ignore_reid_fc=ignore_reid_fc)

    def load_state_dict(self, sd, ignore_reid=False, ignore_reid_fc=False):
        if ignore_reid:
            raise NotImplementedError("ignore_reid=True needs separately loaded ReID weights (model_feats.pth); "
                                      "pass them in the same dict")
        keep = {}
        for k, v in sd.items():
            if "reid_encoder.model.fc." in k or "reid_encoder.model.fc_person." in k:
                continue                                  # the classifier head is never used on this path
            if k == "cls_token":
                print("WARNING: Loading a model with a cls_token, but the current model does not have a cls_token. "
                      "The cls_token will be ignored")
                continue
            if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
                continue                                  # train-mode BN never reads them (network.py:553-556)
            keep[k] = v
        self.engine.load_state_dict(keep)

    # ---- crops: get_image_crops (network.py:492-507) ----------------------------------------------------
    def _ensure_frame(self, image: np.ndarray, boxes: Optional[np.ndarray] = None):
        """Make sure the frame in HBM shows, inside ``boxes``, exactly the pixels of ``image``.

        The adapters call get_image_crops 3 + T times per frame with the same image (byte_tracker.py:278-282, 468-479);
        re-uploading 6 MB each time would dominate.  The library keeps a page-locked host mirror of the frame it holds and
        compares, byte for byte, the rows the crops read (the union of the boxes for a single-box call, else the whole
        frame): equal -> the device frame is valid for these crops by construction (no sampling, no reliance on buffer
        identity); different -> upload."""
        self.engine.sync_frame(image, boxes if boxes is not None and 0 < len(boxes) <= 8 else None)

    def ingest_frame(self, detector_tensor, means, std, height=None, width=None, to_host=True):
        """What the evaluators do before tracker.update (adapters/ByteTrack/yolox/evaluators/mot_evaluator.py:198-204), on the device:
        the detector's normalised RGB CHW float32 input becomes the uint8 BGR frame the crops read - already in HBM, so the following
        get_image_crops calls upload nothing.  ``detector_tensor``: numpy [3,H,W] float32, or the integer device address of such a tensor
        on this engine's GPU (``imgs[0].data_ptr()``, with height / width).  Returns the host copy the adapter hands on as
        ``current_frame`` (``to_host=False``: None; pass the frame's shape with ``device_frame_shape`` instead)."""
        return self.engine.ingest_frame(detector_tensor, means, std, height, width, to_host=to_host)

    def _verify_patches(self, tracks, dets, kalman, mem_slots, det_slots, kal_slots, L, broader):
        """Debug aid (args.verify_patches / BUSCA_VERIFY_PATCHES=1): the patch bank is addressed through the HOST arrays the adapter keeps in
        ``images_mem`` (a crop is uploaded / gathered once, then found again by its address), which assumes the adapter never rewrites a
        stored crop in place - none of the shipped ones does.  This check downloads every slot the call is about to read and compares it
        with the bytes of the host array it stands for; a mismatch raises instead of silently associating on stale pixels."""
        pairs = []
        for t, track in enumerate(tracks):
            sel = self._memory_indices(len(track.images_mem), L, broader)
            if len(sel) == L:
                pairs += [(int(mem_slots[t, k]), track.images_mem[j]) for k, j in enumerate(sel)]
        pairs += [(int(s), d.images_mem[-1]) for s, d in zip(det_slots, dets)]
        if kal_slots is not None:
            pairs += [(int(s), k.images_mem[-1]) for s, k in zip(kal_slots, kalman)]
        if not pairs:
            return
        self.engine.sync()
        got = self.engine.bank_download(np.array([s for s, _ in pairs], np.int32))
        for (slot, host), dev in zip(pairs, got):
            if not np.array_equal(np.asarray(host), dev):
                raise RuntimeError(f"patch-bank slot {slot} no longer matches the host crop it was registered for: a stored crop was modified in place "
                                   "(busca_b200 addresses device patches through the arrays in images_mem; replace the array instead of writing into it)")

    def sync(self):
        """Wait for everything enqueued on this tracker's stream (with ``defer_crop_copies``: the crops' host bytes are valid afterwards)."""
        self.engine.sync()

    def set_frame(self, image: np.ndarray):
        """Explicitly (re)upload the current frame."""
        self.engine.upload_frame(image)

    def get_image_crops(self, image, bboxes, output_size=None, normalize=True):
        if output_size is None:
            output_size = (self.expected_image_size[1], self.expected_image_size[0])
        if tuple(output_size) != (128, 384):
            raise NotImplementedError("crops are fixed at 128x384")
        if isinstance(bboxes, np.ndarray) and bboxes.ndim == 2 and bboxes.shape[1] == 4:
            boxes = np.ascontiguousarray(bboxes, dtype=np.float64)
        elif len(bboxes) == 1:                                # the adapters' per-track call (byte_tracker.py:468-479): one box in a list
            boxes = np.array(bboxes[0], dtype=np.float64).reshape(1, 4)
        else:
            rows = [np.asarray(b, dtype=np.float64).reshape(4) for b in bboxes]
            boxes = np.stack(rows) if rows else np.zeros((0, 4))
        if len(boxes) == 0:
            return np.zeros([0, output_size[0], output_size[1], 3])       # float64, dims swapped - as network.py:503
        assert image is not None, "Image is None"
        self._ensure_frame(image, boxes)
        slots = self.engine.alloc_slots(len(boxes))
        crops, owner = self.engine.crop_owned(boxes, slots)
        self._registry.register(crops, slots, owner)
        if normalize:
            return np.stack([tracking.normalize_crop(c) for c in crops], axis=0)
        return crops

    # ---- memory sampling: _get_track_mem (network.py:247-279) --------------------------------------------
    @staticmethod
    def _memory_indices(n_obs: int, seq_len: int, use_broader_memory: bool) -> List[int]:
        if use_broader_memory and not (seq_len == 1 and n_obs >= 1) and n_obs >= seq_len:
            sep = float(n_obs - 1) / float(seq_len - 1)
            return [int(i * sep) for i in range(seq_len)]
        return list(range(max(0, n_obs - seq_len), n_obs))

    def _get_track_mem(self, track, seq_len, use_broader_memory):
        sel = self._memory_indices(len(track.images_mem), seq_len, use_broader_memory)
        mem = [track.images_mem[j] for j in sel]
        boxes = np.array([track.tlwh_mem[j] for j in sel]) * track.scale if sel else np.zeros((0, 4))
        return mem, boxes

    # ---- association: associate_embeddings (network.py:282-429) -------------------------------------------
    def associate_embeddings(self, tracks_embeddings, dets_embeddings, dists_matrix, seq_len, num_candidates, use_broader_memory,
                             select_highest_candidate, highest_candidate_minimum_thresh=None, keep_highest_value=False,
                             extra_kalman_candidates=[], plot_results=False, normalize_ims=False):
        T, D, K = len(tracks_embeddings), len(dets_embeddings), len(extra_kalman_candidates)
        if T == 0:
            return None, None
        if D == 0 and K == 0:
            return None, None
        if not normalize_ims:
            raise NotImplementedError("busca_b200 expects uint8 crops (normalize_ims=True), as every adapter passes")
        if plot_results:
            raise NotImplementedError("plot_results needs the debug GUI, which is out of scope")
        if K not in (0, T):
            raise ValueError("extra_kalman_candidates must be empty or hold one entry per track")
        L, C = int(seq_len), int(num_candidates)
        temp_slots: List[int] = []
        pending = []                                       # (patch array, slot) uploads for arrays we have never seen

        lookup = self._registry.lookup_cached
        by_id = self._registry._by_id

        def slot_for(patch) -> int:
            hit = by_id.get(id(patch))                     # same array object as in an earlier frame: one dict probe
            if hit is not None and hit[0]() is patch:
                return hit[1]
            s = lookup(patch)
            if s is None:
                arr = np.ascontiguousarray(np.asarray(patch), dtype=np.uint8)
                if arr.shape != PATCH_SHAPE:
                    raise ValueError("images_mem entries must be uint8 [384,128,3] crops, got %s" % (arr.shape,))
                s = int(self.engine.alloc_slots(1)[0])
                temp_slots.append(s)
                pending.append((arr, s))
            return s

        try:
            mem_slots = np.full((T, L), -1, np.int32)
            mem_ltwh = np.empty((T, L, 4), np.float64)
            reliable = np.zeros(T, dtype=bool)
            sel_cache = {}                                     # history length -> sampled indices (same L / flag for every track)
            rel_rows, rel_boxes, rel_scales = [], [], []       # reliable tracks: one array build for all their boxes
            for t, track in enumerate(tracks_embeddings):
                ims = track.images_mem
                n_obs = len(ims)
                sel = sel_cache.get(n_obs)
                if sel is None:
                    sel = sel_cache[n_obs] = self._memory_indices(n_obs, L, use_broader_memory)
                if len(sel) == L:
                    reliable[t] = True
                    boxes = track.tlwh_mem
                    row = []
                    for j in sel:
                        patch = ims[j]
                        hit = by_id.get(id(patch))             # same array object as in an earlier frame: one dict probe
                        row.append(hit[1] if hit is not None and hit[0]() is patch else slot_for(patch))
                    mem_slots[t] = row
                    rel_rows.append(t)
                    rel_boxes.extend([boxes[j] for j in sel])
                    rel_scales.append(track.scale)
                else:                                      # incomplete history: zero images + filler box (network.py:304-308)
                    mem_ltwh[t] = np.array([250.0, 250.0, 500.0, 500.0])
            if rel_rows:
                mem_ltwh[rel_rows] = np.array(rel_boxes, dtype=np.float64).reshape(len(rel_rows), L, 4) \
                    * np.array(rel_scales, dtype=np.float64).reshape(-1, 1, 1)
            det_slots = np.array([slot_for(det.images_mem[-1]) for det in dets_embeddings], np.int32).reshape(D)
            det_ltwh = np.array([det.tlwh_mem[-1] for det in dets_embeddings], np.float64).reshape(D, 4) \
                * np.array([det.scale for det in dets_embeddings], np.float64).reshape(D, 1)
            kal_slots = kal_ltwh = None
            if K:
                kal_slots = np.array([slot_for(kd.images_mem[-1]) for kd in extra_kalman_candidates], np.int32)
                kal_ltwh = np.array([kd.tlwh for kd in extra_kalman_candidates], np.float64).reshape(T, 4) \
                    * np.array([kd.scale for kd in extra_kalman_candidates], np.float64).reshape(T, 1)
            if pending:
                self.engine.bank_upload(np.stack([p for p, _ in pending]), np.array([s for _, s in pending], np.int32))
            if self.verify_patches:
                self._verify_patches(tracks_embeddings, dets_embeddings, extra_kalman_candidates, mem_slots, det_slots, kal_slots, L, use_broader_memory)
            dists = np.asarray(dists_matrix, np.float64).reshape(T, D) if D else None
            want = ("probs", "cand") + (("cand_rows", "mem_logits") if self.store_logits else ())
            out = self.engine.associate(mem_slots, mem_ltwh, det_slots if D else None, det_ltwh if D else None, dists,
                                        kal_slots, kal_ltwh, L, C, want=want)
        finally:
            if temp_slots:
                self.engine.free_slots(temp_slots)
        if self.store_logits:
            self.logits, self.mem_logits = out["cand_rows"], out["mem_logits"]
        probs, cand = out["probs"], out["cand"]

        # scatter into the global matrix (network.py:407-425)
        n_avail = min(D + 1, C) if K else min(D, C)
        probs_matrix = np.zeros([T, D + K])
        rows = np.arange(T)
        if select_highest_candidate:
            best = np.argmax(probs, axis=1)                     # first maximum, as np.argmax in the reference loop
            top = probs[rows, best]
            q = np.zeros_like(probs)
            thr = highest_candidate_minimum_thresh
            on = np.ones(T, bool) if (thr is None or thr == 0) else ((top >= thr) if thr > 0.0 else np.zeros(T, bool))
            q[rows[on], best[on]] = top[on] if keep_highest_value else 1.0
            probs = q
        # the n_avail candidate columns of a track are distinct (detection ids + its own Kalman column)
        probs_matrix[rows[:, None], cand[:, :n_avail]] = probs[:, :n_avail]
        return probs_matrix, reliable

    @staticmethod
    def ltwh_to_ltrb(ltwh):
        ret = np.array(ltwh, copy=True)
        ret[..., 2:] += ret[..., :2]
        return ret


class ReID_Encoder:
    """Name kept for API parity (network.py:510-575).  The encoder is part of the library; this wrapper embeds
    uint8 BGR crops as ONE BatchNorm batch (train-mode statistics, like the reference's domain adaptation)."""
    PRETRAINED_SIZE = (384, 128)

    def __init__(self, busca: BUSCA):
        self.busca = busca

    def embed_patches(self, patches_u8: np.ndarray) -> np.ndarray:
        patches_u8 = np.ascontiguousarray(patches_u8, dtype=np.uint8).reshape(-1, *PATCH_SHAPE)
        eng = self.busca.engine
        slots = eng.alloc_slots(len(patches_u8))
        try:
            eng.bank_upload(patches_u8, slots)
            return eng.reid_embed(slots)
        finally:
            eng.free_slots(slots)
