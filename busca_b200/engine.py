"""Thin Python owner of one ``busca_ctx`` (one per tracker, like the reference's one BUSCA per
BYTETracker, byte_tracker.py:217-221).  Everything numeric happens in libbusca_b200.so."""
from __future__ import annotations

import ctypes as C
import json
import os
import weakref
from typing import Dict, Iterable, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import AssocArgs, Config, StepArgs, check

PATCH_SHAPE = (384, 128, 3)
PATCH_BYTES = 384 * 128 * 3

MEAN_BGR = np.array([0.406, 0.456, 0.485])
STD_BGR = np.array([0.225, 0.224, 0.299])    # sic: the reference's "ghost_normalize" constants (network.py:471-472)


def _ptr(a: Optional[np.ndarray]):
    # the address as a plain int (argtypes are c_void_p): a third of the cost of a.ctypes.data_as(...) - the adapters make hundreds of
    # small calls per frame
    return None if a is None else a.__array_interface__["data"][0]


def normalize_lut() -> np.ndarray:
    """The 256x3 table of BUSCA._normalize_embeddings_batch (network.py:470-478): fp32 division by 255,
    then float64 subtract / divide rounded back to fp32 after each step (numpy in-place semantics)."""
    v = (np.arange(256, dtype=np.float32) / np.float32(255.0)).astype(np.float32)
    lut = np.empty((256, 3), np.float32)
    for c in range(3):
        a = (v.astype(np.float64) - MEAN_BGR[c]).astype(np.float32)
        lut[:, c] = (a.astype(np.float64) / STD_BGR[c]).astype(np.float32)
    return lut


def pe_tables(d_model: int = 512):
    """Separable factors of the reference's 211x211x61xd fp16 table (encodings.py:23-32 with
    positional_encodings 6.0.x): sin/cos interleaved, computed in fp32 with torch exactly as the
    reference does, then cast to fp16.  2.6 GiB in the reference, 166 KB here."""
    import torch
    ch = int(np.ceil(d_model / 6) * 2)
    ch += ch % 2
    inv_freq = 1.0 / (10000 ** (torch.arange(0, ch, 2).float() / ch))

    def code(n):
        s = torch.einsum("i,j->ij", torch.arange(n, dtype=torch.float32), inv_freq)
        return torch.flatten(torch.stack((s.sin(), s.cos()), dim=-1), -2, -1)

    return (code(211).to(torch.float16).numpy(), code(211).to(torch.float16).numpy(),
            code(61)[:, : d_model - 2 * ch].to(torch.float16).numpy())


class _PinnedBlock:
    """Owner of one page-locked host allocation, visible to numpy through ``__array_interface__``.  Every array (and
    every row view an adapter keeps in ``track.images_mem``) made from it holds a reference to it, so the memory goes
    back to the engine's pool exactly when the last view dies."""
    __slots__ = ("ptr", "nbytes", "__array_interface__", "__weakref__")

    def __init__(self, ptr: int, nbytes: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


_last_engine = None        # weak reference to the most recently created Engine (busca_b200.tracking.default_engine)


def last_engine():
    e = _last_engine() if _last_engine is not None else None
    return e if e is not None and getattr(e, "h", None) else None


class _PinnedSlice:
    """A few patches carved out of a page-locked arena.  It is the ``base`` of the array handed to the caller (and of every row
    view of it), so it dies exactly when the last view dies; it keeps its arena alive, and the arena goes back to the pool when
    its last slice is gone."""
    __slots__ = ("arena", "__array_interface__", "__weakref__")

    def __init__(self, arena: _PinnedBlock, offset: int, nbytes: int):
        self.arena = arena
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (arena.ptr + offset, False), "version": 3}


ARENA_PATCHES = 64      # single-box crop calls (byte_tracker.py:468-479) are served from 9.4 MB arenas instead of one cudaHostAlloc each


class Engine:
    def __init__(self, device: int = 0, d_model: int = 512, nhead: int = 4, ff_size: int = 1024, num_layers: int = 4,
                 activation: str = "relu", precision: str = "fp32", sentinel_fp64: bool = True, bank_slots: int = 2048):
        self.L = _lib.load()
        cfg = Config(device=device, d_model=d_model, nhead=nhead, ff_size=ff_size, num_layers=num_layers,
                     activation={"relu": 0, "gelu": 1}[activation], precision={"fp32": 0, "bf16": 1}[precision],
                     sentinel_fp64=int(bool(sentinel_fp64)), bank_slots=bank_slots)
        h = C.c_void_p()
        check(self.L.busca_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.precision = precision
        self.d_model = d_model
        self._free = list(range(bank_slots - 1, -1, -1))
        self._cap = bank_slots
        # page-locked pool for the crop arrays handed to the caller: size class (patches, power of two) -> free pointers
        self._pin_free: Dict[int, List[int]] = {}
        self._pin_total = 0
        # page-locking costs ~0.75 ms per MB (measured on the B200 host, r02n), more than a pageable D2H of the same bytes: blocks are
        # recycled through the pool, and NEW ones are only locked below this cap (adapters keep rows of the big per-frame arrays alive
        # for the life of a track, so those blocks rarely come back)
        self._pin_max = int(float(os.environ.get("BUSCA_PINNED_MAX_GB", "1")) * (1 << 30))
        self._arena: Optional[_PinnedBlock] = None
        self._frame_key = None                              # sync_frame: the last validated frame object
        self._frame_args = None
        self._frame_keep = None
        self._frame_up = C.c_int32(0)
        self._arena_used = 0
        self.device = device
        global _last_engine
        _last_engine = weakref.ref(self)

    def close(self):
        if getattr(self, "h", None):
            for ptrs in self._pin_free.values():
                for p in ptrs:
                    self.L.busca_host_free(self.h, p)
            self._pin_free = {}
            self.L.busca_destroy(self.h)
            self.h = None

    # ---- page-locked host pool ------------------------------------------------------------
    def _pin_get(self, n_patches: int) -> Optional[_PinnedBlock]:
        cls = 1 << max(0, int(n_patches - 1).bit_length())
        nbytes = cls * PATCH_BYTES
        free = self._pin_free.get(cls)
        if free:
            ptr = free.pop()
        else:
            if self._pin_total + nbytes > self._pin_max:
                return None                                 # over the cap: the caller falls back to pageable memory
            ptr = self.L.busca_host_alloc(self.h, nbytes)
            if not ptr:
                return None
            self._pin_total += nbytes
        blk = _PinnedBlock(ptr, nbytes)
        weakref.finalize(blk, Engine._pin_release, weakref.ref(self), self.L, ptr, cls)
        return blk

    @staticmethod
    def _pin_release(self_ref, lib, ptr, cls):
        self = self_ref()
        if self is not None and self.h is not None:
            self._pin_free.setdefault(cls, []).append(ptr)
        else:
            lib.busca_host_free(None, ptr)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights --------------------------------------------------------------------------
    def load_state_dict(self, sd: Dict[str, np.ndarray]):
        """Feed every entry of a model_busca.pth-layout state dict (numpy or torch tensors)."""
        for k, v in sd.items():
            a = v.detach().cpu().numpy() if hasattr(v, "detach") else np.asarray(v)
            if a.dtype == np.int64:
                dt = 2
            elif a.dtype == np.float16:
                dt = 1
            else:
                a = np.ascontiguousarray(a, dtype=np.float32)
                dt = 0
            a = np.ascontiguousarray(a)
            shape = (C.c_int64 * max(1, a.ndim))(*a.shape)
            check(self.L.busca_load_tensor(self.h, k.encode(), _ptr(a), dt, a.ndim, shape))
        tx, ty, tz = pe_tables(self.d_model)
        extra = {"pe.tab_xy": tx, "pe.tab_size": ty, "pe.tab_t": tz, "norm.lut": normalize_lut()}
        for k, a in extra.items():
            a = np.ascontiguousarray(a)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            check(self.L.busca_load_tensor(self.h, k.encode(), _ptr(a), 1 if a.dtype == np.float16 else 0, a.ndim, shape))
        check(self.L.busca_finalize(self.h))

    # ---- patch bank -----------------------------------------------------------------------
    def alloc_slots(self, n: int) -> np.ndarray:
        if n > len(self._free):
            new_cap = max(self._cap * 2, self._cap + n)
            check(self.L.busca_bank_reserve(self.h, new_cap))
            self._free = list(range(new_cap - 1, self._cap - 1, -1)) + self._free
            self._cap = new_cap
        if n == 1:
            return np.array([self._free.pop()], dtype=np.int32)
        # ascending: crop i of a call lands in slot out[i], and the device<->host copies of a call go one DMA per RUN of consecutive slots
        # (busca_crop / busca_bank_*) - recycled slots come back from the free list in arbitrary order, and 300 separate 147 KB copies
        # run at 24 GB/s where one 44 MB copy runs at 55 (tests/probe_pcie.py)
        take = self._free[-n:]
        del self._free[-n:]
        return np.sort(np.array(take, dtype=np.int32))

    def slots_in_use(self) -> int:
        """Patch-bank slots currently handed out (crops some track still references)."""
        return self._cap - len(self._free)

    def free_slots(self, slots: Iterable[int]):
        self._free.extend(int(s) for s in slots)

    def upload_frame(self, image: np.ndarray):
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise ValueError("frame must be uint8 [H,W,3] BGR")
        if image.strides[2] != 1 or image.strides[1] != 3:
            image = np.ascontiguousarray(image)
        check(self.L.busca_upload_frame(self.h, _ptr(image), image.shape[0], image.shape[1], image.strides[0]))

    def ingest_frame(self, chw, mean, std, H: Optional[int] = None, W: Optional[int] = None, to_host: bool = True) -> Optional[np.ndarray]:
        """mot_evaluator.py:198-204 on the device: ``chw`` is the detector's input, either a float32 numpy array [3,H,W] or an integer
        DEVICE address of one (``tensor.data_ptr()``; give H and W) - RGB, normalised with ``mean`` / ``std``.  The de-normalised uint8 BGR
        frame becomes the engine's current frame; returned on the host as well unless ``to_host`` is False."""
        mean = np.ascontiguousarray(mean, np.float32).reshape(3)
        std = np.ascontiguousarray(std, np.float32).reshape(3)
        if isinstance(chw, (int, np.integer)):
            ptr, on_dev = C.c_void_p(int(chw)), 1
        else:
            chw = np.ascontiguousarray(chw, np.float32)
            if chw.ndim != 3 or chw.shape[0] != 3:
                raise ValueError("detector tensor must be float32 [3,H,W]")
            H, W = chw.shape[1], chw.shape[2]
            ptr, on_dev = _ptr(chw), 0
        out = np.empty((H, W, 3), np.uint8) if to_host else None
        check(self.L.busca_ingest_frame(self.h, ptr, on_dev, int(H), int(W), _ptr(mean), _ptr(std), _ptr(out)))
        return out

    def camera_motion(self, previous: Optional[np.ndarray], current: Optional[np.ndarray], shape=None, iterations: int = 100,
                      eps: float = 1e-5) -> Tuple[np.ndarray, float, int]:
        """cv2.findTransformECC(gray(previous), gray(current), eye(2,3), MOTION_EUCLIDEAN, (EPS | COUNT, iterations, eps)) on the device
        -> (warp 2x3 float32, rho, iterations run).  previous None: the current frame of the last call; current None: the frame in HBM
        (give ``shape``).  Raises where cv2 raises (uncorrelated images)."""
        def prep(a):
            if a is None:
                return None
            if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
                raise ValueError("frame must be uint8 [H,W,3] BGR")
            return a if (a.strides[2] == 1 and a.strides[1] == 3) else np.ascontiguousarray(a)
        previous, current = prep(previous), prep(current)
        ref = current if current is not None else previous
        H, W = (ref.shape[0], ref.shape[1]) if ref is not None else (int(shape[0]), int(shape[1]))
        if previous is not None and current is not None and previous.strides[0] != current.strides[0]:
            previous, current = np.ascontiguousarray(previous), np.ascontiguousarray(current)
        stride = ref.strides[0] if ref is not None else W * 3
        warp = np.zeros((2, 3), np.float32)
        rho, it = C.c_double(0.0), C.c_int32(0)
        check(self.L.busca_camera_motion(self.h, _ptr(previous), _ptr(current), H, W, stride, int(iterations), float(eps), _ptr(warp),
                                         C.byref(rho), C.byref(it)))
        return warp, float(rho.value), int(it.value)

    def sync_frame(self, image: np.ndarray, boxes: Optional[np.ndarray] = None) -> bool:
        """Upload ``image`` unless the pixels ``boxes`` read (all pixels without boxes) already are in HBM (busca_sync_frame)."""
        ai = image.__array_interface__
        key = (id(image), ai["data"][0], ai["shape"], ai["strides"])
        if key != self._frame_key:                          # validate once per frame object; the PIXELS are compared by the library every call
            if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
                raise ValueError("frame must be uint8 [H,W,3] BGR")
            if image.strides[2] != 1 or image.strides[1] != 3:
                image = np.ascontiguousarray(image)
                self._frame_key = None
            else:
                self._frame_key = key
            self._frame_args = (image.__array_interface__["data"][0], image.shape[0], image.shape[1], image.strides[0])
            self._frame_keep = image
        up = self._frame_up
        nb = 0 if boxes is None else len(boxes)
        check(self.L.busca_sync_frame(self.h, *self._frame_args, _ptr(boxes) if nb else None, nb, C.byref(up)))
        return bool(up.value)

    def crop(self, boxes: np.ndarray, slots: np.ndarray, to_host: bool = True) -> Optional[np.ndarray]:
        boxes = np.ascontiguousarray(boxes, dtype=np.float64).reshape(-1, 4)
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        out = np.empty((len(boxes),) + PATCH_SHAPE, np.uint8) if to_host else None
        check(self.L.busca_crop(self.h, _ptr(boxes), len(boxes), _ptr(slots), _ptr(out)))
        return out

    def crop_owned(self, boxes: np.ndarray, slots: np.ndarray) -> Tuple[np.ndarray, object]:
        """Crops to bank slots AND to the host, into page-locked memory from the pool.  Returns ``(array, owner)``:
        ``owner`` is the object that dies when the array and all of its views are gone (the pinned block, or the
        array itself on the pageable fallback) - hang slot-recycling finalizers on it."""
        if boxes.dtype != np.float64 or not boxes.flags["C_CONTIGUOUS"]:
            boxes = np.ascontiguousarray(boxes, dtype=np.float64)
        if slots.dtype != np.int32 or not slots.flags["C_CONTIGUOUS"]:
            slots = np.ascontiguousarray(slots, dtype=np.int32)
        n = boxes.size // 4
        blk = None
        if n * 8 <= ARENA_PATCHES:                          # small call: a slice of the current arena
            if self._arena is None or self._arena_used + n > ARENA_PATCHES:
                self._arena, self._arena_used = self._pin_get(ARENA_PATCHES), 0
            if self._arena is not None:
                blk = _PinnedSlice(self._arena, self._arena_used * PATCH_BYTES, n * PATCH_BYTES)
                self._arena_used += n
        else:
            blk = self._pin_get(n)
        if blk is None:
            out = np.empty((n,) + PATCH_SHAPE, np.uint8)
            owner = out
        else:
            out = np.asarray(blk)[: n * PATCH_BYTES].reshape((n,) + PATCH_SHAPE)
            owner = blk
        check(self.L.busca_crop(self.h, _ptr(boxes), n, _ptr(slots), _ptr(out)))
        return out, owner

    def bank_upload(self, patches: np.ndarray, slots: np.ndarray):
        patches = np.ascontiguousarray(patches, dtype=np.uint8).reshape(-1, *PATCH_SHAPE)
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        check(self.L.busca_bank_upload(self.h, _ptr(patches), len(slots), _ptr(slots)))

    def bank_download(self, slots: np.ndarray) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.int32)
        out = np.empty((len(slots),) + PATCH_SHAPE, np.uint8)
        check(self.L.busca_bank_download(self.h, _ptr(slots), len(slots), _ptr(out)))
        return out

    # ---- geometry -------------------------------------------------------------------------
    def center_distance(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 4)
        b = np.ascontiguousarray(b, np.float64).reshape(-1, 4)
        out = np.zeros((len(a), len(b)), np.float64)
        check(self.L.busca_center_distance(self.h, _ptr(a), len(a), _ptr(b), len(b), _ptr(out)))
        return out

    def detection_coverage(self, boxes: np.ndarray, H: int, W: int) -> Tuple[int, np.ndarray]:
        """Union area in pixels of the filled int()-truncated rectangles on an H x W canvas, and the per-box relative areas."""
        boxes = np.ascontiguousarray(boxes, np.float64).reshape(-1, 4)
        areas = np.zeros(len(boxes), np.float64)
        cnt = C.c_int64(0)
        check(self.L.busca_detection_coverage(self.h, _ptr(boxes), len(boxes), int(H), int(W), C.byref(cnt), _ptr(areas)))
        return int(cnt.value), areas

    # ---- host-tracker rounds on the device ---------------------------------------------------------
    def kalman_predict(self, mean: np.ndarray, cov: np.ndarray, tracked: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
        """KalmanFilter.multi_predict (mean and covariance) for [n,8] / [n,8,8] float64; bit-identical to numpy."""
        mean = np.ascontiguousarray(mean, np.float64).reshape(-1, 8)
        cov = np.ascontiguousarray(cov, np.float64).reshape(-1, 8, 8)
        mo, co = np.empty_like(mean), np.empty_like(cov)
        tr = None if tracked is None else np.ascontiguousarray(tracked, np.uint8)
        check(self.L.busca_kalman_predict(self.h, _ptr(mean), _ptr(cov), None if tr is None else _ptr(tr), len(mean), _ptr(mo), _ptr(co)))
        return mo, co

    def kalman_update(self, mean: np.ndarray, cov: np.ndarray, xyah: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """KalmanFilter.update for n independent (track, measurement) pairs."""
        mean = np.ascontiguousarray(mean, np.float64).reshape(-1, 8)
        cov = np.ascontiguousarray(cov, np.float64).reshape(-1, 8, 8)
        z = np.ascontiguousarray(xyah, np.float64).reshape(-1, 4)
        mo, co = np.empty_like(mean), np.empty_like(cov)
        check(self.L.busca_kalman_update(self.h, _ptr(mean), _ptr(cov), _ptr(z), len(mean), _ptr(mo), _ptr(co)))
        return mo, co

    def match_round(self, a_tlbr: np.ndarray, b_tlbr: np.ndarray, b_score: Optional[np.ndarray], cost_limit: float, want_cost: bool = False):
        """iou_distance (+ fuse_score) + linear_assignment in one call: (x [na], y [nb], cost or None)."""
        a = np.ascontiguousarray(a_tlbr, np.float64).reshape(-1, 4)
        b = np.ascontiguousarray(b_tlbr, np.float64).reshape(-1, 4)
        sc = None if b_score is None else np.ascontiguousarray(b_score, np.float64).reshape(-1)
        x, y = np.empty(len(a), np.int32), np.empty(len(b), np.int32)
        cost = np.empty((len(a), len(b)), np.float64) if want_cost else None
        check(self.L.busca_match_round(self.h, _ptr(a), len(a), _ptr(b), len(b), None if sc is None else _ptr(sc), float(cost_limit),
                                       _ptr(x), _ptr(y), None if cost is None else _ptr(cost)))
        return x, y, cost

    def linear_assignment(self, cost: np.ndarray, cost_limit: float) -> Tuple[np.ndarray, np.ndarray]:
        cost = np.ascontiguousarray(cost, np.float64)
        n, m = cost.shape
        x, y = np.empty(n, np.int32), np.empty(m, np.int32)
        check(self.L.busca_linear_assignment(self.h, _ptr(cost), n, m, float(cost_limit), _ptr(x), _ptr(y)))
        return x, y

    def duplicate_tracks(self, a_tlbr, a_age, b_tlbr, b_age, thresh: float = 0.15) -> Tuple[np.ndarray, np.ndarray]:
        a = np.ascontiguousarray(a_tlbr, np.float64).reshape(-1, 4)
        b = np.ascontiguousarray(b_tlbr, np.float64).reshape(-1, 4)
        ga, gb = np.ascontiguousarray(a_age, np.int32), np.ascontiguousarray(b_age, np.int32)
        da, db = np.zeros(len(a), np.uint8), np.zeros(len(b), np.uint8)
        check(self.L.busca_duplicate_tracks(self.h, _ptr(a), _ptr(ga), len(a), _ptr(b), _ptr(gb), len(b), float(thresh), _ptr(da), _ptr(db)))
        return da.astype(bool), db.astype(bool)

    def iou(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, np.float64).reshape(-1, 4)
        b = np.ascontiguousarray(b, np.float64).reshape(-1, 4)
        out = np.zeros((len(a), len(b)), np.float64)
        check(self.L.busca_iou(self.h, _ptr(a), len(a), _ptr(b), len(b), _ptr(out)))
        return out

    def motion_proposals(self, mean: np.ndarray, tracked: Optional[np.ndarray] = None):
        mean = np.ascontiguousarray(mean, np.float64).reshape(-1, 8)
        n = len(mean)
        trk = None if tracked is None else np.ascontiguousarray(tracked, np.uint8)
        mo, tlwh, tlbr = np.empty((n, 8)), np.empty((n, 4)), np.empty((n, 4))
        check(self.L.busca_motion_proposals(self.h, _ptr(mean), _ptr(trk), n, _ptr(mo), _ptr(tlwh), _ptr(tlbr)))
        return mo, tlwh, tlbr

    def frame_geometry(self, mean, tracked, det_tlbr, C_: int, use_kalman: bool = True):
        mean = np.ascontiguousarray(mean, np.float64).reshape(-1, 8)
        det = np.ascontiguousarray(det_tlbr, np.float64).reshape(-1, 4)
        T, D = len(mean), len(det)
        trk = None if tracked is None else np.ascontiguousarray(tracked, np.uint8)
        tlwh, tlbr = np.empty((T, 4)), np.empty((T, 4))
        dist, iou = np.zeros((T, D)), np.zeros((T, D))
        cand = np.empty((T, C_), np.int32)
        check(self.L.busca_frame_geometry(self.h, _ptr(mean), _ptr(trk), T, _ptr(det), D, C_, int(use_kalman),
                                          _ptr(tlwh), _ptr(tlbr), _ptr(dist), _ptr(iou), _ptr(cand)))
        return dict(tlwh=tlwh, tlbr=tlbr, dist=dist, iou=iou, cand=cand)

    def frame_geometry_batch(self, mean, tracked, det_tlbr, C_: int, use_kalman: bool = True):
        """frame_geometry for B frames in one launch: mean [B,T,8], tracked [B,T] or None, det_tlbr [B,D,4]."""
        mean = np.ascontiguousarray(mean, np.float64)
        det = np.ascontiguousarray(det_tlbr, np.float64)
        B, T, D = mean.shape[0], mean.shape[1], det.shape[1]
        trk = None if tracked is None else np.ascontiguousarray(tracked, np.uint8)
        tlwh, tlbr = np.empty((B, T, 4)), np.empty((B, T, 4))
        dist, iou = np.zeros((B, T, D)), np.zeros((B, T, D))
        cand = np.empty((B, T, C_), np.int32)
        check(self.L.busca_frame_geometry_batch(self.h, B, _ptr(mean), _ptr(trk), T, _ptr(det), D, C_, int(use_kalman),
                                                _ptr(tlwh), _ptr(tlbr), _ptr(dist), _ptr(iou), _ptr(cand)))
        return dict(tlwh=tlwh, tlbr=tlbr, dist=dist, iou=iou, cand=cand)

    # ---- network --------------------------------------------------------------------------
    def reid_embed(self, slots: np.ndarray) -> np.ndarray:
        slots = np.ascontiguousarray(slots, dtype=np.int32).reshape(-1)
        out = np.empty((len(slots), 512), np.float32)
        check(self.L.busca_reid_embed(self.h, _ptr(slots), len(slots), _ptr(out)))
        return out

    def transformer(self, mem_emb, can_emb, mem_ltwh, can_ltwh, want=("logits", "probs", "pe_index")):
        mem_emb = np.ascontiguousarray(mem_emb, np.float32)
        can_emb = np.ascontiguousarray(can_emb, np.float32)
        T, L, _ = mem_emb.shape
        C_ = can_emb.shape[1]
        S = L + 2 * (C_ + 2)
        mem_ltwh = np.ascontiguousarray(mem_ltwh, np.float64).reshape(T, L, 4)
        can_ltwh = np.ascontiguousarray(can_ltwh, np.float64).reshape(T, C_, 4)
        bufs = dict(logits=np.empty((T, C_ + 2), np.float32), probs=np.empty((T, C_ + 2), np.float32),
                    pe_index=np.empty((T, S, 3), np.int32), cand_rows=np.empty((T, C_ + 2, 512), np.float32),
                    input_seq=np.empty((T, S, 512), np.float32))
        g = lambda k: _ptr(bufs[k]) if k in want else None
        check(self.L.busca_transformer(self.h, T, L, C_, _ptr(mem_emb), _ptr(can_emb), _ptr(mem_ltwh), _ptr(can_ltwh),
                                       g("logits"), g("probs"), g("pe_index"), g("cand_rows"), g("input_seq")))
        return {k: bufs[k] for k in want}

    def associate(self, mem_slots, mem_ltwh, det_slots, det_ltwh, dists, kal_slots, kal_ltwh, L: int, C_: int,
                  want=("probs", "cand")) -> Dict[str, np.ndarray]:
        mem_slots = np.ascontiguousarray(mem_slots, np.int32)
        T = mem_slots.shape[0]
        D = 0 if det_slots is None else len(det_slots)
        S = L + 2 * (C_ + 2)
        mem_ltwh = np.ascontiguousarray(mem_ltwh, np.float64).reshape(T, L, 4)
        use_kal = kal_slots is not None
        keep = [mem_slots, mem_ltwh]
        a = AssocArgs(T=T, D=D, L=L, C=C_, use_kalman=int(use_kal))
        a.mem_slots = mem_slots.ctypes.data_as(_lib.c_i32p)
        a.mem_ltwh = mem_ltwh.ctypes.data_as(_lib.c_f64p)
        if D:
            ds = np.ascontiguousarray(det_slots, np.int32)
            db = np.ascontiguousarray(det_ltwh, np.float64).reshape(D, 4)
            dd = np.ascontiguousarray(dists, np.float64).reshape(T, D)
            keep += [ds, db, dd]
            a.det_slots = ds.ctypes.data_as(_lib.c_i32p)
            a.det_ltwh = db.ctypes.data_as(_lib.c_f64p)
            a.dists = dd.ctypes.data_as(_lib.c_f64p)
        if use_kal:
            ks = np.ascontiguousarray(kal_slots, np.int32)
            kb = np.ascontiguousarray(kal_ltwh, np.float64).reshape(T, 4)
            keep += [ks, kb]
            a.kal_slots = ks.ctypes.data_as(_lib.c_i32p)
            a.kal_ltwh = kb.ctypes.data_as(_lib.c_f64p)
        shapes = dict(probs=((T, C_ + 2), np.float32), logits=((T, C_ + 2), np.float32), cand=((T, C_), np.int32),
                      pe_index=((T, S, 3), np.int32), mem_emb=((T, L, 512), np.float32), can_emb=((T, C_, 512), np.float32),
                      cand_rows=((T, C_ + 2, 512), np.float32), mem_logits=((T, 512), np.float32),
                      input_seq=((T, S, 512), np.float32))
        out = {}
        for k in want:
            shp, dt = shapes[k]
            out[k] = np.empty(shp, dt)
            setattr(a, k, out[k].ctypes.data_as(_lib.c_i32p if dt == np.int32 else _lib.c_f32p))
        check(self.L.busca_associate(self.h, C.byref(a)))
        return out

    # ---- device-resident path ---------------------------------------------------------------
    def dev_alloc(self, nbytes: int) -> int:
        p = self.L.busca_dev_alloc(self.h, nbytes)
        if not p:
            raise _lib.BuscaError(self.L.busca_last_error().decode())
        return p

    def dev_free(self, p: int):
        self.L.busca_dev_free(self.h, p)

    def to_dev(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a)
        p = self.dev_alloc(max(a.nbytes, 16))
        check(self.L.busca_memcpy_h2d(self.h, p, _ptr(a), a.nbytes))
        return p

    def h2d(self, p: int, a: np.ndarray):
        a = np.ascontiguousarray(a)
        check(self.L.busca_memcpy_h2d(self.h, p, _ptr(a), a.nbytes))

    def from_dev(self, p: int, shape, dtype) -> np.ndarray:
        out = np.empty(shape, dtype)
        check(self.L.busca_memcpy_d2h(self.h, _ptr(out), p, out.nbytes))
        return out

    def frame_step_dev(self, args: StepArgs):
        check(self.L.busca_frame_step_dev(self.h, C.byref(args)))

    def sync(self):
        check(self.L.busca_sync(self.h))

    @property
    def stream(self) -> int:
        return self.L.busca_stream(self.h)

    @property
    def launches(self) -> int:
        return int(self.L.busca_kernel_launches(self.h))

    def set_option(self, name: str, value: int):
        check(self.L.busca_set_option(self.h, name.encode(), int(value)))

    def counter(self, name: str) -> int:
        return int(self.L.busca_counter(self.h, name.encode()))

    def set_profiling(self, on: bool):
        check(self.L.busca_set_profiling(self.h, int(on)))

    def last_profile(self) -> dict:
        return json.loads(self.L.busca_last_profile(self.h).decode() or "{}")
