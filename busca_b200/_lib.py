"""ctypes binding of libbusca_b200.so (include/busca_b200.h).

There is NO CPU fallback: if the shared library is missing or no sm_100 device is present the
product path raises.  (``oracle/`` is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_lib = None

c_i32p = C.POINTER(C.c_int32)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("d_model", C.c_int32), ("nhead", C.c_int32), ("ff_size", C.c_int32),
                ("num_layers", C.c_int32), ("activation", C.c_int32), ("precision", C.c_int32),
                ("sentinel_fp64", C.c_int32), ("bank_slots", C.c_int64)]


class AssocArgs(C.Structure):
    _fields_ = [("T", C.c_int32), ("D", C.c_int32), ("L", C.c_int32), ("C", C.c_int32), ("use_kalman", C.c_int32),
                ("mem_slots", c_i32p), ("mem_ltwh", c_f64p), ("det_slots", c_i32p), ("det_ltwh", c_f64p),
                ("dists", c_f64p), ("kal_slots", c_i32p), ("kal_ltwh", c_f64p),
                ("probs", c_f32p), ("logits", c_f32p), ("cand", c_i32p), ("pe_index", c_i32p),
                ("mem_emb", c_f32p), ("can_emb", c_f32p), ("cand_rows", c_f32p), ("mem_logits", c_f32p),
                ("input_seq", c_f32p)]


class StepArgs(C.Structure):
    _fields_ = [("T", C.c_int32), ("D", C.c_int32), ("L", C.c_int32), ("C", C.c_int32),
                ("track_mean_dev", C.c_void_p), ("tracked_dev", C.c_void_p), ("det_tlbr_dev", C.c_void_p),
                ("mem_slots_dev", C.c_void_p), ("mem_ltwh_dev", C.c_void_p), ("det_slots_dev", C.c_void_p),
                ("kal_slots_dev", C.c_void_p), ("busca_thresh", C.c_float), ("reliable_dev", C.c_void_p),
                ("probs_dev", C.c_void_p), ("keep_dev", C.c_void_p), ("cand_dev", C.c_void_p),
                ("select_highest", C.c_int32), ("highest_min_thresh", C.c_float), ("keep_highest_value", C.c_int32),
                ("frame_dev", C.c_void_p), ("frame_H", C.c_int32), ("frame_W", C.c_int32)]


class DebugConvArgs(C.Structure):
    _fields_ = [("conv_index", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("use_tc", C.c_int32),
                ("mode", C.c_int32), ("in_bf16", C.c_void_p), ("in_scale", C.c_void_p), ("in_shift", C.c_void_p),
                ("e_scale", C.c_void_p), ("e_shift", C.c_void_p), ("idt_bf16", C.c_void_p), ("ds_index", C.c_int32),
                ("ds_H", C.c_int32), ("ds_W", C.c_int32), ("ds_in_bf16", C.c_void_p), ("ds_scale", C.c_void_p),
                ("ds_shift", C.c_void_p), ("out_bf16", C.c_void_p), ("stats_out", C.c_void_p), ("img_w", C.c_void_p)]


EXPORTS = [
    "busca_version", "busca_last_error", "busca_create", "busca_destroy", "busca_load_tensor", "busca_finalize",
    "busca_upload_frame", "busca_sync_frame", "busca_ingest_frame", "busca_camera_motion", "busca_bank_reserve", "busca_bank_capacity", "busca_crop", "busca_bank_upload",
    "busca_bank_download", "busca_center_distance", "busca_iou", "busca_detection_coverage", "busca_kalman_predict", "busca_kalman_update", "busca_match_round", "busca_linear_assignment", "busca_duplicate_tracks", "busca_motion_proposals", "busca_frame_geometry", "busca_frame_geometry_batch",
    "busca_reid_embed", "busca_associate", "busca_transformer", "busca_frame_step_dev", "busca_dev_alloc",
    "busca_dev_free", "busca_host_alloc", "busca_host_free", "busca_memcpy_h2d", "busca_memcpy_d2h", "busca_sync", "busca_stream", "busca_kernel_launches",
    "busca_set_profiling", "busca_last_profile", "busca_set_option", "busca_counter", "busca_debug_conv", "busca_debug_conv_ex", "busca_conv_info", "busca_debug_stem", "busca_debug_umma_rowshift", "busca_debug_maxpool", "busca_debug_gram",
]


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if the sources are newer and nvcc is present) and type the C ABI."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and os.path.exists(_build.NVCC):
        _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m busca_b200.build` (needs nvcc). "
                           "busca_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.busca_version.restype = C.c_char_p
    L.busca_last_error.restype = C.c_char_p
    L.busca_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.busca_destroy.argtypes = [vp]
    L.busca_destroy.restype = None
    L.busca_load_tensor.argtypes = [vp, C.c_char_p, vp, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]
    L.busca_finalize.argtypes = [vp]
    L.busca_upload_frame.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int64]
    L.busca_sync_frame.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int64, vp, C.c_int32, vp]
    L.busca_bank_reserve.argtypes = [vp, C.c_int64]
    L.busca_bank_capacity.argtypes = [vp]
    L.busca_bank_capacity.restype = C.c_int64
    L.busca_crop.argtypes = [vp, vp, C.c_int32, vp, vp]
    L.busca_bank_upload.argtypes = [vp, vp, C.c_int32, vp]
    L.busca_bank_download.argtypes = [vp, vp, C.c_int32, vp]
    L.busca_center_distance.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, vp]
    L.busca_iou.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, vp]
    L.busca_detection_coverage.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64), vp]
    L.busca_ingest_frame.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    L.busca_camera_motion.argtypes = [vp, vp, vp, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_double, vp, vp, vp]
    L.busca_kalman_predict.argtypes = [vp, vp, vp, vp, C.c_int32, vp, vp]
    L.busca_kalman_update.argtypes = [vp, vp, vp, vp, C.c_int32, vp, vp]
    L.busca_match_round.argtypes = [vp, vp, C.c_int32, vp, C.c_int32, vp, C.c_double, vp, vp, vp]
    L.busca_linear_assignment.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_double, vp, vp]
    L.busca_duplicate_tracks.argtypes = [vp, vp, vp, C.c_int32, vp, vp, C.c_int32, C.c_double, vp, vp]
    L.busca_motion_proposals.argtypes = [vp, vp, vp, C.c_int32, vp, vp, vp]
    L.busca_frame_geometry.argtypes = [vp, vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    L.busca_frame_geometry_batch.argtypes = [vp, C.c_int32, vp, vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp]
    L.busca_reid_embed.argtypes = [vp, vp, C.c_int32, vp]
    L.busca_associate.argtypes = [vp, C.POINTER(AssocArgs)]
    L.busca_transformer.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.busca_frame_step_dev.argtypes = [vp, C.POINTER(StepArgs)]
    L.busca_dev_alloc.argtypes = [vp, C.c_int64]
    L.busca_dev_alloc.restype = vp
    L.busca_dev_free.argtypes = [vp, vp]
    L.busca_dev_free.restype = None
    L.busca_host_alloc.argtypes = [vp, C.c_int64]
    L.busca_host_alloc.restype = vp
    L.busca_host_free.argtypes = [vp, vp]
    L.busca_host_free.restype = None
    L.busca_memcpy_h2d.argtypes = [vp, vp, vp, C.c_int64]
    L.busca_memcpy_d2h.argtypes = [vp, vp, vp, C.c_int64]
    L.busca_sync.argtypes = [vp]
    L.busca_stream.argtypes = [vp]
    L.busca_stream.restype = vp
    L.busca_kernel_launches.argtypes = [vp]
    L.busca_kernel_launches.restype = C.c_int64
    L.busca_set_profiling.argtypes = [vp, C.c_int32]
    L.busca_last_profile.argtypes = [vp]
    L.busca_last_profile.restype = C.c_char_p
    L.busca_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    L.busca_counter.argtypes = [vp, C.c_char_p]
    L.busca_counter.restype = C.c_int64
    L.busca_debug_conv.argtypes = [vp, C.c_int32, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp]
    L.busca_debug_conv_ex.argtypes = [vp, C.POINTER(DebugConvArgs)]
    L.busca_conv_info.argtypes = [vp, C.c_int32, vp]
    L.busca_debug_stem.argtypes = [vp, vp, C.c_int32, C.c_int32, vp, vp]
    L.busca_debug_umma_rowshift.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, vp]
    L.busca_debug_gram.argtypes = [vp, vp, C.c_int32, vp]
    L.busca_debug_maxpool.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, C.c_int32, vp, vp, vp]
    _lib = L
    return L


class BuscaError(RuntimeError):
    pass


def check(rc: int):
    if rc != 0:
        raise BuscaError(f"libbusca_b200 error {rc}: {load().busca_last_error().decode()}")
