"""ctypes binding of include/spfe.h.  There is no fallback: if libspfe.so is
missing or no B200 is present the calls raise."""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

DESC_DIM = 256
EMIT_HEAT, EMIT_COV, MATCH_PREV, EMIT_HEAT_INV, LAZY_HEAT, DESC_F16, EXACT = 1, 2, 4, 8, 16, 32, 64
OK, ERR_INVALID, ERR_EMPTY, ERR_WEIGHTS, ERR_NO_DEVICE, ERR_CUDA, ERR_STATE = 0, -1, -2, -3, -4, -5, -6

EXPORTS = ["spfe_default_config", "spfe_create", "spfe_destroy", "spfe_last_error", "spfe_extract", "spfe_submit",
           "spfe_wait", "spfe_last_d2h_bytes", "spfe_fetch_heat", "spfe_submit_pinned", "spfe_host_alloc", "spfe_host_free", "spfe_submit_device", "spfe_slot_sync", "spfe_match_mutual_nn", "spfe_match_knn2", "spfe_desc_set_create", "spfe_desc_set_destroy", "spfe_desc_set_size", "spfe_desc_set_upload", "spfe_desc_set_from_frame", "spfe_match_mutual_nn_sets", "spfe_match_knn2_sets", "spfe_search_guided", "spfe_search_guided_sets", "spfe_guided_last_rounds", "spfe_dust_pose_optimize", "spfe_dust_pose_optimize_batch", "spfe_dust_linearize", "spfe_set_score_threshold", "spfe_reset_stream", "spfe_timer_start",
           "spfe_timer_stop", "spfe_check_weights", "spfe_l2", "spfe_debug_read", "spfe_launch_count", "spfe_dom_timing", "spfe_dom_time", "spfe_profile_device"]


class Config(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("max_keypoints", C.c_int32),
                ("score_thresh", C.c_float), ("nms_radius", C.c_int32), ("border", C.c_int32), ("device_id", C.c_int32),
                ("max_batch", C.c_int32), ("num_slots", C.c_int32), ("flags", C.c_uint32), ("weights_path", C.c_char_p)]


_FP = C.POINTER(C.c_float)


class FrameOut(C.Structure):
    _fields_ = [("n", C.c_int32), ("kp_xy", _FP), ("kp_score", _FP), ("kp_response", _FP), ("desc", _FP),
                ("occ_grid", C.POINTER(C.c_int16)), ("dense_dust", _FP), ("semi_dust", _FP), ("heat", _FP),
                ("heat_inv", _FP), ("cov2", _FP), ("cov2_inv", _FP), ("n_prev", C.c_int32),
                ("match_prev", C.POINTER(C.c_int32)), ("match_dist", _FP), ("desc_f16", C.POINTER(C.c_uint16))]


class GuidedSearch(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("mode", C.c_int32), ("m", C.c_int32), ("n", C.c_int32),
                ("qdesc", C.c_void_p), ("qxy", C.c_void_p), ("qradius", C.c_void_p), ("qvalid", C.c_void_p),
                ("qblocks", C.c_void_p), ("kdesc", C.c_void_p), ("kp_un", C.c_void_p), ("occ_grid", C.c_void_p),
                ("grid_rows", C.c_int32), ("grid_cols", C.c_int32), ("kp_taken", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("best_init", C.c_float), ("th_le", C.c_float),
                ("th_lt", C.c_float), ("c2_adaptive", C.c_float)]


GUIDED_AREA, GUIDED_DUST_CELLS = 0, 1


class DustPose(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n", C.c_int32), ("Xw", C.c_void_p), ("dust", C.c_void_p),
                ("rows", C.c_int32), ("cols", C.c_int32), ("slot", C.c_int32), ("frame", C.c_int32),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("huber_delta", C.c_double), ("chi2_inlier", C.c_double), ("iterations", C.c_int32), ("reserved", C.c_int32)]


class StageTime(C.Structure):
    _fields_ = [("name", C.c_char * 24), ("ms", C.c_float), ("flop", C.c_double), ("bytes", C.c_double)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load libspfe.so (building it in-tree first if the sources are newer)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and _build.needs_build():
        _build.build_lib()
    if not os.path.exists(_build.LIB):
        raise RuntimeError(f"{_build.LIB} is missing: run `python -m sp_orb_slam_b200.build` (no CPU fallback exists)")
    L = C.CDLL(_build.LIB)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.spfe_default_config.argtypes = [C.POINTER(Config), i32, i32, i32]
    L.spfe_default_config.restype = None
    L.spfe_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.spfe_destroy.argtypes = [vp]
    L.spfe_destroy.restype = None
    L.spfe_last_error.argtypes = [vp]
    L.spfe_last_error.restype = C.c_char_p
    L.spfe_extract.argtypes = [vp, vp, C.c_size_t, C.POINTER(FrameOut)]
    L.spfe_submit.argtypes = [vp, i32, C.POINTER(vp), i32, C.c_size_t]
    L.spfe_wait.argtypes = [vp, i32, C.POINTER(FrameOut)]
    L.spfe_last_d2h_bytes.argtypes = [vp, i32]
    L.spfe_last_d2h_bytes.restype = i64
    L.spfe_fetch_heat.argtypes = [vp, i32, i32, vp, vp]
    L.spfe_submit_pinned.argtypes = [vp, i32, vp, i32]
    L.spfe_host_alloc.argtypes = [C.c_size_t]
    L.spfe_host_alloc.restype = vp
    L.spfe_host_free.argtypes = [vp]
    L.spfe_host_free.restype = None
    L.spfe_submit_device.argtypes = [vp, i32, vp, i32]
    L.spfe_slot_sync.argtypes = [vp, i32]
    L.spfe_match_mutual_nn.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.spfe_match_knn2.argtypes = [vp, vp, i32, vp, i32, vp, vp]
    L.spfe_desc_set_create.argtypes = [vp, i32, C.POINTER(vp)]
    L.spfe_desc_set_destroy.argtypes = [vp, vp]
    L.spfe_desc_set_destroy.restype = None
    L.spfe_desc_set_size.argtypes = [vp]
    L.spfe_desc_set_upload.argtypes = [vp, vp, vp, i32]
    L.spfe_desc_set_from_frame.argtypes = [vp, vp, i32, i32, vp, i32]
    L.spfe_match_mutual_nn_sets.argtypes = [vp, vp, vp, vp, vp]
    L.spfe_match_knn2_sets.argtypes = [vp, vp, vp, vp, vp]
    L.spfe_search_guided.argtypes = [vp, C.POINTER(GuidedSearch), vp, vp, vp]
    L.spfe_search_guided_sets.argtypes = [vp, C.POINTER(GuidedSearch), vp, vp, vp, vp, vp]
    L.spfe_guided_last_rounds.argtypes = [vp]
    L.spfe_dust_pose_optimize.argtypes = [vp, C.POINTER(DustPose), vp, vp, vp, C.POINTER(i32), C.POINTER(i32), vp]
    L.spfe_dust_pose_optimize_batch.argtypes = [vp, C.POINTER(DustPose), i32, vp, vp, vp, vp, vp]
    L.spfe_dust_linearize.argtypes = [vp, C.POINTER(DustPose), vp, vp, vp, vp, vp, vp]
    L.spfe_set_score_threshold.argtypes = [vp, C.c_float]
    L.spfe_reset_stream.argtypes = [vp, i32]
    L.spfe_timer_start.argtypes = [vp, i32]
    L.spfe_timer_stop.argtypes = [vp, i32, C.POINTER(C.c_float)]
    L.spfe_check_weights.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
    L.spfe_check_weights.restype = i64
    L.spfe_l2.argtypes = [vp, vp]
    L.spfe_l2.restype = C.c_float
    L.spfe_debug_read.argtypes = [vp, i32, C.c_char_p, vp, C.c_size_t]
    L.spfe_debug_read.restype = i64
    L.spfe_launch_count.argtypes = [vp]
    L.spfe_launch_count.restype = i64
    L.spfe_dom_timing.argtypes = [vp, i32]
    L.spfe_dom_time.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(i32)]
    L.spfe_profile_device.argtypes = [vp, i32, vp, i32, C.POINTER(StageTime), i32]
    _lib = L
    return L
