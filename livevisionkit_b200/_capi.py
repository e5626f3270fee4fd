"""ctypes binding of include/lvkb200.h (the C-ABI of liblvkb200.so).  No torch, no cv2, no oracle."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblvkb200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_NO_MODEL, ERR_CAPACITY = range(6)
BGR, BGRA, RGB, RGBA, YUV, GRAY, UNKNOWN = range(7)
MEM_HOST, MEM_DEVICE = 0, 1
STAGE_COUNT = 6
STAGE_NAMES = ("ingest", "pyramid", "fast", "lk", "estimate", "remap")

(DBG_DETECTION_IMAGE, DBG_DETECTED, DBG_LK_MATCHED, DBG_LK_STATUS, DBG_TRACKED, DBG_MATCHED, DBG_INLIERS,
 DBG_HOMOGRAPHY, DBG_MOTION, DBG_CORRECTION, DBG_WARP_TRANSFORM, DBG_PROPAGATED, DBG_FAST_COUNTS, DBG_MESH_ITERATIONS) = range(14)


class Settings(C.Structure):
    """lvkb200_settings — flat mirror of lvk::StabilizationFilterSettings."""
    _fields_ = [
        ("detection_resolution_width", C.c_int32), ("detection_resolution_height", C.c_int32),
        ("detection_regions_width", C.c_int32), ("detection_regions_height", C.c_int32),
        ("force_detection", C.c_int32),
        ("max_feature_density", C.c_float), ("min_feature_density", C.c_float), ("accumulation_rate", C.c_float),
        ("motion_resolution_width", C.c_int32), ("motion_resolution_height", C.c_int32),
        ("track_local_motions", C.c_int32),
        ("temporal_smoothing", C.c_float), ("local_smoothing", C.c_float),
        ("min_motion_samples", C.c_uint64),
        ("acceptance_threshold", C.c_float), ("uniformity_threshold", C.c_float),
        ("predictive_samples", C.c_uint64),
        ("corrective_limits_width", C.c_float), ("corrective_limits_height", C.c_float),
        ("smoothing_steps", C.c_float), ("response_rate", C.c_float),
        ("background_colour", C.c_double * 4),
        ("crop_to_stable_region", C.c_int32), ("stabilize_output", C.c_int32),
        ("min_scene_quality", C.c_float), ("min_tracking_quality", C.c_float),
    ]


class Result(C.Structure):
    _fields_ = [
        ("has_output", C.c_int32), ("out_timestamp", C.c_uint64), ("out_format", C.c_int32),
        ("tracking_stability", C.c_float), ("scene_quality", C.c_float), ("trust_factor", C.c_float),
        ("feature_count", C.c_int32), ("has_motion", C.c_int32),
    ]


class DeblockSettings(C.Structure):
    """lvkb200_deblock_settings — lvk::DeblockingFilterSettings."""
    _fields_ = [("detection_levels", C.c_uint32), ("block_size", C.c_uint32), ("filter_size", C.c_uint32),
                ("filter_scaling", C.c_float)]


class ScalingSettings(C.Structure):
    """lvkb200_scaling_settings — lvk::ScalingFilterSettings."""
    _fields_ = [("output_width", C.c_int32), ("output_height", C.c_int32), ("sharpness", C.c_float),
                ("yuv_input", C.c_int32)]


class ObsFrame(C.Structure):
    """lvkb200_obs_frame — the fields of obs_source_frame the ingest reads."""
    _fields_ = [("data", C.c_void_p * 4), ("linesize", C.c_uint32 * 4), ("width", C.c_uint32), ("height", C.c_uint32),
                ("format", C.c_int32), ("timestamp", C.c_uint64)]


# lvkb200_video_format
VIDEO_FORMATS = {"I420": 0, "I422": 1, "I444": 2, "I40A": 3, "I42A": 4, "YUVA": 5, "NV12": 6, "YVYU": 7, "YUY2": 8,
                 "UYVY": 9, "AYUV": 10, "Y800": 11, "BGR3": 12}


class KeyPoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("response", C.c_float), ("class_id", C.c_int32)]


# Every symbol include/lvkb200.h declares: name -> (restype, argtypes)
_vp, _sz, _i, _u8p, _fp, _dp = C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_double)
SYMBOLS = {
    "lvkb200_abi_version": (C.c_int, []),
    "lvkb200_device_count": (C.c_int, []),
    "lvkb200_last_error": (C.c_char_p, []),
    "lvkb200_status_string": (C.c_char_p, [C.c_int]),
    "lvkb200_set_assert_handler": (None, [_vp]),
    "lvkb200_settings_default": (None, [C.POINTER(Settings)]),
    "lvkb200_settings_obs_homography": (None, [C.POINTER(Settings)]),
    "lvkb200_settings_obs_field": (None, [C.POINTER(Settings)]),
    "lvkb200_stream_create": (C.c_int, [_i, C.POINTER(Settings), C.POINTER(_vp)]),
    "lvkb200_stream_destroy": (None, [_vp]),
    "lvkb200_stream_configure": (C.c_int, [_vp, C.POINTER(Settings)]),
    "lvkb200_stream_get_settings": (C.c_int, [_vp, C.POINTER(Settings)]),
    "lvkb200_stream_restart": (C.c_int, [_vp]),
    "lvkb200_stream_reset_context": (C.c_int, [_vp]),
    "lvkb200_stream_ready": (C.c_int, [_vp]),
    "lvkb200_stream_frame_delay": (C.c_uint64, [_vp]),
    "lvkb200_stream_stable_region": (C.c_int, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "lvkb200_stream_submit": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, C.c_uint64, _i, _vp, _sz, _i, C.POINTER(Result)]),
    "lvkb200_stream_sync": (C.c_int, [_vp]),
    "lvkb200_stream_prefetch": (C.c_int, [_vp, _vp, _sz, _i, _i]),
    "lvkb200_stream_prefetch_frame": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _i]),
    "lvkb200_stream_submit_async": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, C.c_uint64, _i, _vp, _sz, _i, C.POINTER(Result),
                                              C.POINTER(C.c_uint64)]),
    "lvkb200_stream_wait_output": (C.c_int, [_vp, C.c_uint64]),
    "lvkb200_stream_submit_batch": (C.c_int, [_vp, C.POINTER(_vp), _sz, _i, _i, _i, C.POINTER(C.c_uint64), _i, C.POINTER(_vp), _sz, _i, _i,
                                              C.POINTER(Result)]),
    "lvkb200_device_synchronize": (C.c_int, []),
    "lvkb200_stream_event_record": (C.c_int, [_vp, _i]),
    "lvkb200_stream_event_elapsed_ms": (C.c_int, [_vp, _i, _i, _fp]),
    "lvkb200_stream_set_debug_capture": (C.c_int, [_vp, _i]),
    "lvkb200_stream_debug_fetch": (C.c_int, [_vp, _i, _vp, _sz, C.POINTER(_sz)]),
    "lvkb200_stream_stage_times_us": (C.c_int, [_vp, _fp]),
    "lvkb200_stream_stage_totals_us": (C.c_int, [_vp, _dp, C.POINTER(C.c_uint64), _i]),
    "lvkb200_stream_set_profiling": (C.c_int, [_vp, _i]),
    "lvkb200_kernel_launch_count": (C.c_uint64, []),
    "lvkb200_set_remap_exact": (None, [_i]),
    "lvkb200_remap_exact": (C.c_int, []),
    "lvkb200_remap_homography": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _vp, _sz, _i, _dp, _u8p, _i]),
    "lvkb200_remap_mesh": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _vp, _sz, _i, _fp, _i, _i, _u8p, _i]),
    "lvkb200_warp_mesh_apply": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _vp, _sz, _i, _fp, _i, _i, _u8p, _i, _dp]),
    "lvkb200_detection_image": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _i, _u8p, _i, _i]),
    "lvkb200_fast_detect": (C.c_int, [_vp, _u8p, _i, _i, _i, _i, _i, _i, _i, C.POINTER(KeyPoint), _i, C.POINTER(_i)]),
    "lvkb200_lk_track": (C.c_int, [_vp, _u8p, _u8p, _i, _i, _fp, _i, _i, _fp, _u8p]),
    "lvkb200_find_homography": (C.c_int, [_vp, _fp, _fp, _i, C.c_float, _dp, _u8p]),
    "lvkb200_estimate_affine_partial": (C.c_int, [_vp, _fp, _fp, _i, C.c_float, _dp, _u8p]),
    "lvkb200_estimate_local_motions": (C.c_int, [_vp, _fp, _fp, _i, _fp, _fp, _u8p]),
    "lvkb200_deblock_settings_default": (None, [C.POINTER(DeblockSettings)]),
    "lvkb200_deblock": (C.c_int, [_vp, C.POINTER(DeblockSettings), _vp, _sz, _i, _i, _i, _i, _vp, _sz, _i]),
    "lvkb200_stream_set_deblocking": (C.c_int, [_vp, C.POINTER(DeblockSettings)]),
    "lvkb200_scaling_settings_default": (None, [C.POINTER(ScalingSettings)]),
    "lvkb200_upscale": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _vp, _sz, _i, _i, _i, _i]),
    "lvkb200_sharpen": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _vp, _sz, _i, C.c_float]),
    "lvkb200_scaling_filter": (C.c_int, [_vp, C.POINTER(ScalingSettings), _vp, _sz, _i, _i, _i, _vp, _sz, _i]),
    "lvkb200_video_format_ocl": (C.c_int, [_i]),
    "lvkb200_frame_upload": (C.c_int, [_vp, C.POINTER(ObsFrame), _i, _vp, _sz, _i]),
    "lvkb200_frame_download": (C.c_int, [_vp, _vp, _sz, _i, _i, _i, _i, C.POINTER(ObsFrame), _i]),
    "lvkb200_stream_submit_obs": (C.c_int, [_vp, C.POINTER(ObsFrame), _i, C.POINTER(ObsFrame), _i, C.POINTER(Result)]),
    "lvkb200_stream_prefetch_obs": (C.c_int, [_vp, C.POINTER(ObsFrame)]),
    "lvkb200_stream_submit_obs_batch": (C.c_int, [_vp, C.POINTER(ObsFrame), C.POINTER(ObsFrame), _i, C.POINTER(Result)]),
    "lvkb200_stream_submit_obs_async": (C.c_int, [_vp, C.POINTER(ObsFrame), C.POINTER(ObsFrame), C.POINTER(Result), C.POINTER(C.c_uint64)]),
}

_lib = None


class LvkB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"lvkb200 status {status}: {message}")
        self.status = status


def load():
    """Loads liblvkb200.so.  Fails loudly if the CUDA extension has not been built — there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                              f"g.build()'` (nvcc, sm_100a). livevisionkit_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(status: int):
    if status != OK:
        raise LvkB200Error(status, load().lvkb200_last_error().decode("utf-8", "replace"))
