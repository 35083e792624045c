"""ctypes binding of librespmon_b200.so (include/respmon_b200.h).

The shared library is the only compute path of this package.  If it is missing or fails to load, importing the
binding raises: there is deliberately no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# RESPMON_B200_LIB selects another build of the same library (developer A/B timing, see build.build_variant)
LIB_PATH = os.environ.get("RESPMON_B200_LIB") or os.path.join(_PKG, "librespmon_b200.so")

RM_OK = 0
RM_U8, RM_F32, RM_F64, RM_BGR8 = 0, 1, 2, 3
CLIP_OK, CLIP_NO_ROI, CLIP_NO_CORNERS, CLIP_TRACK_LOST, CLIP_NO_PEAKS = range(5)
CLIP_STATUS_NAMES = ("OK", "NO_ROI", "NO_CORNERS", "TRACK_LOST", "NO_PEAKS")


class RmParams(C.Structure):
    _fields_ = [
        ("pyramid_levels", C.c_int32), ("skip_levels_at_top", C.c_int32),
        ("freq_min", C.c_double), ("freq_max", C.c_double), ("amplification", C.c_double),
        ("temporal_threshold", C.c_double), ("threshold", C.c_int32),
        ("max_corners", C.c_int32), ("quality_level", C.c_double), ("min_distance", C.c_int32),
        ("block_size", C.c_int32), ("lk_win", C.c_int32), ("lk_max_level", C.c_int32), ("lk_max_iter", C.c_int32),
        ("lk_eps", C.c_double), ("lk_min_eig", C.c_double), ("gaussian_cutoff", C.c_double),
        ("filter_order", C.c_int32), ("measure_buffer_len", C.c_int32), ("measure_init_len", C.c_int32),
        ("peak_threshold", C.c_double),
    ]


class RmResult(C.Structure):
    _fields_ = [("bpm", C.c_double), ("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("status", C.c_int32), ("n_peaks", C.c_int32)]


class RmClipDesc(C.Structure):
    """rm_clip_desc (include/respmon_b200.h): one clip of a ragged batch."""
    _fields_ = [("frame_offset", C.c_int64), ("W", C.c_int32), ("H", C.c_int32), ("T", C.c_int32), ("row_stride", C.c_int32)]


class RmClipSpec(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("n_frames", C.c_int32), ("seed", C.c_int32),
                ("x0", C.c_int32), ("y0", C.c_int32), ("w0", C.c_int32), ("h0", C.c_int32)]


_H = C.c_void_p     # rm_handle*
_P = C.c_void_p     # device / host pointer
_S = C.c_void_p     # cudaStream_t
_i32, _i64, _f64, _sz = C.c_int32, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes).  Every symbol include/respmon_b200.h declares must be listed here
# (tests/test_cabi_symbols.py cross-checks the header against this table and the built library).
SIGNATURES = {
    "rm_version": (_i32, []),
    "rm_default_params": (_i32, [C.POINTER(RmParams)]),
    "rm_create": (_i32, [C.POINTER(RmParams), _i32, C.POINTER(_H)]),
    "rm_destroy": (_i32, [_H]),
    "rm_last_error": (C.c_char_p, [_H]),
    "rm_level_sizes": (_i32, [_i32, _i32, _i32, C.POINTER(_i32)]),
    "rm_temporal_bounds": (_i32, [_i32, _f64, _f64, _f64, C.POINTER(_i32), C.POINTER(_i32)]),
    "rm_butter_lowpass": (_i32, [_i32, _f64, C.POINTER(_f64), C.POINTER(_f64)]),
    "rm_lossy_u8_lut": (_i32, [C.POINTER(C.c_uint8)]),
    "rm_synth_clips": (_i32, [_H, _P, _P, _i32, _P, _S]),
    "rm_bgr_to_gray": (_i32, [_H, _P, _P, _i64, _S]),
    "rm_to_f64": (_i32, [_H, _P, _i32, _P, _i64, _S]),
    "rm_f64_to_u8": (_i32, [_H, _P, _P, _i64, _S]),
    "rm_pyr_down_f64": (_i32, [_H, _P, _P, _i64, _i32, _i32, _S]),
    "rm_pyr_up_f64": (_i32, [_H, _P, _P, _P, _i32, _i64, _i32, _i32, _i32, _i32, _S]),
    "rm_lap_record_len": (_i32, [_H, _i32, _i32, C.POINTER(_i64)]),
    "rm_pyramid_workspace_bytes": (_i32, [_H, _i32, _i32, _i64, C.POINTER(_sz)]),
    "rm_heatmap_workspace_bytes": (_i32, [_H, _i32, _i32, _i32, _i32, C.POINTER(_sz)]),
    "rm_pyramid_build": (_i32, [_H, _P, _i32, _i64, _i32, _i32, _P, _P, _sz, _S]),
    "rm_pyramid_build_clips": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _P, _P, _sz, _S]),
    "rm_temporal_bandpass": (_i32, [_H, _P, _P, _i32, _i32, _i64, _f64, _S]),
    "rm_heatmap": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _P, _P, _sz, _S]),
    "rm_roi_workspace_bytes": (_i32, [_H, _i32, _i32, _i32, C.POINTER(_sz)]),
    "rm_roi_select": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _P, _P, _sz, _S]),
    "rm_volume_clip_mean": (_i32, [_H, _P, _P, _P, _P, _i32, _i64, _f64, _P, _S]),
    "rm_measure_workspace_bytes": (_i32, [_H, _i32, _i32, _i32, _i32, C.POINTER(_sz)]),
    "rm_measure_flow": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _i32, _i32, _i32, _i32, _P, _P, _P, _P, _P, _sz, _S]),
    "rm_measure_flow_debug": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _i32, _i32, _i32, _i32, _P, _P, _P, _P, _P, _P, _sz,
                                     _S]),
    "rm_measure_average": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _i32, _i32, _P, _S]),
    "rm_signal_bpm": (_i32, [_H, _P, _i32, _i32, _f64, _P, _P, _P, _P, _P, _S]),
    "rm_measure_signal": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _i32, _i32, _i32, _i32, _f64, _P, _P, _P, _P, _P, _P,
                                 _P, _P, _P, _sz, _S]),
    "rm_measure_signal_stream": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _i32, _i32, _i32, _i32, _i32, _f64, _P, _P, _P,
                                        _P, _P, _P, _P, _P, _P, _sz, _S]),
    "rm_measure_average_stream": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _i32, _i32, _i32, _f64, _P, _P, _P, _P, _P, _P,
                                         _S]),
    "rm_crop_frames": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _i32, _P, _i32, _i32, _P, _i32, _i32, _S]),
    "rm_crop_frames_ragged": (_i32, [_H, _P, _i32, _P, _i32, _P, _i32, _i32, _P, _i32, _i32, _S]),
    "rm_crop_to_ring": (_i32, [_H, _P, _i32, _i32, _i32, _i32, _P, _P, _i32, _i32, _i32, _i32, _S]),
    "rm_pack_results": (_i32, [_H, _P, _P, _P, _P, _i32, _i32, _P, _S]),
    "rm_pack_results_stream": (_i32, [_H, _P, _P, _P, _P, _i32, _i32, _i32, _P, _S]),
    "rm_launch_count": (_i64, [_H]),
    "rm_set_option": (_i32, [_H, C.c_char_p, _i64]),
    "rm_profile_enable": (_i32, [_H, _i32]),
    "rm_profile_reset": (_i32, [_H]),
    "rm_profile_collect": (_i32, [_H]),
    "rm_profile_slot": (_i32, [_H, _i32, C.POINTER(C.c_char_p), C.POINTER(_f64), C.POINTER(_f64)]),
    "rm_profile_entry": (_i32, [_H, _i32, C.POINTER(C.c_char_p), C.POINTER(_f64), C.POINTER(_i64)]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library (once).  Raises if it has not been built -- there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "respmon_b200: %s is missing. Build it with `python -m respmon_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


class RmError(RuntimeError):
    pass


def check(handle, rc: int, what: str):
    if rc != RM_OK:
        msg = lib().rm_last_error(handle) if handle else b""
        raise RmError("%s failed (rc=%d): %s" % (what, rc, (msg or b"").decode(errors="replace")))
