"""ctypes binding of libwbg.so (include/wbg.h).  There is no CPU fallback: if the library is missing or the
call fails the error is raised to the caller."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwbg.so")

WBG_OK, WBG_EINVAL, WBG_ECAP, WBG_ECUDA, WBG_ENOMEM = 0, -1, -2, -3, -4
WBG_U8, WBG_F32 = 0, 1
WBG_CH_GRAD_HIST, WBG_CH_GRAD_MAG, WBG_CH_GRAD_MAG_HIST, WBG_CH_FPGA_HIST4_U1, WBG_CH_FPGA_MAG_U1 = 0, 1, 2, 3, 4
WBG_MAX_BINS, WBG_MAX_NORM, WBG_MAX_CHANNELS = 16, 8, 17
ABI_VERSION = 4


class ChannelOpts(C.Structure):
    _fields_ = [("shrink", C.c_int32), ("n_per_oct", C.c_int32), ("smooth", C.c_int32), ("kind", C.c_int32),
                ("n_bins", C.c_int32), ("full", C.c_int32), ("bias", C.c_float), ("norm", C.c_int32),
                ("eps", C.c_float), ("max_levels", C.c_int32),
                ("cos_t", C.c_double * WBG_MAX_BINS), ("sin_t", C.c_double * WBG_MAX_BINS)]


class Level(C.Structure):
    _fields_ = [("octave", C.c_int32), ("src_h", C.c_int32), ("src_w", C.c_int32), ("nh", C.c_int32),
                ("nw", C.c_int32), ("u", C.c_int32), ("v", C.c_int32), ("win_rows", C.c_int32),
                ("win_cols", C.c_int32), ("skipped", C.c_int32), ("chn_off", C.c_int64), ("win_off", C.c_int64),
                ("scale", C.c_double)]


class PlanInfo(C.Structure):
    _fields_ = [("H", C.c_int32), ("W", C.c_int32), ("n_levels", C.c_int32), ("n_octaves", C.c_int32),
                ("channels", C.c_int32), ("win_m", C.c_int32), ("win_n", C.c_int32), ("reserved", C.c_int32),
                ("chn_floats", C.c_int64), ("octave_elems", C.c_int64), ("windows", C.c_int64), ("n_loc", C.c_int64)]


class ModelDesc(C.Structure):
    _fields_ = [("win_m", C.c_int32), ("win_n", C.c_int32), ("channels", C.c_int32), ("n_stages", C.c_int32),
                ("max_nodes", C.c_int32), ("reserved", C.c_int32),
                ("n_nodes", C.c_void_p), ("feature", C.c_void_p), ("threshold", C.c_void_p), ("left", C.c_void_p),
                ("right", C.c_void_p), ("prediction", C.c_void_p), ("theta", C.c_void_p)]


# numpy view of wbg_hit (36 bytes, no padding)
HIT_DTYPE = np.dtype([("frame", "<i4"), ("level", "<i4"), ("r", "<i4"), ("c", "<i4"), ("score", "<f4"),
                      ("x1", "<f4"), ("y1", "<f4"), ("x2", "<f4"), ("y2", "<f4")])
assert HIT_DTYPE.itemsize == 36

# name -> (restype, argtypes); every symbol declared in include/wbg.h
_P, _I32, _I64, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_size_t
SYMBOLS = {
    "wbg_abi_version": (C.c_int, []),
    "wbg_last_error": (C.c_char_p, []),
    "wbg_device_count": (C.c_int, []),
    "wbg_plan_create": (C.c_int, [_I32, _I32, C.POINTER(ChannelOpts), _I32, _I32, _I32, C.POINTER(_P)]),
    "wbg_plan_create_levels": (C.c_int, [_I32, _I32, C.POINTER(ChannelOpts), _I32, _I32, _I32, C.POINTER(_I32), _I32, C.POINTER(_P)]),
    "wbg_plan_create_bands": (C.c_int, [_I32, _I32, C.POINTER(ChannelOpts), _I32, _I32, _I32, C.POINTER(_I32), _I32, C.POINTER(_P)]),
    "wbg_cascade_tile": (C.c_int, [_I32, _I32, _I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "wbg_plan_destroy": (None, [_P]),
    "wbg_plan_get_info": (C.c_int, [_P, C.POINTER(PlanInfo)]),
    "wbg_plan_get_levels": (C.c_int, [_P, C.POINTER(Level), _I32]),
    "wbg_pyramid_workspace_bytes": (_SZ, [_P, _I32, _I32]),
    "wbg_channel_pyramid": (C.c_int, [_P, _P, _I32, _I32, _P, _P, _SZ, _P]),
    "wbg_avg_pool_2": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "wbg_max_pool_2": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "wbg_smooth_image_3d": (C.c_int, [_P, _I32, _I32, _I32, _P, _P]),
    "wbg_gradients": (C.c_int, [_P, _I32, _I32, _P, _P, _P, _P]),
    "wbg_separable_convolve": (C.c_int, [_P, _I32, _I32, _P, _I32, _P, _I32, _P, _P, _P]),
    "wbg_model_create": (C.c_int, [C.POINTER(ModelDesc), C.POINTER(_P)]),
    "wbg_model_destroy": (None, [_P]),
    "wbg_cascade_workspace_bytes": (_SZ, [_P, _I32]),
    "wbg_cascade_scan": (C.c_int, [_P, _P, _P, _I32, _P, _I64, _P, _P, _P, _P, _SZ, _P]),
    "wbg_predict_workspace_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "wbg_predict_on_image": (C.c_int, [_P, _P, _I32, _I32, _P, _I64, _P, _P, _P, _SZ, _P]),
    "wbg_cascade_trace": (C.c_int, [_P, _P, _I32, _I32, _P, _P, _I64, _P, _P, _P]),
    "wbg_predict_samples": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "wbg_gather_samples": (C.c_int, [_P, _I32, _I32, _I32, _P, _P, _I64, _I32, _I32, _P, _P]),
    "wbg_profile_enable": (C.c_int, [_I32]),
    "wbg_profile_read": (C.c_int, [C.POINTER(C.c_double), C.POINTER(_I64)]),
    "wbg_cascade_counters_enable": (C.c_int, [_I32]),
    "wbg_cascade_counters_read": (C.c_int, [C.POINTER(C.c_uint64)]),
}
PROF_KINDS = ("level_kernel", "cascade_kernel")


class WbgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libwbg error {code}: {msg}")
        self.code = code
        self.msg = msg


_lib = None


def lib():
    """Load libwbg.so (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -m waldboost_b200.build` "
                              "(waldboost_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.wbg_abi_version() != ABI_VERSION:
            raise ImportError(f"{LIB_PATH}: ABI version {L.wbg_abi_version()} != {ABI_VERSION}; rebuild it")
        _lib = L
    return _lib


def last_error():
    return lib().wbg_last_error().decode("utf-8", "replace")


def check(code):
    if code != WBG_OK:
        raise WbgError(code, last_error())
