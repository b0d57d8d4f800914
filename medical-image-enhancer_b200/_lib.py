"""ctypes binding of libmdimg_b200.so (C ABI declared in ``include/mdimg_b200.h``).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is present,
the first compute call raises.  Loading the library itself needs no GPU (symbol checks run on
the CPU-only build box).
"""

from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

LIB_NAME = "libmdimg_b200.so"
LIB_PATH = Path(__file__).resolve().parent / LIB_NAME

MDIMG_OK = 0
ERR_INVALID, ERR_CUDA, ERR_WORKSPACE, ERR_NO_DEVICE = 1, 2, 3, 4
METRIC_COLS = 24
VALIDATION_COLS = 2 * METRIC_COLS + 2

OP_NORMALIZE, OP_METRICS, OP_SIGMA, OP_QUALITY, OP_FULLREF, OP_WAVELET, OP_CLAHE, OP_GAMMA, \
    OP_UNSHARP, OP_LIGHT_DENOISE, OP_BILATERAL, OP_TV, OP_MINMAX, OP_VALIDATION = range(1, 15)

_p = C.c_void_p
_i = C.c_int
_d = C.c_double
_sz = C.c_size_t

# name -> (restype, argtypes).  Kept in one table so the CPU test-suite can check that every
# symbol declared in include/mdimg_b200.h is exported.
_IMG = [_i, _i, _i, _p, _i]          # n, h, w, sel, n_sel
_WS = [_p, _sz, _p]                  # ws, ws_bytes, stream
PROTOTYPES = {
    "mdimg_last_error": (C.c_char_p, []),
    "mdimg_version": (_i, []),
    "mdimg_launch_count": (C.c_ulonglong, []),
    "mdimg_selftest_div16": (_i, [C.POINTER(C.c_ulonglong), _p]),
    "mdimg_init": (_i, [_i]),
    "mdimg_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz), C.POINTER(_sz)]),
    "mdimg_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mdimg_minmax_f32": (_i, [_p, *_IMG, _p, *_WS]),
    "mdimg_mosaic_u8": (_i, [_p, _p, _p, *_IMG, _i, _i, *_WS]),
    "mdimg_normalize_u16": (_i, [_p, _p, *_IMG, *_WS]),
    "mdimg_normalize_f32": (_i, [_p, _p, *_IMG, *_WS]),
    "mdimg_ingest_u16": (_i, [_p, _p, *_IMG, _d, _d, _i, _i, _i, *_WS]),
    "mdimg_metrics": (_i, [_p, *_IMG, _i, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                           C.POINTER(C.c_float), _p, *_WS]),
    "mdimg_estimate_sigma": (_i, [_p, *_IMG, _p, *_WS]),
    "mdimg_quality": (_i, [_p, *_IMG, _i, _p, *_WS]),
    "mdimg_fullref": (_i, [_p, _p, *_IMG, _p, *_WS]),
    "mdimg_validation": (_i, [_p, _p, *_IMG, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                              _p, *_WS]),
    "mdimg_wavelet_denoise": (_i, [_p, _p, *_IMG, _i, _p, _d, _p, *_WS]),
    "mdimg_clahe": (_i, [_p, _p, *_IMG, _d, _i, _p, *_WS]),
    "mdimg_clahe_gamma": (_i, [_p, _p, *_IMG, _d, _i, _d, _p, *_WS]),
    "mdimg_gamma": (_i, [_p, _p, *_IMG, _d, _i, _p, *_WS]),
    "mdimg_unsharp": (_i, [_p, _p, *_IMG, C.POINTER(_d), _i, _d, _i, *_WS]),
    "mdimg_light_denoise": (_i, [_p, _p, *_IMG, _d, _p, *_WS]),
    "mdimg_bilateral": (_i, [_p, _p, *_IMG, _i, C.POINTER(_d), _d, _p]),
    "mdimg_tv_chambolle": (_i, [_p, _p, *_IMG, _d, _d, _i, _p, *_WS]),
    "mdimg_axpby": (_i, [_p, _p, _p, *_IMG, _d, _d, _i, _p]),
    "mdimg_clip01": (_i, [_p, _p, *_IMG, _p]),
    "mdimg_copy": (_i, [_p, _p, *_IMG, _p]),
    "mdimg_export_u16": (_i, [_p, _p, *_IMG, _p]),
}



MAX_PLAN_OPS = 64        # MDIMG_MAX_PLAN_OPS


class EnhancePlan(C.Structure):
    """mdimg_enhance_plan (include/mdimg_b200.h)."""
    _fields_ = [("n_ops", C.c_int32), ("ops", C.c_int32 * MAX_PLAN_OPS),
                ("clahe_clip_limit", _d), ("clahe_tile_size", C.c_int32), ("gamma", _d),
                ("unsharp_radius", _d), ("unsharp_amount", _d), ("denoise_hard", C.c_int32),
                ("post_denoise_strength", _d), ("bilateral_d", C.c_int32),
                ("bilateral_sigma_color", _d), ("bilateral_sigma_space", _d), ("tv_denoise_weight", _d)]


class EnhanceTables(C.Structure):
    """mdimg_enhance_tables (include/mdimg_b200.h)."""
    _fields_ = [("gauss_radius", C.c_int32), ("gauss_taps", _d * 13),
                ("bilateral_d_eff", C.c_int32), ("bilateral_spatial", _d * 81),
                ("pct_lo", C.c_int32 * 5), ("pct_hi", C.c_int32 * 5), ("pct_gamma", C.c_float * 5)]


class ValidationScalars(C.Structure):
    """mdimg_validation_scalars (include/mdimg_b200.h)."""
    _fields_ = [("ssim", _d), ("psnr", _d), ("quality_improvement", _d),
                ("meets_ssim", C.c_int32), ("meets_psnr", C.c_int32), ("meets_improvement", C.c_int32), ("passes", C.c_int32),
                ("niqe_before", _d), ("niqe_after", _d), ("niqe_improved", C.c_int32),
                ("contrast_gain", _d), ("sharpness_gain", _d), ("noise_change", _d),
                ("entropy_change", _d), ("snr_change", _d), ("cnr_change", _d), ("edge_density_change", _d),
                ("histogram_spread_change", _d), ("edge_ratio", _d), ("local_contrast_change", _d),
                ("gradient_strength_change", _d), ("gradient_entropy_change", _d)]


STEP_NAMES = ("denoise", "clahe", "gamma", "unsharp", "post_denoise", "bilateral", "tv_denoise")
FLAG_HALO, FLAG_NOISE_GUARD, FLAG_OVER_PROCESSED, FLAG_ERR_CLAHE_RANGE, FLAG_ERR_GAMMA_NEG = 1, 2, 4, 8, 16
OP_ENHANCE = 15
ISSUE_BITS = {"noise": 1, "blur": 2, "low_contrast": 4, "clipping_low": 8, "clipping_high": 16}

PROTOTYPES.update({
    "mdimg_plan_clamp": (_i, [C.POINTER(EnhancePlan)]),
    "mdimg_enhance_tables_default": (_i, [C.POINTER(EnhancePlan), _i, _i, C.POINTER(EnhanceTables)]),
    "mdimg_enhance": (_i, [_p, _p, _i, _i, _i, C.POINTER(EnhancePlan), C.POINTER(EnhanceTables), _p, _p,
                           C.POINTER(C.c_int32), C.POINTER(C.c_int32), *_WS]),
    "mdimg_detect_issues": (_i, [C.POINTER(_d)]),
    "mdimg_validation_scalars_of": (_i, [C.POINTER(_d), C.POINTER(ValidationScalars)]),
    "mdimg_objective_score": (_i, [C.POINTER(ValidationScalars), C.POINTER(_d), C.POINTER(_d)]),
    "mdimg_enhance_issues": (_i, [_p, _p, _i, _i, _i, _i, C.POINTER(EnhanceTables), _p, C.POINTER(C.c_int32), *_WS]),
})

_lock = threading.Lock()
_lib = None
_initialised_devices: set[int] = set()


class MdimgError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def load_library() -> C.CDLL:
    """dlopen libmdimg_b200.so and attach prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA extension has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
                "mdimg_b200 has no CPU fallback."
            )
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)       # AttributeError here means the ABI and the binding drifted
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load_library().mdimg_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int) -> None:
    if rc == MDIMG_OK:
        return
    msg = last_error()
    if rc == ERR_INVALID:
        raise ValueError(msg)
    raise MdimgError(rc, msg or f"mdimg_b200 call failed with status {rc}")


def ensure_device(index: int) -> None:
    """mdimg_init once per device; raises (no fallback) when the device is not an sm_100 GPU."""
    if index in _initialised_devices:
        return
    lib = load_library()
    check(lib.mdimg_init(int(index)))
    _initialised_devices.add(index)
