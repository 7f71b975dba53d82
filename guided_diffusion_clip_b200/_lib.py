"""ctypes binding of libgd_b200.so (include/gd_b200.h).  This is the ONLY compute backend of the package:
there is no CPU or eager-PyTorch fallback, and a missing library is a hard error."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GD_B200_LIB: alternative build of the same ABI (A/B timing experiments only)
LIB_PATH = os.environ.get("GD_B200_LIB") or os.path.join(_HERE, "libgd_b200.so")

# enums (keep in sync with include/gd_b200.h)
RES_NONE, RES_SAME, RES_UPSAMPLE2, RES_AVGPOOL2 = 0, 1, 2, 3
OUT_NHWC_F16, OUT_NCHW_F32 = 0, 1
GN_SAME, GN_AVGPOOL2, GN_UPSAMPLE2 = 0, 1, 2
CONV_GN_OFF, CONV_GN_SAME, CONV_GN_UPSAMPLE2 = 0, 1, 2
QKV_LEGACY, QKV_NEW = 0, 1
VAR_LEARNED_RANGE, VAR_FIXED, VAR_LEARNED = 0, 1, 2
MEAN_EPSILON, MEAN_START_X = 0, 1
(COEF_SQRT_RECIP_ACP, COEF_SQRT_RECIPM1_ACP, COEF_POST_MEAN1, COEF_POST_MEAN2, COEF_LOG_BETA, COEF_POST_LOGVAR,
 COEF_FIXED_VAR, COEF_FIXED_LOGVAR, COEF_ACP, COEF_ACP_PREV, COEF_NONZERO, COEF_ACP_NEXT) = range(12)
DDIM_REVERSE = 2
COEF_STRIDE = 12

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ConvDesc(C.Structure):
    _fields_ = [
        ("a0", vp), ("c0", i32), ("ld0", i32), ("taps", i32),
        ("a1", vp), ("c1", i32), ("ld1", i32),
        ("n", i32), ("h", i32), ("w", i32),
        ("wpack", vp), ("k_total", i32), ("n_pad", i32),
        ("bias", vp), ("cout", i32),
        ("res", vp), ("ld_res", i32), ("res_mode", i32),
        ("out", vp), ("ld_out", i32), ("out_mode", i32),
        ("bn", i32), ("out_scale", f32),
        ("stats_out", vp),
        ("gn_mode", i32), ("gn_silu", i32), ("gn_coef", vp),
        ("splitk_ws", vp), ("splitk_ws_bytes", i64),
    ]


class ConvInDesc(C.Structure):
    _fields_ = [
        ("x", vp), ("wpack", vp), ("bias", vp), ("out", vp), ("stats_out", vp),
        ("n", i32), ("cin", i32), ("h", i32), ("w", i32), ("cout", i32), ("ld_out", i32),
    ]


class PosteriorDesc(C.Structure):
    _fields_ = [
        ("x", vp), ("model_out", vp), ("grad", vp), ("noise", vp), ("sample", vp), ("pred_xstart", vp),
        ("mean_out", vp), ("var_out", vp), ("logvar_out", vp),
        ("coef", vp), ("t", vp),
        ("n", i32), ("c", i32), ("hw", i32),
        ("var_type", i32), ("mean_type", i32), ("clip_denoised", i32), ("ddim", i32), ("eta", f32),
        ("num_timesteps", i32),
    ]


# name -> (restype, argtypes); every symbol declared in include/gd_b200.h
SIGNATURES = {
    "gd_last_error": (C.c_char_p, []),
    "gd_version": (C.c_int, []),
    "gd_launch_count": (i64, []),
    "gd_launch_count_reset": (None, []),
    "gd_conv_igemm": (C.c_int, [C.POINTER(ConvDesc), vp]),
    "gd_conv_stats_rows": (i64, [i32, i32, i32, C.POINTER(i32)]),
    "gd_conv_gn_fusable": (C.c_int, [i32, i32]),
    "gd_conv_splitk_ws_bytes": (i64, [C.POINTER(ConvDesc)]),
    "gd_groupnorm_finalize_partials": (C.c_int, [vp, i32, i32, vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, i32, vp,
                                                 vp]),
    "gd_groupnorm_coef": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, vp, vp]),
    "gd_im2col3x3_small_cin": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "gd_groupnorm_stats": (C.c_int, [vp, i32, i32, i32, i32, f32, vp, vp, vp]),
    "gd_groupnorm_ws_floats": (i64, [i32, i32, i32]),
    "gd_groupnorm_apply": (C.c_int, [vp, i32, vp, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "gd_groupnorm_bwd": (C.c_int, [vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, i32, i32, vp, i32, vp, i32, i32, i32,
                                   i32, i32, i32, vp]),
    "gd_attention_fwd": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, i32, i32, vp]),
    "gd_attention_bwd": (C.c_int, [vp, i32, vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "gd_attention_fwd_masked": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "gd_attention_bwd_masked": (C.c_int, [vp, i32, vp, i32, vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "gd_layernorm_fwd": (C.c_int, [vp, i32, vp, vp, f32, vp, i32, vp, i32, i32, vp]),
    "gd_layernorm_bwd": (C.c_int, [vp, i32, vp, vp, vp, i32, vp, i32, vp, i32, i32, i32, vp]),
    "gd_quickgelu_fwd": (C.c_int, [vp, i32, vp, i32, i32, i32, vp]),
    "gd_quickgelu_bwd": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, i32, vp]),
    "gd_clip_preprocess_fwd": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, i32, vp]),
    "gd_clip_preprocess_bwd": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
    "gd_clip_head": (C.c_int, [vp, i32, vp, vp, i32, f32, f32, vp, vp, i32, i32, i32, i32, vp]),
    "gd_timestep_embedding": (C.c_int, [vp, vp, i32, i32, vp]),
    "gd_linear_f32": (C.c_int, [vp, i32, vp, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, vp]),
    "gd_embedding_gather": (C.c_int, [vp, vp, vp, i32, i32, i32, vp]),
    "gd_attnpool_ws_floats": (i64, [i32, i32, i32]),
    "gd_attnpool_fwd": (C.c_int, [vp, i32, vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "gd_attnpool_bwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
    "gd_logsoftmax_select_bwd": (C.c_int, [vp, vp, vp, i32, i32, f32, vp]),
    "gd_posterior_step": (C.c_int, [C.POINTER(PosteriorDesc), vp]),
    "gd_to_uint8_nhwc": (C.c_int, [vp, vp, i32, i32, i32, i32, vp]),
    "gd_conv_in3x3": (C.c_int, [C.POINTER(ConvInDesc), vp]),
    "gd_im2col3x3_s2_nhwc": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "gd_upsample2_nhwc": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "gd_col2im3x3_s2_nhwc": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "gd_add_emb_nhwc": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "gd_attention_fwd_hd": (C.c_int, [vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, vp]),
    "gd_tap_gather3x3": (C.c_int, [vp, i32, vp, vp, i32, i32, i32, i32, C.c_float, vp]),
    "gd_nchw_f32_to_nhwc_f16": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "gd_nhwc_f16_to_nchw_f32": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "gd_bilinear_upsample_nchw": (C.c_int, [vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp]),
}

# development hooks (include/gd_b200_devtools.h), bound only if the library was built with them
DEV_SIGNATURES = {
    "gd_debug_set": (None, [C.c_int, C.c_int]),
    "gd_bw_probe": (C.c_int, [i32, i32, vp, vp, i64, vp]),
}

_lib = None


class GdError(RuntimeError):
    pass


def load():
    """Load the kernel library (building nothing: see _build.py / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GdError(
            f"{LIB_PATH} is missing. Build it with `python -m guided_diffusion_clip_b200._build` "
            "(nvcc, sm_100a). There is no fallback compute path."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header disagree
        fn.restype = res
        fn.argtypes = args
    for name, (res, args) in DEV_SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is not None:
            fn.restype = res
            fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().gd_last_error().decode("utf-8", "replace")
        raise GdError(f"{what or 'gd call'} failed (rc={rc}): {msg}")
