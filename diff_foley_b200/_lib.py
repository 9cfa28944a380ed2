"""ctypes binding of libdfb.so (the C ABI declared in include/dfb.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is missing
(`python -m diff_foley_b200.build` was not run) or a call fails, a RuntimeError is raised with the
library's own message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libdfb.so")

_lib = None


class UnetCfg(C.Structure):
    _fields_ = [
        ("in_channels", C.c_int32),
        ("model_channels", C.c_int32),
        ("out_channels", C.c_int32),
        ("num_res_blocks", C.c_int32),
        ("n_channel_mult", C.c_int32),
        ("channel_mult", C.c_int32 * 8),
        ("n_attention_resolutions", C.c_int32),
        ("attention_resolutions", C.c_int32 * 8),
        ("num_heads", C.c_int32),
        ("context_dim", C.c_int32),
        ("latent_h", C.c_int32),
        ("latent_w", C.c_int32),
        ("max_context_len", C.c_int32),
        ("max_batch", C.c_int32),
    ]


class OpInfo(C.Structure):
    _fields_ = [("kind", C.c_char * 24), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("splits", C.c_int32), ("ctas", C.c_int32), ("flops", C.c_double), ("bytes", C.c_double),
                ("ms", C.c_float)]


_vp, _i, _f, _fp, _sz = C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_size_t
_i64p = C.POINTER(C.c_int64)

# name -> (restype, argtypes); every symbol include/dfb.h declares
SIGNATURES = {
    "dfb_last_error": (C.c_char_p, []),
    "dfb_version": (C.c_char_p, []),
    "dfb_unet_create": (_i, [C.POINTER(UnetCfg), _i, C.POINTER(_vp)]),
    "dfb_unet_set_weight": (_i, [_vp, C.c_char_p, _fp, _i64p, _i]),
    "dfb_unet_finalize": (_i, [_vp]),
    "dfb_unet_num_weights": (_i, [_vp]),
    "dfb_unet_weight_name": (C.c_char_p, [_vp, _i]),
    "dfb_unet_set_context": (_i, [_vp, _fp, _i, _i, _vp]),
    "dfb_unet_forward": (_i, [_vp, _fp, _i, _vp, _i, _fp, _i, _fp, _i, _vp]),
    "dfb_ddim_sample": (_i, [_vp, _fp, _fp, _fp, _i, _i, _f, _i, _i64p, C.POINTER(_f), C.POINTER(_f),
                             C.POINTER(_f), C.POINTER(_f), _fp, _fp, _fp, _vp]),
    "dfb_dpm_solver_sample": (_i, [_vp, _fp, _fp, _fp, _i, _i, _f, _i, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f),
                                   C.POINTER(_f), C.POINTER(_f), C.POINTER(_f), C.POINTER(C.c_int32), _fp, _vp]),
    "dfb_groupnorm_bwd": (_i, [_fp, _i, _i, _i, _fp, _fp, _f, _i, _fp, _fp, _fp, _vp, _vp]),
    "dfb_layernorm_bwd": (_i, [_fp, _i, _i, _fp, _f, _fp, _fp, _fp, _vp, _vp]),
    "dfb_attention_bwd": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _fp, _i, _i, _i, _i, _i, _i, _f, _vp, _i, _vp, _i,
                               _vp, _i, _fp, _fp, _vp]),
    "dfb_geglu_fwd": (_i, [_fp, C.c_longlong, _i, _vp, _vp]),
    "dfb_geglu_bwd": (_i, [_fp, _fp, C.c_longlong, _i, _vp, _vp]),
    "dfb_col2im_s2": (_i, [_fp, _i, _i, _i, _i, _fp, _fp, _vp, _vp]),
    "dfb_classifier_head": (_i, [_fp, _i, _i, _i, _fp, _fp, _f, _fp, _vp, _vp]),
    "dfb_scale_f32": (_i, [_fp, _f, C.c_longlong, _vp]),
    "dfb_cast_f16": (_i, [_fp, _vp, C.c_longlong, _vp]),
    "dfb_stem_conv": (_i, [_fp, _i, _i, _i, _i, _fp, _fp, _i, _fp, _vp]),
    "dfb_head_conv": (_i, [_vp, _i, _i, _i, _i, _fp, _fp, _i, _fp, _vp]),
    "dfb_frames_resize": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _vp, _i, _i, _vp, _fp, _vp, _vp]),
    "dfb_comm_unique_id": (_i, [_vp]),
    "dfb_comm_init": (_i, [_vp, _i, _i, _vp]),
    "dfb_comm_destroy": (_i, [_vp]),
    "dfb_unet_profile": (_i, [_vp, _fp, _i, _vp, _i, _fp, _i, _i, C.POINTER(OpInfo), _i, C.POINTER(_i), _vp]),
    "dfb_unet_trace": (_i, [_vp, _fp, _i, _vp, _i, _fp, _i, C.POINTER(C.c_ulonglong), _i, C.POINTER(_i), _vp]),
    "dfb_debug_igemm_force": (None, [_i, _i]),
    "dfb_debug_igemm_pair": (None, [_i]),
    "dfb_unet_debug_num_taps": (_i, [_vp, _i]),
    "dfb_unet_debug_tap": (_i, [_vp, _i, _i, C.c_char_p, _i, C.POINTER(C.c_int32), _fp, _vp]),
    "dfb_unet_last_launch_count": (C.c_longlong, [_vp]),
    "dfb_unet_destroy": (_i, [_vp]),
    "dfb_gemm": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, _i, _fp, _vp, _i, _vp]),
    "dfb_gemm_stats": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, _fp, _vp, _i, _vp, C.POINTER(_i), _vp]),
    "dfb_gemm_ln": (_i, [_vp, _vp, _i, _i, _i, _fp, _fp, _vp, _i, _f, _i, _fp, _vp, _i, _vp]),
    "dfb_conv3x3": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _fp, _fp, _fp, _i, _fp, _vp, _i, _vp]),
    "dfb_conv3x3_cat": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _fp, _fp, _fp, _fp, _vp, _i, _vp]),
    "dfb_conv_taps": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _fp, _vp, _i, _fp, _vp, _i, _vp]),
    "dfb_im2col_f16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dfb_pool2d_f16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "dfb_groupnorm": (_i, [_fp, _i, _fp, _i, _i, _i, _fp, _fp, _f, _i, _vp, _vp, _vp]),
    "dfb_softmax_rows": (_i, [_fp, _i, _i, _f, _vp, _vp]),
    "dfb_layernorm": (_i, [_fp, _i, _i, _fp, _fp, _f, _vp, _vp]),
    "dfb_attention": (_i, [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _f, _vp]),
    "dfb_temb": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "dfb_upsample2x_f16": (_i, [_fp, _vp, _i, _i, _i, _i, _vp]),
    "dfb_im2col_s2": (_i, [_fp, _vp, _i, _i, _i, _i, _vp]),
    "dfb_ddim_step": (_i, [_fp, _fp, _fp, _fp, _f, _f, _f, _f, _f, _f, _fp, _fp, _sz, _vp]),
}


def lib():
    """Loads libdfb.so once; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -m diff_foley_b200.build` "
                "(there is no fallback implementation)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dfb_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"libdfb {what} failed (code {rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None) as a void*."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
