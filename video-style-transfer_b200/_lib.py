"""ctypes binding of csrc/libvst_b200.so (include/vst_b200.h).

The library is the product: if it is missing or a call fails this module raises - there is no
eager-PyTorch or CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB

_lib = None

c_f = C.POINTER(C.c_float)
vp = C.c_void_p
i32 = C.c_int
f32 = C.c_float
sz = C.c_size_t


class NetDesc(C.Structure):
    _fields_ = [("net", i32), ("in_ch", i32), ("c1", i32), ("c2", i32), ("c3", i32), ("d1", i32), ("d2", i32),
                ("N", i32), ("H", i32), ("W", i32), ("flags", i32)]


class ActDesc(C.Structure):
    _fields_ = [("H", i32), ("W", i32), ("C", i32), ("pad", i32), ("kind", i32), ("parity", i32)]


_i8x96 = C.c_byte * 96


class TapGemmDesc(C.Structure):
    _fields_ = [("a", vp), ("a_C", i32), ("a_X", i32), ("a_Y", i32), ("a_N", i32), ("a_P", i32),
                ("b", vp), ("b_K", i32), ("b_rows", i32), ("b_img_rows", i32),
                ("BK", i32), ("kb_per_tap", i32), ("n_taps", i32), ("n_phase", i32), ("n_ntile", i32), ("N_mma", i32),
                ("grid_h", i32), ("grid_w", i32), ("out_mul", i32),
                ("Hout", i32), ("Wout", i32), ("Cout", i32), ("out_cstride", i32),
                ("epi_mode", i32), ("act", i32), ("relu", i32),
                ("rc_k", i32), ("rc_co", i32), ("tile_step_x", i32), ("TW", i32), ("TH", i32), ("MT", i32),
                ("out", vp), ("out_u8", vp), ("bias", vp), ("stats", vp),
                ("tap_dx", _i8x96), ("tap_dy", _i8x96), ("tap_pl", _i8x96), ("ph_oy", C.c_byte * 4), ("ph_ox", C.c_byte * 4)]


class TapGemmPlanInfo(C.Structure):
    _fields_ = [(n, i32) for n in ("stream", "dyshare", "n_cols", "dy_max", "box_rows", "TW", "TH", "MT", "tiles_x", "tiles_y", "cta2")] + \
               [(n, C.c_byte * 48) for n in ("col_dx", "col_dy0", "col_pl", "col_n", "col_t0", "col_ts")]


class PcGemmDesc(C.Structure):
    _fields_ = [("a", vp), ("a_C", i32), ("a_X", i32), ("a_Y", i32), ("a_N", i32), ("a_P", i32),
                ("b", vp), ("b_C", i32), ("b_X", i32), ("b_Y", i32), ("b_N", i32), ("b_P", i32),
                ("n_img", i32), ("grid_h", i32), ("grid_w", i32), ("n_taps", i32), ("M", i32), ("N", i32),
                ("per_image", i32), ("k_splits", i32), ("scale", f32), ("out", vp),
                ("a_dx", _i8x96), ("a_dy", _i8x96), ("a_pl", _i8x96), ("b_dx", _i8x96), ("b_dy", _i8x96), ("b_pl", _i8x96)]


# name -> (restype, argtypes); every symbol include/vst_b200.h declares
PROTOTYPES = {
    "vst_abi_version": (i32, []),
    "vst_last_error": (C.c_char_p, []),
    "vst_device_arch": (i32, []),
    "vst_conv2d_f32": (i32, [vp, vp, vp, vp] + [i32] * 11 + [vp]),
    "vst_conv_transpose2d_f32": (i32, [vp, vp, vp, vp] + [i32] * 5 + [vp]),
    "vst_instance_norm_f32": (i32, [vp] * 7 + [i32, i32, i32, f32, i32, vp]),
    "vst_maxpool2_f32": (i32, [vp, vp, i32, i32, i32, vp]),
    "vst_vgg_normalize_f32": (i32, [vp, vp, i32, i32, i32, vp]),
    "vst_pack_bgr_u8": (i32, [vp, vp, i32, i32, i32, vp]),
    "vst_warp_f32": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, vp]),
    "vst_flow_warp_mask_f32": (i32, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "vst_resize_bilinear_f32": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "vst_motion_mask_f32": (i32, [vp, vp, C.c_size_t, vp]),
    "vst_gram_f32": (i32, [vp, vp, i32, i32, i32, f32, vp]),
    "vst_reduce_scratch_floats": (sz, []),
    "vst_feature_temporal_f32": (i32, [vp] * 6 + [i32] * 6 + [vp]),
    "vst_output_temporal_f32": (i32, [vp] * 8 + [i32] * 4 + [vp]),
    "vst_sqdiff_sum_f32": (i32, [vp, vp, vp, vp, sz, vp]),
    "vst_frame_diff_sqsum_f32": (i32, [vp, vp, vp, vp, f32, f32, vp, vp, sz, vp]),
    "vst_tv_f32": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "vst_weight_flip_transpose_f32": (i32, [vp, vp, i32, i32, i32, vp]),
    "vst_conv_transpose_gather_f32": (i32, [vp, vp, vp, vp] + [i32] * 10 + [vp]),
    "vst_fold_pad_f32": (i32, [vp, vp] + [i32] * 8 + [vp]),
    "vst_conv2d_wgrad_f32": (i32, [vp, vp, vp] + [i32] * 10 + [vp]),
    "vst_channel_sum_f32": (i32, [vp, vp, i32, i32, i32, vp]),
    "vst_act_bwd_f32": (i32, [vp, vp, vp, sz, i32, vp]),
    "vst_instance_norm_bwd_f32": (i32, [vp] * 9 + [i32] * 4 + [vp]),
    "vst_maxpool2_bwd_f32": (i32, [vp, vp, vp, i32, i32, i32, vp]),
    "vst_vgg_normalize_bwd_f32": (i32, [vp, vp, i32, i32, vp]),
    "vst_warp_bwd_f32": (i32, [vp, vp, vp, i32, i32, i32, i32, vp]),
    "vst_feature_temporal_bwd_f32": (i32, [vp] * 4 + [f32, vp, vp, vp] + [i32] * 6 + [vp]),
    "vst_output_temporal_bwd_f32": (i32, [vp] * 6 + [f32, vp, vp, vp] + [i32] * 4 + [vp]),
    "vst_sqdiff_bwd_f32": (i32, [vp, vp, f32, vp, vp, sz, vp]),
    "vst_tv_bwd_f32": (i32, [vp, f32, vp, i32, i32, i32, i32, vp]),
    "vst_gram_bwd_f32": (i32, [vp, vp, vp, i32, i32, i32, f32, vp]),
    "vst_axpy_f32": (i32, [vp, vp, f32, sz, vp]),
    "vst_loss_terms_f32": (i32, [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(f32), C.POINTER(f32), C.POINTER(i32),
                                 i32, i32, vp, vp, vp]),
    "vst_adam_f32": (i32, [vp, vp, vp, vp, sz, f32, f32, f32, f32, i32, f32, vp, vp]),
    "vst_tc_tapgemm": (i32, [C.POINTER(TapGemmDesc), vp]),
    "vst_tc_tapgemm_plan": (i32, [C.POINTER(TapGemmDesc), C.POINTER(TapGemmPlanInfo)]),
    "vst_tc_pcgemm": (i32, [C.POINTER(PcGemmDesc), vp]),
    "vst_gather_sum_f32": (i32, [vp, vp, i32, vp, sz, i32, vp]),
    "vst_tc_nchw_to_act": (i32, [vp, i32, vp, ActDesc, i32, vp]),
    "vst_tc_act_to_nchw": (i32, [vp, ActDesc, i32, vp, vp]),
    "vst_tc_act_to_nchw_first": (i32, [vp, ActDesc, i32, i32, vp, vp]),
    "vst_tc_prologue_x9": (i32, [vp, vp, i32, i32, i32, i32, i32, vp]),
    "vst_tc_prologue_x27": (i32, [vp, vp, i32, i32, i32, vp]),
    "vst_tc_rowconv_expand": (i32, [vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "vst_tc_in_apply": (i32, [vp, vp, vp, vp, vp, ActDesc, vp, ActDesc, i32, f32, i32, vp]),
    "vst_tc_in_bwd_reduce": (i32, [vp, ActDesc, vp, vp, vp, vp, vp, vp, i32, f32, i32, vp]),
    "vst_tc_in_bwd_apply": (i32, [vp, ActDesc, vp, vp, vp, vp, vp, vp, vp, ActDesc, vp, i32, f32, i32, vp]),
    "vst_tc_in_param_grads": (i32, [vp, vp, vp, i32, i32, vp]),
    "vst_tc_maxpool2": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "vst_tc_relu_pool_bwd": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, vp]),
    "vst_tc_sqdiff_sum_bf16": (i32, [vp, vp, vp, vp, sz, vp]),
    "vst_tc_sqdiff_bwd_bf16": (i32, [vp, vp, f32, vp, sz, vp]),
    "vst_tc_gram_grad_weights": (i32, [vp, vp, i32, f32, vp, i32, i32, vp]),
    "vst_tc_add_bf16": (i32, [vp, vp, sz, vp]),
    "vst_plan_arena_bytes": (sz, [C.POINTER(NetDesc)]),
    "vst_plan_create": (i32, [C.POINTER(NetDesc), C.POINTER(vp), i32, vp, sz, vp, C.POINTER(vp)]),
    "vst_plan_destroy": (None, [vp]),
    "vst_plan_forward": (i32, [vp, vp, vp, vp, vp, vp]),
    "vst_plan_forward_bgr8": (i32, [vp, vp, vp, vp, vp, vp]),
    "vst_plan_forward_pair": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "vst_plan_launches": (i32, [vp]),
    "vst_plan_set_timing": (i32, [vp, i32]),
    "vst_plan_get_timing": (i32, [vp, C.POINTER(C.c_float), C.POINTER(i32)]),
    "vst_plan_set_stop_after": (i32, [vp, i32]),
    "vst_plan_debug_activation": (i32, [vp, i32, vp, sz, vp]),
    "vst_tc_conv_workspace_bytes": (sz, [i32] * 6),
    "vst_tc_conv3x3_f32io": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]),
}


class VstError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library (once) and attach prototypes.  Raises if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise VstError(f"{LIB} not found - run `python -m vst_b200.build` (or __graft_entry__.build()); "
                           "there is no fallback path")
        L = C.CDLL(LIB)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        if L.vst_abi_version() != 1:
            raise VstError("libvst_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = lib().vst_last_error().decode(errors="replace")
        raise VstError(f"{what} failed with code {code}: {msg}")
