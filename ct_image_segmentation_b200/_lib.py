"""ctypes binding of libb200seg.so (include/b200seg.h).

The product path has NO fallback: if the library is missing or a call fails a
``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libb200seg.so")

F32, BF16 = 0, 1
LABEL_U8, LABEL_I64 = 0, 1
W_CONV_FPROP, W_CONV_DGRAD, W_CONVTR_FPROP, W_CONVTR_DGRAD = 0, 1, 2, 3
CONV_ACCUMULATE, CONV_FORCE_GENERIC, CONV_PADDED_CHANNELS, CONV_NO_SLIDE, CONV_SPLIT_K, CONV_NO_SPLIT_K = 1, 2, 4, 8, 16, 32
PACK_TC_ONLY = 0x100


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n", "cin", "cout", "in_d", "in_h", "in_w", "out_d", "out_h", "out_w",
        "kd", "kh", "kw", "sd", "sh", "sw", "pd", "ph", "pw",
        "x_ld", "y_ld", "r_ld", "dtype", "flags")]


class PackEntry(C.Structure):
    _fields_ = [("w", C.c_uint64), ("packed", C.c_uint64), ("tc_offset", C.c_uint64),
                ("taps", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32), ("kind", C.c_int32)]


class NormDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("c", C.c_int32), ("spatial", C.c_int64),
                ("x_ld", C.c_int32), ("y_ld", C.c_int32), ("r_ld", C.c_int32),
                ("dtype", C.c_int32), ("eps", C.c_float)]


class DiceDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("c", C.c_int32), ("spatial", C.c_int64),
                ("ld", C.c_int32), ("dtype", C.c_int32), ("label_dtype", C.c_int32),
                ("include_background", C.c_int32)]


_P = C.c_void_p
_CD, _ND, _DD = C.POINTER(ConvDesc), C.POINTER(NormDesc), C.POINTER(DiceDesc)

# name -> (restype, argtypes); every symbol include/b200seg.h declares
SIGNATURES = {
    "b200seg_version": (C.c_int, []),
    "b200seg_last_error": (C.c_char_p, []),
    "b200seg_launch_count": (C.c_longlong, []),
    "b200seg_tc_launch_count": (C.c_longlong, []),
    "b200seg_last_launch": (C.c_char_p, []),
    "b200seg_check_device": (C.c_int, [C.c_int]),
    "b200seg_packed_weight_bytes": (C.c_size_t, [_CD, C.c_int]),
    "b200seg_pack_weight": (C.c_int, [_CD, C.c_int, _P, _P, _P]),
    "b200seg_packed_weight_tc_offset": (C.c_size_t, [_CD, C.c_int]),
    "b200seg_softmax_loss_workspace_bytes": (C.c_size_t, [_P]),
    "b200seg_softmax_loss_fwd": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, C.c_size_t, _P]),
    "b200seg_softmax_loss_bwd": (C.c_int, [_P, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P]),
    "b200seg_softmax_boundary_loss_workspace_bytes": (C.c_size_t, [_P]),
    "b200seg_softmax_boundary_loss_fwd": (C.c_int, [_P, _P, _P, _P, C.c_float, _P, _P, C.c_size_t, _P]),
    "b200seg_softmax_boundary_loss_bwd": (C.c_int, [_P, _P, _P, _P, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "b200seg_adam_step": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int64, _P]),
    "b200seg_dice_loss_epilogue": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_int32, _P, _P, _P, _P]),
    "b200seg_conv_fprop_partials": (C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, C.c_size_t, _P, _P, _P]),
    "b200seg_instnorm_prelu_fwd_partials": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "b200seg_im2col": (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    "b200seg_pack_weights_batched": (C.c_int, [_P, C.c_int32, _P]),
    "b200seg_conv_fprop": (C.c_int, [_CD, _P, _P, _P, _P, _P, _P]),
    "b200seg_conv_fprop_stats_workspace_bytes": (C.c_size_t, [_CD]),
    "b200seg_convtr_fprop_stats_workspace_bytes": (C.c_size_t, [_CD]),
    "b200seg_conv_fprop_stats": (C.c_int, [_CD, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_float, _P, C.c_size_t, _P]),
    "b200seg_convtr_fprop_stats": (C.c_int, [_CD, _P, _P, _P, _P, _P, _P, C.c_int32, C.c_float, _P, C.c_size_t, _P]),
    "b200seg_conv_dgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P]),
    "b200seg_conv_dgrad_instnorm_partials_bytes": (C.c_size_t, [_CD]),
    "b200seg_conv_dgrad_instnorm_partials": (C.c_int, [_CD, _P, _P, _P, _P, _P, C.c_int32, _P, _P, C.c_int32, _P, _P,
                                                       C.c_size_t, _P, _P]),
    "b200seg_instnorm_prelu_bwd_from_partials": (C.c_int, [_ND, _P, _P, _P, _P, _P, _P, C.c_int64, _P, _P, _P,
                                                           C.c_size_t, _P]),
    "b200seg_conv_wgrad_workspace_bytes": (C.c_size_t, [_CD]),
    "b200seg_conv_wgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200seg_convtr_fprop": (C.c_int, [_CD, _P, _P, _P, _P, _P, _P]),
    "b200seg_convtr_dgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P]),
    "b200seg_convtr_wgrad_workspace_bytes": (C.c_size_t, [_CD]),
    "b200seg_convtr_wgrad": (C.c_int, [_CD, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200seg_instnorm_workspace_bytes": (C.c_size_t, [_ND]),
    "b200seg_instnorm_stats": (C.c_int, [_ND, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200seg_instnorm_prelu_fwd": (C.c_int, [_ND, _P, _P, _P, _P, _P, _P, _P]),
    "b200seg_instnorm_prelu_bwd": (C.c_int, [_ND, _P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200seg_softmax_dice_workspace_bytes": (C.c_size_t, [_DD]),
    "b200seg_softmax_dice_fwd": (C.c_int, [_DD, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200seg_softmax_dice_metric_workspace_bytes": (C.c_size_t, [_DD]),
    "b200seg_softmax_dice_metric_fwd": (C.c_int, [_DD, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "b200seg_softmax_dice_bwd": (C.c_int, [_DD, _P, _P, _P, _P, _P, _P]),
    "b200seg_argmax_dice_counts": (C.c_int, [_DD, _P, _P, _P, _P, _P]),
    "b200seg_label_dice_counts": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, _P, _P, C.c_int32, _P, _P]),
    "b200seg_squash_masks": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, _P, _P, _P]),
    "b200seg_hu_window_norm": (C.c_int, [C.c_int64, C.c_int32, _P, _P, _P, _P, _P, _P, C.c_int32,
                                         C.c_int32, _P]),
    "b200seg_window_accumulate": (C.c_int, [C.c_int32, _P, C.c_int32, _P, _P] + [C.c_int32] * 10 + [_P]),
    "b200seg_window_accumulate_weighted": (C.c_int, [C.c_int32, _P, C.c_int32, _P, _P, _P] + [C.c_int32] * 10 + [_P]),
    "b200seg_accum_argmax": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "b200seg_crop_window_norm": (C.c_int, [C.c_int32, _P, _P, _P, C.c_int32, _P, _P] + [C.c_int32] * 6 +
                                 [C.c_float] * 4 + [C.c_int32, _P]),
}

_lib = None


def load():
    """Load libb200seg.so and bind every declared symbol.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -m ct_image_segmentation_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().b200seg_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")
