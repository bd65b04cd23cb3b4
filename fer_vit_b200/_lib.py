"""ctypes binding of libfervit_b200.so (the C ABI declared in include/fervit_b200.h).

The library is built in-tree by ``python -m fer_vit_b200.build`` (or ``__graft_entry__.build()``); loading fails
loudly when it is missing — there is no CPU or PyTorch fallback for the compute path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfervit_b200.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
SITE_INPUT, SITE_HEAD = 0xFFFF0, 0xFFFF1

# slot ids (mirror of the enums in fervit_b200.h)
(G_IN_W, G_IN_B, G_CLS, G_POS, G_HEAD_LN_W, G_HEAD_LN_B, G_HEAD_W, G_HEAD_B, G_SPE_GROUP, G_SPE_LAYER,
 G_LWN_GAMMA, G_LWN_BETA, G_LWN_GATE, G_LEAM_W, G_SPE_GROUPS) = range(15)
NUM_GLOBAL = 16
(B_LN1_W, B_LN1_B, B_QKV_W, B_QKV_B, B_PROJ_W, B_PROJ_B, B_LN2_W, B_LN2_B, B_FC1_W, B_FC1_B, B_FC2_W, B_FC2_B,
 B_AD1_W, B_AD1_B, B_AD2_W, B_AD2_B, B_ALPHA) = range(17)
NUM_BLOCK = 17


def bslot(block: int, s: int) -> int:
    return NUM_GLOBAL + block * NUM_BLOCK + s


class Config(C.Structure):
    _fields_ = [
        ("mode", C.c_int), ("input_kind", C.c_int), ("L", C.c_int), ("Din", C.c_int), ("E", C.c_int),
        ("depth", C.c_int), ("H", C.c_int), ("F", C.c_int), ("C", C.c_int), ("norm_first", C.c_int),
        ("act", C.c_int), ("eps_block", C.c_float), ("eps_head", C.c_float), ("adapter_dim", C.c_int),
        ("dropout", C.c_float), ("head_dropout", C.c_float), ("input_dropout", C.c_int),
        ("img_c", C.c_int), ("img_h", C.c_int), ("img_w", C.c_int), ("patch", C.c_int),
        ("use_spe", C.c_int), ("use_lwn", C.c_int), ("use_lwn_res", C.c_int), ("use_leam", C.c_int),
        ("eps_lwn", C.c_float),
    ]


class PreModules(C.Structure):
    _fields_ = [
        ("use_spe", C.c_int), ("use_lwn", C.c_int), ("use_lwn_res", C.c_int), ("use_leam", C.c_int),
        ("group_embed", C.c_void_p), ("layer_embed", C.c_void_p), ("groups", C.c_void_p),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("gate", C.c_void_p), ("leam_w", C.c_void_p),
        ("eps", C.c_float),
    ]


class LatentAugmentParams(C.Structure):
    _fields_ = [("noise_std", C.c_float), ("use_scale", C.c_int), ("scale_min", C.c_float),
                ("scale_max", C.c_float), ("mask_prob", C.c_float)]


_p, _i, _ll, _f, _u64, _u32 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_ulonglong, C.c_uint

# name -> (restype, argtypes); every symbol include/fervit_b200.h declares
SIGNATURES = {
    "fervit_abi_version": (_i, []),
    "fervit_last_error": (C.c_char_p, []),
    "fervit_launch_count": (_u64, []),
    "fervit_profile_enable": (_i, [_i]),
    "fervit_profile_read": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_ll)]),
    "fervit_plan_create": (_i, [C.POINTER(Config), C.POINTER(_p)]),
    "fervit_plan_destroy": (None, [_p]),
    "fervit_plan_num_slots": (_i, [_p]),
    "fervit_plan_slot_numel": (_ll, [_p, _i]),
    "fervit_plan_set_params": (_i, [_p, C.POINTER(_p), _i]),
    "fervit_plan_set_ln_fold": (_i, [_p, C.POINTER(_i), C.POINTER(_i), _i]),
    "fervit_plan_wcache_bytes": (_ll, [_p]),
    "fervit_plan_set_wcache": (_i, [_p, _p, _ll]),
    "fervit_plan_refresh_wcache": (_i, [_p, C.POINTER(_i), _i, _p]),
    "fervit_plan_workspace_bytes": (_ll, [_p, _i, _i]),
    "fervit_plan_forward": (_i, [_p, _p, _i, _p, _ll, _i, _i, _u64, _p, _p, _p]),
    "fervit_plan_num_stages": (_i, [_p]),
    "fervit_plan_backward": (_i, [_p, _p, _i, _p, _ll, _i, _u64, _p, _p, C.POINTER(_p), _i, _i, _i, _p]),
    "fervit_plan_saved_buffer": (_i, [_p, _p, _i, _i, _i, C.POINTER(_p), C.POINTER(_ll)]),
    "fervit_cross_entropy": (_i, [_p, _p, _p, _f, _i, _i, _p, _f, _p, _p, _p, _p]),
    "fervit_cross_entropy_mixup": (_i, [_p, _p, _p, _p, _f, _i, _i, _f, _p, _f, _p, _p, _p]),
    "fervit_latent_batch": (_i, [_p, _p, _ll, _p, _i, _ll, C.POINTER(LatentAugmentParams), _u64, _p, _p, C.c_double, _p, _p,
                                _p, _p, _p]),
    "fervit_latent_decompose": (_i, [_p, _p, _i, _i, _ll, _i, _i, _f, _p, _p, _p]),
    "fervit_premodules_forward": (_i, [C.POINTER(PreModules), _p, _i, _i, _i, _p, _p]),
    "fervit_premodules_scratch_floats": (_ll, [_i, _i, _i]),
    "fervit_premodules_backward": (_i, [C.POINTER(PreModules), _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fervit_linear_forward": (_i, [_i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _i, _p]),
    "fervit_gemm_scratch_bytes": (_ll, []),
    "fervit_set_gemm_scratch": (_i, [_p, _ll]),
    "fervit_debug_gemm_clock": (_i, [_p, _p]),
    "fervit_debug_gemm_timeline": (_i, [_p, _i]),
    "fervit_debug_adapter_timeline": (_i, [_p, _i]),
    "fervit_gemm_prof": (_i, [_i, _p]),
    "fervit_gemm_prof_read": (_i, [_p, _p, _p, _p, _i]),
    "fervit_adapter_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "fervit_adapter_backward_input": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "fervit_adamw_scratch_floats": (_ll, [_i, _p]),
    "fervit_adamw_step": (_i, [_i, _p, _p, _p, _p, _p, _p, _p, _p, _f, _p, _p]),
    "fervit_linear_dgrad": (_i, [_i, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _i, _p]),
    "fervit_linear_wgrad_scratch_floats": (_ll, [_i, _i, _i]),
    "fervit_linear_wgrad": (_i, [_i, _p, _p, _i, _i, _i, _f, _p, _p, _p]),
    "fervit_linear_wgrad_bias_scratch_floats": (_ll, [_i, _i, _i]),
    "fervit_linear_wgrad_bias": (_i, [_i, _p, _p, _i, _i, _i, _f, _p, _p, _p, _p]),
    "fervit_layernorm_forward": (_i, [_i, _p, _p, _p, _f, _i, _i, _p, _p, _p, _p, _p]),
    "fervit_layernorm_scratch_floats": (_ll, [_i, _i]),
    "fervit_layernorm_backward": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "fervit_attention_forward": (_i, [_i, _p, _i, _i, _i, _i, _f, _u64, _u32, _p, _p, _p]),
    "fervit_attention_backward": (_i, [_i, _p, _p, _p, _p, _i, _i, _i, _i, _f, _u64, _u32, _p, _p]),
    "fervit_dropout_mask": (_i, [_p, _ll, _f, _u64, _u32, _p]),
    "fervit_cast_bf16": (_i, [_p, _p, _ll, _p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the shared library once; raise if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -m fer_vit_b200.build). "
                "fer_vit_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if handle.fervit_abi_version() != 1:
            raise RuntimeError("libfervit_b200.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(status: int) -> None:
    """Convert a non-zero C status into RuntimeError carrying the library's message (SURVEY.md §8b errors)."""
    if status != 0:
        msg = lib().fervit_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"fervit_b200: {msg}")


def launch_count() -> int:
    return int(lib().fervit_launch_count())
