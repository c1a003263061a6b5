"""ctypes binding of ``libvimoclip_b200.so`` (the C-ABI declared in ``include/vimoclip_b200.h``).

There is no CPU fallback: if the shared library is missing this module raises at first use, and
every op raises on non-CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvimoclip_b200.so")

# enums of include/vimoclip_b200.h
SRC_U8, SRC_U8_WRAP, SRC_F32_WRAP, SRC_F32_NORM = 0, 1, 2, 3
DST_U8, DST_F32_NCHW, DST_BF16_PATCH = 0, 1, 2
ACT_NONE, ACT_QUICKGELU, ACT_GELU_ERF, ACT_RELU = 0, 1, 2, 3
ABI_VERSION = 3
OPT_GEMM_IMPL, OPT_ATTN_IMPL, OPT_PROLOGUE_IMPL, OPT_LN_FUSE, OPT_ATTN_BWD_IMPL, OPT_LAST_BLOCK_CLS = 0, 1, 2, 3, 4, 5
OPT_ATTN_PREFETCH = 6


class GemmEpilogue(C.Structure):
    _fields_ = [
        ("bias", C.c_void_p),
        ("resid", C.c_void_p),
        ("ldr", C.c_longlong),
        ("out", C.c_void_p),
        ("ldo", C.c_longlong),
        ("out_bf16", C.c_int),
        ("act", C.c_int),
        ("alpha", C.c_float),
        ("row_group", C.c_int),
        ("ln_eps", C.c_float),
        ("raw16_out", C.c_void_p),
        ("raw16_ld", C.c_longlong),
        ("stats_out", C.c_void_p),
        ("stats_in", C.c_void_p),
        ("stats_parts", C.c_int),
        ("stats_ld", C.c_longlong),
        ("colsum", C.c_void_p),
        ("resid_bf16", C.c_int),
    ]


class VitLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "ln2_g", "ln2_b", "w_qkv", "b_qkv", "w_out", "b_out", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
        "w_qkv_f", "b_qkv_f", "cs_qkv", "w_fc1_f", "b_fc1_f", "cs_fc1")]


class VitModel(C.Structure):
    _fields_ = [
        ("image", C.c_int), ("patch", C.c_int), ("width", C.c_int), ("layers", C.c_int), ("heads", C.c_int),
        ("out_dim", C.c_int), ("ld_patch", C.c_int),
        ("w_patch", C.c_void_p), ("cls_pos0", C.c_void_p), ("pos", C.c_void_p),
        ("ln_pre_g", C.c_void_p), ("ln_pre_b", C.c_void_p), ("ln_post_g", C.c_void_p), ("ln_post_b", C.c_void_p),
        ("w_proj", C.c_void_p), ("layer", C.POINTER(VitLayer)),
        ("ln_mode", C.c_int), ("last_block_cls", C.c_int), ("attn_impl", C.c_int),
    ]


class TfamLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("b_sin", "b_sout", "b_cin", "b_cout", "b_f1", "b_f2", "ns_g", "ns_b", "nc_g", "nc_b",
                                          "nf_g", "nf_b")] + [("ns_eps", C.c_float), ("nc_eps", C.c_float), ("nf_eps", C.c_float)]


class TfamModel(C.Structure):
    _fields_ = [
        ("d_model", C.c_int), ("nhead", C.c_int), ("dim_ff", C.c_int), ("layers", C.c_int), ("num_classes", C.c_int),
        ("hidden", C.c_int), ("act", C.c_int),
        ("wstream", C.c_void_p), ("layer", C.POINTER(TfamLayer)),
        ("cls_ln_g", C.c_void_p), ("cls_ln_b", C.c_void_p), ("cls_ln_eps", C.c_float),
        ("w1t", C.c_void_p), ("b1", C.c_void_p), ("w2t", C.c_void_p), ("b2", C.c_void_p),
    ]


TFAM_MAX_LAYERS = 8

_SIGNATURES = {
    "vmc_last_error": (C.c_char_p, []),
    "vmc_abi_version": (C.c_int, []),
    "vmc_set_option": (C.c_int, [C.c_int, C.c_longlong]),
    "vmc_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "vmc_launch_count": (C.c_longlong, []),
    "vmc_reset_launch_count": (None, []),
    "vmc_profile_begin": (None, []),
    "vmc_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.c_int]),
    "vmc_prologue": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_frame_diff_prologue": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_resize_geometry": (C.c_int, [C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "vmc_resize_center_crop": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_resize_geometry_ex": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int] + [C.POINTER(C.c_int)] * 4),
    "vmc_resize_center_crop_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_gemm_bf16": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.POINTER(GemmEpilogue), C.c_void_p]),
    "vmc_gemm_bf16_ex": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(GemmEpilogue), C.c_void_p]),
    "vmc_layernorm": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vmc_layernorm_stats": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "vmc_layernorm_ex": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "vmc_gemm_stats_parts": (C.c_int, [C.c_int, C.c_int]),
    "vmc_attention_vit": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_attention_vit_short_mma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_attention_cls": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_attention_vit_impl": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_attention_masked": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_transpose_split": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_transpose_cast": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "vmc_cast_f32": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "vmc_colsum": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vmc_colsum_slices": (C.c_int, [C.c_int, C.c_int]),
    "vmc_layernorm_bwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "vmc_eltwise": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p]),
    "vmc_qgelu_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "vmc_layernorm_bwd_fused_blocks": (C.c_int, [C.c_int]),
    "vmc_layernorm_bwd_fused": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong,
                                          C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vmc_cast_colsum_slices": (C.c_int, [C.c_int, C.c_int]),
    "vmc_cast_colsum": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vmc_attention_vit_bwd_short": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_broadcast_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "vmc_attention_masked_train": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_attention_masked_bwd": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vmc_cast_bf16": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_mean_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_cosine_distill_loss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "vmc_student_heads": (C.c_int, [C.c_void_p] * 5 + [C.c_float] + [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p]),
    "vmc_tfam_head": (C.c_int, [C.c_void_p] * 3 + [C.c_float] + [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_void_p]),
    "vmc_tfam_wstream_bytes": (C.c_longlong, [C.c_int]),
    "vmc_tfam_fused_supported": (C.c_int, [C.POINTER(TfamModel), C.c_int, C.c_int, C.c_int]),
    "vmc_tfam_forward": (C.c_int, [C.POINTER(TfamModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "vmc_vit_workspace_bytes": (C.c_longlong, [C.POINTER(VitModel), C.c_int]),
    "vmc_vit_forward": (C.c_int, [C.POINTER(VitModel), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None
_lock = threading.Lock()


class VmcError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the shared library once (thread-safe: DataParallel calls forward from several threads)."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise VmcError(
                        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(or `make -C vimo-clip_b200/csrc`). There is no CPU fallback."
                    )
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype = res
                    fn.argtypes = args
                if handle.vmc_abi_version() != ABI_VERSION:
                    raise VmcError("libvimoclip_b200.so ABI version mismatch; rebuild the library")
                _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vmc_last_error().decode("utf-8", "replace")
        raise VmcError(f"{what} failed (code {rc}): {msg}")
