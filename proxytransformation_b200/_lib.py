"""ctypes binding of csrc/libptpreshape.so (the C ABI declared in include/pt_preshape.h).

There is no CPU fallback: if the library is missing or a call fails the product path raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libptpreshape.so")

PT_DTYPE_F32, PT_DTYPE_BF16, PT_DTYPE_F16 = 0, 1, 2
PT_POOL_VARIANT_MMA, PT_POOL_VARIANT_UMMA = 0, 1


class PtError(RuntimeError):
    pass


class ProxyBlockParams(Structure):
    _fields_ = [(n, c_void_p) for n in (
        "ln1_w", "ln1_b", "pos_bias", "qkv_w", "qkv_b", "pp_w", "pp_b", "proj_w", "proj_b", "ln2_w", "ln2_b", "fc1_w", "fc1_b",
        "fc2_w", "fc2_b", "lno_w", "lno_b", "qkv_w_split", "proj_w_split", "fc1_w_split", "fc2_w_split", "pp_w_split")]


class GemmTcDesc(Structure):
    _fields_ = [("M", c_int), ("N", c_int), ("K", c_int), ("batch", c_int),
                ("a_split", c_void_p), ("a_rows", c_int), ("a_cols", c_int), ("lda", c_int), ("a_koff_z", c_int),
                ("w_split", c_void_p), ("w_rows", c_int), ("ldw", c_int), ("w_row_z", c_int),
                ("bias", c_void_p), ("bias_off_z", ctypes.c_longlong), ("residual", c_void_p), ("act", c_int),
                ("C", c_void_p), ("ldc", c_int), ("c_off_z", ctypes.c_longlong),
                ("c_split", c_void_p), ("cs_plane", ctypes.c_longlong), ("ldcs", c_int), ("cs_off_z", ctypes.c_longlong),
                ("bn", c_int)]


class ImgPoolParams(Structure):
    _fields_ = [(n, c_void_p) for n in ("w_qc", "q0", "w_kc", "g_k", "w_vc", "h_v", "cproj_w", "cproj_b", "ln_w", "ln_b",
                                        "w_qc_split", "wk_pad_split", "gk_pad_split", "wv_cat_split", "cproj_split")] + \
               [("variant", c_int)]


_P = c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "pt_abi_version": (c_int, []),
    "pt_last_error_string": (c_char_p, []),
    "pt_launch_count": (c_int64, []),
    "pt_profile_enable": (c_int, [c_int]),
    "pt_profile_num_tags": (c_int, []),
    "pt_profile_tag_name": (c_char_p, [c_int]),
    "pt_profile_read": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(c_int64)]),
    "pt_profile_timeline": (c_int, [POINTER(c_int), POINTER(ctypes.c_double), POINTER(ctypes.c_double), c_int, POINTER(c_int)]),
    "pt_minmax_ws_bytes": (c_size_t, [c_int, c_int]),
    "pt_minmax_centres": (c_int, [_P, c_int, c_int, c_int, _P, c_float, _P, _P, _P, _P, c_size_t, _P]),
    "pt_ball_query_firstk": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P]),
    "pt_offset_net_fused": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_float, _P, _P, _P]),
    "pt_cluster_dropout": (c_int, [_P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, _P]),
    "pt_point_encoder_fused": (c_int, [_P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "pt_proxy_block_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "pt_proxy_block_fused": (c_int, [_P, _P, _P, POINTER(ProxyBlockParams), c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "pt_position_bias": (c_int, [_P, _P, _P, c_int, c_int, _P, _P]),
    "pt_heads": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "pt_cluster_conv_bn_stats": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P]),
    "pt_linear_bn_stats": (c_int, [_P, _P, _P, c_int, c_int, c_int, _P, _P]),
    "pt_bn_batch_affine": (c_int, [_P, ctypes.c_longlong, _P, _P, _P, _P, c_float, c_float, c_int, _P, _P, _P]),
    "pt_img_attnpool_ws_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "pt_img_attnpool": (c_int, [_P, c_int, POINTER(ImgPoolParams), c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "pt_img_attnpool_stage": (c_int, [_P, c_int, POINTER(ImgPoolParams), c_int, c_int, c_int, c_int, c_int, _P, _P, c_size_t, c_int, _P]),
    "pt_debug_pool_trace": (c_int, [POINTER(ctypes.c_ulonglong), c_int]),
    "pt_debug_umma_trace": (c_int, [POINTER(ctypes.c_ulonglong), c_int]),
    "pt_debug_pool_events": (c_int, [POINTER(ctypes.c_longlong), c_int]),
    "pt_scatter_ws_bytes": (c_size_t, [c_int, c_int]),
    "pt_affine_scatter_compact": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t, _P]),
    "pt_affine_scatter_compact_stage": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_size_t, c_int, _P]),
    "pt_proxy_attention_tc": (c_int, [_P, ctypes.c_longlong, c_int, _P, ctypes.c_longlong, ctypes.c_longlong, _P, ctypes.c_longlong, _P,
                                      c_int, c_int, c_int, c_int, c_int, _P, _P, ctypes.c_longlong, _P]),
    "pt_aggregate_sample": (c_int, [_P, _P, c_int, _P, _P, ctypes.c_longlong, _P, _P]),
    "pt_sparse_collate": (c_int, [_P, _P, c_int, c_int, ctypes.c_float, c_int, _P, _P, _P, _P]),
    "pt_gemm_ws_bytes": (c_size_t, [c_int, c_int, c_int]),
    "pt_gemm_nt": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P, _P, c_size_t, _P]),
    "pt_gemm_tc": (c_int, [POINTER(GemmTcDesc), _P]),
    "pt_split_bf16": (c_int, [_P, c_int64, _P, _P]),
    "pt_layernorm": (c_int, [_P, _P, _P, _P, c_int, c_int, c_int, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """dlopen the library and attach signatures.  Raises PtError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PtError(f"{LIB_PATH} is missing: build it with `python -m proxytransformation_b200.build_ext` "
                      "(there is no CPU fallback for the preshape path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().pt_last_error_string()
        raise PtError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().pt_launch_count())


def profile_enable(on: bool):
    check(load().pt_profile_enable(1 if on else 0), "pt_profile_enable")


def profile_timeline(max_records: int = 4096):
    """-> [(kernel kind, start ms, end ms)] of every launch recorded since profile_enable(True), on one time axis."""
    L = load()
    tags = (c_int * max_records)()
    t0 = (ctypes.c_double * max_records)()
    t1 = (ctypes.c_double * max_records)()
    n = c_int(0)
    check(L.pt_profile_timeline(tags, t0, t1, max_records, ctypes.byref(n)), "pt_profile_timeline")
    return [(L.pt_profile_tag_name(tags[i]).decode(), t0[i], t1[i]) for i in range(n.value)]


def profile_read() -> dict:
    """-> {kernel kind: (total ms, launches)} for everything launched since profile_enable(True)."""
    L = load()
    out = {}
    for t in range(L.pt_profile_num_tags()):
        ms, n = ctypes.c_double(0.0), c_int64(0)
        check(L.pt_profile_read(t, ctypes.byref(ms), ctypes.byref(n)), "pt_profile_read")
        if n.value:
            out[L.pt_profile_tag_name(t).decode()] = (ms.value, n.value)
    return out
