"""ctypes binding of libvscb200.so (the C ABI declared in include/vscb200.h).

The library is built in-tree by ``vsc22_submission_b200.build`` and MUST be present: there is no
CPU or PyTorch fallback anywhere in this package -- a missing or unloadable library raises.
Loading performs no CUDA call (DataLoader workers may fork before CUDA init,
reference vsc/baseline/inference.py:1-17).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSCB200_LIB") or os.path.join(_HERE, "libvscb200.so")   # VSCB200_LIB: A/B builds (tools/variants.py)

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
ACT_NONE, ACT_QUICK_GELU, ACT_GELU = -1, 0, 1
TAIL_TOKENS, TAIL_GEM_LINEAR, TAIL_GEM_CONV_LINEAR = 0, 1, 2
EPI_BF16, EPI_F32, EPI_RESIDUAL_F32 = 0, 1, 2


class VitSpecC(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("img", "patch", "width", "layers", "heads", "patch_bias", "pre_norm", "act",
                                       "tail", "out_dim", "gem_hidden")] + [("ln_eps", C.c_float), ("gem_p", C.c_float),
                                                                           ("precision", C.c_int)]


class SwinSpecC(C.Structure):
    _fields_ = [("img", C.c_int), ("patch", C.c_int), ("embed", C.c_int), ("n_stages", C.c_int), ("depths", C.c_int * 4),
                ("heads", C.c_int * 4), ("window", C.c_int), ("pretrained_windows", C.c_int * 4), ("out_dim", C.c_int),
                ("ln_eps", C.c_float), ("gem_p", C.c_float), ("precision", C.c_int)]


PRECISION = {"bf16": 0, "fp32": 1}


class Vscb200Error(RuntimeError):
    pass


_p, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> (restype, argtypes); every symbol include/vscb200.h declares
SIGNATURES = {
    "vscb200_last_error": (C.c_char_p, []),
    "vscb200_version": (_i, []),
    "vscb200_launch_count": (_i64, []),
    "vscb200_device_count": (_i, []),
    "vscb200_set_device": (_i, [_i]),
    "vscb200_trim": (_i, []),
    "vscb200_prof_enable": (_i, [_i]),
    "vscb200_prof_collect": (_i, [_p, _p, _p, _i]),
    "vscb200_index_create": (_i, [_i, _i, C.POINTER(_p)]),
    "vscb200_index_destroy": (None, [_p]),
    "vscb200_index_add": (_i, [_p, _p, _i64, _p]),
    "vscb200_index_add_sn": (_i, [_p, _p, _i64, _i, _i, _p, _i, _f, _p, _p]),
    "vscb200_index_add_host": (_i, [_p, _p, _i64]),
    "vscb200_index_reset": (_i, [_p]),
    "vscb200_index_ntotal": (_i64, [_p]),
    "vscb200_index_dim": (_i, [_p]),
    "vscb200_index_metric": (_i, [_p]),
    "vscb200_index_set_id_offset": (_i, [_p, _i64]),
    "vscb200_index_search": (_i, [_p, _p, _i64, _i, _p, _p, _p]),
    "vscb200_index_search_host": (_i, [_p, _p, _i64, _i, _p, _p]),
    "vscb200_index_last_fallbacks": (_i64, [_p]),
    "vscb200_index_reconstruct_n": (_i, [_p, _i64, _i64, _p]),
    "vscb200_index_range_search_host": (_i, [_p, _p, _i64, _f, _p, C.POINTER(_p), C.POINTER(_p)]),
    "vscb200_free": (None, [_p]),
    "vscb200_index_scores": (_i, [_p, _p, _i64, _p, _i64, _p]),
    "vscb200_index_global_search": (_i, [_p, _p, _i64, _i64, _i, _f, C.POINTER(_i64), _p]),
    "vscb200_index_global_results": (_i, [_p, _p, _p, _p, _p]),
    "vscb200_index_global_video_pairs": (_i, [_p, _p, _i64, _p, _i64, C.POINTER(_i64), _p]),
    "vscb200_index_video_pair_results": (_i, [_p, _p, _p, _p, _p]),
    "vscb200_sn_transform": (_i, [_p, _i64, _i, _i, _i, _f, _p, _p, _p]),
    "vscb200_sn_transform_dev": (_i, [_p, _i64, _i, _p, _i, _f, _p, _p, _p]),
    "vscb200_low_var_dim_dev": (_i, [_p, _i64, _i, _p, _p]),
    "vscb200_low_var_dim": (_i, [_p, _i64, _i, C.POINTER(_i), _p]),
    "vscb200_sn_bias": (_i, [_p, _i64, _i, _i, _f, _p, _p]),
    "vscb200_sn2_adapt": (_i, [_p, _p, _p, _i64, _i, _i, _f, _i, _p, _p]),
    "vscb200_col_sums": (_i, [_p, _i64, _i, _p, C.c_double, _p, _p]),
    "vscb200_var_argmin_dev": (_i, [_p, _i, _p, _p]),
    "vscb200_col_moments_local": (_i, [_p, _i64, _i, _p, _p]),
    "vscb200_var_argmin_moments": (_i, [_p, C.c_double, _i, _p, _p]),
    "vscb200_topk_pack": (_i, [_p, _p, _i64, _i, _i, _p, _p]),
    "vscb200_topk_pack_cols": (_i, [_p, _p, _i64, _i, _i, _p, _i, _i, _p]),
    "vscb200_topk_merge": (_i, [_p, _i, _i64, _i, _i, _i, _p, _p, _p]),
    "vscb200_topk_merge_cols": (_i, [_p, _i, _i64, _i, _i, _i, _i, _i, _p, _p, _p, _p]),
    "vscb200_vit_create": (_i, [C.POINTER(VitSpecC), _i, C.POINTER(_p)]),
    "vscb200_vit_destroy": (None, [_p]),
    "vscb200_vit_set_param": (_i, [_p, C.c_char_p, _p, _i64, _p]),
    "vscb200_vit_forward": (_i, [_p, _p, _i64, _p, _p]),
    "vscb200_vit_forward_host": (_i, [_p, _p, _i64, _p]),
    "vscb200_vit_out_elems_per_frame": (_i64, [_p]),
    "vscb200_swin_create": (_i, [C.POINTER(SwinSpecC), _i, C.POINTER(_p)]),
    "vscb200_swin_destroy": (None, [_p]),
    "vscb200_swin_set_param": (_i, [_p, C.c_char_p, _p, _i64, _p]),
    "vscb200_swin_forward": (_i, [_p, _p, _i64, _p, _p]),
    "vscb200_swin_forward_host": (_i, [_p, _p, _i64, _p]),
    "vscb200_swin_out_dim": (_i, [_p]),
    "vscb200_ensemble_pca": (_i, [_p, _p, _i, _i64, _p, _p, _i, _p, _p]),
    "vscb200_near_dup_keep": (_i, [_p, _i64, _i, C.c_double, _p, _p]),
    "vscb200_pair_sims": (_i, [_p, _p, _i, _i64, _p, _p, _p, _p, _p, _f, _p, _p]),
    "vscb200_pair_topk": (_i, [_p, _i64, _p, _p, _p, _p, _i, _p, _p, _p]),
    "vscb200_pair_segment_images": (_i, [_p, _i64, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    "vscb200_resize_normalize": (_i, [_p, _i64, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float), _p, _p, _p]),
    "vscb200_jpeg_decode": (_i, [_p, _p, _i64, _p, _p, _p, _p]),
    "vscb200_tn_align": (_i, [_p, _p, _i, _i64, _p, _p, _p, _i, _i, _i, C.c_double, C.c_double, C.c_double, _p, _p, _p]),
    "vscb200_tn_box_scores": (_i, [_p, _p, _p, _i64, _p, _p, _i, _f, _p, _p]),
    "vscb200_gemm_bf16": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i64, _i64, _i64, _i, _i, _p]),
    "vscb200_gemm_split": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _i64, _i64, _i64, _i, _i, _p]),
    "vscb200_split_f32_bf16": (_i, [_p, _p, _p, _i64, _p]),
    "vscb200_attention_fp32": (_i, [_p, _p, _p, _p, _i64, _i, _i, _i, _p]),
    "vscb200_layernorm": (_i, [_p, _p, _p, _p, _i64, _i, _f, _i, _p]),
    "vscb200_attention": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vscb200_cast_f32_bf16": (_i, [_p, _p, _i64, _p]),
}

_lib = None


def lib() -> C.CDLL:
    """The loaded library; raises Vscb200Error when it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Vscb200Error(
                f"{LIB_PATH} not found: build it with `python -m vsc22_submission_b200.build` "
                "(or __graft_entry__.build()). This package has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here means header/library drift
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().vscb200_last_error().decode(errors="replace")
        raise Vscb200Error(f"{what or 'vscb200'} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(lib().vscb200_launch_count())


PROF_KINDS = ("gemm", "attention", "layernorm", "vit_other", "scores", "select")


def prof_enable(on: bool):
    lib().vscb200_prof_enable(1 if on else 0)


def prof_collect():
    """-> {kind: {"ms": device ms, "launches": n, "work": algorithmic flops or bytes}} since the last call."""
    n = len(PROF_KINDS)
    ms, ln, wk = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)()
    check(lib().vscb200_prof_collect(ms, ln, wk, n), "prof_collect")
    return {k: {"ms": ms[i], "launches": int(ln[i]), "work": wk[i]} for i, k in enumerate(PROF_KINDS)}
