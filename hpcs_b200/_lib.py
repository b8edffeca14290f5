"""ctypes binding of ``libhpcs_b200.so`` (the C ABI declared in ``include/hpcs_b200.h``).

There is deliberately no CPU or eager-PyTorch fallback: if the library cannot be loaded, or a
tensor is not on a CUDA device of compute capability 10.x, the call raises.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Optional

import torch

from . import build as _build

_c = ctypes
_LOCK = threading.Lock()
_LIB: Optional[ctypes.CDLL] = None
_DEVICE_OK = set()

ABI_VERSION = 4

# name -> (restype, argtypes); mirrors include/hpcs_b200.h one to one
_P, _I, _L, _F, _Z = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_size_t
SIGNATURES = {
    "hpcs_abi_version": (_I, []),
    "hpcs_last_error": (_c.c_char_p, []),
    "hpcs_device_check": (_I, []),
    "hpcs_launch_count": (_c.c_uint64, []),
    "hpcs_knn_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "hpcs_knn_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "hpcs_knn_ffma_f32": (_I, [_P, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "hpcs_knn_fallback_rows": (_I, [_P, _Z, _I, _I, _I, _I, _P, _c.POINTER(_c.c_int)]),
    "hpcs_knn_path_stats": (_I, [_P, _Z, _I, _I, _I, _I, _P, _c.POINTER(_c.c_int)]),
    "hpcs_edge_feat_fwd_f32": (_I, [_P, _P, _I, _I, _I, _I, _I, _P, _P]),
    "hpcs_edge_feat_bwd_workspace_bytes": (_Z, [_I, _I, _I]),
    "hpcs_edge_feat_bwd_is_fast": (_I, [_P, _I, _I, _I]),
    "hpcs_edge_feat_bwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "hpcs_edge_rev_build": (_I, [_P, _I, _I, _I, _P, _Z, _P]),
    "hpcs_edge_feat_bwd_prebuilt_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "hpcs_hyp_triplet_workspace_bytes": (_Z, [_L, _I]),
    "hpcs_hyp_triplet_fwd_f32": (_I, [_P, _L, _I, _P, _P, _P, _L, _P, _F, _I, _F, _I, _P, _P, _P, _Z, _P]),
    "hpcs_hyp_triplet_fwd_i32_f32": (_I, [_P, _L, _I, _P, _P, _P, _L, _P, _F, _I, _F, _I, _P, _P, _P, _Z, _P]),
    "hpcs_triplet_filter_i32_f32": (_I, [_P, _L, _I, _P, _P, _P, _L, _I, _F, _P, _P, _Z, _P]),
    "hpcs_hyp_triplet_bwd_f32": (_I, [_P, _P, _L, _I, _P, _P, _Z, _P, _P, _P]),
    "hpcs_triplet_filter_f32": (_I, [_P, _L, _I, _P, _P, _P, _L, _I, _F, _P, _P, _Z, _P]),
    "hpcs_triplet_sample_i32": (_I, [_P, _L, _P, _I, _L, _c.c_uint64, _P, _P, _P, _P]),
    "hpcs_triplet_sample_state_i32": (_I, [_P, _L, _P, _I, _L, _P, _P, _P, _P, _P]),
    "hpcs_hyp_lca_fwd_f32": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "hpcs_hyp_lca_bwd_f32": (_I, [_P, _P, _P, _L, _I, _I, _P, _P, _P]),
    "hpcs_expmap0_fwd_f32": (_I, [_P, _L, _I, _P, _P]),
    "hpcs_expmap0_bwd_f32": (_I, [_P, _P, _L, _I, _P, _P]),
    "hpcs_leaves_f32": (_I, [_P, _L, _I, _P, _P, _P]),
    "hpcs_linkage_workspace_bytes": (_Z, [_I, _I, _I, _I]),
    "hpcs_linkage_debug_counters_offset": (_Z, [_I, _I, _I]),
    "hpcs_linkage_f64": (_I, [_P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "hpcs_fcluster_maxclust_i32": (_I, [_P, _I, _I, _P, _I, _I, _P, _P]),
    "hpcs_cut_scores_f64": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P]),
    "hpcs_rotate_points_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "hpcs_one_hot_f32": (_I, [_P, _L, _I, _P, _P]),
    "hpcs_cosface_logits_f32": (_I, [_P, _P, _P, _L, _I, _I, _F, _F, _P, _P]),
    "hpcs_edgeconv_bn_fold_f32": (_I, [_P, _L, _P, _P, _P, _P, _P, _F, _F, _I, _P, _I, _P]),
    "hpcs_edgeconv_bn_sums_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P]),
    "hpcs_edgeconv_bn_sums_finish_f32": (_I, [_P, _L, _I, _P, _I, _P, _P, _P]),
    "hpcs_edgeconv_coef_floats": (_I, []),
    "hpcs_edgeconv_scratch_floats": (_Z, [_I, _I, _I]),
    "hpcs_vn_point_linear_f32": (_I, [_P, _P, _I, _I, _I, _P, _P, _P]),
    "hpcs_vn_point_linear_bwd_f32": (_I, [_P, _P, _P, _P, _I, _I, _I, _P, _P, _P]),
    "hpcs_edgeconv_fwd_f32": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P, _I, _P, _P, _P, _P, _P]),
    "hpcs_edgeconv_bwd_stage2_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "hpcs_edgeconv_bwd_stage1_f32": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P]),
    "hpcs_peak_probe": (_I, [_I, _I, _P, _Z, _P, _c.POINTER(_c.c_double), _P]),
}


def lib_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building in-tree with nvcc if the .so is missing) and type the C ABI."""
    global _LIB
    if _LIB is not None:
        return _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = _build.LIB_PATH
        if not os.path.exists(path):
            path = _build.build()
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if lib.hpcs_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libhpcs_b200.so ABI {lib.hpcs_abi_version()} != expected {ABI_VERSION}")
        _LIB = lib
    return _LIB


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().hpcs_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().hpcs_launch_count())


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """All tensors on one CUDA device that the library supports; raise otherwise (no fallback)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("hpcs_b200 runs on CUDA (sm_100a) only; got a CPU tensor and there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is None:
        raise RuntimeError("no tensor given")
    if dev.index not in _DEVICE_OK:
        with torch.cuda.device(dev):
            check(load().hpcs_device_check(), "hpcs_device_check")
        _DEVICE_OK.add(dev.index)
    return dev


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    """Caller-owned scratch from the caching allocator (stream-ordered, 256-byte aligned)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
