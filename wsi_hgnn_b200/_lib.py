"""ctypes binding of the C-ABI library (include/wsi_hgnn.h).  There is NO fallback: if the
library is missing or fails to load, importing an op raises."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libwsi_hgnn.so")
ABI_VERSION = 11

_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float

# name -> (restype, argtypes); must list every symbol include/wsi_hgnn.h declares (checked by tests)
PROTOTYPES = {
    "wsi_abi_version": (_I, []),
    "wsi_last_error": (c_char_p, []),
    "wsi_num_sms": (_I, []),
    "wsi_set_device": (_I, [_I]),
    "wsi_launch_count": (_L, []),
    "wsi_typed_linear_workspace_bytes": (_L, [_L, _I, _I, _I, _I]),
    "wsi_typed_linear_f32": (_I, [_P, _L, _P, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P, _L, _P, _P, _P, _L, _I, _P, _L, _P]),
    "wsi_hetero_attn_fwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _L, _P, _P]),
    "wsi_hetero_attn_work_fwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P, _L, _P, _P, _P, _P, _P,
                                      _P, _L, _L, _P, _P, _P, _L, _P, _P]),
    "wsi_typed_linear_tc_ok": (_I, [_L, _I, _I]),
    "wsi_split_bf16": (_I, [_P, _L, _L, _I, _P, _P]),
    "wsi_typed_linear_split": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P, _L, _P, _P, _P, _L, _P, _P]),
    "wsi_hetero_attn_bwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P, _L, _P, _L, _P, _L, _P, _L,
                                 _P, _P]),
    "wsi_hetero_attn_seg_fwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _L, _I, _I, _I, _P, _L, _P]),
    "wsi_head_perm": (_I, [_I, _I, _P]),
    "wsi_rel_transform": (_I, [_P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _P, _L, _P]),
    "wsi_segment_combine": (_I, [_P, _L, _P, _P, _L, _I, _P, _L, _P]),
    "wsi_typed_layernorm": (_I, [_P, _L, _P, _P, _P, _I, _I, _F, _P, _L, _P]),
    "wsi_plan_workspace_bytes": (_L, [_L, _L]),
    "wsi_plan_build_csr": (_I, [_P, _P, _P, _P, _P, _I, _L, _L, _P, _P, _P, _P, _P, _P, _P, _L, _P]),
    "wsi_plan_attn_work_count": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _L, _P]),
    "wsi_plan_attn_work_fill": (_I, [_P, _P, _L, _I, _P, _P, _L, _L, _P, _P, _P, _P, _P, _P, _P]),
    "wsi_segment_pool_workspace_bytes": (_L, [_L, _L, _I]),
    "wsi_segment_pool_fwd": (_I, [_P, _L, _P, _L, _L, _I, _I, _P, _L, _P, _L, _P]),
    "wsi_segment_pool_affine_workspace_bytes": (_L, [_L, _L, _I]),
    "wsi_segment_pool_affine_fwd": (_I, [_P, _L, _P, _I, _I, _L, _I, _I, _P, _P, _P, _P, _I, _I, _P, _L, _P, _L, _P]),
    "wsi_knn_workspace_bytes": (_L, [_L, _I, _I, _L, _L]),
    "wsi_knn_topk": (_I, [_P, _L, _I, _I, _L, _L, _P, _P, _P, _L, _P]),
    "wsi_edge_pearson": (_I, [_P, _L, _I, _P, _P, _L, _P, _P, _P]),
}

_lib = None


class WsiError(RuntimeError):
    pass


def load():
    """Load libwsi_hgnn.so once; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise WsiError(f"{LIB_PATH} not found - build it with `python -m wsi_hgnn_b200.build` "
                       "(there is no CPU / PyTorch fallback for the hot path)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    v = lib.wsi_abi_version()
    if v != ABI_VERSION:
        raise WsiError(f"libwsi_hgnn.so ABI {v} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().wsi_last_error().decode("utf-8", "replace")
        if rc == -3:
            raise NotImplementedError(f"{what}: {msg}")
        raise WsiError(f"{what}: {msg} (code {rc})")
