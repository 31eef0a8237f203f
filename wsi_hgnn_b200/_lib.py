"""ctypes binding of the C-ABI library (include/wsi_hgnn.h).  There is NO fallback: if the
library is missing or fails to load, importing an op raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libwsi_hgnn.so")
ABI_VERSION = 22

_P, _I, _L, _F = c_void_p, c_int, c_int64, c_float



class HeatGraph(Structure):
    """struct wsi_heat_graph (include/wsi_hgnn.h)."""
    _fields_ = [("n_rows", c_int64), ("T", c_int32), ("B", c_int32), ("type_ptr_host", c_void_p), ("seg_ptr", c_void_p),
                ("e_src", c_void_p), ("e_sim", c_void_p), ("e_rel", c_void_p), ("node_inv_r", c_void_p),
                ("items", c_void_p), ("n_items", c_int64), ("split_row", c_void_p), ("split_ptr", c_void_p),
                ("part_rel", c_void_p), ("part_split", c_void_p), ("split_cnt", c_void_p), ("sched", c_void_p),
                ("n_split", c_int64), ("n_part", c_int64)]


class HeatParams(Structure):
    """struct wsi_heat_params (include/wsi_hgnn.h)."""
    _fields_ = [("F", c_int32), ("D", c_int32), ("H", c_int32), ("L", c_int32), ("opf", c_int32), ("w_in_split", c_void_p),
                ("b_in", c_void_p), ("w_kvq_split", POINTER(c_void_p)), ("b_kvq", POINTER(c_void_p)),
                ("w_a_split", POINTER(c_void_p)), ("b_a", POINTER(c_void_p)), ("skip", POINTER(c_void_p)),
                ("e_w", POINTER(c_void_p)), ("e_b", POINTER(c_void_p)), ("pool_op", c_int32), ("n_out", c_int32),
                ("M", c_void_p), ("c", c_void_p), ("b_total", c_void_p), ("seg_scale", c_void_p)]


class SlideDesc(Structure):
    """struct wsi_slide_desc (include/wsi_hgnn.h)."""
    _fields_ = [("feat", c_void_p), ("ldf", c_int64), ("feat_is_op", c_int32), ("src", c_void_p), ("dst", c_void_p), ("sim", c_void_p),
                ("rel_table", c_void_p), ("seg_ptr", c_void_p), ("node_inv_r", c_void_p), ("type_ptr_host", c_void_p),
                ("n_nodes", c_int64), ("n_edges", c_int64), ("T", c_int32), ("R", c_int32), ("chunk", c_int32)]


class StreamSlide(Structure):
    """struct wsi_stream_slide (include/wsi_hgnn.h)."""
    _fields_ = [("blob_host", c_void_p), ("nbytes", c_int64), ("off_feat", c_int64), ("off_src", c_int64), ("off_dst", c_int64),
                ("off_sim", c_int64), ("n_nodes", c_int64), ("n_edges", c_int64), ("T", c_int32), ("R", c_int32), ("F", c_int32),
                ("feat_is_op", c_int32), ("nodes_per_type_host", c_void_p), ("edges_per_rel_host", c_void_p),
                ("rel_src_type_host", c_void_p), ("rel_dst_type_host", c_void_p)]


# name -> (restype, argtypes); must list every symbol include/wsi_hgnn.h declares (checked by tests)
PROTOTYPES = {
    "wsi_abi_version": (_I, []),
    "wsi_last_error": (c_char_p, []),
    "wsi_num_sms": (_I, []),
    "wsi_set_device": (_I, [_I]),
    "wsi_launch_count": (_L, []),
    "wsi_dev_set": (_I, [c_char_p, _I]),
    "wsi_typed_linear_workspace_bytes": (_L, [_L, _I, _I, _I, _I, _I]),
    "wsi_typed_linear_f32": (_I, [_P, _L, _P, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P, _L, _P, _P, _P, _L, _I, _I, _P, _L, _P]),
    "wsi_hetero_attn_fwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _L, _P, _P]),
    "wsi_hetero_attn_work_fwd": (_I, [_P, _L, _P, _L, _I, _P, _I, _L, _P, _P, _P, _P, _P, _P, _L, _L, _I, _I, _P, _L, _P, _P, _P, _P, _P,
                                      _P, _L, _L, _P, _P, _P, _L, _P, _I, _P]),
    "wsi_typed_linear_tc_ok": (_I, [_L, _I, _I]),
    "wsi_to_operand": (_I, [_P, _L, _L, _I, _I, _P, _P]),
    "wsi_gather_to_operand": (_I, [_P, _L, _P, _L, _I, _I, _P, _P]),
    "wsi_to_operand_colsum_workspace_bytes": (_L, [_I, _P, _I]),
    "wsi_to_operand_colsum": (_I, [_P, _L, _I, _P, _I, _P, _P, _P, _L, _P]),
    "wsi_gather_rows16": (_I, [_P, _L, _P, _L, _I, _P, _P]),
    "wsi_typed_linear_op": (_I, [_P, _P, _P, _I, _I, _P, _I, _I, _P, _P, _L, _P, _L, _P, _P, _P, _L, _P, _I, _P]),
    "wsi_hetero_attn_bwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _P, _L, _P, _L, _P, _L, _P, _L,
                                 _P, _P, _P, _P, _P, _L, _P, _P]),
    "wsi_hetero_attn_seg_fwd": (_I, [_P, _L, _P, _L, _I, _P, _I, _L, _P, _P, _P, _P, _L, _I, _I, _I, _P, _L, _P, _L, _P, _I, _P]),
    "wsi_head_perm": (_I, [_I, _I, _P]),
    "wsi_rel_transform": (_I, [_P, _L, _P, _P, _P, _P, _I, _I, _I, _I, _P, _L, _P]),
    "wsi_segment_combine": (_I, [_P, _I, _L, _P, _P, _P, _L, _I, _P, _L, _P, _I, _P]),
    "wsi_typed_layernorm": (_I, [_P, _L, _P, _P, _P, _P, _I, _I, _F, _P, _L, _P, _I, _P]),
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
    "wsi_heat_forward_workspace_bytes": (_L, [_L, _I, _I, _L, _I, _I]),
    "wsi_heat_forward": (_I, [_P, _L, _I, POINTER(HeatGraph), POINTER(HeatParams), _P, _L, _P, _L, _P, _L, _P]),
    "wsi_stream_slot_bytes": (_L, [_L, _L, _L, _I, _I, _I, _I, _I]),
    "wsi_stream_host_slot_bytes": (_L, [_L, _I, _I]),
    "wsi_stream_forward": (_I, [POINTER(StreamSlide), _L, POINTER(HeatParams), _P, _I, _P, _L, _P, _L, _P]),
    "wsi_skip_mix_bwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _P, _P, _I, _I, _P, _L, _P, _L, _P, _P]),
    "wsi_typed_wgrad_supported": (_I, [_L, _I, _I, _I]),
    "wsi_typed_wgrad_workspace_bytes": (_L, [_I, _I, _P, _I]),
    "wsi_typed_wgrad": (_I, [_P, _P, _I, _I, _P, _I, _P, _P, _L, _P]),
    "wsi_adam_step_masked": (_I, [_P, _P, _P, _P, _L, _I, _P, _P, _P, _P, _F, _F, _F, _F, _F, _F, _I, _P]),
    "wsi_adam_step": (_I, [_P, _P, _P, _P, _L, _L, _F, _F, _F, _F, _F, _F, _I, _P]),
    "wsi_slide_forward_workspace_bytes": (_L, [_L, _L, _I, _I, _I, _L]),
    "wsi_slide_forward": (_I, [POINTER(SlideDesc), POINTER(HeatParams), _L, _P, _P, _L, _P, _L, _P, _P]),
    "wsi_slide_plan": (_I, [POINTER(SlideDesc), POINTER(HeatParams), _L, _P, _P, _L, _P]),
    "wsi_slide_run": (_I, [POINTER(SlideDesc), POINTER(HeatParams), _L, _P, _P, _L, _P, _L, _P, _P]),
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
