"""ctypes binding of the C-ABI (include/wn_b200.h). Fails loudly if the CUDA library is missing: there is no fallback."""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WN_B200_LIB") or os.path.join(_HERE, "lib", "libwn_b200.so")  # the override is for A/B experiments

WN_OK = 0
WN_QUERY_DEFAULT = 0
WN_QUERY_PRESORTED = 1
WN_QUERY_NO_TILING = 2
WN_QUERY_OUT_BITS = 8
WN_HIERARCHY = {"lbvh": 0, "kd": 1, "kd_sah": 2, "reference": 3}
WN_RADIUS_BOX_CORNER = 0
WN_RADIUS_VERTEX = 1


class WnError(RuntimeError):
    """Error reported by libwn_b200 (status code + wn_last_error())."""

    def __init__(self, status, message):
        super().__init__(f"libwn_b200 status {status}: {message}")
        self.status = status


class wn_options(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("device", ctypes.c_int32), ("accuracy_scale", ctypes.c_float), ("order", ctypes.c_int32),
        ("leaf_size", ctypes.c_int32), ("morton_bits", ctypes.c_int32), ("radius_mode", ctypes.c_int32),
        ("approximate_single_triangles", ctypes.c_int32), ("keep_build_data", ctypes.c_int32), ("hierarchy", ctypes.c_int32), ("reserved", ctypes.c_int32 * 6),
    ]


class wn_info(ctypes.Structure):
    _fields_ = [
        ("struct_size", ctypes.c_uint32), ("device", ctypes.c_int32), ("num_vertices", ctypes.c_int64), ("num_triangles", ctypes.c_int64),
        ("num_tree_nodes", ctypes.c_int64), ("num_entries", ctypes.c_int64), ("num_leaf_entries", ctypes.c_int64),
        ("tree_bytes", ctypes.c_int64), ("build_scratch_bytes", ctypes.c_int64), ("build_ms", ctypes.c_float),
        ("build_ms_morton", ctypes.c_float), ("build_ms_sort", ctypes.c_float), ("build_ms_hierarchy", ctypes.c_float),
        ("build_ms_moments", ctypes.c_float), ("build_ms_pack", ctypes.c_float), ("max_depth", ctypes.c_int32), ("width", ctypes.c_int32),
        ("accuracy_scale", ctypes.c_float), ("order", ctypes.c_int32),
    ]


class wn_query_stats(ctypes.Structure):
    _fields_ = [("node_tests", ctypes.c_uint64), ("far_field_evals", ctypes.c_uint64), ("exact_triangles", ctypes.c_uint64),
                ("lane_slots", ctypes.c_uint64)]


_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_f = ctypes.c_float
_u32 = ctypes.c_uint32
_i32 = ctypes.c_int32
_f3 = ctypes.c_float * 3
_l3 = ctypes.c_int64 * 3

# name -> (restype, argtypes). Must list every WN_API symbol of include/wn_b200.h (tests/test_capi_symbols.py checks).
SIGNATURES = {
    "wn_last_error": (ctypes.c_char_p, []),
    "wn_version": (ctypes.c_char_p, []),
    "wn_options_init": (ctypes.c_int, [ctypes.POINTER(wn_options)]),
    "wn_create": (ctypes.c_int, [_vp, _i64, _vp, _i64, ctypes.POINTER(wn_options), ctypes.POINTER(_vp)]),
    "wn_create_from_topology": (ctypes.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, ctypes.POINTER(wn_options), ctypes.POINTER(_vp)]),
    "wn_destroy": (ctypes.c_int, [_vp]),
    "wn_get_info": (ctypes.c_int, [_vp, ctypes.POINTER(wn_info)]),
    "wn_solid_angle": (ctypes.c_int, [_vp, _vp, _i64, _f, _u32, _vp, _vp]),
    "wn_is_inside": (ctypes.c_int, [_vp, _vp, _i64, _f, _u32, _vp, _vp]),
    "wn_query_grid": (ctypes.c_int, [_vp, _f3, _f3, _l3, _i64, _i64, _f, _u32, _vp, _vp, _vp]),
    "wn_query_grid_strided": (ctypes.c_int, [_vp, _f3, _f3, _l3, _i64, _i64, _f, _u32, _vp, _vp, _vp]),
    "wn_query_grid_sharded": (ctypes.c_int, [_vp, _f3, _f3, _l3, _i32, _i32, _f, _u32, _vp, _vp, _vp]),
    "wn_grid_shard_layout": (ctypes.c_int, [_l3, _i32, _i32, ctypes.POINTER(_i32), ctypes.POINTER(_i32), ctypes.POINTER(_i64), ctypes.POINTER(_i64),
                                            ctypes.POINTER(_i64)]),
    "wn_query_stats_points": (ctypes.c_int, [_vp, _vp, _i64, _f, _u32, ctypes.POINTER(wn_query_stats), _vp]),
    "wn_query_stats_grid": (ctypes.c_int, [_vp, _f3, _f3, _l3, _i64, _i64, _f, _u32, ctypes.POINTER(wn_query_stats), _vp]),
    "wn_exact": (ctypes.c_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    "wn_exact_grid": (ctypes.c_int, [_vp, _f3, _f3, _l3, _i64, _i64, _vp, _vp, _vp]),
    "wn_sdf_grid": (ctypes.c_int, [_vp, _f3, _f3, _l3, _f, _f, _u32, _vp, ctypes.POINTER(_i64), _vp]),
    "wn_sdf_grid_sparse": (ctypes.c_int, [_vp, _f3, _f3, _l3, _f, _f, _u32, _i64, _vp, _vp, _vp, ctypes.POINTER(_i64), _vp]),
    "wn_closest_point": (ctypes.c_int, [_vp, _vp, _i64, _f, _u32, _vp, _vp, _vp, _vp]),
    "wn_tree_packed_size": (ctypes.c_int, [_vp, ctypes.POINTER(_i64)]),
    "wn_tree_pack": (ctypes.c_int, [_vp, _vp, _i64, _vp]),
    "wn_create_from_packed": (ctypes.c_int, [_vp, _i64, ctypes.POINTER(wn_options), ctypes.POINTER(_vp)]),
    "wn_replicate": (ctypes.c_int, [_vp, ctypes.POINTER(_i32), _i32, ctypes.POINTER(_vp)]),
    "wn_query_grid_multi": (ctypes.c_int, [ctypes.POINTER(_vp), _i32, _f3, _f3, _l3, _f, _u32, _vp, _vp]),
    "wn_debug_node_moments": (ctypes.c_int, [_vp, _i64, _i64, _vp]),
    "wn_debug_topology": (ctypes.c_int, [_vp, _vp, _i64, ctypes.POINTER(_i64)]),
    "wn_debug_last_plan": (ctypes.c_int, [_vp, _vp, _i64, ctypes.POINTER(_i64)]),
    "wn_debug_sort_pairs_u64": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32]),
    "wn_debug_sort_pairs_u32": (ctypes.c_int, [_vp, _vp, _i64, _i32, _i32]),
    "wn_debug_fma_peak": (ctypes.c_int, [_i32, _i32, ctypes.POINTER(_f), ctypes.POINTER(_f)]),
}

_lib = None


def lib() -> ctypes.CDLL:
    """Load libwn_b200.so. Raises ImportError if it has not been built (python -m lagrange_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA engine has not been built (run `python -m lagrange_b200.build`). "
                "lagrange_b200 has no CPU fallback.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status != WN_OK:
        raise WnError(status, lib().wn_last_error().decode("utf-8", "replace"))
