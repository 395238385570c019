"""Loader for libb200icp.so (the C-ABI in include/b200icp.h).

The library is built in-tree by `make -C norlab_icp_mapper_b200/csrc` (or `__graft_entry__.build()`).
There is NO fallback: if the shared object is missing this raises, and if no B200 is present
`b200icp_create` fails with B200ICP_ERR_CUDA and the wrappers raise.
"""
import ctypes as C
import os

from . import _abi
from ._abi import Config, Result, Timing

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200ICP_LIB: another build of the same library next to this file (development: the stamped build of tools/gpu_stamps.py)
SO_PATH = os.path.join(_HERE, os.environ.get("B200ICP_LIB", "libb200icp.so"))

# every symbol include/b200icp.h declares (tests check the export list against the header)
SYMBOLS = [
    "b200icp_abi_version", "b200icp_config_default", "b200icp_create", "b200icp_destroy",
    "b200icp_last_error", "b200icp_stream", "b200icp_set_profiling", "b200icp_get_timing", "b200icp_set_sm_share",
    "b200icp_set_map", "b200icp_set_map_device", "b200icp_map_size", "b200icp_register",
    "b200icp_register_device", "b200icp_register_normals", "b200icp_register_descriptors", "b200icp_register_batch", "b200icp_match", "b200icp_knn", "b200icp_transform",
    "b200icp_transform_device", "b200icp_get_map_mean", "b200icp_get_grid_info",
    "b200icp_set_trace", "b200icp_get_trace", "b200icp_map_insert_point_distance",
    "b200icp_map_surface_normals", "b200icp_cloud_surface_normals", "b200icp_map_window", "b200icp_map_commit", "b200icp_map_counts",
    "b200icp_map_has_normals", "b200icp_map_download", "b200icp_map_set_prob", "b200icp_map_has_prob",
    "b200icp_map_download_prob", "b200icp_map_append", "b200icp_map_reserve", "b200icp_filter_cloud", "b200icp_scan_upload", "b200icp_scan_size", "b200icp_scan_transform", "b200icp_scan_register",
    "b200icp_scan_insert_point_distance", "b200icp_scan_download", "b200icp_scan_set_descriptors", "b200icp_scan_info", "b200icp_scan_filter",
    "b200icp_scan_add_prob", "b200icp_scan_surface_normals", "b200icp_scan_select_extra", "b200icp_scan_append", "b200icp_scan_octree",
    "b200icp_scan_dynamic_points", "b200icp_scan_download_descriptors", "b200icp_map_set_extra", "b200icp_map_extra_rows",
    "b200icp_map_select_extra", "b200icp_map_download_extra", "b200icp_map_replace_local", "b200icp_map_insert_point_distance_prob",
    "b200icp_map_evict_parked", "b200icp_map_append_cloud", "b200icp_scan_snapshot", "b200icp_map_begin_update", "b200icp_map_end_update", "b200icp_map_update_in_progress",
    "b200icp_map_octree", "b200icp_map_cut_at_threshold", "b200icp_map_dynamic_points",
]

_lib = None


class B200ICPError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"[b200icp status {status}] {message}")
        self.status = status


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `make -C norlab_icp_mapper_b200/csrc` "
            "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.b200icp_abi_version.restype = i32
    L.b200icp_config_default.argtypes = [C.POINTER(Config), i32]
    L.b200icp_config_default.restype = None
    L.b200icp_create.argtypes = [C.POINTER(Config), i32, C.POINTER(vp)]
    L.b200icp_destroy.argtypes = [vp]
    L.b200icp_destroy.restype = None
    L.b200icp_last_error.argtypes = [vp]
    L.b200icp_last_error.restype = C.c_char_p
    L.b200icp_stream.argtypes = [vp]
    L.b200icp_stream.restype = vp
    L.b200icp_set_profiling.argtypes = [vp, i32]
    L.b200icp_get_timing.argtypes = [vp, C.POINTER(Timing)]
    L.b200icp_set_sm_share.argtypes = [vp, i32]
    L.b200icp_set_map.argtypes = [vp, vp, i32, vp, i64]
    L.b200icp_set_map_device.argtypes = [vp, vp, i32, vp, i64]
    L.b200icp_map_size.argtypes = [vp]
    L.b200icp_map_size.restype = i64
    L.b200icp_register.argtypes = [vp, vp, i32, i64, vp, vp, C.POINTER(Result)]
    L.b200icp_register_device.argtypes = [vp, vp, i32, i64, vp, vp, C.POINTER(Result)]
    L.b200icp_register_normals.argtypes = [vp, vp, i32, i64, vp, vp, vp, C.POINTER(Result)]
    L.b200icp_register_descriptors.argtypes = [vp, vp, i32, i64, vp, vp, vp, vp, C.POINTER(Result)]
    L.b200icp_register_batch.argtypes = [C.POINTER(vp), i32, C.POINTER(_abi.Pair), i64, C.POINTER(_abi.PairResult)]
    L.b200icp_match.argtypes = [vp, vp, i32, i64, vp, vp]
    L.b200icp_knn.argtypes = [vp, vp, i32, i64, vp, i32, i64, i32, i32, f32, vp, vp]
    L.b200icp_transform.argtypes = [vp, vp, i32, vp, i64, vp]
    L.b200icp_transform_device.argtypes = [vp, vp, i32, vp, i64, vp]
    L.b200icp_get_map_mean.argtypes = [vp, vp]
    L.b200icp_get_grid_info.argtypes = [vp, C.POINTER(f32), vp]
    L.b200icp_set_trace.argtypes = [vp, i32]
    L.b200icp_get_trace.argtypes = [vp, vp, i32]
    L.b200icp_map_insert_point_distance.argtypes = [vp, vp, i32, i64, vp, f32, C.POINTER(i64), vp]
    L.b200icp_map_surface_normals.argtypes = [vp, i32]
    L.b200icp_map_window.argtypes = [vp, i32, vp, C.POINTER(i64)]
    L.b200icp_map_commit.argtypes = [vp]
    L.b200icp_map_counts.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    L.b200icp_map_has_normals.argtypes = [vp]
    L.b200icp_map_download.argtypes = [vp, i32, vp, vp, i64, C.POINTER(i64)]
    L.b200icp_map_set_prob.argtypes = [vp, vp, f32]
    L.b200icp_map_has_prob.argtypes = [vp]
    L.b200icp_map_download_prob.argtypes = [vp, i32, vp, i64]
    L.b200icp_map_reserve.argtypes = [vp, i64, i32]
    L.b200icp_map_append.argtypes = [vp, vp, i32, i64, vp, vp, C.POINTER(i64)]
    L.b200icp_filter_cloud.argtypes = [vp, vp, i32, C.POINTER(i64), vp, i32]
    L.b200icp_map_octree.argtypes = [vp, vp, i32, i64, vp, vp, f32, i32, i32, C.POINTER(i64)]
    L.b200icp_map_cut_at_threshold.argtypes = [vp, f32, i32, C.POINTER(i64)]
    L.b200icp_map_dynamic_points.argtypes = [vp, vp, i32, i64, vp, vp, vp]
    L.b200icp_cloud_surface_normals.argtypes = [vp, vp, i32, i64, i32, vp]
    L.b200icp_map_insert_point_distance_prob.argtypes = [vp, vp, i32, i64, vp, vp, f32, C.POINTER(i64), vp]
    L.b200icp_map_set_extra.argtypes = [vp, vp, i32]
    L.b200icp_map_extra_rows.argtypes = [vp]
    L.b200icp_map_select_extra.argtypes = [vp, vp, i32]
    L.b200icp_map_download_extra.argtypes = [vp, i32, vp, i64]
    L.b200icp_map_replace_local.argtypes = [vp, vp, i32, i64, vp, vp, vp, i32]
    L.b200icp_scan_upload.argtypes = [vp, vp, i32, i64]
    L.b200icp_scan_set_descriptors.argtypes = [vp, vp, vp, vp, i32, vp, i32]
    L.b200icp_scan_size.argtypes = [vp]
    L.b200icp_scan_size.restype = i64
    L.b200icp_scan_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.b200icp_scan_filter.argtypes = [vp, vp, i32, C.POINTER(i64)]
    L.b200icp_scan_add_prob.argtypes = [vp, f32]
    L.b200icp_scan_surface_normals.argtypes = [vp, i32]
    L.b200icp_scan_select_extra.argtypes = [vp, vp, i32]
    L.b200icp_scan_transform.argtypes = [vp, vp]
    L.b200icp_scan_register.argtypes = [vp, vp, vp, C.POINTER(Result)]
    L.b200icp_scan_insert_point_distance.argtypes = [vp, f32, C.POINTER(i64)]
    L.b200icp_scan_append.argtypes = [vp, C.POINTER(i64)]
    L.b200icp_scan_octree.argtypes = [vp, f32, i32, i32, C.POINTER(i64)]
    L.b200icp_scan_dynamic_points.argtypes = [vp, vp, vp]
    L.b200icp_scan_download.argtypes = [vp, vp, i64, C.POINTER(i64)]
    L.b200icp_scan_download_descriptors.argtypes = [vp, vp, vp, vp, i64]
    L.b200icp_map_evict_parked.argtypes = [vp, vp, vp, vp, vp, i64, C.POINTER(i64)]
    L.b200icp_map_append_cloud.argtypes = [vp, vp, i32, i64, vp, vp, vp, i32, C.POINTER(i64)]
    for name in ("b200icp_scan_snapshot", "b200icp_map_begin_update", "b200icp_map_end_update", "b200icp_map_update_in_progress"):
        getattr(L, name).argtypes = [vp]
    if L.b200icp_abi_version() != _abi.ABI_VERSION:
        raise ImportError("libb200icp.so ABI version mismatch")
    _lib = L
    return L
