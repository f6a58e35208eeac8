"""ctypes binding of the C ABI (include/mrmp_b200.h).

The product path has no CPU fallback: if the CUDA shared library is missing or no
device is present, calls raise MRMPError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

OK, ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_NOMEM, ERR_UNSUPPORTED = 0, -1, -2, -3, -4, -5
MODEL_POINT2D, MODEL_POINT3D, MODEL_ARM22 = 0, 1, 2
MODEL_LINE2D, MODEL_SNAKE2D, MODEL_ARM33, MODEL_CAPSULE3D = 3, 4, 5, 6
CROSS_COUNT, CROSS_FIRST, CROSS_FLAG_CBS = 2, 3, 1
STATE_DIM = {MODEL_POINT2D: 2, MODEL_POINT3D: 3, MODEL_ARM22: 2, MODEL_LINE2D: 3, MODEL_SNAKE2D: 6,
             MODEL_ARM33: 6, MODEL_CAPSULE3D: 6}
WORK_DIM = {MODEL_POINT2D: 2, MODEL_POINT3D: 3, MODEL_ARM22: 2, MODEL_LINE2D: 2, MODEL_SNAKE2D: 2,
            MODEL_ARM33: 3, MODEL_CAPSULE3D: 3}

# every buffer argument is declared void* so that numpy arrays (ndarray.ctypes.data_as),
# torch pinned tensors (data_ptr()) and raw device pointers can all be passed
_vp = C.c_void_p
_dp = _ip = _bp = _lp = C.c_void_p


class MRMPError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"mrmp_b200 error {code}: {msg}")
        self.code = code


_SIGS = {
    "mrmp_last_error": (C.c_char_p, []),
    "mrmp_version": (C.c_char_p, []),
    "mrmp_kernel_launch_count": (C.c_int64, []),
    "mrmp_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mrmp_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_int64]),
    "mrmp_host_free": (C.c_int, [_vp]),
    "mrmp_ctx_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, _dp, _dp, _dp, C.c_int, _dp,
                                  C.c_double, C.c_double, C.c_double, C.c_int]),
    "mrmp_ctx_destroy": (None, [_vp]),
    "mrmp_ctx_state_dim": (C.c_int, [_vp]),
    "mrmp_ctx_sync": (C.c_int, [_vp]),
    "mrmp_ctx_set_stream": (C.c_int, [_vp, _vp]),
    "mrmp_connect_point": (C.c_int, [_vp, C.c_int, _dp, _bp]),
    "mrmp_connect_points": (C.c_int, [_vp, C.c_int64, _ip, _dp, _bp]),
    "mrmp_connect_edge": (C.c_int, [_vp, C.c_int, _dp, _dp, _bp]),
    "mrmp_connect_edges": (C.c_int, [_vp, C.c_int64, _ip, _dp, _dp, _bp]),
    "mrmp_collide_pair": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _bp]),
    "mrmp_collide_pairs": (C.c_int, [_vp, C.c_int64, _ip, _ip, _dp, _dp, _dp, _dp, _bp]),
    "mrmp_collide_static_pairs": (C.c_int, [_vp, C.c_int64, _ip, _ip, _dp, _dp, _bp]),
    "mrmp_collide_configs": (C.c_int, [_vp, C.c_int64, _dp, _dp, _bp, _lp]),
    "mrmp_collide_config_moves": (C.c_int, [_vp, C.c_int, _dp, C.c_int64, _dp, _bp]),
    "mrmp_connect_points_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "mrmp_connect_edges_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "mrmp_collide_pairs_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mrmp_collide_static_pairs_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mrmp_collide_configs_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    "mrmp_collide_config_moves_dev": (C.c_int, [_vp, C.c_int, _vp, C.c_int64, _vp, _vp, _vp]),
    "mrmp_roadmaps_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "mrmp_roadmaps_destroy": (None, [_vp]),
    "mrmp_roadmaps_set_nodes": (C.c_int, [_vp, C.c_int, _dp]),
    "mrmp_roadmaps_sample": (C.c_int, [_vp, C.c_int, C.c_uint64, _dp, _dp, _lp]),
    "mrmp_roadmaps_build": (C.c_int, [_vp, _dp]),
    "mrmp_roadmaps_build_async": (C.c_int, [_vp, _dp]),
    "mrmp_roadmaps_finish": (C.c_int, [_vp, _lp]),
    "mrmp_roadmaps_nnz": (C.c_int, [_vp, _lp]),
    "mrmp_roadmaps_used_grid": (C.c_int, [_vp]),
    "mrmp_roadmaps_get_nodes": (C.c_int, [_vp, _dp]),
    "mrmp_roadmaps_get_csr": (C.c_int, [_vp, _ip, _ip, C.c_int64, _lp]),
    "mrmp_roadmaps_read_async": (C.c_int, [_vp, _dp, _ip, _ip, C.c_int64, _lp]),
    "mrmp_roadmaps_read_wait": (C.c_int, [_vp]),
    "mrmp_host_register": (C.c_int, [_vp, C.c_size_t]),
    "mrmp_host_unregister": (C.c_int, [_vp]),
    "mrmp_roadmaps_device_ptrs": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(_vp),
                                            C.POINTER(_vp), C.POINTER(_vp), _lp]),
    "mrmp_roadmaps_set_nodes_dev": (C.c_int, [_vp, C.c_int, _vp, _vp]),
    "mrmp_roadmaps_bind_nodes": (C.c_int, [_vp, C.c_int, _vp]),
    "mrmp_prm_build": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip, _ip, C.c_int64,
                                 _lp]),
    "mrmp_prms": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_uint64, _dp, _dp, _dp, _ip,
                            _ip, C.c_int64, _lp]),
    "mrmp_collide_edges_indexed": (C.c_int, [_vp, C.c_int64, _ip, _ip, _ip, _ip, _ip, _ip, _bp]),
    "mrmp_collide_edges_indexed_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp,
                                                 _vp, _vp]),
    "mrmp_collide_cross": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int,
                                     _vp]),
    "mrmp_collide_cross_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp,
                                         C.c_int, _vp, _vp]),
    "mrmp_collide_cross_reduce": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp,
                                            _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp]),
    "mrmp_collide_cross_reduce_dev": (C.c_int, [_vp, C.c_int64, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp,
                                                _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int,
                                                _vp, _vp]),
    "mrmp_roadmaps_edge_conflicts": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int, _vp,
                                               C.c_int64]),
    "mrmp_roadmaps_edge_conflicts_dev": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int,
                                                   _vp, _vp]),
    "mrmp_roadmaps_edge_conflicts_async": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp, C.c_int,
                                                     _vp, C.c_int64, _vp]),
    "mrmp_roadmaps_adopt_csr_shards_async": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64, _vp]),
    "mrmp_roadmaps_push_table_rows": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, C.c_int64, _vp]),
    "mrmp_roadmaps_strong_step": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_uint64, _dp, _dp, _dp,
                                            C.c_int, C.c_int64, C.c_int64, _vp, _vp, _vp, C.c_int, _vp,
                                            C.c_int, C.c_int64, C.c_int, C.c_int]),
    "mrmp_exchange_push_counted": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _vp, C.c_int64, C.c_int64, _vp]),
    "mrmp_roadmaps_set_csr_dev": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp]),
    "mrmp_roadmaps_pack_csr_shard_dev": (C.c_int, [_vp, C.c_int, C.c_int64, _vp, _vp]),
    "mrmp_roadmaps_set_csr_shards_dev": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int64, _vp, _vp, _lp]),
    "mrmp_exchange_handle_bytes": (C.c_int, []),
    "mrmp_exchange_create": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _vp, C.POINTER(_vp), _vp]),
    "mrmp_exchange_destroy": (None, [_vp]),
    "mrmp_exchange_connect": (C.c_int, [_vp, _vp]),
    "mrmp_exchange_connect_ptrs": (C.c_int, [_vp, _vp]),
    "mrmp_exchange_local_buffer": (_vp, [_vp]),
    "mrmp_exchange_ptr": (_vp, [_vp, C.c_int, C.c_int]),
    "mrmp_exchange_slots": (_vp, [_vp, C.c_int]),
    "mrmp_exchange_slot_bytes": (C.c_int64, [_vp, C.c_int]),
    "mrmp_exchange_push": (C.c_int, [_vp, C.c_int, C.c_int, _vp, C.c_int64, _vp]),
    "mrmp_exchange_signal": (C.c_int, [_vp, C.c_int, C.c_uint, _vp]),
    "mrmp_exchange_wait": (C.c_int, [_vp, C.c_int, C.c_uint, _vp]),
    "mrmp_exchange_advance": (C.c_int, [_vp, C.c_int]),
    "mrmp_exchange_error": (C.c_int, [_vp]),
    "mrmp_roadmaps_push_csr_shard": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64, _vp]),
    "mrmp_roadmaps_adopt_csr_shards": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int64, _vp,
                                                 _lp]),
    "mrmp_ctx_stats": (C.c_int, [_vp, _vp, C.c_int]),
    "mrmp_ctx_stats8": (C.c_int, [_vp, _vp, C.c_int]),
    "mrmp_sample_states": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_uint64, C.c_int64, _dp]),
    "mrmp_sample_free": (C.c_int, [_vp, C.c_int, C.c_uint64, C.c_int64, _dp, _lp]),
    "mrmp_state_dist": (C.c_int, [_vp, C.c_int64, _dp, _dp, _dp]),
    "mrmp_mid_status": (C.c_int, [_vp, C.c_int64, _dp, _dp, _dp]),
    "mrmp_rrt_extend": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_double,
                                  C.POINTER(C.c_int32), C.POINTER(C.c_int32), _dp]),
    "mrmp_nearest": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int, _ip, _dp]),
    "mrmp_validate_solution": (C.c_int, [_vp, C.c_int64, _dp, C.POINTER(C.c_int32), _lp,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                         C.POINTER(C.c_int32)]),
    "mrmp_first_conflict": (C.c_int, [_vp, C.c_int64, _dp, _lp, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "mrmp_distance_tables": (C.c_int, [_vp, C.c_int, _ip, _dp, _ip, _ip, _ip, _dp]),
    "mrmp_distance_tables_g": (C.c_int, [_vp, C.c_int, _ip, _dp, _ip, _ip, _ip, C.c_double, _dp]),
    "mrmp_sssp_create": (C.c_int, [_vp, C.c_int, C.POINTER(_vp)]),
    "mrmp_sssp_destroy": (None, [_vp]),
    "mrmp_sssp_set_roadmap": (C.c_int, [_vp, C.c_int, C.c_int, _dp]),
    "mrmp_sssp_nv": (C.c_int, [_vp, C.c_int, C.POINTER(C.c_int)]),
    "mrmp_sssp_get_nodes": (C.c_int, [_vp, C.c_int, _dp, C.c_int]),
    "mrmp_sssp_last_call_us": (C.c_int, [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mrmp_sssp_expand": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, C.c_uint64, C.c_uint64,
                                   C.c_int, C.c_double, C.c_double, _ip, C.c_int, _ip, C.c_int,
                                   C.POINTER(C.c_int32), _dp, C.POINTER(C.c_int32), _vp, _vp,
                                   C.c_int64, C.POINTER(C.c_int32), _ip, _bp]),
}

_LIB = None


def load(path: str | None = None) -> C.CDLL:
    """Load libmrmp_b200.so (building it in-tree if the sources are newer)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or os.environ.get("MRMP_B200_LIB") or _build.LIB_PATH
    if path is None and "MRMP_B200_LIB" not in os.environ:
        if _build._stale():
            # Sources differ from what the shipped .so was built from: rebuild or fail loudly.
            # A stale library would be called through argtypes it does not have.
            try:
                p = _build.build()
            except Exception as e:
                raise MRMPError(ERR_CUDA, "libmrmp_b200.so is missing or older than csrc/ and "
                                          f"cannot be rebuilt: {e}")
    lib = C.CDLL(p)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _LIB = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise MRMPError(rc, load().mrmp_last_error().decode("utf-8", "replace"))


def f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def kernel_launch_count() -> int:
    return int(load().mrmp_kernel_launch_count())


def device_count() -> int:
    n = C.c_int(0)
    rc = load().mrmp_device_count(C.byref(n))
    return int(n.value) if rc == OK else 0
