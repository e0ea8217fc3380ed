"""ctypes binding of libstuffing_b200.so (include/stuffing_b200.h).

The product path has NO CPU fallback: if the shared library is missing, or no CUDA device can be
opened, every entry point raises.  This file is the Python twin of julia/StuffingB200.jl.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("STUFFING_B200_LIB", os.path.join(_HERE, "libstuffing_b200.so"))  # (the same override as julia/StuffingB200.jl)

STF_OK, STF_E_CUDA, STF_E_ARG, STF_E_CAPACITY, STF_E_BADCODE, STF_E_RANGE, STF_E_STATE = 0, -1, -2, -3, -4, -5, -6

ITEM_DT = np.dtype([("i1", "<i4"), ("i2", "<i4"), ("l", "<i4"), ("a", "<i4"), ("b", "<i4")])
CELL_DT = np.dtype([("l", "<i4"), ("a", "<i4"), ("b", "<i4"), ("label", "<i4")])

EXPORTS = {
    # name: (restype, argtypes)
    "stf_create": (C.c_int32, [C.c_int32, C.POINTER(C.c_void_p)]),
    "stf_destroy": (None, [C.c_void_p]),
    "stf_last_error": (C.c_char_p, [C.c_void_p]),
    "stf_stream": (C.c_void_p, [C.c_void_p]),
    "stf_upload_objects": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stf_upload_packed": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stf_num_objects": (C.c_int32, [C.c_void_p]),
    "stf_num_levels": (C.c_int32, [C.c_void_p]),
    "stf_get_level_info": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "stf_download_level": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "stf_download_trees": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_read_nodes": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stf_get_shifts": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stf_set_shifts": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "stf_set_level_shift": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "stf_build": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32]),
    "stf_collision": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "stf_collision_bfs": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "stf_collision_dfs": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "stf_collide_items": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_filter_pairs": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_totalcollisions_native": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_locate": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_totalcollisions_spacial": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_set_keep_items": (C.c_int32, [C.c_void_p, C.c_int32]),
    "stf_fetch_items": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_fetch_result": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_linked_reset": (C.c_int32, [C.c_void_p]),
    "stf_linked_locate": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "stf_linked_clear": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "stf_linked_collect": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_partialcollisions": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_void_p]),
    "stf_grad2d": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stf_set_optimiser": (C.c_int32, [C.c_void_p, C.c_int32, C.c_double, C.c_double]),
    "stf_reset_velocity": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "stf_get_velocity": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "stf_batchsteps": (C.c_int32, [C.c_void_p, C.c_int64, C.c_void_p, C.c_uint64, C.c_uint64]),
    "stf_fit_round": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "stf_fit_round_moved": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "stf_fit_round_positions": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stf_fit_rounds": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stf_fit_epoch_e": (C.c_int32, [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "stf_ground_compose": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "stf_ground_level": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "stf_shard_collide": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "stf_shard_pack": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int64]),
    "stf_shard_step": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_uint64, C.c_void_p]),
    "stf_group_create": (C.c_int32, [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "stf_group_destroy": (None, [C.c_void_p]),
    "stf_group_size": (C.c_int32, [C.c_void_p]),
    "stf_group_ctx": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "stf_group_last_error": (C.c_char_p, [C.c_void_p]),
    "stf_group_set_shifts": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "stf_group_reset_velocity": (C.c_int32, [C.c_void_p, C.c_int32, C.c_void_p]),
    "stf_group_fit_round": (C.c_int32, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]),
    "stf_group_nvlink_bytes": (C.c_int64, [C.c_void_p]),
    "stf_counters": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "stf_launch_count": (C.c_int64, [C.c_void_p]),
    "stf_profile_enable": (C.c_int32, [C.c_void_p, C.c_int32]),
    "stf_profile_build_kernels": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "stf_profile_read": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
}
STAGES = ("shift_rebuild", "locate", "sort", "emit", "narrow", "result", "step_prep", "step", "build")

COUNTER_FIELDS = ("records", "items", "direct", "hits", "overflow", "moved", "visited", "reserved")

_lib = None


class StuffingError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"stuffing_b200 error {code}: {msg}")
        self.code = code


def load():
    """dlopen the library and declare every export; raises if it is missing (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            f = getattr(L, name)  # AttributeError if the export is missing
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def i32(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.int32))
