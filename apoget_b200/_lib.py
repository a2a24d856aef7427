"""ctypes binding of libapt_b200.so (the C ABI declared in include/apt_b200.h).

The library is the product; there is no Python or CPU fallback: if it is missing, or if no sm_100
device is visible when a tree or pool is created, the call fails with an exception.
"""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libapt_b200.so")

u32p = ctypes.POINTER(ctypes.c_uint32)
i32p = ctypes.POINTER(ctypes.c_int32)
u64p = ctypes.POINTER(ctypes.c_uint64)
f32p = ctypes.POINTER(ctypes.c_float)
f64p = ctypes.POINTER(ctypes.c_double)
u8p = ctypes.POINTER(ctypes.c_uint8)
vp = ctypes.c_void_p


class AptParams(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("words", ctypes.c_uint32), ("rset_size", ctypes.c_uint32),
                ("full_continuous", ctypes.c_int32), ("env_crit", ctypes.c_int32),
                ("compat_thr", ctypes.c_double), ("sim_thr", ctypes.c_double), ("avg_thr", ctypes.c_double),
                ("trim_period", ctypes.c_uint32), ("trim_thr", ctypes.c_float), ("trim_delay", ctypes.c_float),
                ("trim_min_delay", ctypes.c_uint32), ("device", ctypes.c_int32), ("flags", ctypes.c_uint32)]


class AptNodeInfo(ctypes.Structure):
    _fields_ = [("sid", ctypes.c_uint32), ("parent_sid", ctypes.c_uint32), ("data", ctypes.c_uint32 * 5),
                ("fill", ctypes.c_uint32), ("n_children", ctypes.c_uint32), ("n_contributors", ctypes.c_uint32)]


# every symbol include/apt_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "apt_params_default": (None, [ctypes.POINTER(AptParams)]),
    "apt_last_error": (ctypes.c_char_p, []),
    "apt_device_available": (ctypes.c_int, []),
    "apt_tree_create": (ctypes.c_int, [ctypes.POINTER(AptParams), ctypes.POINTER(vp)]),
    "apt_tree_destroy": (None, [vp]),
    "apt_tree_add_genomes": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32p]),
    "apt_tree_add_genomes_dev": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, u32p]),
    "apt_tree_parallel_for": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp]),
    "apt_host_alloc": (ctypes.c_int, [ctypes.c_uint64, ctypes.POINTER(vp)]),
    "apt_host_free": (None, [vp]),
    "apt_tree_del_genome": (ctypes.c_int, [vp, ctypes.c_uint32]),
    "apt_tree_del_genomes": (ctypes.c_int, [vp, ctypes.c_uint32, vp]),
    "apt_tree_register_candidate": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int]),
    "apt_tree_set_step": (ctypes.c_int, [vp, ctypes.c_uint32]),
    "apt_tree_step": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.c_uint32]),
    "apt_tree_stats": (ctypes.c_int, [vp, u32p]),
    "apt_tree_reset_stats": (ctypes.c_int, [vp]),
    "apt_tree_event_count": (ctypes.c_uint32, [vp]),
    "apt_tree_events": (ctypes.c_uint32, [vp, vp, ctypes.c_uint32, ctypes.c_int]),
    "apt_tree_to_json": (ctypes.c_uint64, [vp, ctypes.c_char_p, ctypes.c_uint64]),
    "apt_tree_from_json": (ctypes.c_int, [vp, ctypes.c_char_p, ctypes.c_uint64]),
    "apt_tree_clone": (ctypes.c_int, [vp, ctypes.POINTER(vp)]),
    "apt_tree_trimmed": (ctypes.c_uint32, [vp, vp, ctypes.c_uint32, ctypes.c_int]),
    "apt_tree_node_info": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.POINTER(AptNodeInfo)]),
    "apt_tree_node_children": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.c_uint32]),
    "apt_tree_node_contributors": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.c_uint32]),
    "apt_tree_node_rep_gids": (ctypes.c_int, [vp, ctypes.c_uint32, vp, ctypes.c_uint32]),
    "apt_tree_node_distances": (ctypes.c_int, [vp, ctypes.c_uint32, vp]),
    "apt_tree_root": (ctypes.c_int, [vp, u32p]),
    "apt_tree_alive_species": (ctypes.c_int, [vp, vp, ctypes.c_uint32, u32p]),
    "apt_tree_counters": (ctypes.c_int, [vp, u64p]),
    "apt_tree_timers": (ctypes.c_int, [vp, f64p]),
    "apt_tree_kernel_span": (ctypes.c_int, [vp, f64p]),
    "apt_pool_create": (ctypes.c_int, [ctypes.POINTER(AptParams), ctypes.c_uint32, vp, vp, vp, ctypes.POINTER(vp)]),
    "apt_pool_destroy": (None, [vp]),
    "apt_score_pairs": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp, ctypes.c_uint32, vp, vp, vp, vp]),
    "apt_classify_dense": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp, vp, vp, vp, vp]),
    "apt_classify_dense_dev": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp, vp, vp, vp, vp, vp]),
    "apt_classify_dense_shard_dev": (ctypes.c_int, [vp, ctypes.c_uint32, vp, vp, ctypes.c_uint32, ctypes.c_uint32,
                                                    ctypes.c_uint32, u64p, vp]),
    "apt_combine_shards_dev": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, vp, vp, vp, vp, vp, vp]),
    "apt_pool_uses_tensor_cores": (ctypes.c_int, [vp]),
    "apt_distance_matrix": (ctypes.c_int, [ctypes.POINTER(AptParams), ctypes.c_uint32, vp, ctypes.c_uint32, vp, vp]),
    "apt_env_contribution": (ctypes.c_int, [ctypes.POINTER(AptParams), ctypes.c_uint32, vp, vp, vp, vp, vp]),
    "apt_mate_compat": (ctypes.c_int, [ctypes.POINTER(AptParams), ctypes.c_uint32, vp, vp, ctypes.c_uint32, vp, vp, vp, vp]),
    "apt_microbench": (ctypes.c_int, [ctypes.c_int, f64p]),
    "apt_microbench_cvt": (ctypes.c_int, [ctypes.c_int, f64p]),
    "apt_pairwise": (ctypes.c_int, [ctypes.POINTER(AptParams), ctypes.c_uint32, vp, ctypes.c_uint32, vp, vp, vp]),
}

_lib = None


class AptError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libapt_b200 error {status}: {message}")
        self.status = status


def load(path: str) -> ctypes.CDLL:
    """Binds every declared symbol of the library at `path` (AttributeError if one is missing)."""
    L = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


def lib() -> ctypes.CDLL:
    """Loads the CUDA extension; raises if it has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = load(LIB_PATH)
    return _lib


def check(status: int) -> None:
    if status:
        raise AptError(status, lib().apt_last_error().decode(errors="replace"))
