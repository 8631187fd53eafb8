"""ctypes binding of libhnsw_b200.so (include/hnsw_b200.h).  No torch types cross this boundary.

The library is the product: if it is missing the import fails loudly — there is no Python or
CPU fallback for any compute entry point.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HNSWB200_LIB") or os.path.join(_HERE, "libhnsw_b200.so")   # env override: kernel tuning builds only

OK, EINVAL, ECUDA, ENOMEM = 0, 1, 2, 3
L2, ANGULAR, IP = 0, 1, 2
MODE_PARITY = 0
FLAVOUR_OHNSW, FLAVOUR_HNSW_BA = 0, 1


class Info(C.Structure):
    _fields_ = [("n", C.c_int64), ("dim", C.c_int32), ("metric", C.c_int32), ("M", C.c_int32),
                ("ef_construction", C.c_int32), ("max_layer", C.c_int32), ("entry_point", C.c_int64),
                ("slots0", C.c_int32), ("slots_upper", C.c_int32), ("flavour", C.c_int32), ("device", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("search_queries", C.c_uint64), ("search_n_dist", C.c_uint64), ("search_n_exp0", C.c_uint64),
                ("search_n_expU", C.c_uint64), ("search_visited_overflows", C.c_uint64),
                ("search_algorithmic_bytes", C.c_double), ("search_kernel_ms", C.c_double),
                ("build_inserts", C.c_uint64), ("build_n_dist", C.c_uint64), ("build_n_exp", C.c_uint64),
                ("build_algorithmic_bytes", C.c_double), ("build_seconds", C.c_double),
                ("gpu_launches", C.c_uint64), ("num_layers", C.c_int32),
                ("layer_nodes", C.c_int64 * 16), ("layer_min_degree", C.c_int32 * 16),
                ("layer_max_degree", C.c_int32 * 16), ("layer_mean_degree", C.c_double * 16),
                ("layer_isolated", C.c_int64 * 16), ("build_visited_overflows", C.c_uint64), ("search_tie_overflows", C.c_uint64),
                ("search_tie_spills", C.c_uint64), ("build_dropped_incoming", C.c_uint64),
                ("search_zero_copy", C.c_uint64)]


class HnswB200Error(RuntimeError):
    pass


_lib = None

# every symbol include/hnsw_b200.h declares: (restype, argtypes)
_vp, _i32, _i64, _u64, _f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
SIGNATURES = {
    "hnswb200_create": (_i32, [C.POINTER(_vp), _i32, _i32, _i32, _i32, _u64, _i32]),
    "hnswb200_set_flavour": (_i32, [_vp, _i32]),
    "hnswb200_set_param": (_i32, [_vp, C.c_char_p, _i64]),
    "hnswb200_destroy": (_i32, [_vp]),
    "hnswb200_build": (_i32, [_vp, _vp, _i64, _vp]),
    "hnswb200_insert": (_i32, [_vp, _vp, _i64, _vp]),
    "hnswb200_search": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "hnswb200_search_device": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hnswb200_search_device_multi": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hnswb200_last_search_counters": (_i32, [_vp, _vp, _i64]),
    "hnswb200_import_graph": (_i32, [_vp, _vp, _i64, _i32, _i32, _i64, _vp, _vp]),
    "hnswb200_export_layer": (_i32, [_vp, _i32, _i32, _vp, _vp, C.POINTER(_i64)]),
    "hnswb200_export_levels": (_i32, [_vp, _vp]),
    "hnswb200_bruteforce_knn": (_i32, [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp, _vp]),
    "hnswb200_bruteforce_last_unproven": (_i64, []),
    "hnswb200_recall": (_i32, [_vp, _vp, _i64, _i32, _f64, C.POINTER(_f64)]),
    "hnswb200_merge_topk_device": (_i32, [_vp, _vp, _i32, _i64, _i32, _i64, _vp, _vp, _vp, _vp]),
    "hnswb200_sharded_create": (_i32, [C.POINTER(_vp), _i32, _i32, _i32, _i32, _u64, _i32, _vp]),
    "hnswb200_sharded_destroy": (_i32, [_vp]),
    "hnswb200_sharded_set_param": (_i32, [_vp, C.c_char_p, _i64]),
    "hnswb200_sharded_set_flavour": (_i32, [_vp, _i32]),
    "hnswb200_sharded_build": (_i32, [_vp, _vp, _i64, _vp]),
    "hnswb200_sharded_search": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "hnswb200_sharded_search_device": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp]),
    "hnswb200_sharded_shard": (_i32, [_vp, _i32, C.POINTER(_vp), C.POINTER(_i64)]),
    "hnswb200_sharded_get_info": (_i32, [_vp, C.POINTER(Info), C.POINTER(_i32)]),
    "hnswb200_sharded_get_stats": (_i32, [_vp, C.POINTER(Stats)]),
    "hnswb200_search_device_sharded": (_i32, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp]),
    "hnswb200_get_info": (_i32, [_vp, C.POINTER(Info)]),
    "hnswb200_get_stats": (_i32, [_vp, C.POINTER(Stats)]),
    "hnswb200_host_register": (_i32, [_vp, _i64]),
    "hnswb200_host_unregister": (_i32, [_vp]),
    "hnswb200_last_error": (C.c_char_p, []),
    "hnswb200_version": (C.c_char_p, []),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C ocaml-hnsw_b200` (or __graft_entry__.build()). "
                "hnsw_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if os.environ.get("HNSWB200_LIB") and not hasattr(L, name):
                continue                       # an older tuning build named by the override: A/B probes only
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = lib().hnswb200_last_error().decode()
    if rc == EINVAL:
        raise ValueError(msg)            # OCaml: Invalid_argument
    if rc == ENOMEM:
        raise MemoryError(msg)           # OCaml: Out_of_memory
    raise HnswB200Error(msg)             # OCaml: Failure


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_mat(a, dim=None):
    """A Lacaml.S.mat (dim x n, Fortran layout) is byte-identical to a C-order float32 [n][dim]."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2:
        raise ValueError("expected a 2-D float32 array [n][dim]")
    if dim is not None and a.shape[1] != dim:
        raise ValueError(f"vector dimension {a.shape[1]} does not match the index dimension {dim}")
    return a


_PINNED = {}          # address -> bytes of the buffers pinned through host_register (this process)


def host_register(a):
    """Pin a caller buffer (the payload of a Bigarray): copies from / to it are asynchronous DMA, and the search
    reads queries from it / stores rows into it from inside the kernel, no copy (include/hnsw_b200.h)."""
    check(lib().hnswb200_host_register(ptr(a), a.nbytes))
    _PINNED[a.ctypes.data] = a.nbytes


def host_unregister(a):
    check(lib().hnswb200_host_unregister(ptr(a)))
    _PINNED.pop(a.ctypes.data, None)


def is_pinned(a):
    """True when the bytes of `a` (a C-contiguous array or slice) lie inside a buffer pinned with host_register."""
    lo = a.ctypes.data
    return a.flags.c_contiguous and any(base <= lo and lo + a.nbytes <= base + n for base, n in _PINNED.items())
