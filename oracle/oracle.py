"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes front-end for oracle/liboracle.so, the CPU restatement of lehy/ocaml-hnsw `lib/ohnsw.ml`
(see ohnsw_oracle.hpp for the parity status and the file:line each function follows).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

METRIC_L2, METRIC_ANGULAR, METRIC_IP = 0, 1, 2
SUM_SEQUENTIAL, SUM_TEAM8 = 0, 1

_lib = None


def build(force=False):
    """Compile liboracle.so with the committed recipe (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_double
        P = C.POINTER
        sig = {
            "orc_last_error": (C.c_char_p, []),
            "orc_vec_create": (vp, [i32, i32, i32]),
            "orc_vec_destroy": (None, [vp]),
            "orc_vec_set_accept_ties": (None, [vp, i32]),
            "orc_vec_set_ba_build": (None, [vp, i32]),
            "orc_vec_build": (i32, [vp, vp, i64, i32, i32, vp]),
            "orc_vec_search": (i32, [vp, vp, i64, i32, i32, vp, vp, vp]),
            "orc_vec_search_mt": (i32, [vp, vp, i64, i32, i32, vp, vp, i32, P(f64), vp]),
            "orc_vec_info": (i32, [vp, P(i64), P(i32), P(i64), P(i32)]),
            "orc_vec_counters": (None, [vp, vp, i32]),
            "orc_vec_levels": (i32, [vp, vp]),
            "orc_vec_layer_nnz": (i64, [vp, i32]),
            "orc_vec_export_layer": (i32, [vp, i32, vp, vp]),
            "orc_vec_import": (i32, [vp, vp, i64, i32, i64, vp, vp]),
            "orc_vec_invariant": (i32, [vp]),
            "orc_vec_distance": (f64, [vp, vp, vp]),
            "orc_work_distance": (C.c_float, [vp, vp, i32, i32, i32]),
            "orc_work_distance_scalar": (C.c_float, [vp, vp, i32, i32]),
            "orc_bruteforce": (i32, [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp, i32]),
            "orc_recall": (i32, [vp, vp, i64, i32, f64, P(f64)]),
            "orc_abs_create": (vp, [vp, i64]),
            "orc_abs_destroy": (None, [vp]),
            "orc_abs_layer_create": (i32, [vp, i32, i64]),
            "orc_abs_layer_create_loop": (i32, [vp, i32]),
            "orc_abs_set_connections": (i32, [vp, i32, i32, vp, i32]),
            "orc_abs_adjacent": (i32, [vp, i32, i32, vp, i32]),
            "orc_abs_add_node": (i32, [vp]),
            "orc_abs_graph_add_node": (i32, [vp, i32]),
            "orc_abs_set_entry_point": (i32, [vp, i64]),
            "orc_abs_entry_point": (i64, [vp]),
            "orc_abs_set_max_layer": (i32, [vp, i32]),
            "orc_abs_max_layer": (i32, [vp]),
            "orc_abs_num_nodes": (i64, [vp]),
            "orc_abs_layer_num_nodes": (i64, [vp, i32]),
            "orc_abs_invariant": (i32, [vp]),
            "orc_abs_graph_invariant": (i32, [vp, i32]),
            "orc_abs_search_one": (i32, [vp, i32, i32, f64]),
            "orc_abs_search_k": (i32, [vp, i32, vp, i32, i32, f64, vp, vp, i32]),
            "orc_abs_select": (i32, [vp, f64, vp, i32, i32, vp, i32]),
            "orc_abs_insert_all": (i32, [vp, i32, i32, vp]),
            "orc_abs_knn": (i32, [vp, f64, i32, vp, vp]),
            "orc_nb_create": (vp, []),
            "orc_nb_destroy": (None, [vp]),
            "orc_nb_add": (None, [vp, i32]),
            "orc_nb_remove": (None, [vp, i32]),
            "orc_nb_length": (i32, [vp]),
            "orc_nb_get": (i32, [vp, vp, i32]),
            "orc_visited_create": (vp, [i64]),
            "orc_visited_destroy": (None, [vp]),
            "orc_visited_mem": (i32, [vp, i64]),
            "orc_visited_add": (None, [vp, i64]),
            "orc_visited_clear": (None, [vp]),
            "orc_visited_card": (i64, [vp]),
            "orc_visited_set_epoch_near_max": (None, [vp, i64]),
            "orc_visited_epoch": (i64, [vp]),
            "orc_num_threads": (i32, []),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _check(rc):
    if rc != 0:
        msg = lib().orc_last_error().decode()
        if rc == 1:
            raise ValueError(msg)      # OCaml Invalid_argument
        if rc == 3:
            raise MemoryError(msg)
        raise RuntimeError(msg)


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2
    return a


class Graph:
    """Flat exchange form of a layered graph: per layer CSR, list order preserved."""

    def __init__(self, n, max_layer, entry, offsets, nbrs, levels=None):
        self.n, self.max_layer, self.entry = n, max_layer, entry
        self.offsets, self.nbrs, self.levels = offsets, nbrs, levels

    def row(self, layer, node):
        o = self.offsets[layer]
        return self.nbrs[layer][o[node]:o[node + 1]]


class VecOracle:
    """Ohnsw over fp32 row vectors ([n][dim] C-order == Lacaml.S.mat dim x n)."""

    def __init__(self, dim, metric=METRIC_L2, order=SUM_TEAM8):
        self.dim, self.metric, self.order = dim, metric, order
        self._h = lib().orc_vec_create(dim, metric, order)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_vec_destroy(self._h)
            self._h = None

    def set_accept_ties(self, on=True):
        """Hnsw.Ba's acceptance rule (lib/hnsw.ml:494-506): candidates that tie with the current maximum enter."""
        lib().orc_vec_set_accept_ties(self._h, 1 if on else 0)
        return self

    def set_ba_build(self, on=True):
        """Hnsw.Ba's build parameters as the GPU's HNSW_BA flavour follows them (see ohnsw_oracle.hpp)."""
        lib().orc_vec_set_ba_build(self._h, 1 if on else 0)
        return self

    # Ohnsw.build_batch_bigarray (ohnsw.ml:840) / repeated Ohnsw.insert (:766)
    def build(self, data, M, efC, levels=None):
        data = _f32(data)
        assert data.shape[1] == self.dim
        lv = None if levels is None else np.ascontiguousarray(levels, dtype=np.int32)
        _check(lib().orc_vec_build(self._h, _ptr(data), data.shape[0], M, efC, _ptr(lv)))
        return self

    # Ohnsw.knn_batch_bigarray (ohnsw.ml:877); ef is the reference's ~k, the first k rows are kept
    def search(self, queries, k, ef=None, counters=False):
        q = _f32(queries)
        ef = k if ef is None else ef
        ids = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        cnt = np.zeros((q.shape[0], 3), np.uint64) if counters else None
        _check(lib().orc_vec_search(self._h, _ptr(q), q.shape[0], k, ef, _ptr(ids), _ptr(dists), _ptr(cnt)))
        return (ids, dists, cnt) if counters else (ids, dists)

    def search_mt(self, queries, k, ef=None, nthreads=None):
        q = _f32(queries)
        ef = k if ef is None else ef
        nthreads = nthreads or lib().orc_num_threads()
        ids = np.empty((q.shape[0], k), np.int32)
        dists = np.empty((q.shape[0], k), np.float32)
        secs = C.c_double(0)
        tot = np.zeros(3, np.uint64)
        _check(lib().orc_vec_search_mt(self._h, _ptr(q), q.shape[0], k, ef, _ptr(ids), _ptr(dists), nthreads,
                                       C.byref(secs), _ptr(tot)))
        return ids, dists, secs.value, tot

    def info(self):
        n, ml, e, d = C.c_int64(), C.c_int(), C.c_int64(), C.c_int()
        lib().orc_vec_info(self._h, C.byref(n), C.byref(ml), C.byref(e), C.byref(d))
        return dict(n=n.value, max_layer=ml.value, entry=e.value, dim=d.value)

    def counters(self, reset=False):
        out = np.zeros(3, np.uint64)
        lib().orc_vec_counters(self._h, _ptr(out), 1 if reset else 0)
        return out

    def export(self):
        inf = self.info()
        n = inf["n"]
        offs, nbrs = [], []
        for l in range(inf["max_layer"] + 1):
            nnz = lib().orc_vec_layer_nnz(self._h, l)
            o = np.empty(n + 1, np.int64)
            a = np.empty(max(nnz, 1), np.int32)
            _check(lib().orc_vec_export_layer(self._h, l, _ptr(o), _ptr(a)))
            offs.append(o)
            nbrs.append(a[:nnz])
        lv = np.empty(n, np.int32)
        lib().orc_vec_levels(self._h, _ptr(lv))
        return Graph(n, inf["max_layer"], inf["entry"], offs, nbrs, lv)

    def import_graph(self, data, g):
        data = _f32(data)
        L = g.max_layer + 1
        offs = [np.ascontiguousarray(o, np.int64) for o in g.offsets]
        nbrs = [np.ascontiguousarray(a if len(a) else np.zeros(1, np.int32), np.int32) for a in g.nbrs]
        po = (C.c_void_p * L)(*[o.ctypes.data for o in offs])
        pn = (C.c_void_p * L)(*[a.ctypes.data for a in nbrs])
        _check(lib().orc_vec_import(self._h, _ptr(data), data.shape[0], g.max_layer, g.entry, po, pn))
        return self

    def invariant(self):
        return bool(lib().orc_vec_invariant(self._h))

    def distance(self, a, b):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b, np.float32)
        return lib().orc_vec_distance(self._h, _ptr(a), _ptr(b))


def work_distance(a, b, metric=METRIC_L2, order=SUM_TEAM8):
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_work_distance(_ptr(a), _ptr(b), a.shape[0], metric, order))


def work_distance_scalar(a, b, metric=METRIC_L2):
    """The scalar definition of the TEAM8 order (pins the AVX2 path)."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return float(lib().orc_work_distance_scalar(_ptr(a), _ptr(b), a.shape[0], metric))


def bruteforce(data, queries, k, metric=METRIC_L2, order=SUM_TEAM8, nthreads=None):
    """benchmark/dataset.ml:15-30 (plus ids)."""
    data, q = _f32(data), _f32(queries)
    ids = np.empty((q.shape[0], k), np.int32)
    dists = np.empty((q.shape[0], k), np.float32)
    _check(lib().orc_bruteforce(_ptr(data), data.shape[0], _ptr(q), q.shape[0], data.shape[1], k, metric, order,
                                _ptr(ids), _ptr(dists), nthreads or lib().orc_num_threads()))
    return ids, dists


def recall(expected, got, epsilon=1e-8):
    """Recall.compute (benchmark/dataset.ml:105-127) on [nq][k] arrays."""
    e, g = _f32(expected), _f32(got)
    if e.shape != g.shape:
        raise ValueError("Recall.compute: arrrays have unequal shapes")   # dataset.ml:112 (sic)
    out = C.c_double()
    lib().orc_recall(_ptr(e), _ptr(g), e.shape[0], e.shape[1], epsilon, C.byref(out))
    return out.value


class AbsOracle:
    """Hgraph over OCaml floats with distance |a-b| (Hgraph.Test, ohnsw.ml:361-362)."""

    def __init__(self, values):
        v = np.ascontiguousarray(values, np.float64)
        self.values = v
        self._h = lib().orc_abs_create(_ptr(v), len(v))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_abs_destroy(self._h)
            self._h = None

    def layer_create(self, layer, n):
        _check(lib().orc_abs_layer_create(self._h, layer, n))

    def layer_create_loop(self, layer=0):
        _check(lib().orc_abs_layer_create_loop(self._h, layer))

    def set_connections(self, layer, node, ids):
        a = np.ascontiguousarray(ids, np.int32)
        _check(lib().orc_abs_set_connections(self._h, layer, node, _ptr(a), len(a)))

    def adjacent(self, layer, node):
        out = np.empty(256, np.int32)
        n = lib().orc_abs_adjacent(self._h, layer, node, _ptr(out), 256)
        return out[:n].tolist()

    def add_node(self):
        return lib().orc_abs_add_node(self._h)

    def graph_add_node(self, layer):
        lib().orc_abs_graph_add_node(self._h, layer)

    def set_entry_point(self, n):
        _check(lib().orc_abs_set_entry_point(self._h, n))

    def entry_point(self):
        e = lib().orc_abs_entry_point(self._h)
        return None if e < 0 else e

    def set_max_layer(self, n):
        lib().orc_abs_set_max_layer(self._h, n)

    def max_layer(self):
        return lib().orc_abs_max_layer(self._h)

    def num_nodes(self):
        return lib().orc_abs_num_nodes(self._h)

    def layer_exists(self, layer):
        return lib().orc_abs_layer_num_nodes(self._h, layer) >= 0

    def invariant(self):
        return bool(lib().orc_abs_invariant(self._h))

    def graph_invariant(self, layer=0):
        return bool(lib().orc_abs_graph_invariant(self._h, layer))

    def search_one(self, layer, start, target):
        return lib().orc_abs_search_one(self._h, layer, start, float(target))

    def search_k(self, layer, start_nodes, k, target):
        s = np.ascontiguousarray(start_nodes, np.int32)
        on, od = np.empty(1024, np.int32), np.empty(1024, np.float64)
        n = lib().orc_abs_search_k(self._h, layer, _ptr(s), len(s), k, float(target), _ptr(on), _ptr(od), 1024)
        return list(zip(on[:n].tolist(), od[:n].tolist()))

    def select(self, target, cands, num):
        c = np.ascontiguousarray(cands, np.int32)
        out = np.empty(1024, np.int32)
        n = lib().orc_abs_select(self._h, float(target), _ptr(c), len(c), num, _ptr(out), 1024)
        return out[:n].tolist()

    def insert_all(self, M, efC, levels=None):
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        _check(lib().orc_abs_insert_all(self._h, M, efC, _ptr(lv)))

    def knn(self, target, k):
        on, od = np.empty(max(k, 1), np.int32), np.empty(max(k, 1), np.float64)
        n = lib().orc_abs_knn(self._h, float(target), k, _ptr(on), _ptr(od))
        if n < 0:
            _check(-n)
        return list(zip(on[:n].tolist(), od[:n].tolist()))


class NeighboursBox:
    def __init__(self):
        self._h = lib().orc_nb_create()

    def __del__(self):
        lib().orc_nb_destroy(self._h)

    def add(self, n):
        lib().orc_nb_add(self._h, n)

    def remove(self, n):
        lib().orc_nb_remove(self._h, n)

    def length(self):
        return lib().orc_nb_length(self._h)

    def list(self):
        out = np.empty(256, np.int32)
        n = lib().orc_nb_get(self._h, _ptr(out), 256)
        return out[:n].tolist()


class VisitedBox:
    def __init__(self, n):
        self._h = lib().orc_visited_create(n)

    def __del__(self):
        lib().orc_visited_destroy(self._h)

    def mem(self, node):
        r = lib().orc_visited_mem(self._h, node)
        if r < 0:
            raise IndexError("index out of bounds")
        return bool(r)

    def add(self, node):
        lib().orc_visited_add(self._h, node)

    def clear(self):
        lib().orc_visited_clear(self._h)

    def card(self):
        return lib().orc_visited_card(self._h)

    def set_epoch_near_max(self, below):
        lib().orc_visited_set_epoch_near_max(self._h, below)
