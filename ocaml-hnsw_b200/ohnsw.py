"""Host-side mirror of the reference's `Ohnsw` module (lib/ohnsw.ml) over the C ABI.

Same entry points, argument meaning and error behaviour as the OCaml values they stand for, so
the parity tests read like the reference's own drivers (benchmark/benchmark.ml:74-98):

    hgraph = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, train, num_connections=16,
                                        num_nodes_search_construction=100)
    ids, distances = Ohnsw.knn_batch_bigarray(hgraph, test, k=10)

Batches are float32 arrays [n][dim] (C order) — the memory of a `Lacaml.S.mat` dim x n.
`distance` is narrowed from a closure to a tag (`distance_l2`, `distance_angular`,
`distance_ip`): an OCaml closure cannot run on the GPU (the one intentional API narrowing).
"""
import ctypes as C

import numpy as np

from . import _capi as capi

distance_l2 = capi.L2              # Ohnsw.distance_l2, lib/ohnsw.ml:899
distance_angular = capi.ANGULAR
distance_ip = capi.IP


class Hgraph:
    """Ohnsw.Hgraph.t (lib/ohnsw.ml:306-351): an opaque handle on a GPU-resident layered graph."""

    def __init__(self, dim, distance=distance_l2, num_connections=16, num_nodes_search_construction=100,
                 seed=0, device=0, flavour=capi.FLAVOUR_OHNSW):
        self._h = C.c_void_p()
        capi.check(capi.lib().hnswb200_create(C.byref(self._h), dim, distance, num_connections,
                                              num_nodes_search_construction, seed, device))
        self.dim = dim
        if flavour != capi.FLAVOUR_OHNSW:
            capi.check(capi.lib().hnswb200_set_flavour(self._h, flavour))

    def close(self):
        if getattr(self, "_borrowed", None) is not None:      # a shard of a MultiGpuHgraph: the owner destroys it
            self._h = None
            return
        if getattr(self, "_h", None) and self._h and capi is not None and getattr(capi, "_lib", None) is not None:
            capi._lib.hnswb200_destroy(self._h)       # (at interpreter shutdown the module may already be gone)
            self._h = None

    __del__ = close

    # -- Hgraph accessors (lib/ohnsw.ml:335-346)
    def info(self):
        out = capi.Info()
        capi.check(capi.lib().hnswb200_get_info(self._h, C.byref(out)))
        return out

    def num_nodes(self):
        return self.info().n

    def max_layer(self):
        return self.info().max_layer

    def entry_point(self):
        e = self.info().entry_point
        return None if e < 0 else e

    def stats(self):
        out = capi.Stats()
        capi.check(capi.lib().hnswb200_get_stats(self._h, C.byref(out)))
        return out

    def set_param(self, name, value):
        capi.check(capi.lib().hnswb200_set_param(self._h, name.encode(), int(value)))

    # -- graph exchange
    def import_graph(self, data, graph, id_base=0):
        data = capi.as_mat(data, self.dim)
        L = graph.max_layer + 1
        offs = [np.ascontiguousarray(o, np.int64) for o in graph.offsets]
        nbrs = [np.ascontiguousarray(a if len(a) else np.zeros(1, np.int32), np.int32) for a in graph.nbrs]
        po = (C.c_void_p * L)(*[o.ctypes.data for o in offs])
        pn = (C.c_void_p * L)(*[a.ctypes.data for a in nbrs])
        capi.check(capi.lib().hnswb200_import_graph(self._h, capi.ptr(data), data.shape[0], id_base,
                                                    graph.max_layer, graph.entry, po, pn))
        return self

    def export_graph(self, id_base=0):
        from .graphio import FlatGraph
        inf = self.info()
        offs, nbrs = [], []
        for l in range(inf.max_layer + 1):
            nnz = C.c_int64()
            capi.check(capi.lib().hnswb200_export_layer(self._h, l, id_base, None, None, C.byref(nnz)))
            o = np.empty(inf.n + 1, np.int64)
            a = np.empty(max(nnz.value, 1), np.int32)
            capi.check(capi.lib().hnswb200_export_layer(self._h, l, id_base, capi.ptr(o), capi.ptr(a), C.byref(nnz)))
            offs.append(o)
            nbrs.append(a[:nnz.value])
        lv = np.empty(inf.n, np.int32)
        capi.check(capi.lib().hnswb200_export_levels(self._h, capi.ptr(lv)))
        return FlatGraph(inf.n, inf.max_layer, inf.entry_point + id_base, offs, nbrs, lv)

    def search_device(self, d_queries, nq, k, ef, d_ids, d_dists, stream=0, mode=capi.MODE_PARITY):
        """hnswb200_search_device: all buffers are device pointers (ints) in this index's GPU memory;
        with a non-zero `stream` (a cudaStream_t) the call only enqueues."""
        capi.check(capi.lib().hnswb200_search_device(self._h, d_queries, nq, k, ef, mode, d_ids, d_dists, stream or None))

    def last_search_counters(self, nq):
        out = np.empty((nq, 3), np.uint32)
        capi.check(capi.lib().hnswb200_last_search_counters(self._h, capi.ptr(out), nq))
        return out


class Visited:
    """Ohnsw.Visited.t (lib/ohnsw.ml:256-268).  The GPU keeps its visited sets in shared memory
    per query; this object exists only so `insert` / `knn` keep the reference's signatures."""

    def __init__(self, n=0):
        self.n = n

    @staticmethod
    def create(n):
        return Visited(n)


def build_batch_bigarray(distance, batch, *, num_connections, num_nodes_search_construction,
                         levels=None, seed=0, device=0):
    """Ohnsw.build_batch_bigarray (lib/ohnsw.ml:840-857)."""
    batch = capi.as_mat(batch)
    h = Hgraph(batch.shape[1], distance, num_connections, num_nodes_search_construction, seed, device)
    lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(batch), batch.shape[0], capi.ptr(lv)))
    return h


def insert(hgraph, target, *, num_connections=None, num_nodes_search_construction=None, level_mult=None,
           visited=None, levels=None):
    """Ohnsw.insert (lib/ohnsw.ml:766-837); `target` may be one vector or a batch [n][dim].
    num_connections / num_nodes_search_construction / level_mult are fixed at index creation
    (the reference passes the same values on every call); they are accepted and checked."""
    inf = hgraph.info()
    if num_connections is not None and num_connections != inf.M:
        raise ValueError("insert: num_connections differs from the index's")
    if num_nodes_search_construction is not None and num_nodes_search_construction != inf.ef_construction:
        raise ValueError("insert: num_nodes_search_construction differs from the index's")
    t = np.asarray(target, np.float32)
    if t.ndim == 1:
        t = t[None, :]
    t = capi.as_mat(t, hgraph.dim)
    lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
    capi.check(capi.lib().hnswb200_insert(hgraph._h, capi.ptr(t), t.shape[0], capi.ptr(lv)))


def knn_batch_bigarray(hgraph, batch, *, k, ef=None, mode=capi.MODE_PARITY, out=None):
    """Ohnsw.knn_batch_bigarray (lib/ohnsw.ml:877-897) -> (ids [nq][k] int32, distances [nq][k] f32).

    ids are -1 and distances NaN where fewer than k were found.  The reference's beam width is
    k itself (lib/ohnsw.ml:873); `ef` > k is the "~k:ef, keep the first k rows" use."""
    batch = capi.as_mat(batch, hgraph.dim)
    nq = batch.shape[0]
    if out is None:
        ids = np.empty((nq, k), np.int32)
        dists = np.empty((nq, k), np.float32)
    else:                                  # caller-owned result buffers (a Lacaml k x nq mat, pinned or not)
        ids, dists = out
        assert ids.shape == (nq, k) and ids.dtype == np.int32 and ids.flags.c_contiguous
        assert dists.shape == (nq, k) and dists.dtype == np.float32 and dists.flags.c_contiguous
    capi.check(capi.lib().hnswb200_search(hgraph._h, capi.ptr(batch), nq, k, k if ef is None else ef, mode,
                                          capi.ptr(ids), capi.ptr(dists)))
    return ids, dists


def knn(hgraph, visited, *, k, target, ef=None):
    """Ohnsw.knn (lib/ohnsw.ml:859-875): the result min-queue as an ascending list of
    (node, distance).  Raises ValueError("knn: empty hgraph") like the reference (:862)."""
    t = np.asarray(target, np.float32)[None, :]
    ids, d = knn_batch_bigarray(hgraph, t, k=k, ef=ef)
    return [(int(i), float(x)) for i, x in zip(ids[0], d[0]) if i >= 0]
