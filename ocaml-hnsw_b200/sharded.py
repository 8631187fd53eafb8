"""Multi-GPU layer: the dataset split into contiguous row shards, one per GPU.

The reference has nothing distributed; this is the sharding BASELINE.json's
north_star prescribes on top of the same `Ohnsw` entry points: every shard builds and searches
the HNSW sub-graph of its own rows `[lo, hi)`, the query batch goes to every shard, and the
per-shard `[nq][k]` result rows are merged into global top-k rows.  Exchange and merge are the tail
of the search kernel itself (csrc/search.cuh, ShardTail): a warp stores its finished row — ids
already global — into a gather block on the home GPU through the NVLink peer mapping and bumps the
query's arrival counter there; the warp that arrives last merges the rows in place.  No all-gather,
no merge launch.  No collective runs during graph traversal or the build.

Two hosts for the same kernels:
  * `MultiGpuHgraph` — ONE process drives every GPU through `hnswb200_sharded_*` (no torch, no NCCL):
    what an OCaml or C host uses (examples/benchmark_c.c --gpus N).
  * `ShardedHgraph` — one process per GPU (torchrun): buffers in torch symmetric memory, one device
    barrier per step; where symmetric memory is unavailable, ONE packed NCCL all-gather followed by
    `hnswb200_merge_topk_device`.  torch is plumbing here: peer-mapped buffers and `torch.distributed`.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from . import ohnsw


def shard_range(n, rank, world):
    """Contiguous rows of shard `rank`: [rank*n/world, (rank+1)*n/world)  (SURVEY.md section 8e)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def shard_offsets(n, world):
    return np.array([shard_range(n, r, world)[0] for r in range(world)], np.int64)


def gather_rows(local, world, group=None):
    """All-gather a per-shard result block -> [world][...] on every rank (any backend)."""
    import torch
    import torch.distributed as dist
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    if world == 1:
        out[0].copy_(local)
    else:
        # concatenated along dim 0 (the layout every backend accepts); `out` is the stacked view of it
        dist.all_gather_into_tensor(out.view((world * local.shape[0],) + tuple(local.shape[1:])), local.contiguous(), group=group)
    return out


class MultiGpuHgraph:
    """`Ohnsw.Hgraph` over every GPU of the box from one process (`hnswb200_sharded_*`, include/hnsw_b200.h)."""

    def __init__(self, dim, distance=ohnsw.distance_l2, num_connections=16, num_nodes_search_construction=100, seed=0,
                 devices=(0,)):
        self._s = C.c_void_p()
        devs = (C.c_int * len(devices))(*devices)
        capi.check(capi.lib().hnswb200_sharded_create(C.byref(self._s), dim, distance, num_connections,
                                                      num_nodes_search_construction, seed, len(devices), devs))
        self.dim, self.n_shards = dim, len(devices)

    def close(self):
        if getattr(self, "_s", None) and self._s and capi is not None and getattr(capi, "_lib", None) is not None:
            capi._lib.hnswb200_sharded_destroy(self._s)
            self._s = None

    __del__ = close

    @staticmethod
    def build_batch_bigarray(distance, batch, *, num_connections, num_nodes_search_construction, devices=(0,), levels=None,
                             seed=0, params=None):
        """Ohnsw.build_batch_bigarray (lib/ohnsw.ml:840-857) over all shards at once."""
        batch = capi.as_mat(batch)
        m = MultiGpuHgraph(batch.shape[1], distance, num_connections, num_nodes_search_construction, seed, devices)
        for name, v in (params or {}).items():
            m.set_param(name, v)
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        capi.check(capi.lib().hnswb200_sharded_build(m._s, capi.ptr(batch), batch.shape[0], capi.ptr(lv)))
        return m

    def set_param(self, name, value):
        capi.check(capi.lib().hnswb200_sharded_set_param(self._s, name.encode(), int(value)))

    def knn_batch_bigarray(self, batch, *, k, ef=None, mode=capi.MODE_PARITY, out=None):
        """Ohnsw.knn_batch_bigarray (lib/ohnsw.ml:877-897): global ids, rows ascending by (distance, id)."""
        batch = capi.as_mat(batch, self.dim)
        nq = batch.shape[0]
        ids, dists = out if out is not None else (np.empty((nq, k), np.int32), np.empty((nq, k), np.float32))
        capi.check(capi.lib().hnswb200_sharded_search(self._s, capi.ptr(batch), nq, k, k if ef is None else ef, mode,
                                                      capi.ptr(ids), capi.ptr(dists)))
        return ids, dists

    def search_device(self, d_queries, nq, k, ef, d_ids, d_dists, stream=0, mode=capi.MODE_PARITY):
        """All buffers on the first shard's device (ints = device pointers)."""
        capi.check(capi.lib().hnswb200_sharded_search_device(self._s, d_queries, nq, k, ef, mode, d_ids, d_dists, stream or None))

    def shard(self, i):
        """(shard i as an Ohnsw.Hgraph borrowed from this handle, its first global row)."""
        h, first = C.c_void_p(), C.c_int64()
        capi.check(capi.lib().hnswb200_sharded_shard(self._s, i, C.byref(h), C.byref(first)))
        g = ohnsw.Hgraph.__new__(ohnsw.Hgraph)
        g._h, g.dim, g._borrowed = h, self.dim, self           # keeps the owner alive; never destroyed through this view
        return g, first.value

    def info(self):
        out, ns = capi.Info(), C.c_int()
        capi.check(capi.lib().hnswb200_sharded_get_info(self._s, C.byref(out), C.byref(ns)))
        return out

    def num_nodes(self):
        return self.info().n

    def stats(self):
        out = capi.Stats()
        capi.check(capi.lib().hnswb200_sharded_get_stats(self._s, C.byref(out)))
        return out


class ShardedHgraph:
    """`Ohnsw.Hgraph` over `world` GPUs, one process each, as S row shards x R replicas (S * R = world).

    Rank r is shard s = r % S of replica group g = r // S.  The rows are cut into S contiguous shards (a replica
    group holds the whole dataset); a batch of queries is cut into R contiguous slices, group g answers slice g —
    every shard of the group searches the slice, rows are exchanged and merged inside the search kernel on the
    group's first rank (ShardTail) and the merged rows are stored straight into EVERY rank's result block.  R = 1
    is the pure row sharding of SURVEY.md section 8e (datasets larger than one GPU); R = world is pure replication
    (an index that fits one GPU many times over: no query is answered twice, no per-shard top-k is wasted)."""

    def __init__(self, local, n_total, rank, world, group=None, peer_exchange=True, replicas=1):
        if world % replicas:
            raise ValueError("replicas must divide the number of ranks")
        self.local, self.n_total, self.rank, self.world, self.group = local, int(n_total), rank, world, group
        self.R, self.S = replicas, world // replicas
        self.shard, self.replica = rank % self.S, rank // self.S
        self.offsets = shard_offsets(n_total, self.S)
        self._buf = {}
        self.peer_exchange = peer_exchange and world > 1     # False: packed NCCL all-gather (pure sharding only)
        self.exchange = "none" if world == 1 else None       # set on first use
        self._step = 0

    @staticmethod
    def rows_of(n_total, rank, world, replicas=1):
        """Rows [lo, hi) this rank indexes."""
        S = world // replicas
        return shard_range(n_total, rank % S, S)

    @staticmethod
    def build(distance, local_rows, n_total, *, num_connections, num_nodes_search_construction, rank=0, world=1,
              group=None, levels=None, seed=0, device=0, params=None, replicas=1):
        """Ohnsw.build_batch_bigarray on this rank's rows."""
        local_rows = capi.as_mat(local_rows)
        lo, hi = ShardedHgraph.rows_of(n_total, rank, world, replicas)
        if local_rows.shape[0] != hi - lo:
            raise ValueError(f"rank {rank}: expected rows [{lo}, {hi}) of the dataset, got {local_rows.shape[0]} rows")
        S = world // replicas
        h = ohnsw.Hgraph(local_rows.shape[1], distance, num_connections, num_nodes_search_construction, seed + rank % S, device)
        for name, v in (params or {}).items():
            h.set_param(name, v)
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(local_rows), local_rows.shape[0], capi.ptr(lv)))
        return ShardedHgraph(h, n_total, rank, world, group, replicas=replicas)

    def query_slice(self, nq, replica=None):
        """Queries [lo, hi) of a batch that replica group `replica` (default: this rank's) answers."""
        return shard_range(nq, self.replica if replica is None else replica, self.R)

    def _peer_buffers(self, nq, k, dev):
        """Two symmetric-memory blocks (double buffered: the rows a call returns stay valid until the call after
        next).  Layout in int32 words: gather rows [S][2 (ids | dists, each [slice][k])] | merged rows [2][nq][k] |
        arrival counters [slice].  Gather rows and counters are used on the first rank of each replica group (its
        home); the merged rows are stored into every rank's block by whichever warp arrives last for a query."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        out = []
        nqk = nq * k
        sl_max = -(-nq // self.R)
        lo, hi = self.query_slice(nq)
        sl = hi - lo
        g_words = self.S * 2 * sl_max * k
        words = g_words + 2 * nqk + sl_max
        for _ in range(2):
            t = symm_mem.empty((words,), dtype=torch.int32, device=dev)
            t.zero_()
            hdl = symm_mem.rendezvous(t, group)
            home = int(hdl.buffer_ptrs[self.replica * self.S])
            fin = g_words * 4
            # every rank's merged-row block, at this group's slice of the batch
            f_ids = (C.c_void_p * self.world)(*[int(hdl.buffer_ptrs[r]) + fin + lo * k * 4 for r in range(self.world)])
            f_d = (C.c_void_p * self.world)(*[int(hdl.buffer_ptrs[r]) + fin + nqk * 4 + lo * k * 4 for r in range(self.world)])
            out.append(dict(t=t, hdl=hdl, g_ids=home, g_d=home + self.S * sl * k * 4, arrive=home + fin + 2 * nqk * 4,
                            f_ids=f_ids, f_d=f_d, out_ids=t[g_words:g_words + nqk].view(nq, k),
                            out_d=t[g_words + nqk:g_words + 2 * nqk].view(torch.float32).view(nq, k),
                            arrive_t=t[g_words + 2 * nqk:]))
        torch.cuda.synchronize()
        out[0]["hdl"].barrier(channel=0)
        torch.cuda.synchronize()
        return out

    def _buffers(self, nq, k):
        import torch
        key = (nq, k)
        if key not in self._buf:
            dev = torch.device("cuda", self.local.info().device)
            peer = None
            layout = f"{self.S} row shard(s) x {self.R} replica(s)"
            if self.peer_exchange and self.world <= 8:
                try:
                    peer = self._peer_buffers(nq, k, dev)
                    self.exchange = (f"{layout}; rows stored into the group home's gather block over NVLink and merged by the last warp "
                                     "to arrive, both inside the search kernel; one device barrier per step")
                except Exception as e:                      # no symmetric memory on this system / build
                    self.exchange = f"{layout}; one packed NCCL all-gather, then the merge kernel (symmetric memory unavailable: {type(e).__name__})"
            elif self.world > 1:
                self.exchange = f"{layout}; one packed NCCL all-gather, then the merge kernel"
            if peer is None and self.R > 1:
                raise capi.HnswB200Error("replica groups need torch symmetric memory (peer-mapped buffers)")
            self._peer = peer
            # one packed block per rank: [0] = ids (int32), [1] = distances (fp32 bits) -> ONE all-gather
            packed = torch.empty((2, nq, k), dtype=torch.int32, device=dev)
            self._buf = {key: dict(
                packed=packed, ids=packed[0], d=packed[1].view(torch.float32),
                gathered=torch.empty((self.world, 2, nq, k), dtype=torch.int32, device=dev),
                out_ids=torch.empty((nq, k), dtype=torch.int32, device=dev),
                out_d=torch.empty((nq, k), dtype=torch.float32, device=dev))}
        return self._buf[key]

    def knn_batch_device(self, q_dev, *, k, ef=None, mode=capi.MODE_PARITY, nq=None):
        """Queries on this rank's GPU -> (ids int32 [nq][k] global, distances float32 [nq][k]) torch tensors on the
        GPU, on every rank.  `q_dev` is the whole batch (float32 [nq][dim], the same on every rank) or, with `nq`
        given, only this rank's slice of it (`query_slice(nq)`; a torch tensor on the GPU or a pinned numpy array).
        Everything is enqueued on torch's current stream."""
        import torch
        sliced = nq is not None
        nq = q_dev.shape[0] if nq is None else nq
        b = self._buffers(nq, k)
        # torch's default stream has handle 0, which the C ABI reads as "the index's own stream, synchronous":
        # name the legacy default stream explicitly (cudaStreamLegacy = 1) so the call only enqueues, in order
        # with the caller's copies and with the barrier / merge that follow on the same stream
        stream = torch.cuda.current_stream().cuda_stream or 1
        if self.world == 1:
            self.local.search_device(q_dev.data_ptr(), nq, k, k if ef is None else ef, b["ids"].data_ptr(), b["d"].data_ptr(),
                                     stream=stream, mode=mode)
            return b["ids"], b["d"]
        import torch.distributed as dist
        if self._peer is not None:
            # fused exchange + merge (ShardTail): rows go to the group home's gather block, the last warp to arrive for
            # a query merges and stores the global row into every rank's block; the barrier ends the step
            pb, nxt = self._peer[self._step & 1], self._peer[(self._step + 1) & 1]
            self._step += 1
            if self.shard == 0:
                nxt["arrive_t"].zero_()            # counters of the NEXT call; ordered before this call's barrier
            lo, hi = self.query_slice(nq)
            q_base = q_dev.ctypes.data if isinstance(q_dev, np.ndarray) else q_dev.data_ptr()    # (a pinned host slice, see knn_batch_bigarray)
            q_ptr = q_base + (0 if sliced else lo * q_dev.shape[1] * 4)
            if hi > lo:
                capi.check(capi.lib().hnswb200_search_device_sharded(
                    self.local._h, q_ptr, hi - lo, k, k if ef is None else ef, mode, self.shard, self.S,
                    int(self.offsets[self.shard]), pb["g_ids"], pb["g_d"], pb["arrive"], self.world, pb["f_ids"], pb["f_d"], stream))
            pb["hdl"].barrier(channel=0)
            return pb["out_ids"], pb["out_d"]
        else:
            self.local.search_device(q_dev.data_ptr(), nq, k, k if ef is None else ef, b["ids"].data_ptr(), b["d"].data_ptr(),
                                     stream=stream, mode=mode)
            g = b["gathered"]
            dist.all_gather_into_tensor(g.view(self.world * 2, nq, k), b["packed"], group=self.group)
        capi.check(capi.lib().hnswb200_merge_topk_device(g.data_ptr(), g.data_ptr() + nq * k * 4, self.world, nq, k,
                                                         2 * nq * k, capi.ptr(self.offsets), b["out_ids"].data_ptr(),
                                                         b["out_d"].data_ptr(), stream))
        return b["out_ids"], b["out_d"]

    def knn_batch_bigarray(self, batch, *, k, ef=None, mode=capi.MODE_PARITY, out=None):
        """Ohnsw.knn_batch_bigarray with host buffers: H2D of the queries this rank's group answers, search with the
        fused exchange + merge, D2H of the merged rows."""
        import torch
        batch = capi.as_mat(batch, self.local.dim)
        if self.world == 1:
            return ohnsw.knn_batch_bigarray(self.local, batch, k=k, ef=ef, mode=mode, out=out)
        dev = torch.device("cuda", self.local.info().device)
        nq = batch.shape[0]
        if self._buffers(nq, k) and self._peer is not None:
            lo, hi = self.query_slice(nq)
            if capi.is_pinned(batch[lo:hi]):
                q_dev = batch[lo:hi]              # pinned: the kernel reads its slice of the queries from host memory, no copy
            else:
                q_dev = torch.from_numpy(batch[lo:hi]).to(dev, non_blocking=True)     # only the slice travels
            ids, d = self.knn_batch_device(q_dev, k=k, ef=ef, mode=mode, nq=nq)
        else:
            q_dev = torch.from_numpy(batch).to(dev, non_blocking=True)
            ids, d = self.knn_batch_device(q_dev, k=k, ef=ef, mode=mode)
        if out is None:
            return ids.cpu().numpy(), d.cpu().numpy()
        torch.from_numpy(out[0]).copy_(ids, non_blocking=True)
        torch.from_numpy(out[1]).copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
