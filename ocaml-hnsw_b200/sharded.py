"""Multi-GPU layer: one process per GPU, the dataset split into contiguous row shards.

The reference has nothing distributed; this is the sharding BASELINE.json's
north_star prescribes on top of the same `Ohnsw` entry points: every rank builds and searches
the HNSW sub-graph of its own rows `[lo, hi)`, the query batch is the same on every rank, and
the per-shard `[nq][k]` result rows are exchanged and merged by `hnswb200_merge_topk_device`
(shard-local ids become global ids inside the merge).  The exchange is fused into the search
kernel: every rank's warps store their finished rows straight into every peer's gather buffer
(torch symmetric memory = peer-mapped HBM over NVLink / NVSwitch), so what follows the search is one
device-side barrier and the merge — no all-gather.  Where symmetric memory is unavailable the
exchange is ONE packed NCCL all-gather.  No collective runs during graph traversal or the build.

torch is plumbing here: device buffers for the exchange and `torch.distributed`.
"""
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


class ShardedHgraph:
    """`Ohnsw.Hgraph` over a row-sharded dataset.  `n_total` rows exist across the group; this rank
    owns `shard_range(n_total, rank, world)`."""

    def __init__(self, local, n_total, rank, world, group=None, peer_exchange=True):
        self.local, self.n_total, self.rank, self.world, self.group = local, int(n_total), rank, world, group
        self.offsets = shard_offsets(n_total, world)
        self._buf = {}
        self.peer_exchange = peer_exchange and world > 1     # False: packed NCCL all-gather
        self.exchange = "none" if world == 1 else None       # set on first use: "peer-store" | "nccl-all-gather"
        self._step = 0

    @staticmethod
    def build(distance, local_rows, n_total, *, num_connections, num_nodes_search_construction, rank=0, world=1,
              group=None, levels=None, seed=0, device=0, params=None):
        """Ohnsw.build_batch_bigarray on this rank's rows."""
        local_rows = capi.as_mat(local_rows)
        lo, hi = shard_range(n_total, rank, world)
        if local_rows.shape[0] != hi - lo:
            raise ValueError(f"rank {rank}: expected rows [{lo}, {hi}) of the dataset, got {local_rows.shape[0]} rows")
        h = ohnsw.Hgraph(local_rows.shape[1], distance, num_connections, num_nodes_search_construction, seed + rank, device)
        for name, v in (params or {}).items():
            h.set_param(name, v)
        lv = None if levels is None else np.ascontiguousarray(levels, np.int32)
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(local_rows), local_rows.shape[0], capi.ptr(lv)))
        return ShardedHgraph(h, n_total, rank, world, group)

    def _peer_buffers(self, nq, k, dev):
        """Two gather buffers [world][2][nq][k] in symmetric memory (double buffered: a rank may start
        writing step i+2 only after the barrier of step i+1, by which every rank has merged step i)."""
        import ctypes as C
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        group = self.group if self.group is not None else dist.group.WORLD
        out = []
        for _ in range(2):
            t = symm_mem.empty((self.world, 2, nq, k), dtype=torch.int32, device=dev)
            hdl = symm_mem.rendezvous(t, group)
            block = 2 * nq * k * 4
            ids_ptrs = (C.c_void_p * self.world)(*[int(hdl.buffer_ptrs[r]) + self.rank * block for r in range(self.world)])
            d_ptrs = (C.c_void_p * self.world)(*[int(hdl.buffer_ptrs[r]) + self.rank * block + nq * k * 4 for r in range(self.world)])
            out.append(dict(t=t, hdl=hdl, ids_ptrs=ids_ptrs, d_ptrs=d_ptrs))
        return out

    def _buffers(self, nq, k):
        import torch
        key = (nq, k)
        if key not in self._buf:
            dev = torch.device("cuda", self.local.info().device)
            peer = None
            if self.peer_exchange and self.world <= 8:
                try:
                    peer = self._peer_buffers(nq, k, dev)
                    self.exchange = "peer-store"
                except Exception as e:                      # no symmetric memory on this system / build
                    self.exchange = f"nccl-all-gather (symmetric memory unavailable: {type(e).__name__})"
            elif self.world > 1:
                self.exchange = "nccl-all-gather"
            self._peer = peer
            # one packed block per rank: [0] = ids (int32), [1] = distances (fp32 bits) -> ONE all-gather
            packed = torch.empty((2, nq, k), dtype=torch.int32, device=dev)
            self._buf = {key: dict(
                packed=packed, ids=packed[0], d=packed[1].view(torch.float32),
                gathered=torch.empty((self.world, 2, nq, k), dtype=torch.int32, device=dev),
                out_ids=torch.empty((nq, k), dtype=torch.int32, device=dev),
                out_d=torch.empty((nq, k), dtype=torch.float32, device=dev))}
        return self._buf[key]

    def knn_batch_device(self, q_dev, *, k, ef=None, mode=capi.MODE_PARITY):
        """Queries already on this rank's GPU (torch float32 [nq][dim], the same on every rank) ->
        (ids int32 [nq][k] global, distances float32 [nq][k]) torch tensors on the GPU, on every rank.
        Everything is enqueued on torch's current stream."""
        import torch
        nq = q_dev.shape[0]
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
            # fused exchange: the search kernel stores every finished row into all peers' buffers
            pb = self._peer[self._step & 1]
            self._step += 1
            capi.check(capi.lib().hnswb200_search_device_multi(self.local._h, q_dev.data_ptr(), nq, k, k if ef is None else ef,
                                                               mode, self.world, pb["ids_ptrs"], pb["d_ptrs"], stream))
            pb["hdl"].barrier(channel=0)
            g = pb["t"]
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
        """Ohnsw.knn_batch_bigarray with host buffers: H2D of the queries, per-shard search,
        all-gather + merge, D2H of the merged rows."""
        import torch
        batch = capi.as_mat(batch, self.local.dim)
        if self.world == 1:
            return ohnsw.knn_batch_bigarray(self.local, batch, k=k, ef=ef, mode=mode, out=out)
        dev = torch.device("cuda", self.local.info().device)
        q_dev = torch.from_numpy(batch).to(dev, non_blocking=True)
        ids, d = self.knn_batch_device(q_dev, k=k, ef=ef, mode=mode)
        if out is None:
            return ids.cpu().numpy(), d.cpu().numpy()
        torch.from_numpy(out[0]).copy_(ids, non_blocking=True)
        torch.from_numpy(out[1]).copy_(d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return out
