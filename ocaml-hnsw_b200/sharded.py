"""Multi-GPU layer: one process per GPU, the dataset split into contiguous row shards.

The reference has nothing distributed; this is the sharding BASELINE.json's
north_star prescribes on top of the same `Ohnsw` entry points: every rank builds and searches
the HNSW sub-graph of its own rows `[lo, hi)`, the query batch is the same on every rank, and
the per-shard `[nq][k]` result rows are exchanged with ONE all-gather (NCCL over NVLink) and
merged by `hnswb200_merge_topk_device` (shard-local ids become global ids inside the merge).
No collective runs during graph traversal or during the build.

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

    def __init__(self, local, n_total, rank, world, group=None):
        self.local, self.n_total, self.rank, self.world, self.group = local, int(n_total), rank, world, group
        self.offsets = shard_offsets(n_total, world)
        self._buf = {}

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

    def _buffers(self, nq, k):
        import torch
        key = (nq, k)
        if key not in self._buf:
            dev = torch.device("cuda", self.local.info().device)
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
        stream = torch.cuda.current_stream().cuda_stream
        self.local.search_device(q_dev.data_ptr(), nq, k, k if ef is None else ef, b["ids"].data_ptr(), b["d"].data_ptr(),
                                 stream=stream, mode=mode)
        if self.world == 1:
            return b["ids"], b["d"]
        import torch.distributed as dist
        g = b["gathered"]
        dist.all_gather_into_tensor(g.view(self.world * 2, nq, k), b["packed"], group=self.group)
        capi.check(capi.lib().hnswb200_merge_topk_device(g.data_ptr(), g.data_ptr() + nq * k * 4, self.world, nq, k,
                                                         2 * nq * k, capi.ptr(self.offsets), b["out_ids"].data_ptr(),
                                                         b["out_d"].data_ptr(), stream or None))
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
