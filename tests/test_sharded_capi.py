"""hnswb200_sharded_* (include/hnsw_b200.h): several row shards driven from ONE process through the C ABI,
the exchange and merge fused into the tail of the search kernel (csrc/search.cuh, ShardTail).

On a one-GPU box every shard is placed on device 0 — the kernels are the ones N GPUs run (stores into the
home gather block, system-scope arrival counters, last-arriver merge); no kernel waits on another, so
sharing a device is safe.  The expected rows are an exact host merge, by (distance, global id), of what the
oracle returns on each shard's exported graph."""
import ctypes as C

import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from ocaml_hnsw_b200.sharded import MultiGpuHgraph, shard_range
from oracle import oracle as O
from tests.util import draw_levels, uniform

pytestmark = pytest.mark.gpu


def _exact_merge(per_shard, offsets, k):
    """per_shard: [(ids [nq][k] local, dists [nq][k])] -> global rows ascending by (distance, id), -1 / NaN padded."""
    nq = per_shard[0][0].shape[0]
    gid = np.concatenate([np.where(i >= 0, i.astype(np.int64) + off, np.int64(1) << 40) for (i, _), off in zip(per_shard, offsets)], 1)
    gd = np.concatenate([np.where(i >= 0, d, np.float32(np.inf)) for i, d in per_shard], 1)
    order = np.lexsort((gid, gd), axis=1)[:, :k]
    ids = np.take_along_axis(gid, order, 1)
    d = np.take_along_axis(gd, order, 1)
    missing = ids >= (np.int64(1) << 40)
    return np.where(missing, -1, ids).astype(np.int32), np.where(missing, np.float32(np.nan), d).astype(np.float32)


def _oracle_per_shard(m, X, Q, k, ef, metric=O.METRIC_L2):
    out, offs = [], []
    for i in range(m.n_shards):
        h, first = m.shard(i)
        g = h.export_graph()
        o = O.VecOracle(X.shape[1], metric)
        o.import_graph(X[first:first + g.n], O.Graph(g.n, g.max_layer, g.entry, g.offsets, g.nbrs, g.levels))
        out.append(o.search(Q, k, ef))
        offs.append(first)
    return out, offs


@pytest.mark.parametrize("n_shards,n,dim,M", [(2, 6001, 128, 16), (3, 5000, 32, 8), (5, 4003, 100, 6), (8, 4000, 128, 16)])
def test_sharded_search_is_the_exact_merge(n_shards, n, dim, M):
    X = H.sift_like(n, dim, seed=11) if dim == 128 else uniform(n, dim, 12)
    Q = H.sift_like(257, dim, seed=13) if dim == 128 else uniform(257, dim, 14)
    lv = draw_levels(n, M)
    m = MultiGpuHgraph.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=M, num_nodes_search_construction=60,
                                            devices=[0] * n_shards, levels=lv)
    assert m.num_nodes() == n
    for i in range(n_shards):
        h, first = m.shard(i)
        lo, hi = shard_range(n, i, n_shards)
        assert (first, h.num_nodes()) == (lo, hi - lo)
        assert np.array_equal(h.export_graph().levels[1:], lv[lo + 1:hi])     # each shard draws from its slice of `levels`
    for k, ef in [(10, 10), (10, 48), (3, 100)]:
        per, offs = _oracle_per_shard(m, X, Q, k, ef)
        want_i, want_d = _exact_merge(per, offs, k)
        for rep in range(3):                                                 # consecutive calls: the arrival counters alternate
            ids, d = m.knn_batch_bigarray(Q, k=k, ef=ef)
            assert np.array_equal(ids, want_i), f"k={k} ef={ef} call {rep}"
            assert np.array_equal(d.view(np.uint32), want_d.view(np.uint32))
    st = m.stats()
    assert st.search_queries == 257 and st.build_inserts == n and st.search_n_dist > 0 and st.search_kernel_ms > 0
    gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
    ids, _ = m.knn_batch_bigarray(Q, k=10, ef=64)
    assert H.Recall.ids(gt, ids) > 0.9


def test_sharded_search_with_pinned_host_buffers():
    """hnswb200_sharded_search with pinned buffers: every shard's kernel reads the batch from host memory and the warp that
    merges a query stores its row into the caller's arrays — the same rows as with pageable buffers, call after call."""
    n, dim = 5000, 64
    X, lv = uniform(n, dim, 41), draw_levels(n, 8)
    m = MultiGpuHgraph.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=8, num_nodes_search_construction=50,
                                            devices=[0, 0, 0], levels=lv)
    Q = np.ascontiguousarray(uniform(300, dim, 42))
    out = (np.full((300, 10), -7, np.int32), np.full((300, 10), -7.0, np.float32))
    want = [m.knn_batch_bigarray(Q[::s].copy() if s == 1 else Q[::-1].copy(), k=10, ef=40) for s in (1, -1)]
    for b in (Q,) + out:
        capi.host_register(b)
    try:
        for rep in range(2):
            m.knn_batch_bigarray(Q, k=10, ef=40, out=out)
            assert np.array_equal(out[0], want[rep][0]) and np.array_equal(out[1].view(np.uint32), want[rep][1].view(np.uint32))
            Q[:] = Q[::-1].copy()                          # new contents in the same pinned buffer for the next call
    finally:
        for b in (Q,) + out:
            capi.host_unregister(b)


def test_sharded_device_buffers_and_streams():
    """hnswb200_sharded_search_device: queries and results in device memory, on a caller's stream, batch contents
    changing from call to call (the calls are ordered with the caller's copies)."""
    import torch
    n, dim = 5000, 64
    X, lv = uniform(n, dim, 21), draw_levels(n, 8)
    m = MultiGpuHgraph.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=8, num_nodes_search_construction=50,
                                            devices=[0, 0, 0], levels=lv)
    stream = torch.cuda.Stream()
    ids = torch.empty((300, 10), dtype=torch.int32, device="cuda")
    d = torch.empty((300, 10), dtype=torch.float32, device="cuda")
    for seed in (1, 2, 3):
        Q = uniform(300, dim, 30 + seed)
        want_i, want_d = m.knn_batch_bigarray(Q, k=10, ef=40)
        qp = torch.from_numpy(Q).pin_memory()
        with torch.cuda.stream(stream):
            q = qp.to("cuda", non_blocking=True)
            m.search_device(q.data_ptr(), 300, 10, 40, ids.data_ptr(), d.data_ptr(), stream=stream.cuda_stream)
            got_i, got_d = ids.cpu(), d.cpu()
        stream.synchronize()
        assert np.array_equal(got_i.numpy(), want_i) and np.array_equal(got_d.numpy().view(np.uint32), want_d.view(np.uint32))
    torch.cuda.synchronize()
    q = torch.from_numpy(Q).cuda()
    m.search_device(q.data_ptr(), 300, 10, 40, ids.data_ptr(), d.data_ptr())              # NULL stream: synchronous
    assert np.array_equal(ids.cpu().numpy(), want_i)


def test_sharded_padding_flavour_and_errors():
    """Fewer than k rows over all shards: -1 / NaN padding; the Hnsw.Ba flavour pads with +inf; argument errors."""
    X = uniform(9, 6, 1)                                    # dim 6: rows padded to 8 floats on the device
    m = MultiGpuHgraph.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=4, num_nodes_search_construction=10,
                                            devices=[0, 0, 0], levels=np.zeros(9, np.int32))
    Q = uniform(5, 6, 2)
    ids, d = m.knn_batch_bigarray(Q, k=12, ef=12)
    assert (ids[:, :9] >= 0).all() and (ids[:, 9:] == -1).all() and np.isnan(d[:, 9:]).all()
    assert all(sorted(r[:9].tolist()) == list(range(9)) for r in ids)                    # every row of every shard, global ids
    ref = np.sqrt(((X[None, :, :] - Q[:, None, :]) ** 2).sum(-1))
    assert np.allclose(np.sort(ref, 1), d[:, :9], rtol=1e-5)
    capi.check(capi.lib().hnswb200_sharded_set_flavour(m._s, capi.FLAVOUR_HNSW_BA))
    _, d = m.knn_batch_bigarray(Q, k=12, ef=12)
    assert np.isinf(d[:, 9:]).all()
    with pytest.raises(ValueError, match="not empty"):
        capi.check(capi.lib().hnswb200_sharded_build(m._s, capi.ptr(X), 9, None))
    e = MultiGpuHgraph(6, devices=[0, 0])
    with pytest.raises(ValueError, match="knn: empty hgraph"):                          # lib/ohnsw.ml:862
        e.knn_batch_bigarray(Q, k=3)
    with pytest.raises(ValueError, match="fewer rows than shards"):
        capi.check(capi.lib().hnswb200_sharded_build(e._s, capi.ptr(X), 1, None))
    s = C.c_void_p()
    assert capi.lib().hnswb200_sharded_create(C.byref(s), 8, 0, 16, 100, 0, 0, None) == capi.EINVAL
    assert capi.lib().hnswb200_sharded_create(C.byref(s), 8, 0, 16, 100, 0, 2, (C.c_int * 2)(0, 99)) != capi.OK


def test_search_device_sharded_tail_with_caller_buffers():
    """hnswb200_search_device_sharded — the entry point the one-process-per-GPU host uses with symmetric-memory
    buffers: here two independent indexes (two ranks' shards) and plain device buffers on one GPU, called one
    after the other; the merged rows land in BOTH final destinations."""
    import torch
    n, dim, k, ef = 4000, 48, 10, 40
    X, Q = uniform(n, dim, 5), uniform(123, dim, 6)
    lv = draw_levels(n, 8)
    hs = []
    for r in range(2):
        lo, hi = shard_range(n, r, 2)
        hs.append(Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X[lo:hi], num_connections=8, num_nodes_search_construction=50,
                                             levels=lv[lo:hi]))
    nq = len(Q)
    q = torch.from_numpy(Q).cuda()
    g_ids = torch.empty((2, nq, k), dtype=torch.int32, device="cuda")
    g_d = torch.empty((2, nq, k), dtype=torch.float32, device="cuda")
    arrive = torch.zeros(nq, dtype=torch.int32, device="cuda")
    fin = torch.full((2, 2, nq, k), -7, dtype=torch.int32, device="cuda")
    f_ids = (C.c_void_p * 2)(fin[0, 0].data_ptr(), fin[1, 0].data_ptr())
    f_d = (C.c_void_p * 2)(fin[0, 1].data_ptr(), fin[1, 1].data_ptr())
    torch.cuda.synchronize()
    for r in range(2):
        capi.check(capi.lib().hnswb200_search_device_sharded(hs[r]._h, q.data_ptr(), nq, k, ef, capi.MODE_PARITY, r, 2,
                                                             shard_range(n, r, 2)[0], g_ids.data_ptr(), g_d.data_ptr(),
                                                             arrive.data_ptr(), 2, f_ids, f_d, None))
    torch.cuda.synchronize()
    per = [Ohnsw.knn_batch_bigarray(h, Q, k=k, ef=ef) for h in hs]
    want_i, want_d = _exact_merge(per, [shard_range(n, r, 2)[0] for r in range(2)], k)
    got = fin.cpu().numpy()
    for r in range(2):
        assert np.array_equal(got[r, 0], want_i) and np.array_equal(got[r, 1].view(np.uint32), want_d.view(np.uint32))
    assert (arrive.cpu().numpy() == 2).all()
