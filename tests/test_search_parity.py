"""GPU search vs the oracle on oracle-built graphs (the 'graph built by the reference, exported
to the GPU layout' check): ids, distances and work counters must be identical, bit for bit.

All calls go through the C ABI (ctypes) via the Ohnsw mirror."""
import json
import os

import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from oracle import oracle as O
from tests.util import assert_same_results, draw_levels, grid36, uniform

pytestmark = pytest.mark.gpu


def _oracle_index(X, M, efC, metric=O.METRIC_L2, seed=7):
    o = O.VecOracle(X.shape[1], metric)
    o.build(X, M, efC, draw_levels(len(X), M, seed))
    return o


def _gpu_from(o, X, M, efC, metric=capi.L2):
    h = Ohnsw.Hgraph(X.shape[1], metric, M, efC)
    h.import_graph(X, o.export())
    return h


def _check(o, h, Q, k, ef):
    ids_o, d_o, cnt_o = o.search(Q, k, ef, counters=True)
    ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q, k=k, ef=ef)
    assert_same_results(ids_g, d_g, ids_o, d_o)
    cnt_g = h.last_search_counters(len(Q))
    assert np.array_equal(cnt_g.astype(np.uint64), cnt_o), "work counters (n_dist, n_exp0, n_expU) differ"
    return ids_g, d_g


@pytest.fixture(scope="module")
def uni2k():
    X = uniform(2000, 128, 1234)
    Q = uniform(300, 128, 4321)
    o = _oracle_index(X, 16, 100)
    return X, Q, o, _gpu_from(o, X, 16, 100)


@pytest.mark.parametrize("k,ef", [(10, 10), (10, 50), (10, 200), (1, 1), (50, 50), (3, 33), (100, 512)])
def test_uniform_128(uni2k, k, ef):
    X, Q, o, h = uni2k
    _check(o, h, Q, k, ef)


def test_graph_roundtrip_through_gpu_layout(uni2k):
    X, Q, o, h = uni2k
    g0, g1 = o.export(), h.export_graph()
    assert (g1.n, g1.max_layer, g1.entry) == (g0.n, g0.max_layer, g0.entry)
    for l in range(g0.max_layer + 1):
        assert np.array_equal(g0.offsets[l], g1.offsets[l]) and np.array_equal(g0.nbrs[l], g1.nbrs[l])
    st = h.stats()
    assert st.num_layers == g0.max_layer + 1 and st.layer_nodes[0] == 2000
    assert st.layer_max_degree[0] <= 32 and st.layer_isolated[0] == 0


def test_visited_spill_is_exact(uni2k):
    """Tiny visited hash: every query outgrows shared memory and continues on the global bitset."""
    X, Q, o, _ = uni2k
    h = _gpu_from(o, X, 16, 100)
    h.set_param("hash_slots", 1024)
    _check(o, h, Q, 10, 200)
    assert h.stats().search_visited_overflows > 0


def test_grid36_fixture():
    """test/test.ml: 36-point 2-D grid, M=3, efC=20, k=3 (dim 2 -> padded rows)."""
    X = grid36()
    Q = uniform(10, 2, 5) * 3 + 2.5
    o = _oracle_index(X, 3, 20)
    h = _gpu_from(o, X, 3, 20)
    _check(o, h, Q, 3, 3)
    _check(o, h, Q, 3, 20)


def test_integer_data_with_exact_ties():
    """SIFT-like integer coordinates: squared distances are exact integers, ties are common;
    exercises the (distance, id) order and the evicted-tie list."""
    rng = np.random.default_rng(11)
    X = rng.integers(0, 4, (3000, 16)).astype(np.float32)      # many equal distances, many duplicates
    Q = rng.integers(0, 4, (200, 16)).astype(np.float32)
    o = _oracle_index(X, 8, 40)
    h = _gpu_from(o, X, 8, 40)
    for k, ef in [(10, 10), (10, 40), (5, 100)]:
        _check(o, h, Q, k, ef)


def test_sift_like_128():
    X = H.sift_like(4000, 128, seed=1234)
    Q = H.sift_like(200, 128, seed=4321)
    o = _oracle_index(X, 16, 100)
    h = _gpu_from(o, X, 16, 100)
    ids, _ = _check(o, h, Q, 10, 64)
    gt, _ = O.bruteforce(X, Q, 10)
    assert H.Recall.ids(gt, ids) > 0.9


@pytest.mark.parametrize("dim,M,metric", [(1, 4, capi.L2), (3, 4, capi.L2), (96, 16, capi.L2), (100, 24, capi.ANGULAR),
                                          (100, 24, capi.IP), (200, 8, capi.L2), (960, 16, capi.L2)])
def test_other_shapes(dim, M, metric):
    n = 1200 if dim < 900 else 600
    X = uniform(n, dim, 21)
    Q = uniform(64, dim, 22)
    if metric != capi.L2:
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    o = _oracle_index(X, M, 60, metric)
    h = _gpu_from(o, X, M, 60, metric)
    _check(o, h, Q, 10, 10)
    _check(o, h, Q, 10, 80)


def test_fewer_than_k_results_are_padded():
    X = uniform(7, 8, 1)
    o = _oracle_index(X, 4, 10)
    h = _gpu_from(o, X, 4, 10)
    ids, d = _check(o, h, uniform(5, 8, 2), 12, 12)
    assert (ids[:, 7:] == -1).all() and np.isnan(d[:, 7:]).all()        # lib/ohnsw.ml:879-881


def test_empty_hgraph_raises():
    h = Ohnsw.Hgraph(8)
    with pytest.raises(ValueError, match="knn: empty hgraph"):          # lib/ohnsw.ml:862
        Ohnsw.knn_batch_bigarray(h, np.zeros((1, 8), np.float32), k=1)
    with pytest.raises(ValueError, match="knn: empty hgraph"):
        Ohnsw.knn(h, Ohnsw.Visited.create(0), k=1, target=np.zeros(8, np.float32))


def test_reference_inline_goldens_on_gpu(golden_dir):
    """TestSearchK (lib/ohnsw.ml:593-644) run through the GPU: 1-D values as dim-1 vectors, the
    start node as entry point of a single-layer imported graph."""
    cases = json.load(open(os.path.join(golden_dir, "ohnsw_inline_tests.json")))["search_k"]
    for case in cases:
        vals = np.array(case["values"], np.float32)[:, None]
        n = len(vals)
        if case["graph"] == "ring":
            a = O.AbsOracle(case["values"]); a.layer_create_loop(0)
            rows = [a.adjacent(0, i) for i in range(n)]
        else:
            rows = [[] for _ in range(n)]
        offs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
        nbrs = np.array([x for r in rows for x in r], np.int32)
        g = H.FlatGraph(n, 0, case["start"], [offs], [nbrs])
        h = Ohnsw.Hgraph(1, capi.L2, 2, 10).import_graph(vals, g)
        k = case["k"]
        ids, d = Ohnsw.knn_batch_bigarray(h, np.array([[case["target"]]], np.float32), k=k)
        got = [(int(i), float(x)) for i, x in zip(ids[0], d[0]) if i >= 0]
        assert [i for i, _ in got] == [i for i, _ in case["expect"]], case
        assert np.allclose([x for _, x in got], [x for _, x in case["expect"]], rtol=1e-6, atol=1e-6)


def test_import_validation():
    h = Ohnsw.Hgraph(4, capi.L2, 4, 10)
    X = uniform(3, 4, 1)
    bad = H.FlatGraph(3, 0, 0, [np.array([0, 2, 2, 2])], [np.array([1, 1], np.int32)])
    with pytest.raises(ValueError, match="duplicate"):
        h.import_graph(X, bad)
    bad = H.FlatGraph(3, 0, 5, [np.array([0, 0, 0, 0])], [np.zeros(0, np.int32)])
    with pytest.raises(ValueError, match="invalid node"):                # lib/ohnsw.ml:343
        h.import_graph(X, bad)
    bad = H.FlatGraph(3, 0, 0, [np.array([0, 1, 1, 1])], [np.array([9], np.int32)])
    with pytest.raises(ValueError, match="out of range"):
        h.import_graph(X, bad)


def test_search_device_multi_writes_every_destination(uni2k):
    """hnswb200_search_device_multi (the fused multi-GPU exchange): the same result rows land in
    every destination block; here all destinations are local buffers."""
    import ctypes as C
    import torch
    X, Q, o, h = uni2k
    nq, k, ef = len(Q), 10, 40
    ids_o, d_o = o.search(Q, k, ef)
    q = torch.from_numpy(Q).cuda()
    blocks = torch.full((3, 2, nq, k), -7, dtype=torch.int32, device="cuda")
    ids_ptrs = (C.c_void_p * 3)(*[blocks[r, 0].data_ptr() for r in range(3)])
    d_ptrs = (C.c_void_p * 3)(*[blocks[r, 1].data_ptr() for r in range(3)])
    torch.cuda.synchronize()
    capi.check(capi.lib().hnswb200_search_device_multi(h._h, q.data_ptr(), nq, k, ef, capi.MODE_PARITY, 3, ids_ptrs, d_ptrs, None))
    torch.cuda.synchronize()
    got = blocks.cpu().numpy()
    for r in range(3):
        assert np.array_equal(got[r, 0], ids_o)
        assert np.array_equal(got[r, 1].view(np.uint32), d_o.view(np.uint32))
    with pytest.raises(ValueError, match="destinations"):
        capi.check(capi.lib().hnswb200_search_device_multi(h._h, q.data_ptr(), nq, k, ef, capi.MODE_PARITY, 9, ids_ptrs, d_ptrs, None))


def test_config_c1_benchmark_ml_shape():
    """BASELINE.json configs[0] / SURVEY.md C1 — the reference's own CPU-runnable case
    (benchmark/benchmark.ml shape): 10k x 128 uniform [-1,1), M=16, efConstruction=100, "~k:50, keep
    10" (Q4).  Search parity on the oracle-built graph, and the literal recipe's first 10 queries."""
    X, Q = uniform(10000, 128, 1234), uniform(1000, 128, 4321)
    o = _oracle_index(X, 16, 100)
    h = _gpu_from(o, X, 16, 100)
    ids, d = _check(o, h, Q, 10, 50)
    _check(o, h, Q[:10], 10, 10)                              # benchmark.ml: num_test = 10, ~k:10
    gt = O.bruteforce(X, Q, 10)[0]
    assert 0.5 < H.Recall.ids(gt, ids) < 0.9                  # ~0.67 at ef=50 on uniform data
    hb = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=100,
                                    levels=draw_levels(len(X), 16))
    ids_b, _ = Ohnsw.knn_batch_bigarray(hb, Q, k=10, ef=50)
    assert H.Recall.ids(gt, ids_b) >= H.Recall.ids(gt, ids) - 0.005


def test_heavy_duplicates_are_exact():
    """Hundreds of copies of a few vectors: far more than 32 evicted candidates tie at the beam's top
    distance.  The tie list continues in a global region (search.cuh), so the PARITY search still takes
    every decision the sequential reference loop takes: ids, distances and work counters equal the
    oracle's (both order equal keys by (distance, id); Core_kernel.Heap leaves that order unspecified)."""
    rng = np.random.default_rng(5)
    base = uniform(40, 16, 7)
    X = np.concatenate([np.repeat(base[:4], 150, axis=0), base[4:], uniform(500, 16, 8)]).astype(np.float32)
    X = X[rng.permutation(len(X))]
    Q = np.concatenate([base[:4] + 1e-3, uniform(20, 16, 9)]).astype(np.float32)
    o = _oracle_index(X, 8, 40)
    h = _gpu_from(o, X, 8, 40)
    for k, ef in [(10, 20), (10, 10), (5, 64)]:
        _check(o, h, Q, k, ef)
    # the list really left shared memory somewhere in this test's runs, and never overflowed its region
    spills = 0
    for ef in (10, 20, 64):
        Ohnsw.knn_batch_bigarray(h, Q, k=5, ef=ef)
        st = h.stats()
        spills += st.search_tie_spills
        assert st.search_tie_overflows == 0
    assert spills == 0          # duplicates are never linked to each other by the heuristic: few of them are visited


def test_tie_list_beyond_shared_memory_is_exact():
    """All 1 120 vectors of {-1,0,1}^8 with four non-zeros are at squared distance 4 from the origin, and at one
    of three distances from (0.5, 0, ..., 0) — exact in fp32.  Closer candidates then evict beam entries that
    tie with the new top one by one, and those stay poppable (strict stop rule, lib/ohnsw.ml:568): with
    ef = 128 far more than the 32 the shared-memory tie list holds.  The list continues in a global region,
    so ids, distances and work counters are still the oracle's, under both acceptance rules."""
    import itertools
    pts = []
    for nz in itertools.combinations(range(8), 4):
        for signs in itertools.product((-1.0, 1.0), repeat=4):
            v = np.zeros(8, np.float32)
            v[list(nz)] = signs
            pts.append(v)
    X = np.array(pts, np.float32)[np.random.default_rng(3).permutation(len(pts))]
    Q = np.zeros((12, 8), np.float32)
    Q[:, 0] = [0.5, -0.5, 0.25, 0.0, 0.5, 0.75, -0.25, 0.5, 0.0, 0.125, 0.5, -0.5]
    Q[4:8, 1] = 0.5
    Q[8:, 2] = [0.25, 0.5, -0.5, 0.125]
    o = _oracle_index(X, 12, 60)
    spills = 0
    h = _gpu_from(o, X, 12, 60)
    for k, ef in [(10, 128), (10, 64), (40, 200), (10, 10)]:
        _check(o, h, Q, k, ef)
        st = h.stats()
        spills += st.search_tie_spills
        assert st.search_tie_overflows == 0
    assert spills > 0, "path B rule: no query outgrew the 32-entry shared tie list"
    hb = Ohnsw.Hgraph(8, capi.L2, 12, 60, flavour=capi.FLAVOUR_HNSW_BA).import_graph(X, o.export())
    o.set_accept_ties(True)
    spills = 0
    for k, ef in [(10, 16), (10, 128), (5, 40)]:
        ids_o, d_o, cnt_o = o.search(Q, k, ef, counters=True)
        ids_g, d_g = Ohnsw.knn_batch_bigarray(hb, Q, k=k, ef=ef)
        assert_same_results(ids_g, d_g, ids_o, np.where(ids_o < 0, np.float32(np.inf), d_o))
        assert np.array_equal(hb.last_search_counters(len(Q)).astype(np.uint64), cnt_o)
        st = hb.stats()
        spills += st.search_tie_spills
        assert st.search_tie_overflows == 0
    assert spills > 0, "accept-ties rule: no query outgrew the 32-entry shared tie list"


def test_device_queries_with_padded_rows():
    """Device-resident queries are dense [nq][dim] even when the index pads its rows (dim = 6 -> 8
    floats, or an explicit row_floats): same rows as the host-buffer call."""
    import torch
    X, Q = uniform(800, 6, 51), uniform(40, 6, 52)
    lv = draw_levels(len(X), 4)
    for rf in (0, 32):
        h = Ohnsw.Hgraph(6, capi.L2, 4, 30)
        if rf:
            h.set_param("row_floats", rf)
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), len(X), capi.ptr(lv)))
        ids_h, d_h = Ohnsw.knn_batch_bigarray(h, Q, k=5, ef=20)
        q = torch.from_numpy(Q).cuda()
        ids = torch.empty((40, 5), dtype=torch.int32, device="cuda")
        d = torch.empty((40, 5), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        h.search_device(q.data_ptr(), 40, 5, 20, ids.data_ptr(), d.data_ptr())
        assert np.array_equal(ids.cpu().numpy(), ids_h) and np.array_equal(d.cpu().numpy().view(np.uint32), d_h.view(np.uint32))
    with pytest.raises(ValueError, match="row_floats"):
        h.set_param("row_floats", 64)                      # not on a built index


@pytest.mark.parametrize("dim,rf,metric", [(100, 128, capi.ANGULAR), (100, 128, capi.L2), (40, 128, capi.L2), (6, 32, capi.L2),
                                           (300, 384, capi.L2)])
def test_wider_row_stride_is_invisible(dim, rf, metric):
    """row_floats: rows stored at a wider stride (128-byte aligned 400-byte rows for the GloVe shape, config C3).  The
    padding is never read and never counted: ids, distances and work counters stay the oracle's, for a search on the
    oracle's graph and for a sequential build (edge for edge)."""
    X, Q = uniform(1000, dim, 61), uniform(64, dim, 62)
    if metric != capi.L2:
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    M, efC = 12, 50
    lv = draw_levels(len(X), M, 7)
    lv[0] = 0
    o = O.VecOracle(dim, metric).build(X, M, efC, lv)
    h = Ohnsw.Hgraph(dim, metric, M, efC)
    h.set_param("row_floats", rf)
    h.import_graph(X, o.export())
    _check(o, h, Q, 10, 10)
    _check(o, h, Q, 10, 70)
    hb = Ohnsw.Hgraph(dim, metric, M, efC)
    hb.set_param("row_floats", rf)
    hb.set_param("build_batch", 1)
    capi.check(capi.lib().hnswb200_build(hb._h, capi.ptr(X), len(X), capi.ptr(lv)))
    g, go = hb.export_graph(), o.export()
    assert (g.n, g.max_layer, g.entry) == (go.n, go.max_layer, go.entry)
    for l in range(go.max_layer + 1):
        assert np.array_equal(g.offsets[l], go.offsets[l]) and np.array_equal(g.nbrs[l], go.nbrs[l])


@pytest.mark.parametrize("dim", [64, 6])
def test_pinned_host_buffers_are_read_and_written_in_place(dim):
    """hnswb200_search with PINNED buffers (hnswb200_host_register): the kernel reads the queries from the caller's
    buffer and stores the rows into it, no copy — same ids, distances and counters as with pageable buffers (and as
    the oracle), `search_zero_copy` says which path ran; rows the index pads (dim = 6) still take the copy."""
    X, Q = uniform(1500, dim, 81), np.ascontiguousarray(uniform(200, dim, 82))
    o = _oracle_index(X, 8, 50)
    h = _gpu_from(o, X, 8, 50)
    ids_o, d_o, cnt_o = o.search(Q, 10, 40, counters=True)
    ids_p, d_p = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=40)                    # pageable
    assert h.stats().search_zero_copy == 0
    out = (np.full((200, 10), -7, np.int32), np.full((200, 10), -7.0, np.float32))
    for b in (Q,) + out:
        capi.host_register(b)
    try:
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=40, out=out)
        assert h.stats().search_zero_copy == (3 if dim % 4 == 0 else 2)
        assert_same_results(out[0], out[1], ids_o, d_o)
        assert np.array_equal(out[0], ids_p) and np.array_equal(out[1].view(np.uint32), d_p.view(np.uint32))
        assert np.array_equal(h.last_search_counters(200).astype(np.uint64), cnt_o)
        Q[:] = Q[::-1].copy()                                                   # new contents, same pinned buffer
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=40, out=out)
        assert np.array_equal(out[0], ids_p[::-1])
        # the device-pointer entry points take a pinned host buffer for the queries too (read in place by the kernel)
        import torch
        ids_t = torch.empty((200, 10), dtype=torch.int32, device="cuda")
        d_t = torch.empty((200, 10), dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        h.search_device(Q.ctypes.data, 200, 10, 40, ids_t.data_ptr(), d_t.data_ptr())
        assert np.array_equal(ids_t.cpu().numpy(), ids_p[::-1])
        h.set_param("host_zero_copy", 0)
        out[0][:] = -7
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=40, out=out)
        assert h.stats().search_zero_copy == 0 and np.array_equal(out[0], ids_p[::-1])
    finally:
        for b in (Q,) + out:
            capi.host_unregister(b)


def test_hnsw_ba_acceptance_rule_on_tie_heavy_data():
    """HNSW_BA flavour: candidates that TIE with the beam's maximum are accepted (lib/hnsw.ml:494-506).
    Integer data makes ties the common case; ids, distances and counters must equal the oracle run
    with the same rule, and differ from the path-B rule somewhere (the rule is really exercised)."""
    rng = np.random.default_rng(12)
    X = np.unique(rng.integers(0, 4, (4000, 12)).astype(np.float32), axis=0)
    X = X[rng.permutation(len(X))]
    Q = rng.integers(0, 4, (300, 12)).astype(np.float32)
    o = _oracle_index(X, 8, 40)
    h = Ohnsw.Hgraph(12, capi.L2, 8, 40, flavour=capi.FLAVOUR_HNSW_BA).import_graph(X, o.export())
    ids_b, d_b = o.search(Q, 10, 30)
    o.set_accept_ties(True)
    differs = False
    for k, ef in [(10, 10), (10, 30), (5, 64)]:
        ids_o, d_o, cnt_o = o.search(Q, k, ef, counters=True)
        ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q, k=k, ef=ef)
        d_o = np.where(ids_o < 0, np.float32(np.inf), d_o)               # Hnsw.Ba pads with +inf (lib/hnsw.ml:770)
        assert_same_results(ids_g, d_g, ids_o, d_o)
        assert np.array_equal(h.last_search_counters(len(Q)).astype(np.uint64), cnt_o), f"k={k} ef={ef}: work counters differ"
        assert h.stats().search_tie_overflows == 0
        if (k, ef) == (10, 30):
            differs = not np.array_equal(ids_o, ids_b)
    assert differs, "the two acceptance rules gave identical results: the test data has no ties at the boundary"


# ---- the reference's literal benchmark shape, and the bulk-copy staged gather ---------------------
@pytest.fixture(scope="module")
def bench_ml():
    """benchmark/benchmark.ml:115-128 as written: N = 5000 (argv default), dim = 784, M = 15,
    efConstruction = 400, uniform [-1, 1) (dataset.ml:47-49).  M = 15 gives 30- / 15-slot adjacency rows
    (120 / 60 bytes: not line aligned, not a multiple of 4 slots) and 3 136-byte vector rows, which
    take the bulk-copy staged gather."""
    X, Q = uniform(5000, 784, 1234), uniform(200, 784, 4321)
    lv = draw_levels(5000, 15)
    o = O.VecOracle(784).build(X, 15, 400, lv)
    return X, Q, lv, o


def test_benchmark_ml_literal_shape_search(bench_ml):
    X, Q, lv, o = bench_ml
    h = _gpu_from(o, X, 15, 400)
    inf = h.info()
    assert (inf.slots0, inf.slots_upper) == (30, 15)
    _check(o, h, Q[:10], 10, 10)                  # num_test = 10, ~k:10 (benchmark.ml:118,127; beam = k, lib/ohnsw.ml:873)
    _check(o, h, Q, 10, 10)
    _check(o, h, Q, 10, 100)
    _check(o, h, Q, 50, 400)


@pytest.mark.parametrize("dim,M", [(256, 16), (300, 5), (784, 15), (960, 16), (2048, 8)])
def test_staged_gather_equals_ldg_gather(dim, M):
    """Rows of >= 1 KB are fetched with cp.async.bulk into a per-warp shared-memory ring (UBLKCP); the
    result must be what the per-lane LDG gather gives (stage_rows = -1) and what the oracle gives,
    whatever the ring depth."""
    n = 700
    X, Q = uniform(n, dim, 61), uniform(48, dim, 62)
    o = _oracle_index(X, M, 50)
    rows = {}
    for groups in (-1, 4, 6, 9, 16):
        h = _gpu_from(o, X, M, 50)
        h.set_param("stage_rows", groups)
        for vm in (1, 2):
            h.set_param("visited_mode", vm)
            rows[(groups, vm)] = _check(o, h, Q, 10, 64)
    base = rows[(-1, 1)]
    for key, r in rows.items():
        assert_same_results(r[0], r[1], base[0], base[1])


@pytest.mark.parametrize("hash_bits", [32, 16, -1])
def test_visited_hash_formats_are_exact(uni2k, hash_bits):
    """The shared-memory visited hash holds 32-bit ids or 16-bit quotiented entries (quotient + displacement,
    half the memory); -1 caps the displacement at one slot so that probe sequences do run out and the set
    moves to the global bitset mid-expansion.  Every variant must give the oracle's rows and counters."""
    X, Q, o, _ = uni2k
    h = _gpu_from(o, X, 16, 100)
    h.set_param("visited_mode", 1)
    h.set_param("hash_bits", hash_bits)
    if hash_bits == -1:
        h.set_param("hash_slots", 1024)                  # 2 000 ids over 1 024 slots: neighbours in the table do collide
        _check(o, h, Q, 10, 10)                          # ~400 visited nodes: far below the load bound of 768
        assert h.stats().search_visited_overflows > 0    # so every spill here is a probe that ran out of reach
        _check(o, h, Q, 10, 64)
        return
    for k, ef in [(10, 10), (10, 64), (10, 200)]:
        _check(o, h, Q, k, ef)
    h.set_param("hash_slots", 1024)                      # and with a table every query outgrows
    _check(o, h, Q, 10, 200)
    assert h.stats().search_visited_overflows > 0


def test_unknown_mode_is_rejected(uni2k):
    X, Q, o, h = uni2k
    with pytest.raises(ValueError, match="unknown mode"):
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=10, mode=1)


@pytest.mark.parametrize("gang", [2, 4])
def test_gang_of_warps_per_query_is_exact(uni2k, gang):
    """Several warps per query (search.cuh, Gang): warp 0 runs the traversal, the distance rounds of every
    expansion are shared among the gang.  Rows and work counters must be the oracle's, whatever the gang."""
    X, Q, o, _ = uni2k
    h = _gpu_from(o, X, 16, 100)
    h.set_param("gang", gang)
    for k, ef in [(10, 10), (10, 50), (10, 200), (100, 512)]:
        _check(o, h, Q, k, ef)
    _check(o, h, Q[:1], 10, 64)                           # one query (Ohnsw.knn): the automatic choice is a gang too
    h.set_param("hash_slots", 1024)                       # visited set spilling under a gang
    _check(o, h, Q, 10, 200)
    # tie-heavy integer data and 48-slot rows (two 32-slot chunks per expansion)
    rng = np.random.default_rng(11)
    Xi = rng.integers(0, 4, (3000, 16)).astype(np.float32)
    Qi = rng.integers(0, 4, (200, 16)).astype(np.float32)
    oi = _oracle_index(Xi, 24, 40)
    hi = _gpu_from(oi, Xi, 24, 40)
    hi.set_param("gang", gang)
    for k, ef in [(10, 10), (10, 40), (5, 100)]:
        _check(oi, hi, Qi, k, ef)


def test_small_batches_pick_a_gang_automatically(uni2k):
    X, Q, o, h = uni2k
    for nq in (1, 7, 300):
        _check(o, h, Q[:nq], 10, 32)
