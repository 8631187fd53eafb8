"""GPU index construction (Ohnsw.build_batch_bigarray / Ohnsw.insert, lib/ohnsw.ml:766-857) vs the
oracle.  A batched GPU build cannot reproduce the sequential build edge for edge, so parity is:

  * structural invariants the reference's own tests hold (symmetric links, Graph.Test.invariant
    lib/ohnsw.ml:217-225; degree bounds :818-823; levels and entry point :832-836),
  * search on the GPU-built graph is id-for-id what the oracle's search returns on the SAME graph
    (the graph is exported, loaded into the oracle, and both are queried),
  * recall@10 within 0.005 of an oracle-built index on identical inputs and levels
    (BASELINE.json north_star), and the reference's distance-threshold recall likewise,
  * determinism: two builds give the same graph bit for bit.

All calls go through the C ABI."""
import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Hnsw, Ohnsw, capi
from oracle import oracle as O
from tests.util import assert_same_results, draw_levels, grid36, uniform

pytestmark = pytest.mark.gpu


def _check_structure(g, levels, M, cap0=None):
    cap0 = 2 * M if cap0 is None else cap0
    n = g.n
    assert np.array_equal(g.levels, levels)
    top = int(levels.max())
    assert g.max_layer == top
    assert g.entry == int(np.argmax(levels == top))           # first node to reach the top layer (:832-836)
    for l in range(g.max_layer + 1):
        deg = g.degree(l)
        assert deg.max() <= (cap0 if l == 0 else M)
        assert (deg[levels < l] == 0).all()                   # rows only up to the node's level
        a = g.nbrs[l]
        assert len(a) == 0 or (a.min() >= 0 and a.max() < n)
        src = np.repeat(np.arange(n), deg)
        assert (a != src).all(), "self link"
        pairs = src.astype(np.int64) * n + a
        assert len(np.unique(pairs)) == len(pairs), "duplicate link"
        assert np.array_equal(np.sort(pairs), np.sort(a.astype(np.int64) * n + src)), f"layer {l} links are not symmetric"


def _to_oracle(X, g, metric=O.METRIC_L2):
    o = O.VecOracle(X.shape[1], metric)
    o.import_graph(X, O.Graph(g.n, g.max_layer, g.entry, g.offsets, g.nbrs, g.levels))
    return o


@pytest.fixture(scope="module")
def built20k():
    n, M, efC = 20000, 16, 100
    X = H.sift_like(n, 128, seed=1234)
    Q = H.sift_like(2000, 128, seed=4321)
    lv = draw_levels(n, M)
    h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=M, num_nodes_search_construction=efC, levels=lv)
    o = O.VecOracle(128).build(X, M, efC, lv)
    gt_ids, gt_d = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
    return X, Q, lv, h, o, gt_ids, gt_d


def test_structure_and_stats(built20k):
    X, Q, lv, h, o, _, _ = built20k
    g = h.export_graph()
    _check_structure(g, lv, 16)
    st = h.stats()
    assert st.build_inserts == len(X) and st.build_n_dist > 0 and st.build_seconds > 0
    assert st.layer_isolated[0] == 0
    # same shape of graph as the sequential build: mean layer-0 degree within 10 %
    go = o.export()
    dego = np.diff(go.offsets[0])
    assert abs(g.degree(0).mean() - dego.mean()) < 0.1 * dego.mean()


@pytest.mark.parametrize("ef", [10, 32, 64, 128])
def test_recall_matches_oracle_built_index(built20k, ef):
    X, Q, lv, h, o, gt_ids, gt_d = built20k
    ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
    ids_o, d_o = o.search_mt(Q, 10, ef)[:2]
    r_g, r_o = H.Recall.ids(gt_ids, ids_g), H.Recall.ids(gt_ids, ids_o)
    assert r_g >= r_o - 0.005, f"ef={ef}: GPU-built recall@10 {r_g:.4f} vs oracle-built {r_o:.4f}"
    # the reference's own recall definition (benchmark/dataset.ml:105-127)
    c_g, c_o = H.Recall.compute(gt_d, d_g, 1e-4), H.Recall.compute(gt_d, d_o, 1e-4)
    assert c_g >= c_o - 0.005


def test_search_on_gpu_built_graph_is_exact(built20k):
    X, Q, lv, h, _, _, _ = built20k
    o2 = _to_oracle(X, h.export_graph())
    assert o2.invariant()
    for k, ef in [(10, 10), (10, 64)]:
        ids_o, d_o, cnt_o = o2.search(Q[:500], k, ef, counters=True)
        ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q[:500], k=k, ef=ef)
        assert_same_results(ids_g, d_g, ids_o, d_o)
        assert np.array_equal(h.last_search_counters(500).astype(np.uint64), cnt_o)


def _same_graph(g, go):
    assert (g.n, g.max_layer, g.entry) == (go.n, go.max_layer, go.entry)
    assert np.array_equal(g.levels, go.levels)
    for l in range(g.max_layer + 1):
        assert np.array_equal(g.offsets[l], go.offsets[l]), f"layer {l}: degrees differ"
        bad = np.nonzero(g.nbrs[l] != go.nbrs[l])[0]
        assert bad.size == 0, f"layer {l}: {bad.size} adjacency slots differ (first at {bad[:3]})"


@pytest.mark.parametrize("kind,n,dim,M,efC", [("uniform", 1500, 32, 6, 30), ("sift", 1200, 128, 16, 60),
                                               ("ints", 1000, 8, 5, 25), ("uniform", 600, 100, 4, 40)])
def test_sequential_gpu_build_is_the_oracle_graph(kind, n, dim, M, efC):
    """build_batch = 1: inserts one at a time through the sequential link kernel — the GPU-built
    graph must be the oracle's (= the reference's, lib/ohnsw.ml:766-837) edge for edge, list order
    included, also on tie-heavy integer data."""
    if kind == "uniform":
        X = uniform(n, dim, 41)
    elif kind == "sift":
        X = H.sift_like(n, dim, seed=42)
    else:
        X = np.random.default_rng(43).integers(0, 5, (n, dim)).astype(np.float32)
        X = np.unique(X, axis=0)                       # duplicate vectors make the heuristic degenerate; keep ties only
        X = X[np.random.default_rng(44).permutation(len(X))]
        n = len(X)
    lv = draw_levels(n, M)
    lv[0] = 0
    o = O.VecOracle(dim).build(X, M, efC, lv)
    h = Ohnsw.Hgraph(dim, Ohnsw.distance_l2, M, efC)
    h.set_param("build_batch", 1)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), n, capi.ptr(lv)))
    _same_graph(h.export_graph(), o.export())
    # distance evaluations: the search part counts exactly like the oracle; selection evaluates
    # kept-vs-candidate distances eight at a time where the reference stops at the first failure
    assert h.stats().build_n_dist >= o.counters()[0] * 0.9


def test_sequential_build_then_inserts_is_the_oracle_graph():
    """Ohnsw.build_batch_bigarray on a prefix, then Ohnsw.insert one vector at a time and in a block
    (lib/ohnsw.ml:766-837), all in sequential mode: still the graph the oracle builds in one go."""
    n, dim, M, efC = 900, 32, 8, 40
    X = uniform(n, dim, 71)
    lv = draw_levels(n, M)
    lv[0] = 0
    o = O.VecOracle(dim).build(X, M, efC, lv)
    h = Ohnsw.Hgraph(dim, Ohnsw.distance_l2, M, efC)
    h.set_param("build_batch", 1)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X[:500]), 500, capi.ptr(lv[:500])))
    for i in range(500, 520):
        Ohnsw.insert(h, X[i], num_connections=M, num_nodes_search_construction=efC, levels=lv[i:i + 1])
    Ohnsw.insert(h, X[520:], levels=lv[520:])
    _same_graph(h.export_graph(), o.export())


def test_sequential_hnsw_ba_flavour_build_is_the_oracle_graph():
    """HNSW_BA flavour, build_batch = 1: ties accepted in the insert search, M links for a new node on
    layer 0 too, small candidate sets kept whole — against the oracle with the same three rules."""
    n, dim, M, efC = 1200, 24, 6, 30
    X = np.unique(np.random.default_rng(60).integers(0, 6, (n, dim)).astype(np.float32), axis=0)
    X = X[np.random.default_rng(61).permutation(len(X))]
    n = len(X)
    lv = draw_levels(n, M)
    lv[0] = 0
    o = O.VecOracle(dim).set_accept_ties(True).set_ba_build(True).build(X, M, efC, lv)
    h = Ohnsw.Hgraph(dim, Ohnsw.distance_l2, M, efC, flavour=capi.FLAVOUR_HNSW_BA)
    h.set_param("build_batch", 1)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), n, capi.ptr(lv)))
    g = h.export_graph()
    _same_graph(g, o.export())
    assert g.degree(0).max() <= 2 * M


def test_committed_fixture_build_and_search(golden_dir):
    """tests/golden/oracle_small.npz: the sequential GPU build reproduces the frozen graph, and the GPU
    search on it the frozen rows and work counters."""
    import os
    f = np.load(os.path.join(golden_dir, "oracle_small.npz"))
    M, efC, k, ef = f["params"].tolist()
    X, Q = f["X"], f["Q"]
    h = Ohnsw.Hgraph(X.shape[1], Ohnsw.distance_l2, M, efC)
    h.set_param("build_batch", 1)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), len(X), capi.ptr(np.ascontiguousarray(f["levels"], np.int32))))
    g = h.export_graph()
    assert (g.entry, g.max_layer) == (int(f["entry"]), int(f["max_layer"]))
    for l in range(g.max_layer + 1):
        assert np.array_equal(g.offsets[l], f[f"offsets{l}"]) and np.array_equal(g.nbrs[l], f[f"nbrs{l}"])
    ids, d = Ohnsw.knn_batch_bigarray(h, Q, k=k, ef=ef)
    assert np.array_equal(ids, f["ids"]) and np.array_equal(d.view(np.uint32), f["dists"].view(np.uint32))
    assert np.array_equal(h.last_search_counters(len(Q)).astype(np.uint64), f["counters"])


def test_build_is_deterministic():
    X = uniform(6000, 64, 5)
    lv = draw_levels(len(X), 8)
    gs = []
    for _ in range(2):
        h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=8, num_nodes_search_construction=50, levels=lv)
        gs.append(h.export_graph())
    for l in range(gs[0].max_layer + 1):
        assert np.array_equal(gs[0].offsets[l], gs[1].offsets[l]) and np.array_equal(gs[0].nbrs[l], gs[1].nbrs[l])


def test_insert_after_build(built20k):
    """Ohnsw.insert (lib/ohnsw.ml:766-837) on an existing index, one vector and a batch."""
    X, Q, lv, _, _, _, _ = built20k
    n0 = 5000
    M, efC = 16, 100
    h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X[:n0], num_connections=M, num_nodes_search_construction=efC, levels=lv[:n0])
    vis = Ohnsw.Visited.create(n0)
    Ohnsw.insert(h, X[n0], num_connections=M, num_nodes_search_construction=efC, visited=vis, levels=lv[n0:n0 + 1])
    assert h.num_nodes() == n0 + 1
    Ohnsw.insert(h, X[n0 + 1:8000], levels=lv[n0 + 1:8000])
    assert h.num_nodes() == 8000
    g = h.export_graph()
    _check_structure(g, lv[:8000], M)
    gt, _ = H.brute_force_knn_l2(X[:8000], Q[:500], 10, return_ids=True)
    ids, _ = Ohnsw.knn_batch_bigarray(h, Q[:500], k=10, ef=64)
    assert H.Recall.ids(gt, ids) > 0.95
    with pytest.raises(ValueError):
        Ohnsw.insert(h, X[0], num_connections=M + 1)


def test_insert_into_imported_graph(built20k):
    """A graph built by the reference/oracle, loaded into the GPU layout, keeps growing on the GPU."""
    X, Q, lv, _, _, _, _ = built20k
    n0 = 3000
    o = O.VecOracle(128).build(X[:n0], 16, 100, lv[:n0])
    h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 100).import_graph(X[:n0], o.export())
    Ohnsw.insert(h, X[n0:6000], levels=lv[n0:6000])
    g = h.export_graph()
    _check_structure(g, lv[:6000], 16)
    gt, _ = H.brute_force_knn_l2(X[:6000], Q[:300], 10, return_ids=True)
    ids, _ = Ohnsw.knn_batch_bigarray(h, Q[:300], k=10, ef=64)
    assert H.Recall.ids(gt, ids) > 0.95


@pytest.mark.parametrize("n,dim,M,efC,metric", [(1, 8, 4, 10, capi.L2), (2, 8, 4, 10, capi.L2), (37, 3, 3, 20, capi.L2),
                                                 (3000, 100, 24, 60, capi.ANGULAR), (1500, 960, 16, 60, capi.L2),
                                                 (4000, 96, 16, 40, capi.L2), (2500, 200, 6, 300, capi.L2)])
def test_other_shapes(n, dim, M, efC, metric):
    X = uniform(n, dim, 31)
    Q = uniform(50, dim, 32)
    if metric != capi.L2:
        X /= np.linalg.norm(X, axis=1, keepdims=True)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    lv = draw_levels(n, M)
    lv[0] = 0
    h = Ohnsw.build_batch_bigarray(metric, X, num_connections=M, num_nodes_search_construction=efC, levels=lv)
    g = h.export_graph()
    _check_structure(g, lv, M)
    o2 = _to_oracle(X, g, {capi.L2: O.METRIC_L2, capi.ANGULAR: O.METRIC_ANGULAR}[metric])
    k = min(10, n)
    ids_o, d_o = o2.search(Q, k, 40)
    ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q, k=k, ef=40)
    assert_same_results(ids_g, d_g, ids_o, d_o)
    if n >= 1000:
        gt, _ = O.bruteforce(X, Q, 10, {capi.L2: O.METRIC_L2, capi.ANGULAR: O.METRIC_ANGULAR}[metric])
        o = O.VecOracle(dim, {capi.L2: O.METRIC_L2, capi.ANGULAR: O.METRIC_ANGULAR}[metric]).build(X, M, efC, lv)
        ids_ref, _ = o.search(Q, 10, 40)
        assert H.Recall.ids(gt, ids_g) >= H.Recall.ids(gt, ids_ref) - 0.03     # 50 queries: loose


def test_grid36_build():
    """test/test.ml:88-127: the 36-point grid, M=3, efC=20, k=3."""
    X = grid36()
    lv = draw_levels(36, 3)
    lv[0] = 0
    h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=3, num_nodes_search_construction=20, levels=lv)
    _check_structure(h.export_graph(), lv, 3)
    ids, d = Ohnsw.knn_batch_bigarray(h, X, k=3)
    assert (ids[:, 0] == np.arange(36)).mean() > 0.9 and np.allclose(d[ids[:, 0] == np.arange(36), 0], 0)


def test_hnsw_ba_flavour_build():
    """Hnsw.Ba.build / knn_batch (lib/hnsw.ml:753-777): M links per new node, caps 2M / M,
    distances only, +inf padding."""
    X = H.sift_like(6000, 128, seed=1)
    Q = H.sift_like(200, 128, seed=2)
    lv = draw_levels(len(X), 16)
    h = Hnsw.Ba.build(X, num_neighbours=16, num_neighbours_build=100, levels=lv)
    g = h.export_graph()
    _check_structure(g, lv, 16)
    d = Hnsw.Ba.knn_batch(h, Q, num_neighbours_search=64, num_neighbours=10)
    gt = H.brute_force_knn_l2(X, Q, 10)
    assert H.Recall.compute(gt, d, 1e-4) > 0.95
    few = Hnsw.Ba.build(X[:4], num_neighbours=4, num_neighbours_build=10, levels=np.zeros(4, np.int32))
    d = Hnsw.Ba.knn_batch(few, Q[:3], num_neighbours_search=8, num_neighbours=8)
    assert np.isinf(d[:, 4:]).all() and np.isfinite(d[:, :4]).all()          # lib/hnsw.ml:770-771
    # Hnsw.Ba.knn (lib/hnsw.ml:763-767): (node, distance) pairs, nodes numbered from 1 (:313-325)
    one = Hnsw.Ba.knn(few, X[2], num_neighbours_search=8, num_neighbours=8)
    assert len(one) == 4 and one[0] == (3, 0.0) and sorted(n for n, _ in one) == [1, 2, 3, 4]
    assert [x for _, x in one] == sorted(x for _, x in one)


def test_build_argument_errors():
    h = Ohnsw.Hgraph(8, Ohnsw.distance_l2, 4, 10)
    X = uniform(10, 8, 1)
    with pytest.raises(ValueError, match="level"):
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), 10, capi.ptr(np.full(10, 99, np.int32))))
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), 10, None))
    with pytest.raises(ValueError, match="not empty"):
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), 10, None))


def test_benchmark_ml_literal_shape_sequential_build():
    """benchmark/benchmark.ml:115-128 as written (dim 784, M = 15, efConstruction = 400, uniform data), in
    sequential mode: the GPU-built graph is the oracle's edge for edge.  Exercises 30- / 15-slot rows and
    the bulk-copy staged gather inside the insert search, the selection heuristic and the link kernel.
    (N = 1500 of the default 5000 keeps 1500 one-insert launch chains within the test budget.)"""
    n, dim, M, efC = 1500, 784, 15, 400
    X = uniform(n, dim, 1234)
    lv = draw_levels(n, M)
    lv[0] = 0
    o = O.VecOracle(dim).build(X, M, efC, lv)
    h = Ohnsw.Hgraph(dim, Ohnsw.distance_l2, M, efC)
    h.set_param("build_batch", 1)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), n, capi.ptr(lv)))
    _same_graph(h.export_graph(), o.export())


def test_benchmark_ml_literal_shape_batched_build():
    """Same shape at the literal N = 5000, batched build: structure, search parity on the built graph, and
    recall@10 at the reference's own beam (ef = k = 10) and wider, against the oracle-built index."""
    n, dim, M, efC = 5000, 784, 15, 400
    X, Q = uniform(n, dim, 1234), uniform(200, dim, 4321)
    lv = draw_levels(n, M)
    lv[0] = 0
    h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=M, num_nodes_search_construction=efC, levels=lv)
    g = h.export_graph()
    _check_structure(g, lv, M)
    o2 = _to_oracle(X, g)
    o = O.VecOracle(dim).build(X, M, efC, lv)
    gt, _ = O.bruteforce(X, Q, 10)
    for ef in (10, 100):
        ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
        ids_2, d_2 = o2.search(Q, 10, ef)
        assert_same_results(ids_g, d_g, ids_2, d_2)
        ids_o, _ = o.search(Q, 10, ef)
        assert abs(H.Recall.ids(gt, ids_g) - H.Recall.ids(gt, ids_o)) <= 0.03      # 200 queries on uniform 784-d data: loose


@pytest.fixture(scope="module")
def built100k():
    """100k x 128 SIFT-like, M = 16, efConstruction = 200 (the bench's parameters at 1/10 of its rows): the GPU's
    batched build runs ~70 batches up to n/64 nodes — the regime where batch members not seeing each
    other matters — against the oracle's sequential build on identical inputs and levels."""
    n, M, efC = 100_000, 16, 200
    X = H.sift_like(n, 128, seed=1234)
    Q = H.sift_like(5000, 128, seed=4321)
    lv = draw_levels(n, M)
    h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=M, num_nodes_search_construction=efC, levels=lv)
    o = O.VecOracle(128).build(X, M, efC, lv)
    gt_ids, gt_d = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
    return X, Q, lv, h, o, gt_ids, gt_d


@pytest.mark.parametrize("ef", [16, 41, 128])
def test_100k_batched_build_recall_two_sided(built100k, ef):
    X, Q, lv, h, o, gt_ids, gt_d = built100k
    ids_g, d_g = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
    ids_o, d_o = o.search_mt(Q, 10, ef)[:2]
    r_g, r_o = H.Recall.ids(gt_ids, ids_g), H.Recall.ids(gt_ids, ids_o)
    assert abs(r_g - r_o) <= 0.005, f"ef={ef}: GPU-built recall@10 {r_g:.4f} vs oracle-built {r_o:.4f}"
    c_g, c_o = H.Recall.compute(gt_d, d_g, 1e-4), H.Recall.compute(gt_d, d_o, 1e-4)
    assert abs(c_g - c_o) <= 0.005


def test_100k_batched_build_drops_no_incoming_link(built100k):
    X, Q, lv, h, o, _, _ = built100k
    st = h.stats()
    assert st.build_dropped_incoming == 0       # rows that received more than LINK_MCAP new nodes in one batch
    _check_structure(h.export_graph(), lv, 16)


def test_gang_build_is_the_one_warp_build():
    """Small build batches run a gang of warps per insert (shared distance rounds).  Nothing an insert computes
    depends on the gang size, so the graph must be the graph of the one-warp-per-insert build, bit for bit."""
    X = H.sift_like(30000, 128, seed=77)
    lv = draw_levels(len(X), 16)
    gs = []
    for gang in (1, 0, 2):
        h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 100)
        h.set_param("gang", gang)
        capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), len(X), capi.ptr(lv)))
        gs.append(h.export_graph())
    for g in gs[1:]:
        _same_graph(g, gs[0])


def _build_with(X, lv, M, efC, **params):
    h = Ohnsw.Hgraph(X.shape[1], Ohnsw.distance_l2, M, efC)
    for k, v in params.items():
        h.set_param(k, v)
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), len(X), capi.ptr(lv)))
    return h


def test_build_kernel_variants_and_order_do_not_change_the_graph():
    """Phase 1 has two instances (the new node's vector in registers / in shared memory only) and takes a large
    batch longest-insert-first; every insert sees the same snapshot whichever warp runs it and whenever, so the
    graph must not depend on either."""
    X = H.sift_like(40000, 128, seed=78)
    lv = draw_levels(len(X), 16)
    gs = [_build_with(X, lv, 16, 100, build_qreg=q).export_graph() for q in (0, 1, 2)]
    for g in gs[1:]:
        _same_graph(g, gs[0])


def test_mates_pass_links_batch_members_and_keeps_the_invariants():
    """10 000 nodes enter a 20 000-node graph as ONE batch.  Without the mates pass members of a batch never link
    to each other and recall drops; with it (build.cuh) the links the sequential loop would have made between
    them are back.  The structural invariants hold either way."""
    n0, n, M, efC = 20000, 30000, 16, 100
    X = H.sift_like(n, 128, seed=79)
    Q = H.sift_like(2000, 128, seed=80)
    lv = draw_levels(n, M)
    lv[n0:] = np.minimum(lv[n0:], lv[:n0].max())          # a node that raises the top layer would close the batch
    gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
    rec, links = {}, {}
    for mates in (0, 1):
        h = _build_with(X[:n0], lv[:n0], M, efC)
        h.set_param("build_ratio", 1); h.set_param("build_ratio_early", 1); h.set_param("build_mates", mates)
        launches = h.stats().gpu_launches
        capi.check(capi.lib().hnswb200_insert(h._h, capi.ptr(X[n0:]), n - n0, capi.ptr(lv[n0:])))
        assert h.stats().gpu_launches - launches < 40      # one batch
        g = h.export_graph()
        _check_structure(g, lv, M)
        rec[mates] = H.Recall.ids(gt, Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=32)[0])
        src = np.repeat(np.arange(n), g.degree(0))
        links[mates] = int(((src >= n0) & (g.nbrs[0] >= n0)).sum())
    assert links[0] == 0 and links[1] > 0, links
    assert rec[1] > rec[0] + 0.003, rec


def test_default_schedule_uses_few_batches_and_matches_fine_batches():
    """Default schedule: a batch is at most 1/64 of the final graph and 1/4 of the graph so far (early rows are
    re-selected many times as the graph grows, so coarse early batches leave no trace).  Against batches of
    1/64 of the graph so far throughout — round 1's schedule — recall is the same and the launch count far lower."""
    n, M, efC = 50000, 16, 100
    X = H.sift_like(n, 128, seed=81)
    Q = H.sift_like(2000, 128, seed=82)
    lv = draw_levels(n, M)
    gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
    h_def = _build_with(X, lv, M, efC)
    h_fine = _build_with(X, lv, M, efC, build_ratio_early=64)
    assert h_def.stats().gpu_launches < h_fine.stats().gpu_launches / 2
    for ef in (16, 64):
        r_def = H.Recall.ids(gt, Ohnsw.knn_batch_bigarray(h_def, Q, k=10, ef=ef)[0])
        r_fine = H.Recall.ids(gt, Ohnsw.knn_batch_bigarray(h_fine, Q, k=10, ef=ef)[0])
        assert abs(r_def - r_fine) <= 0.005, (ef, r_def, r_fine)
