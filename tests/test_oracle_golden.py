"""The oracle against every known-answer test lehy/ocaml-hnsw holds for the hot path.

Each test names the reference inline test (lib/ohnsw.ml line) it reproduces.  The fixture
tests/golden/ohnsw_inline_tests.json is a hand transcription of the reference's expect tests.
"""
import json
import os

import pytest

from oracle import oracle as O


@pytest.fixture(scope="module")
def golden(golden_dir):
    with open(os.path.join(golden_dir, "ohnsw_inline_tests.json")) as f:
        return json.load(f)


def _graph(case):
    h = O.AbsOracle(case["values"])
    if case.get("graph", "empty") == "ring":
        h.layer_create_loop(0)
    else:
        h.layer_create(0, len(case["values"]))
    return h


def test_search_one_golden(golden):       # TestSearchOne, ohnsw.ml:514-534
    for case in golden["search_one"]:
        h = _graph(case)
        assert h.search_one(0, case["start"], case["target"]) == case["expect"], case


def test_search_k_golden(golden):         # TestSearchK, ohnsw.ml:593-644
    for case in golden["search_k"]:
        h = _graph(case)
        got = h.search_k(0, [case["start"]], case["k"], case["target"])
        assert [n for n, _ in got] == [n for n, _ in case["expect"]], case
        for (_, d), (_, e) in zip(got, case["expect"]):
            assert d == pytest.approx(e, abs=1e-12), case   # sexp prints 1.1 for 3.1-2.


def test_select_neighbours_golden(golden):   # TestSelectNeighbours, ohnsw.ml:665-764
    for case in golden["select_neighbours"]:
        h = O.AbsOracle(case["values"])
        h.layer_create(0, len(case["values"]))
        got = h.select(case["target"], case["candidates"], case["n"])
        assert sorted(got) == case["expect"], case


# ---- containers ---------------------------------------------------------------------------------
def test_neighbours():                    # Neighbours.Test, ohnsw.ml:138-156
    a = O.NeighboursBox()
    assert a.length() == 0
    a.add(42); a.add(53)
    assert a.length() == 2 and a.list() == [53, 42]          # add prepends (:116-118)
    a.remove(42)
    assert a.length() == 1
    b = O.NeighboursBox(); b.add(42); b.add(53); b.remove(47)
    assert b.length() == 2 and b.list() == [42, 53]          # remove rebuilds reversed (:119-124)
    assert all(x in (42, 53) for x in b.list())
    assert not all(x == 42 for x in b.list())


def test_graph_symmetric_invariant():     # Graph.Test, ohnsw.ml:227-250
    h = O.AbsOracle([0.0] * 13)
    h.layer_create(0, 0)
    assert h.num_nodes() == 0 and h.graph_invariant(0)
    h.layer_create(0, 42)
    assert h.num_nodes() == 42
    h.graph_add_node(0)
    assert h.num_nodes() == 43 and h.graph_invariant(0)
    h.layer_create(0, 12)
    h.graph_add_node(0)
    h.set_connections(0, 1, [1, 2, 3, 11])
    for b in (1, 2, 3, 11):
        assert b in h.adjacent(0, 1)
    assert h.graph_invariant(0)
    h.set_connections(0, 1, [])
    assert h.graph_invariant(0) and 2 not in h.adjacent(0, 1) and 1 not in h.adjacent(0, 2)
    ring = O.AbsOracle([0., 1., 2., 3., 4.])
    ring.layer_create_loop(0)
    assert ring.graph_invariant(0)
    assert sorted(ring.adjacent(0, 0)) == [1, 4]


def test_visited():                       # Visited.Test, ohnsw.ml:270-296
    v = O.VisitedBox(42)
    assert v.card() == 0
    v.add(41); v.add(0)
    assert v.card() == 2 and v.mem(41) and v.mem(0)
    v = O.VisitedBox(42); v.add(0); v.clear()
    assert v.card() == 0 and not v.mem(0)
    v = O.VisitedBox(3)
    assert not v.mem(0) and not v.mem(1) and not v.mem(2)
    with pytest.raises(IndexError):
        v.mem(3)
    v.add(1); v.add(1)
    assert not v.mem(0) and v.mem(1) and not v.mem(2) and v.card() == 1
    v.clear()
    assert not v.mem(1) and v.card() == 0
    v = O.VisitedBox(3)                    # epoch overflow wrap, :285-295
    v.set_epoch_near_max(10)
    for _ in range(16):
        for _ in range(2):
            v.add(1); v.clear()
            assert not v.mem(1) and v.card() == 0


def test_hgraph_bookkeeping():            # Hgraph.Test, ohnsw.ml:366-399
    h = O.AbsOracle([1., 2., 3.])
    assert h.invariant() and h.num_nodes() == 0 and h.max_layer() == 0 and h.entry_point() is None
    added = h.add_node()
    h.set_entry_point(0)
    assert h.invariant() and added == 0 and h.entry_point() == 0 and h.num_nodes() == 1
    with pytest.raises(ValueError, match="Hgraph.set_entry_point: invalid node"):
        h.set_entry_point(1)
    assert h.layer_exists(0) and not h.layer_exists(1)
    h = O.AbsOracle([1., 2., 3.])
    a1 = h.add_node()
    h.set_max_layer(3)
    a2 = h.add_node()
    assert h.invariant() and (a1, a2) == (0, 1)
    assert all(h.layer_exists(i) for i in range(4)) and not h.layer_exists(4)
    assert h.num_nodes() == 2


def test_knn_empty_raises():              # ohnsw.ml:862
    h = O.AbsOracle([1., 2., 3.])
    with pytest.raises(ValueError, match="knn: empty hgraph"):
        h.knn(1.0, 2)


def test_insert_then_knn_1d():
    # 1-D end-to-end: 20 points on a line, every insert at level 0; exact answers are obvious.
    vals = [float(i) for i in range(20)]
    h = O.AbsOracle(vals)
    h.insert_all(M=3, efC=20, levels=[0] * 20)
    assert h.invariant()
    got = h.knn(7.2, 3)
    assert [n for n, _ in got] == [7, 8, 6]
    # Q2: a node whose level exceeds max_layer becomes entry point with no links up there
    h = O.AbsOracle(vals)
    h.insert_all(M=3, efC=20, levels=[0, 0, 2] + [0] * 17)
    assert h.max_layer() == 2 and h.entry_point() == 2
    assert h.adjacent(1, 2) == [] and h.adjacent(2, 2) == []
    assert [n for n, _ in h.knn(15.1, 2)] == [15, 16]


def test_oracle_reproduces_committed_fixture(golden_dir):
    """tests/golden/oracle_small.npz (made by tests/golden/make_oracle_fixture.py): the oracle's build
    and search on a small seeded data set, frozen."""
    import os
    import numpy as np
    from oracle import oracle as O
    f = np.load(os.path.join(golden_dir, "oracle_small.npz"))
    M, efC, k, ef = f["params"].tolist()
    o = O.VecOracle(f["X"].shape[1]).build(f["X"], M, efC, f["levels"])
    g = o.export()
    assert (g.entry, g.max_layer) == (int(f["entry"]), int(f["max_layer"]))
    for l in range(g.max_layer + 1):
        assert np.array_equal(g.offsets[l], f[f"offsets{l}"]) and np.array_equal(g.nbrs[l], f[f"nbrs{l}"])
    ids, d, cnt = o.search(f["Q"], k, ef, counters=True)
    assert np.array_equal(ids, f["ids"]) and np.array_equal(d.view(np.uint32), f["dists"].view(np.uint32))
    assert np.array_equal(cnt, f["counters"])
