"""Structural properties of the oracle's insert / build / knn — the part of the path the reference's own tests
leave unpinned (`insert`, `build_batch_bigarray`, `knn`, `knn_batch_bigarray` have no
golden vectors).  Each property is a statement the reference's code makes true by construction; the line that
makes it true is cited.  These run on the CPU and keep the checker honest between GPU runs.
"""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle as O
from tests.util import draw_levels, uniform


def _build(n, dim, M, efC, seed, ba=False, levels=None):
    X = uniform(n, dim, seed)
    lv = draw_levels(n, M, seed + 1) if levels is None else levels
    o = O.VecOracle(dim)
    if ba:
        o.set_accept_ties().set_ba_build()
    o.build(X, M, efC, lv)
    lv = lv.copy()
    lv[0] = 0                  # the first node draws no level: it is put on layer 0 (lib/ohnsw.ml:773-778)
    return X, lv, o, o.export()


def _rows(g, layer):
    return [g.row(layer, i).tolist() for i in range(g.n)]


@settings(max_examples=25, deadline=None)
@given(n=st.integers(1, 160), dim=st.sampled_from([1, 2, 7, 16]), M=st.sampled_from([2, 4, 8]),
       efC=st.sampled_from([1, 4, 20]), seed=st.integers(0, 10_000), ba=st.booleans())
def test_built_graph_is_symmetric_bounded_and_layered(n, dim, M, efC, seed, ba):
    X, lv, o, g = _build(n, dim, M, efC, seed, ba)
    assert o.invariant()                                     # Graph.invariant, lib/ohnsw.ml:204-225
    assert g.n == n and np.array_equal(g.levels, lv)
    # max layer / entry point: the first node of the highest level drawn so far (lib/ohnsw.ml:832-836)
    assert g.max_layer == lv.max()
    assert g.entry == int(np.argmax(lv == lv.max()))
    for l in range(g.max_layer + 1):
        rows = _rows(g, l)
        cap = 2 * M if l == 0 else M                         # lib/ohnsw.ml:818-823 (path B); lib/hnsw.ml:753-758 (Ba)
        for i, r in enumerate(rows):
            assert len(r) <= cap
            assert len(set(r)) == len(r) and i not in r      # no duplicates, no self link
            if lv[i] < l:
                assert r == []                               # a node has links only up to its level (:806-830)
            for j in r:
                assert i in rows[j]                          # symmetric (Q7, lib/ohnsw.ml:182-196)
                assert lv[j] >= l


@settings(max_examples=20, deadline=None)
@given(n=st.integers(2, 120), dim=st.sampled_from([2, 8]), seed=st.integers(0, 10_000), k=st.integers(1, 12),
       ef=st.integers(1, 40))
def test_knn_rows_are_ascending_unique_and_padded(n, dim, seed, k, ef):
    X, lv, o, g = _build(n, dim, 4, 16, seed)
    Q = uniform(9, dim, seed + 5)
    ef = max(ef, k)
    ids, d = o.search(Q, k, ef)
    for row_i, row_d in zip(ids, d):
        m = int((row_i >= 0).sum())
        assert (row_i[:m] >= 0).all() and (row_i[m:] == -1).all() and np.isnan(row_d[m:]).all()   # Q10, :879-897
        assert len(set(row_i[:m].tolist())) == m
        assert np.all(np.diff(row_d[:m]) >= 0)               # results pop in ascending order (:870-875)
    # the distance reported is the distance to that row (lib/ohnsw.ml:899), fp32-rounded
    for qi in range(len(Q)):
        for i, x in zip(ids[qi], d[qi]):
            if i >= 0:
                assert np.float32(o.distance(X[i], Q[qi])) == x
    # a wider beam never returns a worse k-th neighbour on the same graph... is NOT a property of HNSW in
    # general; what is: ef = k keeps exactly the first k rows of the ef = k beam (Q4)
    a, _ = o.search(Q, k, k)
    b, _ = o.search(Q, k)
    assert np.array_equal(a, b)


def test_beam_as_wide_as_a_connected_graph_is_exact():
    """With ef >= n on a graph whose layer 0 is connected, search_k visits every node and the result is the
    exact k-NN (lib/ohnsw.ml:543-588: nothing is ever rejected while the result holds fewer than ef)."""
    checked = 0
    for seed in range(12):
        n = 90
        X, lv, o, g = _build(n, 4, 6, 40, seed)
        rows = _rows(g, 0)
        seen, todo = {g.entry}, [g.entry]
        while todo:
            for j in rows[todo.pop()]:
                if j not in seen:
                    seen.add(j)
                    todo.append(j)
        if len(seen) != n:
            continue
        checked += 1
        Q = uniform(20, 4, seed + 100)
        ids, d = o.search(Q, 10, n)
        _, bd = O.bruteforce(X, Q, 10)
        assert np.array_equal(d, bd)
    assert checked >= 6


def test_counters_count_distance_evaluations_and_expansions():
    X, lv, o, g = _build(300, 8, 4, 20, 3)
    Q = uniform(17, 8, 9)
    ids, d, cnt = o.search(Q, 5, 12, counters=True)
    ids_mt, d_mt, _, tot = o.search_mt(Q, 5, 12, nthreads=3)
    assert np.array_equal(ids, ids_mt) and np.array_equal(d.view(np.uint32), d_mt.view(np.uint32))
    assert np.array_equal(cnt.sum(axis=0), tot)              # the threaded arm does the same work, query by query
    assert (cnt[:, 0] >= cnt[:, 1]).all() and (cnt[:, 1] >= 1).all()      # every query expands its entry point
    # every distance evaluation is for a distinct node on layer 0, plus the upper-layer scans: never more than
    # (expansions on layer 0) * 2M + (upper expansions) * M + 1
    assert (cnt[:, 0] <= cnt[:, 1] * 8 + cnt[:, 2] * 4 + 1).all()


def test_rebuilding_from_the_same_inputs_is_the_same_graph():
    _, _, _, a = _build(400, 6, 4, 24, 11)
    _, _, _, b = _build(400, 6, 4, 24, 11)
    assert a.max_layer == b.max_layer and a.entry == b.entry
    for l in range(a.max_layer + 1):
        assert np.array_equal(a.offsets[l], b.offsets[l]) and np.array_equal(a.nbrs[l], b.nbrs[l])


def test_export_import_round_trip_keeps_results():
    X, lv, o, g = _build(500, 8, 6, 30, 21)
    o2 = O.VecOracle(8).import_graph(X, g)
    Q = uniform(40, 8, 22)
    a = o.search(Q, 10, 25)
    b = o2.search(Q, 10, 25)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    g2 = o2.export()
    for l in range(g.max_layer + 1):
        assert np.array_equal(g.nbrs[l], g2.nbrs[l])


@pytest.mark.parametrize("M", [2, 16, 24])
def test_level_formula_is_round_to_nearest(M):
    """Q1 (lib/ohnsw.ml:781): level = floor(-ln U / ln M + 0.5), so P(level >= l) = M^-(l - 1/2)."""
    lv = draw_levels(400_000, M, 5)
    for l in (1, 2):
        p = (lv >= l).mean()
        expect = float(M) ** -(l - 0.5)
        assert abs(p - expect) < 4 * np.sqrt(expect / len(lv)) + 1e-4
