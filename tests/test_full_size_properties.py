"""BASELINE.json's full size (config C2: 1M x 128 fp32, 10k queries, M=16, efConstruction=200, k=10),
where the oracle cannot run in test time: size-independent properties of the build + search path.

  * rows ascending, ids valid and distinct, distances = exact L2 of the returned ids (recomputed in numpy)
  * idempotence: the same call twice returns the same bytes; host-buffer and device-buffer calls agree
  * a vector of the data set finds itself first at distance 0
  * the graph invariants the reference tests hold (symmetric links, degree bounds, (almost) no isolated node on layer 0)
  * recall@10 >= 0.95 at ef = 48 against the exact scan, and monotone in ef
  * brute force: tensor-core path == fp32 path, bit for bit
"""
import os

import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from tests.util import draw_levels

pytestmark = pytest.mark.gpu

N, DIM, NQ, M, EFC, K = 1_000_000, 128, 10_000, 16, 200, 10


@pytest.fixture(scope="module")
def full():
    X = H.sift_like(N, DIM, seed=1234)
    Q = H.sift_like(NQ, DIM, seed=4321)
    lv = draw_levels(N, M)
    h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=M, num_nodes_search_construction=EFC, levels=lv)
    gt_ids, gt_d = H.brute_force_knn_l2(X, Q, K, return_ids=True)
    assert capi.lib().hnswb200_bruteforce_last_unproven() == 0          # tensor-core path, proven exact
    return X, Q, lv, h, gt_ids, gt_d


def test_result_rows_are_well_formed_and_exact(full):
    X, Q, lv, h, gt_ids, gt_d = full
    ids, d = Ohnsw.knn_batch_bigarray(h, Q, k=K, ef=48)
    assert ids.min() >= 0 and ids.max() < N
    assert (np.diff(d, axis=1) >= 0).all(), "rows must ascend"
    assert (np.sort(ids, axis=1)[:, 1:] != np.sort(ids, axis=1)[:, :-1]).all(), "duplicate id in a row"
    sel = np.random.default_rng(0).choice(NQ, 500, replace=False)
    exact = np.sqrt(((X[ids[sel]].astype(np.float64) - Q[sel, None, :].astype(np.float64)) ** 2).sum(-1))
    assert np.allclose(d[sel], exact, rtol=1e-6, atol=0)               # integer-valued data: fp32 sums are exact
    assert H.Recall.ids(gt_ids, ids) >= 0.95
    assert H.Recall.compute(gt_d, d, 1e-4) >= 0.95                      # the reference's recall definition


def test_idempotent_and_buffer_kinds_agree(full):
    import torch
    X, Q, lv, h, _, _ = full
    a = Ohnsw.knn_batch_bigarray(h, Q, k=K, ef=41)
    b = Ohnsw.knn_batch_bigarray(h, Q, k=K, ef=41)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    h.set_param("host_chunks", 3)                 # copy / search overlap in three pieces: same bytes
    c = Ohnsw.knn_batch_bigarray(h, Q, k=K, ef=41)
    h.set_param("host_chunks", 1)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1].view(np.uint32), c[1].view(np.uint32))
    q = torch.from_numpy(Q).cuda()
    ids = torch.empty((NQ, K), dtype=torch.int32, device="cuda")
    d = torch.empty((NQ, K), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    h.search_device(q.data_ptr(), NQ, K, 41, ids.data_ptr(), d.data_ptr())
    assert np.array_equal(ids.cpu().numpy(), a[0]) and np.array_equal(d.cpu().numpy().view(np.uint32), a[1].view(np.uint32))


def test_recall_is_monotone_in_ef(full):
    X, Q, lv, h, gt_ids, _ = full
    rec = [H.Recall.ids(gt_ids[:2000], Ohnsw.knn_batch_bigarray(h, Q[:2000], k=K, ef=ef)[0]) for ef in (10, 16, 32, 64, 128, 256)]
    assert all(b >= a - 1e-3 for a, b in zip(rec, rec[1:])), rec
    assert rec[-1] > 0.999


def test_data_vectors_find_themselves(full):
    X, Q, lv, h, _, _ = full
    sel = np.random.default_rng(1).choice(N, 2000, replace=False)
    ids, d = Ohnsw.knn_batch_bigarray(h, X[sel], k=1, ef=32)
    hit = ids[:, 0] == sel
    assert hit.mean() > 0.97                     # duplicates of a vector may answer for it
    assert (d[:, 0] == 0).mean() > 0.99


def test_graph_invariants_at_scale(full):
    X, Q, lv, h, _, _ = full
    st = h.stats()
    # path B has no do_not_isolate rule: symmetric pruning may strand a node; it must stay rare
    assert st.layer_nodes[0] == N and st.layer_isolated[0] <= N // 100_000
    assert st.layer_max_degree[0] <= 2 * M and all(st.layer_max_degree[l] <= M for l in range(1, st.num_layers))
    assert st.num_layers == int(lv.max()) + 1
    g = h.export_graph()
    for l in range(g.max_layer + 1):
        deg = g.degree(l)
        src = np.repeat(np.arange(N, dtype=np.int64), deg)
        fwd = src * N + g.nbrs[l]
        assert np.array_equal(np.sort(fwd), np.sort(g.nbrs[l].astype(np.int64) * N + src)), f"layer {l} links not symmetric"


def test_bruteforce_paths_agree_at_scale(full):
    X, Q, lv, h, gt_ids, gt_d = full
    os.environ["HNSWB200_BRUTEFORCE"] = "fp32"
    try:
        ids, d = H.brute_force_knn_l2(X, Q[:2000], K, return_ids=True)
    finally:
        os.environ.pop("HNSWB200_BRUTEFORCE")
    assert capi.lib().hnswb200_bruteforce_last_unproven() == -1
    assert np.array_equal(ids, gt_ids[:2000]) and np.array_equal(d.view(np.uint32), gt_d[:2000].view(np.uint32))
