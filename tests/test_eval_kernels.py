"""Brute-force ground truth (benchmark/dataset.ml:15-30) and the shard merge, on the GPU."""
import numpy as np
import pytest

import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import capi
from oracle import oracle as O
from tests.util import uniform

pytestmark = pytest.mark.gpu

REL = 1e-5      # fp32 summation-order tolerance north_star states


def _same_up_to_ties(ids_a, d_a, ids_b, d_b):
    assert np.allclose(d_a, d_b, rtol=REL, atol=1e-6, equal_nan=True)
    for i in np.nonzero((ids_a != ids_b).any(axis=1))[0]:
        for j in np.nonzero(ids_a[i] != ids_b[i])[0]:
            # a differing id must sit in a group of (near-)equal distances
            assert abs(d_a[i, j] - d_b[i, j]) <= REL * max(d_a[i, j], 1e-6)


@pytest.mark.parametrize("n,nq,dim,k", [(5000, 100, 128, 10), (777, 13, 7, 5), (3000, 70, 100, 100), (90, 4, 960, 64)])
def test_bruteforce_matches_oracle(n, nq, dim, k):
    X, Q = uniform(n, dim, 3), uniform(nq, dim, 4)
    ids_o, d_o = O.bruteforce(X, Q, k)
    ids_g, d_g = H.brute_force_knn_l2(X, Q, k, return_ids=True)
    _same_up_to_ties(ids_g, d_g, ids_o, d_o)
    assert H.Recall.compute(d_o, d_g, epsilon=1e-4) > 0.999


@pytest.mark.parametrize("kind,n,nq,dim,k", [("uniform", 20000, 300, 128, 10), ("sift", 30000, 500, 128, 10), ("uniform", 5000, 130, 100, 16),
                                              ("normal", 8000, 64, 960, 10), ("uniform", 1000, 257, 20, 1)])
def test_bruteforce_tensor_core_path_is_exact(kind, n, nq, dim, k):
    """L2, k <= 16, n >= 256 runs on tcgen05 (bf16 hi/lo split GEMM ranks, fp32 re-ranks and proves):
    the result must be the oracle's, ids and distances bit for bit (re-rank uses the oracle's order)."""
    if kind == "sift":
        X, Q = H.sift_like(n, dim, seed=5), H.sift_like(nq, dim, seed=6)
    elif kind == "normal":
        X = np.random.default_rng(7).standard_normal((n, dim)).astype(np.float32) * 3
        Q = np.random.default_rng(8).standard_normal((nq, dim)).astype(np.float32) * 3
    else:
        X, Q = uniform(n, dim, 3), uniform(nq, dim, 4)
    ids_g, d_g = H.brute_force_knn_l2(X, Q, k, return_ids=True)
    unproven = capi.lib().hnswb200_bruteforce_last_unproven()
    assert unproven >= 0, "tensor-core path was not taken"
    ids_o, d_o = O.bruteforce(X, Q, k)
    if unproven == 0:
        assert np.array_equal(ids_g, ids_o) and np.array_equal(d_g.view(np.uint32), d_o.view(np.uint32))
    else:
        _same_up_to_ties(ids_g, d_g, ids_o, d_o)
    assert unproven <= nq // 50, f"{unproven} of {nq} queries fell back to fp32"


def test_bruteforce_pads_when_k_exceeds_n():
    X, Q = uniform(20, 16, 3), uniform(3, 16, 4)
    ids_g, d_g = H.brute_force_knn_l2(X, Q, 32, return_ids=True)
    ids_o, d_o = O.bruteforce(X, Q, 32)
    assert (ids_g[:, 20:] == -1).all() and np.isnan(d_g[:, 20:]).all()
    _same_up_to_ties(ids_g, d_g, ids_o, d_o)


def test_merge_topk_packed_blocks():
    """The layout of the single all-gather: every shard contributes one [ids | dists] block."""
    import torch
    rng = np.random.default_rng(10)
    S, nq, k = 3, 50, 7
    d = np.sort(rng.random((S, nq, k), dtype=np.float32), axis=2)
    ids = rng.integers(0, 1000, (S, nq, k)).astype(np.int32)
    packed = np.stack([ids, d.view(np.int32)], axis=1)             # [S][2][nq][k]
    t = torch.from_numpy(np.ascontiguousarray(packed)).cuda()
    o_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    offs = np.arange(S, dtype=np.int64) * 1000
    capi.check(capi.lib().hnswb200_merge_topk_device(t.data_ptr(), t.data_ptr() + nq * k * 4, S, nq, k, 2 * nq * k,
                                                     capi.ptr(offs), o_ids.data_ptr(), o_d.data_ptr(), None))
    got = o_ids.cpu().numpy()
    for q in range(nq):
        cand = sorted((d[s, q, j], ids[s, q, j] + offs[s]) for s in range(S) for j in range(k))
        assert got[q].tolist() == [c[1] for c in cand[:k]]
    with pytest.raises(ValueError, match="shard_stride"):
        capi.check(capi.lib().hnswb200_merge_topk_device(t.data_ptr(), t.data_ptr(), S, nq, k, 3, None, o_ids.data_ptr(),
                                                         o_d.data_ptr(), None))


def test_bruteforce_integer_data_is_exact():
    rng = np.random.default_rng(5)
    X = rng.integers(0, 219, (4000, 128)).astype(np.float32)
    Q = rng.integers(0, 219, (50, 128)).astype(np.float32)
    ids_o, d_o = O.bruteforce(X, Q, 10)
    ids_g, d_g = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
    assert np.array_equal(ids_o, ids_g) and np.array_equal(d_o, d_g)


def test_merge_topk():
    import ctypes as C
    import torch
    rng = np.random.default_rng(9)
    S, nq, k = 5, 300, 10
    d = np.sort(rng.random((S, nq, k), dtype=np.float32), axis=2)
    ids = rng.integers(0, 1 << 20, (S, nq, k)).astype(np.int32)
    ids[0, :, 7:] = -1; d[0, :, 7:] = np.nan                       # a shard with fewer than k results
    ids[1, 0, :] = -1; d[1, 0, :] = np.nan
    t_ids, t_d = torch.from_numpy(ids).cuda(), torch.from_numpy(d).cuda()
    o_ids = torch.empty((nq, k), dtype=torch.int32, device="cuda")
    o_d = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    offs = np.arange(S, dtype=np.int64) * (1 << 20)                # shard s holds global rows [s * 2^20, ...)
    capi.check(capi.lib().hnswb200_merge_topk_device(t_ids.data_ptr(), t_d.data_ptr(), S, nq, k, 0, capi.ptr(offs),
                                                     o_ids.data_ptr(), o_d.data_ptr(), None))
    got_ids, got_d = o_ids.cpu().numpy(), o_d.cpu().numpy()
    for q in range(nq):
        cand = [(d[s, q, j], ids[s, q, j] + offs[s]) for s in range(S) for j in range(k) if ids[s, q, j] >= 0]
        cand.sort()
        assert got_ids[q].tolist() == [c[1] for c in cand[:k]]
        assert got_d[q].tolist() == [c[0] for c in cand[:k]]
