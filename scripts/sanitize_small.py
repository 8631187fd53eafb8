"""Small end-to-end run for compute-sanitizer: build (batched + sequential), search (hash, spill, bitset), brute force (both), merge."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from tests.util import draw_levels, uniform
X = H.sift_like(3000, 128, seed=1); Q = H.sift_like(200, 128, seed=2)
lv = draw_levels(len(X), 16)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=60, levels=lv)
for ef in (10, 40, 200):
    ids, d = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
h.set_param("hash_slots", 1024); Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=100); h.set_param("hash_slots", 0)
h.set_param("visited_mode", 2); Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=20); h.set_param("visited_mode", 0)
Ohnsw.insert(h, X[:50] + 1.0)
hs = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 8, 30); hs.set_param("build_batch", 1)
capi.check(capi.lib().hnswb200_build(hs._h, capi.ptr(X[:300]), 300, capi.ptr(lv[:300])))
X2 = uniform(700, 100, 3)
h2 = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X2, num_connections=6, num_nodes_search_construction=300, levels=draw_levels(700, 6))
Ohnsw.knn_batch_bigarray(h2, uniform(50, 100, 4), k=5, ef=64)
X3 = uniform(400, 960, 5)
h3 = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X3, num_connections=8, num_nodes_search_construction=40, levels=draw_levels(400, 8))
Ohnsw.knn_batch_bigarray(h3, uniform(20, 960, 6), k=5, ef=30)
gt = H.brute_force_knn_l2(X, Q, 10)
os.environ["HNSWB200_BRUTEFORCE"] = "fp32"; gt2 = H.brute_force_knn_l2(X, Q, 10)
assert np.array_equal(gt.view(np.uint32), gt2.view(np.uint32))
print("sanitize run ok", H.Recall.compute(gt, d, 1e-4))
