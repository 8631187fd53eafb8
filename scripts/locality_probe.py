"""Does ordering the query batch by spatial cell raise the L2 hit rate enough to matter? (tuning probe)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
def t(Qx, ef=41):
    ms = []
    for _ in range(5):
        Ohnsw.knn_batch_bigarray(h, Qx, k=10, ef=ef); ms.append(h.stats().search_kernel_ms)
    return min(ms)
print("original order", t(Q), flush=True)
rng = np.random.default_rng(0)
for ncell in (100, 1000, 4000):
    C = X[rng.choice(n, ncell, replace=False)]
    d = (Q * Q).sum(1)[:, None] + (C * C).sum(1)[None, :] - 2 * Q @ C.T
    cell = d.argmin(1)
    order = np.argsort(cell, kind="stable")
    print(f"sorted by nearest of {ncell} random centroids", t(np.ascontiguousarray(Q[order])), flush=True)
