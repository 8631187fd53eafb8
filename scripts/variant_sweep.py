"""Run one tuning build of the library (HNSWB200_LIB) on the bench workload; prints kernel ms per ef/mode."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
out = [os.path.basename(os.environ.get("HNSWB200_LIB", "default"))]
for mode in (1, 2):
    h.set_param("visited_mode", mode)
    for ef in (32, 48, 64):
        ms = []
        for _ in range(4):
            Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef); ms.append(h.stats().search_kernel_ms)
        out.append(f"m{mode}/ef{ef}={min(ms):.3f}")
print(" ".join(out), flush=True)
