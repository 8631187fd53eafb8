"""A/B of two library builds on the bench workload: search kernel ms per ef (HNSWB200_LIB selects the build)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
out = [os.path.basename(os.environ.get("HNSWB200_LIB", "HEAD")), f"build={h.stats().build_seconds:.2f}s"]
for ef in (16, 32, 41, 48, 64, 96, 128, 256):
    ms = []
    for _ in range(5):
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef); ms.append(h.stats().search_kernel_ms)
    out.append(f"ef{ef}={min(ms):.3f}")
print(" ".join(out), flush=True)
