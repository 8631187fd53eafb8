"""Search kernel time vs resident warps per SM (tuning probe)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
h.set_param("hash_slots", 0)
for w in (12, 16, 20, 24):
    h.set_param("max_warps_per_sm", w)
    ms = []
    for _ in range(4):
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=48); ms.append(h.stats().search_kernel_ms)
    print(f"warps/SM<={w} kernel_ms={min(ms):.3f}", flush=True)
