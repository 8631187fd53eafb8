"""A/B of two library builds on the bench workload (HNSWB200_LIB selects the build): search kernel time at the bench's ef."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
efs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [41]
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
t = time.time()
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
print(f"{os.environ.get('HNSWB200_LIB', 'default lib')}: build {time.time() - t:.2f}s", flush=True)
for ef in efs:
    ms = []
    for _ in range(12):
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
        s = h.stats(); ms.append(s.search_kernel_ms)
    print(f"  ef={ef}: kernel_ms min {min(ms):.4f} median {sorted(ms)[len(ms)//2]:.4f}  ndist/q {s.search_n_dist/10000:.1f}  GB/s {s.search_algorithmic_bytes/min(ms)/1e6:.0f}", flush=True)
