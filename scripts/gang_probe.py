"""Gang of warps per query / insert: search kernel time for small batches (a replica's share of the queries) with
1, 2 and 4 warps per query, and build time with and without gangs for the small early batches."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
lv = draw_levels(n, 16, 7)
for gang in (1, 0):
    h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 200)
    h.set_param("gang", gang)
    t = time.time()
    H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), n, H.capi.ptr(lv)))
    print(f"build gang={'auto' if gang == 0 else 1}: wall {time.time() - t:.2f}s lib {h.stats().build_seconds:.2f}s", flush=True)
for ef in (41,):
    for nq in (10000, 5000, 2500, 1250, 625, 100, 1):
        for gang in (1, 2, 4):
            h.set_param("gang", gang)
            ms = []
            for _ in range(8):
                Ohnsw.knn_batch_bigarray(h, Q[:nq], k=10, ef=ef)
                ms.append(h.stats().search_kernel_ms)
            print(f"ef={ef} nq={nq} gang={gang}: kernel_ms min {min(ms):.4f}  -> {nq / min(ms) / 1e3:.2f} M q/s", flush=True)
