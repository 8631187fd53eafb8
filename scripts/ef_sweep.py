"""Search kernel time across the ef sweep on the bench workload (tuning probe)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
t = time.time()
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
st = h.stats()
print(f"build {time.time()-t:.2f}s lib {st.build_seconds:.2f}s ndist/ins {st.build_n_dist/n:.0f} spills {st.build_visited_overflows}", flush=True)
gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
for mode in (0,):
    h.set_param("visited_mode", mode)
    for ef in (41, 56, 64, 72, 80, 128):
        ms = []
        for _ in range(3):
            ids, _ = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef); s = h.stats(); ms.append(s.search_kernel_ms)
        print(f"mode={mode} ef={ef} kernel_ms={min(ms):.3f} recall={H.Recall.ids(gt, ids):.4f} spills={s.search_visited_overflows} ndist/q={s.search_n_dist/10000:.0f} GB/s={(s.search_algorithmic_bytes)/min(ms)/1e6:.0f}", flush=True)
