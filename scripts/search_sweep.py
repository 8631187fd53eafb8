"""Tuning probe on the bench workload: search kernel time vs visited-hash size / CTA shape / ef."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
efc = int(sys.argv[2]) if len(sys.argv) > 2 else 200
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
t = time.time()
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=efc, levels=draw_levels(n, 16, 7))
st = h.stats()
print(f"build {time.time()-t:.2f}s lib {st.build_seconds:.2f}s ndist/ins {st.build_n_dist/n:.0f} spills {st.build_visited_overflows}", flush=True)
def run(ef, reps=5):
    ms = []
    for _ in range(reps):
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
        s = h.stats(); ms.append(s.search_kernel_ms)
    return min(ms), s
for ef in (48,):
    for hs in (0, 1536, 2048, 2304, 2560, 3072, 4096):
        for wpc in (0, 2, 4):
            h.set_param("hash_slots", hs); h.set_param("warps_per_cta", wpc)
            ms, s = run(ef)
            print(f"ef={ef} hash={hs} wpc={wpc} kernel_ms={ms:.3f} spills={s.search_visited_overflows} GB/s={(s.search_algorithmic_bytes)/ms/1e6:.0f}", flush=True)
h.set_param("hash_slots", 0); h.set_param("warps_per_cta", 0)
for ef in (10, 16, 32, 64, 96, 128, 256, 512):
    ms, s = run(ef, 3)
    print(f"ef={ef} kernel_ms={ms:.3f} spills={s.search_visited_overflows} ndist/q={s.search_n_dist/10000:.0f} GB/s={(s.search_algorithmic_bytes)/ms/1e6:.0f}", flush=True)
