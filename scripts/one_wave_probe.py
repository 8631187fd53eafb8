"""One-wave batches (fewer queries than resident warps: a replica's share at N = 4 / 8): kernel time vs CTA shape.
    python scripts/one_wave_probe.py [rows] [ef]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 29
X = H.sift_like(1000000, 128, seed=1234)[:n].copy(); Q = H.sift_like(10000, 128, seed=4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
h.set_param("gang", 1)
for nq in (2500, 3000, 1250):
    for w in (0, 1, 2, 4):
        h.set_param("warps_per_cta", w)
        ms = []
        for _ in range(10):
            Ohnsw.knn_batch_bigarray(h, Q[:nq], k=10, ef=ef)
            ms.append(h.stats().search_kernel_ms)
        print(f"rows={n} ef={ef} nq={nq} warps_per_cta={w}: kernel_ms min {min(ms):.4f} median {sorted(ms)[5]:.4f}", flush=True)
