"""Build schedule vs graph quality on the bench workload (sift-like n x 128, M=16, efC=200): recall@10 of the GPU-built
index at several ef against the exact scan, and the build time.  argv: n [ratio:batch:mates[:ratio_early] ...]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
runs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(64, 16384, 0)]
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
lv = draw_levels(n, 16, 7)
gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
efs = (10, 16, 24, 32, 41, 48, 64, 128)
print("ef     ", " ".join(f"{e:6d}" for e in efs))
for r in runs:
    h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 200)
    h.set_param("build_ratio", r[0]); h.set_param("build_batch", r[1]); h.set_param("build_mates", r[2])
    if len(r) > 3: h.set_param("build_ratio_early", r[3])
    H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), n, H.capi.ptr(lv)))
    st = h.stats()
    rec = " ".join(f"{H.Recall.ids(gt, Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)[0]):.4f}" for ef in efs)
    print(f"{rec}  run={r} build_s={st.build_seconds:.2f} ndist/ins={st.build_n_dist/n:.0f} deg0={st.layer_mean_degree[0]:.2f}", flush=True)
    h.close()
