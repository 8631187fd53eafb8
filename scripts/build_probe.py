"""Build-quality probe: GPU-built (several batch schedules) vs oracle-built recall, and build time.
argv: n with_oracle(0/1) generator efC [ratio:batch[:mates[:ratio_early]] ...]"""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from oracle import oracle as O
from tests.util import draw_levels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
with_oracle = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
gen = sys.argv[3] if len(sys.argv) > 3 else "sift"
nq = 2000
M, efC = 16, (int(sys.argv[4]) if len(sys.argv) > 4 else 100)
if gen == "sift":
    X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(nq, 128, seed=4321)
else:
    X = (np.random.default_rng(1234).random((n, 128), dtype=np.float32) * 2 - 1)
    Q = (np.random.default_rng(4321).random((nq, 128), dtype=np.float32) * 2 - 1)
lv = draw_levels(n, M)
gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
efs = (10, 16, 32, 64, 128)
if with_oracle:
    t = time.time(); o = O.VecOracle(128).build(X, M, efC, lv); print("oracle build s", time.time() - t, flush=True)
    print("oracle ", " ".join(f"{H.Recall.ids(gt, o.search_mt(Q, 10, ef)[0]):.4f}" for ef in efs), flush=True)
runs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[5:]] or \
    [(16, 16384), (32, 16384), (48, 16384), (64, 16384), (128, 16384), (16, 1024), (10**9, 1)][: (7 if n <= 20000 else 5)]
for r in runs:
    ratio, batch = r[0], r[1]
    h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, M, efC)
    h.set_param("build_ratio", ratio); h.set_param("build_batch", batch)
    if len(r) > 2: h.set_param("build_mates", r[2])
    if len(r) > 3: h.set_param("build_ratio_early", r[3])
    t = time.time()
    H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), n, H.capi.ptr(lv)))
    dt = time.time() - t
    st = h.stats()
    rec = " ".join(f"{H.Recall.ids(gt, Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)[0]):.4f}" for ef in efs)
    print(f"run={r} build_s={dt:.2f} ndist/ins={st.build_n_dist/n:.0f} launches={st.gpu_launches} deg0={st.layer_mean_degree[0]:.2f} recall {rec}", flush=True)
