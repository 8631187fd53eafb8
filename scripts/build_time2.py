import sys, os, time
os.environ["HNSWB200_BUILD_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(2000, 128, seed=4321)
lv = draw_levels(n, 16, 7)
gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
for rep in range(2):
    h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 200)
    t = time.time()
    H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), n, H.capi.ptr(lv)))
    ids, _ = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=41)
    print(f"wide={os.environ.get('HNSWB200_BUILD_WIDE','auto')} lib {h.stats().build_seconds:.3f}s recall@ef41 {H.Recall.ids(gt, ids):.4f} checksum {int(h.export_graph().nbrs[0].astype(np.int64).sum())}", flush=True)
    h.close()
