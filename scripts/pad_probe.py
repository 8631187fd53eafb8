"""dim = 100 (GloVe shape): does padding vector rows to a whole number of 128-byte lines pay? (tuning probe)"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from bench import draw_levels, make_data
class A: pass
a = A(); a.metric = "angular"; a.dim = 100
n = 1183514
X = make_data(a, n, 1234); Q = make_data(a, 10000, 4321)
lv = draw_levels(n, 24, 7)
res = {}
for rf in (100, 128):
    h = Ohnsw.Hgraph(100, Ohnsw.distance_angular, 24, 200)
    h.set_param("row_floats", rf)
    t = time.time()
    capi.check(capi.lib().hnswb200_build(h._h, capi.ptr(X), n, capi.ptr(lv)))
    bs = h.stats().build_seconds
    for ef in (32, 65):
        ms = []
        for _ in range(4):
            ids, d = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef); ms.append(h.stats().search_kernel_ms)
        res[(rf, ef)] = ids
        print(f"row_floats={rf} build {bs:.2f}s ef={ef} kernel_ms={min(ms):.3f}", flush=True)
    h.close()
print("same ids:", all(np.array_equal(res[(100, ef)], res[(128, ef)]) for ef in (32, 65)))
