import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from bench import draw_levels
n = 1000000
X = H.sift_like(n, 128, seed=1234); Q = np.ascontiguousarray(H.sift_like(10000, 128, seed=4321))
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
out = (np.empty((10000, 10), np.int32), np.empty((10000, 10), np.float32))
for b in (Q,) + out: capi.host_register(b)
for C, zc in ((1, 1), (1, 0), (4, 0), (8, 0), (1, 1), (1, 0)):
    h.set_param("host_chunks", C)
    h.set_param("host_zero_copy", zc)
    for _ in range(3): Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=41, out=out)
    t = time.perf_counter()
    for _ in range(20): Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=41, out=out)
    dt = (time.perf_counter() - t) / 20
    print(f"host_chunks={C} zero_copy={zc} (path bits {h.stats().search_zero_copy}, kernel {h.stats().search_kernel_ms:.3f} ms): {dt*1e3:.3f} ms per call, {10000/dt/1e6:.2f} M queries/s", flush=True)
