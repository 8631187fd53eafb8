"""One GPU build of the bench workload's first n rows (for ncu captures of the build kernels). argv: n [ratio]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300000
X = H.sift_like(n, 128, seed=1234)
lv = draw_levels(n, 16, 7)
h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 200)
if len(sys.argv) > 2: h.set_param("build_ratio", int(sys.argv[2]))
t = time.time()
H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), n, H.capi.ptr(lv)))
print(f"n={n} wall {time.time()-t:.2f}s lib {h.stats().build_seconds:.2f}s", flush=True)
