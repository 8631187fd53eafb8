"""Build-time probe on the bench workload with the per-phase trace (HNSWB200_BUILD_TRACE=1: phase totals and
phase-1 time per batch-size bucket).  argv: n, then any number of ratio:batch[:gang[:build_qreg[:build_mates[:ratio_early]]]] tuples."""
import sys, os, time
os.environ["HNSWB200_BUILD_TRACE"] = "1"
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
runs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(64, 16384), (64, 16384), (32, 16384)]
X = H.sift_like(n, 128, seed=1234)
lv = draw_levels(n, 16, 7)
for r in runs:
    ratio, batch = r[0], r[1]
    h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, 16, 200)
    h.set_param("build_ratio", ratio); h.set_param("build_batch", batch)
    if len(r) > 2: h.set_param("gang", r[2])
    if len(r) > 3: h.set_param("build_qreg", r[3])
    if len(r) > 4: h.set_param("build_mates", r[4])
    if len(r) > 5: h.set_param("build_ratio_early", r[5])
    t = time.time()
    H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), n, H.capi.ptr(lv)))
    print(f"ratio={ratio} batch={batch} wall {time.time()-t:.2f}s lib {h.stats().build_seconds:.2f}s", flush=True)
    h.close()
