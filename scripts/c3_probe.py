"""Config C3 (GloVe shape, 1.18M x 100 angular, M = 24): search-kernel time at the bench's ef for several planner knobs on
ONE built index.  `python scripts/c3_probe.py [ef] [once]` — `once` runs the default plan only (for an ncu capture)."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
import bench
a = types.SimpleNamespace(**bench.CONFIGS["c3"], k=10)
ef = int(sys.argv[1]) if len(sys.argv) > 1 else 64
once = len(sys.argv) > 2
X = bench.make_data(a, a.n, 1234); Q = bench.make_data(a, a.nq, 4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_angular, X, num_connections=a.M, num_nodes_search_construction=a.efc,
                               levels=bench.draw_levels(a.n, a.M, 7))
def run(label, reps=4):
    ms = []
    for _ in range(reps):
        Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
        ms.append(h.stats().search_kernel_ms)
    st = h.stats()
    print(f"{label:34s} kernel ms {min(ms):.3f}  evals/q {st.search_n_dist / a.nq:.0f}  exp/q {st.search_n_exp0 / a.nq:.1f}  spills {st.search_visited_overflows}", flush=True)
os.environ["HNSWB200_TRACE"] = "1"
run("default", 1 if once else 4)
if once:
    sys.exit(0)
for name, vals, reset in (("hash_bits", (32, 16), 0), ("hash_slots", (2048, 4096, 6144), 0), ("visited_mode", (2,), 0),
                          ("max_warps_per_sm", (12, 16, 20), 0), ("warps_per_cta", (1, 2), 0)):
    for v in vals:
        h.set_param(name, v)
        run(f"{name}={v}")
    h.set_param(name, reset)
for e in (32, 48, 96, 128):
    ef = e
    run(f"ef={e}")
