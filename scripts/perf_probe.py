"""Kernel-level probe: oracle-built graph -> GPU layout -> timed search at several ef."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from oracle import oracle as O
from tests.util import draw_levels

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
nq = 10000
X = H.sift_like(n, 128, seed=1234)
Q = H.sift_like(nq, 128, seed=4321)
M, efC = 16, 100
t = time.time(); o = O.VecOracle(128).build(X, M, efC, draw_levels(n, M)); print("oracle build s", time.time() - t, flush=True)
h = Ohnsw.Hgraph(128, Ohnsw.distance_l2, M, efC).import_graph(X, o.export())
gt, _ = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
for ef in (10, 16, 32, 64, 128, 256):
    for rep in range(3):
        t = time.time(); ids, d = Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef); dt = time.time() - t
    st = h.stats()
    print(f"ef={ef} recall={H.Recall.ids(gt, ids):.4f} e2e_qps={nq/dt:.0f} kernel_ms={st.search_kernel_ms:.3f} "
          f"kernel_qps={nq/st.search_kernel_ms*1e3:.0f} ndist/q={st.search_n_dist/nq:.0f} nexp0/q={st.search_n_exp0/nq:.1f} "
          f"GB/s={st.search_algorithmic_bytes/st.search_kernel_ms/1e6:.1f} spills={st.search_visited_overflows}", flush=True)
