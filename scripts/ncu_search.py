"""Short run for `ncu --set full -k regex:search_kernel`: build the bench index, search 10k queries a few times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from bench import draw_levels
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
ef = int(sys.argv[2]) if len(sys.argv) > 2 else 41
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(nq, 128, seed=4321)
h = Ohnsw.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200, levels=draw_levels(n, 16, 7))
for _ in range(4):
    Ohnsw.knn_batch_bigarray(h, Q, k=10, ef=ef)
print("kernel ms", h.stats().search_kernel_ms)
