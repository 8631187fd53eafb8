import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
ids, d = H.brute_force_knn_l2(X, Q, 10, return_ids=True)
print(int(ids.astype(np.int64).sum()))
