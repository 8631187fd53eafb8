"""Brute-force ground truth: tensor-core path vs fp32 CUDA-core path, wall time of the C-ABI call (H2D included)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
for kind in ("sift", "uniform"):
    if kind == "sift":
        X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(10000, 128, seed=4321)
    else:
        X = (np.random.default_rng(1).random((n, 128), dtype=np.float32) * 2 - 1); Q = (np.random.default_rng(2).random((10000, 128), dtype=np.float32) * 2 - 1)
    for mode in ("tc", "fp32", "tc", "fp32"):
        if mode == "fp32": os.environ["HNSWB200_BRUTEFORCE"] = "fp32"
        else: os.environ.pop("HNSWB200_BRUTEFORCE", None)
        t = time.time(); ids, d = H.brute_force_knn_l2(X, Q, 10, return_ids=True); dt = time.time() - t
        print(kind, mode, f"{dt*1e3:.1f} ms unproven={capi.lib().hnswb200_bruteforce_last_unproven()} checksum={int(ids.astype(np.int64).sum())}", flush=True)
