"""Generates tests/golden/oracle_small.npz: a small seeded data set, the graph the oracle builds on
it (= what lib/ohnsw.ml:766-857 builds, sequential inserts) and the rows its knn_batch returns.
The fixture freezes today's oracle: tests/test_oracle_golden.py checks the oracle still reproduces
it, the gpu tests check that the CUDA build (sequential mode) and search reproduce it bit for bit.

    python tests/golden/make_oracle_fixture.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O          # noqa: E402
from tests.util import draw_levels, uniform   # noqa: E402

n, dim, nq, M, efC, k, ef = 400, 24, 32, 6, 40, 8, 24
X = np.round(uniform(n, dim, 101) * 64) / 64        # dyadic values: exact in fp32 whatever the summation order
Q = np.round(uniform(nq, dim, 102) * 64) / 64
lv = draw_levels(n, M, seed=103)
lv[0] = 0
o = O.VecOracle(dim).build(X, M, efC, lv)
g = o.export()
ids, d, cnt = o.search(Q, k, ef, counters=True)
out = dict(X=X.astype(np.float32), Q=Q.astype(np.float32), levels=lv.astype(np.int32), params=np.array([M, efC, k, ef], np.int32),
           entry=np.int64(g.entry), max_layer=np.int32(g.max_layer), ids=ids, dists=d, counters=cnt)
for l in range(g.max_layer + 1):
    out[f"offsets{l}"] = g.offsets[l].astype(np.int64)
    out[f"nbrs{l}"] = g.nbrs[l].astype(np.int32)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "oracle_small.npz"), **out)
print("wrote oracle_small.npz:", n, "nodes,", g.max_layer + 1, "layers,", int(sum(len(a) for a in g.nbrs)), "links")
