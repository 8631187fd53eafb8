"""Tuning probe for the non-headline configs (bench.py --config c3|c4|c5 shapes): build once, find the ef for
recall@10 >= 0.95, then time the search kernel under a list of parameter settings.

    python scripts/config_probe.py c4 [rows] [name=v,name=v ...]      e.g.  c4 200000 stage_rows=4 stage_rows=8,max_warps_per_sm=6
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw


class A:
    pass


cfg = sys.argv[1]
a = A()
a.latent, a.k = 16, 10
for k, v in bench.CONFIGS[cfg].items():
    setattr(a, k, v)
settings = []
build_params = {}
for arg in sys.argv[2:]:
    if arg.isdigit():
        a.n = int(arg)
    elif arg.startswith("build:"):
        build_params.update({kv.split("=")[0]: int(kv.split("=")[1]) for kv in arg[6:].split(",")})
    else:
        settings.append({kv.split("=")[0]: int(kv.split("=")[1]) for kv in arg.split(",")})
metric = Ohnsw.distance_l2 if a.metric == "l2" else Ohnsw.distance_angular
t = time.time()
X, Q = bench.make_data(a, a.n, 1234), bench.make_data(a, a.nq, 4321)
print(f"data {time.time() - t:.1f}s", flush=True)
h = Ohnsw.Hgraph(a.dim, metric, a.M, a.efc)
for name, v in build_params.items():
    h.set_param(name, v)
t = time.time()
H.capi.check(H.capi.lib().hnswb200_build(h._h, H.capi.ptr(X), a.n, H.capi.ptr(bench.draw_levels(a.n, a.M, 7))))
st = h.stats()
print(f"build {time.time() - t:.2f}s lib {st.build_seconds:.2f}s ndist/insert {st.build_n_dist / a.n:.0f} "
      f"dropped_incoming {st.build_dropped_incoming} GB/s {st.build_algorithmic_bytes / st.build_seconds / 1e9:.0f}", flush=True)
t = time.time()
gt, _ = H.brute_force_knn_l2(X, Q, a.k, return_ids=True, metric=metric)
print(f"ground truth {time.time() - t:.2f}s", flush=True)
ef_star = None
lo, hi = a.k, None
for ef in bench.EF_SWEEP:
    if ef < a.k:
        continue
    ids, _ = Ohnsw.knn_batch_bigarray(h, Q, k=a.k, ef=ef)
    r = H.Recall.ids(gt, ids)
    if r >= 0.95:
        hi = ef
        break
    lo = ef
while hi is not None and hi - lo > 1:
    mid = (lo + hi) // 2
    ids, _ = Ohnsw.knn_batch_bigarray(h, Q, k=a.k, ef=mid)
    if H.Recall.ids(gt, ids) >= 0.95:
        hi = mid
    else:
        lo = mid
ef_star = hi or bench.EF_SWEEP[-1]
print(f"ef* = {ef_star}", flush=True)
Ohnsw.knn_batch_bigarray(h, Q, k=a.k, ef=ef_star)
c = h.last_search_counters(a.nq)[:, 0].astype(np.float64)
print(f"ndist per query: mean {c.mean():.0f} p50 {np.percentile(c, 50):.0f} p99 {np.percentile(c, 99):.0f} max {c.max():.0f} (max/mean {c.max() / c.mean():.2f})", flush=True)


def run(reps=5):
    ms = []
    for _ in range(reps):
        Ohnsw.knn_batch_bigarray(h, Q, k=a.k, ef=ef_star)
        s = h.stats()
        ms.append(s.search_kernel_ms)
    return min(ms), float(np.mean(ms)), s


known = ("stage_rows", "max_warps_per_sm", "visited_mode", "hash_slots", "warps_per_cta", "hash_bits")
defaults = {"stage_ahead": -1}
for setting in [{}] + settings:
    for name in known + tuple(defaults):
        h.set_param(name, setting.get(name, defaults.get(name, 0)))
    ms, avg, s = run()
    gbs = s.search_algorithmic_bytes / ms / 1e6
    print(f"{setting or 'default'}: kernel_ms min {ms:.3f} avg {avg:.3f}  {gbs:.0f} GB/s  frac {gbs / 6553:.3f}  "
          f"ndist/q {s.search_n_dist / a.nq:.0f} spills {s.search_visited_overflows}", flush=True)
