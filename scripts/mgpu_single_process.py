"""ONE process, N GPUs through hnswb200_sharded_* (no torch.distributed): exactness of the fused exchange + merge
against a host merge of per-shard searches, then step times (host-buffer call and device-resident queries).

    python scripts/mgpu_single_process.py [n_gpus] [rows] [nq] [ef]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from ocaml_hnsw_b200.sharded import MultiGpuHgraph
from bench import draw_levels

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
ef = int(sys.argv[4]) if len(sys.argv) > 4 else 32
k = 10
X = H.sift_like(n, 128, seed=1234)
Q = H.sift_like(nq, 128, seed=4321)
t = time.perf_counter()
m = MultiGpuHgraph.build_batch_bigarray(Ohnsw.distance_l2, X, num_connections=16, num_nodes_search_construction=200,
                                        devices=list(range(G)), levels=draw_levels(n, 16, 7))
print(f"build on {G} GPUs: {time.perf_counter() - t:.2f} s (slowest shard {m.stats().build_seconds:.2f} s)", flush=True)
ids, d = m.knn_batch_bigarray(Q, k=k, ef=ef)
per, offs = [], []
for i in range(G):
    h, first = m.shard(i)
    per.append(Ohnsw.knn_batch_bigarray(h, Q, k=k, ef=ef))
    offs.append(first)
gid = np.concatenate([np.where(i >= 0, i.astype(np.int64) + o, np.int64(1) << 40) for (i, _), o in zip(per, offs)], 1)
gd = np.concatenate([np.where(i >= 0, dd, np.float32(np.inf)) for i, dd in per], 1)
order = np.lexsort((gid, gd), axis=1)[:, :k]
want = np.take_along_axis(gid, order, 1)
print("merged ids equal the exact host merge:", bool(np.array_equal(ids.astype(np.int64), want)), flush=True)
assert np.array_equal(ids.astype(np.int64), want)
gt, _ = H.brute_force_knn_l2(X, Q, k, return_ids=True)
print(f"recall@10 at ef={ef}: {H.Recall.ids(gt, ids):.4f}", flush=True)
out = (np.empty((nq, k), np.int32), np.empty((nq, k), np.float32))
for buf in (Q,) + out:
    capi.host_register(buf)
for path in (0, 1):
    m.set_param("query_path", path)
    for _ in range(3):
        m.knn_batch_bigarray(Q, k=k, ef=ef, out=out)
    t = time.perf_counter()
    for _ in range(20):
        m.knn_batch_bigarray(Q, k=k, ef=ef, out=out)
    dt = (time.perf_counter() - t) / 20
    print(f"host-buffer call, query_path={path}: {dt * 1e3:.3f} ms/step = {nq / dt / 1e6:.2f} M queries/s; device step {m.stats().search_kernel_ms:.3f} ms", flush=True)
import torch
q = torch.from_numpy(Q).cuda(0)
ti = torch.empty((nq, k), dtype=torch.int32, device="cuda:0")
td = torch.empty((nq, k), dtype=torch.float32, device="cuda:0")
s = torch.cuda.Stream(device=0)
with torch.cuda.stream(s):
    for _ in range(3):
        m.search_device(q.data_ptr(), nq, k, ef, ti.data_ptr(), td.data_ptr(), stream=s.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        m.search_device(q.data_ptr(), nq, k, ef, ti.data_ptr(), td.data_ptr(), stream=s.cuda_stream)
    e1.record()
s.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"device-resident queries: {ms:.3f} ms/step = {nq / ms / 1e3:.2f} M queries/s", flush=True)
assert np.array_equal(ti.cpu().numpy(), ids)
