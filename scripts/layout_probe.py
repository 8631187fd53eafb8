"""torchrun probe: S row shards x R replicas for every R dividing the world size — build seconds, ef for recall@10 >= 0.95,
device-timed step (queries resident), on the bench workload.   torchrun --nproc-per-node N scripts/layout_probe.py [rows]"""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw, capi
from ocaml_hnsw_b200.sharded import ShardedHgraph, gather_rows
from bench import draw_levels, EF_SWEEP

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nq, k = 10_000, 10
Xall = H.sift_like(n, 128, seed=1234); Q = H.sift_like(nq, 128, seed=4321)
gt, _ = H.brute_force_knn_l2(Xall, Q, k, device=lr, return_ids=True)
q_dev = torch.from_numpy(Q).to(dev)
stream = torch.cuda.Stream(device=dev)
for R in [r for r in (1, 2, 4, 8) if world % r == 0 and r <= world]:
    S = world // R
    lo, hi = ShardedHgraph.rows_of(n, rank, world, R)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    sh = ShardedHgraph.build(Ohnsw.distance_l2, Xall[lo:hi].copy(), n, num_connections=16, num_nodes_search_construction=200,
                             rank=rank, world=world, levels=draw_levels(hi - lo, 16, 7 + rank % S), device=lr, replicas=R)
    torch.cuda.synchronize(); dist.barrier()
    build_s = time.perf_counter() - t0
    def rec_at(ef):
        with torch.cuda.stream(stream):
            ids, _ = sh.knn_batch_device(q_dev, k=k, ef=ef)
        stream.synchronize()
        return H.Recall.ids(gt, ids.cpu().numpy())
    lo_ef, hi_ef = k - 1, None
    for ef in EF_SWEEP:
        if rec_at(ef) >= 0.95: hi_ef = ef; break
        lo_ef = ef
    while hi_ef - lo_ef > 1:
        mid = (lo_ef + hi_ef) // 2
        if rec_at(mid) >= 0.95: hi_ef = mid
        else: lo_ef = mid
    rec = rec_at(hi_ef)
    for _ in range(3): rec_at(hi_ef)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record()
        for _ in range(20): sh.knn_batch_device(q_dev, k=k, ef=hi_ef)
        e1.record()
    stream.synchronize(); dist.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / 20
    if rank == 0:
        print(f"world={world} S={S} R={R}: build {build_s:.2f}s ef={hi_ef} recall={rec:.4f} step {ms:.3f} ms -> {nq / ms / 1e3:.2f} M q/s", flush=True)
    del sh
dist.barrier(); dist.destroy_process_group()
