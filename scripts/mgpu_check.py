"""torchrun check of the sharded path: peer-store exchange vs packed NCCL all-gather give identical
merged rows, and those equal an exact merge of per-shard results done in numpy."""
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ocaml_hnsw_b200 as H
from ocaml_hnsw_b200 import Ohnsw
from ocaml_hnsw_b200.sharded import ShardedHgraph, shard_range
from bench import draw_levels

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
n, nq, k, ef = 200003, 3000, 10, 32
X = H.sift_like(n, 128, seed=1234); Q = H.sift_like(nq, 128, seed=4321)
lo, hi = shard_range(n, rank, world)
sh = ShardedHgraph.build(Ohnsw.distance_l2, X[lo:hi].copy(), n, num_connections=16, num_nodes_search_construction=100,
                         rank=rank, world=world, levels=draw_levels(hi - lo, 16, 7 + rank), device=lr)
q_dev = torch.from_numpy(Q).to(dev)
res = {}
for mode in ("peer", "nccl"):
    s2 = ShardedHgraph(sh.local, n, rank, world, peer_exchange=(mode == "peer"))
    for rep in range(3):                                   # several steps: exercises the double buffering
        ids, d = s2.knn_batch_device(q_dev, k=k, ef=ef)
        torch.cuda.synchronize()
    res[mode] = (ids.cpu().numpy().copy(), d.cpu().numpy().copy(), s2.exchange)
# reference merge: gather every shard's local result on the host
ids_l, d_l = Ohnsw.knn_batch_bigarray(sh.local, Q, k=k, ef=ef)
gi = [torch.empty((nq, k), dtype=torch.int32, device=dev) for _ in range(world)]
gd = [torch.empty((nq, k), dtype=torch.float32, device=dev) for _ in range(world)]
dist.all_gather(gi, torch.from_numpy(ids_l).to(dev)); dist.all_gather(gd, torch.from_numpy(d_l).to(dev))
allid = np.stack([g.cpu().numpy().astype(np.int64) + shard_range(n, r, world)[0] for r, g in enumerate(gi)], 1).reshape(nq, -1)
alld = np.stack([g.cpu().numpy() for g in gd], 1).reshape(nq, -1)
order = np.lexsort((allid, alld), axis=1)[:, :k]
want = np.take_along_axis(allid, order, axis=1)
ok = True
for mode, (ids, d, ex) in res.items():
    same = np.array_equal(ids.astype(np.int64), want)
    ok &= same
    if rank == 0: print(f"exchange={ex}: merged ids equal the exact merge: {same}", flush=True)
assert np.array_equal(res["peer"][0], res["nccl"][0]) and np.array_equal(res["peer"][1].view(np.uint32), res["nccl"][1].view(np.uint32))
gt, _ = H.brute_force_knn_l2(X, Q, k, device=lr, return_ids=True)
if rank == 0: print("recall@10 of the sharded search:", H.Recall.ids(gt, res["peer"][0]), "all ok" if ok else "MISMATCH", flush=True)
dist.barrier(); dist.destroy_process_group()
assert ok
