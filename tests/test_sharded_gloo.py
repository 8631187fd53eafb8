"""Host-side logic of the multi-GPU path (ocaml-hnsw_b200/sharded.py) on CPU: world_size 2, gloo.

What runs on the GPU in production (per-shard search, merge kernel) is replaced HERE, in the
test only, by a numpy exact top-k per shard and a numpy merge — the point is the partition, the
all-gather layout `[world][nq][k]` the merge kernel consumes, and the shard offsets that turn
shard-local ids into global ids."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.environ["REPO_ROOT"])
    from ocaml_hnsw_b200.sharded import shard_range, shard_offsets, gather_rows

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n, nq, dim, k = 1001, 37, 12, 5                      # odd n: ragged shards
    rng = np.random.default_rng(0)
    X = rng.random((n, dim), dtype=np.float32)
    Q = rng.random((nq, dim), dtype=np.float32)
    lo, hi = shard_range(n, rank, world)
    assert shard_offsets(n, world)[rank] == lo
    cover = [shard_range(n, r, world) for r in range(world)]
    assert cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
    d = ((Q[:, None, :] - X[None, lo:hi, :]) ** 2).sum(-1)
    ids_l = np.argsort(d, axis=1, kind="stable")[:, :k].astype(np.int32)            # shard-local ids
    d_l = np.take_along_axis(d, ids_l, axis=1).astype(np.float32)
    all_ids = gather_rows(torch.from_numpy(ids_l), world).numpy()
    all_d = gather_rows(torch.from_numpy(d_l), world).numpy()
    assert all_ids.shape == (world, nq, k) and all_d.shape == (world, nq, k)
    assert np.array_equal(all_ids[rank], ids_l) and np.array_equal(all_d[rank], d_l)
    # what the merge kernel does: global id = local id + shard offset, k best by (distance, id)
    offs = shard_offsets(n, world)
    gid = all_ids.astype(np.int64) + offs[:, None, None]
    flat_d = all_d.transpose(1, 0, 2).reshape(nq, -1); flat_i = gid.transpose(1, 0, 2).reshape(nq, -1)
    order = np.lexsort((flat_i, flat_d), axis=1)[:, :k]
    merged = np.take_along_axis(flat_i, order, axis=1)
    dfull = ((Q[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    want = np.argsort(dfull, axis=1, kind="stable")[:, :k]
    assert np.array_equal(merged, want), "merged shard results differ from the exact global top-k"
    dist.barrier()
    dist.destroy_process_group()
    print("rank", rank, "ok")
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_shard_gather_merge_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), REPO_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    outs = [p.communicate(timeout=240)[0].decode() for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o}"
        assert f"rank {r} ok" in o


def test_shard_ranges_partition():
    from ocaml_hnsw_b200.sharded import shard_offsets, shard_range
    for n in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            rs = [shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert max(hi - lo for lo, hi in rs) - min(hi - lo for lo, hi in rs) <= 1
            assert shard_offsets(n, world).tolist() == [lo for lo, _ in rs]


def test_sharded_build_rejects_wrong_rows():
    from ocaml_hnsw_b200.sharded import ShardedHgraph
    with pytest.raises(ValueError, match="expected rows"):
        ShardedHgraph.build(0, np.zeros((10, 4), np.float32), 100, num_connections=4, num_nodes_search_construction=10,
                            rank=0, world=2)


def test_shards_times_replicas_partition():
    """S row shards x R replicas (sharded.py): every replica group covers the rows exactly once, the replica groups
    cover the queries exactly once, and the first rank of a group is its home."""
    import sys
    sys.path.insert(0, ROOT)
    from ocaml_hnsw_b200.sharded import ShardedHgraph, shard_range

    class FakeLocal:
        pass
    for world in (1, 2, 4, 8):
        for R in (r for r in (1, 2, 4, 8) if world % r == 0):
            S = world // R
            n, nq = 1001, 37
            rows = [ShardedHgraph.rows_of(n, r, world, R) for r in range(world)]
            for g in range(R):
                group = rows[g * S:(g + 1) * S]
                assert group[0][0] == 0 and group[-1][1] == n and all(a[1] == b[0] for a, b in zip(group, group[1:]))
                assert group == rows[:S]                                 # every group holds the same shards
            hs = [ShardedHgraph(FakeLocal(), n, r, world, replicas=R) for r in range(world)]
            slices = sorted({h.query_slice(nq) for h in hs})
            assert slices[0][0] == 0 and slices[-1][1] == nq and all(a[1] == b[0] for a, b in zip(slices, slices[1:]))
            assert len(slices) == R
            for r, h in enumerate(hs):
                assert (h.shard, h.replica) == (r % S, r // S) and h.query_slice(nq) == shard_range(nq, r // S, R)
                assert len(h.offsets) == S and h.offsets[h.shard] == rows[r][0]
    with pytest.raises(ValueError):
        ShardedHgraph(FakeLocal(), 10, 0, 4, replicas=3)
