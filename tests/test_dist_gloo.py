"""CPU, world_size 2, gloo: the N > 1 plumbing (single arena broadcast, batch sharding, max-over-ranks timing)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from consistentid_b200 import dist as cdist
    r, l, w = cdist.init_from_env(backend="gloo")
    arena = torch.arange(1000, dtype=torch.float16) if r == 0 else torch.zeros(1000, dtype=torch.float16)
    cdist.broadcast_arena(arena, src=0)
    ok_bcast = bool(torch.equal(arena, torch.arange(1000, dtype=torch.float16)))
    s, e = cdist.shard_batch(5, r, w)
    mx = cdist.max_over_ranks(float(10 + r), device="cpu")
    cdist.barrier()
    q.put((r, ok_bcast, (s, e), mx))
    dist.destroy_process_group()


def test_two_rank_plumbing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    assert [r[1] for r in res] == [True, True]                  # weights bit-identical after the single broadcast
    assert [r[2] for r in res] == [(0, 3), (3, 5)]              # contiguous, disjoint, covering shards
    assert [r[3] for r in res] == [11.0, 11.0]                  # max over ranks


def test_shard_edges():
    from consistentid_b200.dist import shard_batch
    assert shard_batch(0, 0, 4) == (0, 0)
    assert [shard_batch(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    assert shard_batch(3, 3, 4) == (3, 3)
