"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard partitioning and the padded
variable-length all-gather used for the per-node result exchange."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_views_balanced_and_contiguous():
    from limap_b200.dist import partition_views
    w = np.array([10, 10, 10, 10, 40, 5, 5, 5, 5, 100], float)
    for world in (1, 2, 3, 4, 8):
        parts = partition_views(w, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == len(w)
        for (a, b), (c, d) in zip(parts[:-1], parts[1:]):
            assert b == c and a <= b
    p2 = partition_views(w, 2)
    assert abs(w[p2[0][0]:p2[0][1]].sum() - 100) <= 10
    assert partition_views(np.ones(100), 4) == [(0, 25), (25, 50), (50, 75), (75, 100)]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from limap_b200.dist import all_gather_padded
    t = torch.arange(3 + 4 * rank, dtype=torch.int64) + 100 * rank
    parts = all_gather_padded(t)
    ok = len(parts) == world
    for r, p in enumerate(parts):
        ok &= torch.equal(p, torch.arange(3 + 4 * r, dtype=torch.int64) + 100 * r)
    e = all_gather_padded(torch.zeros(0, dtype=torch.int64) if rank == 0 else torch.ones(2, dtype=torch.int64))
    ok &= e[0].numel() == 0 and e[1].numel() == 2
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_all_gather_padded_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
