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


def test_partition_by_cost_is_balanced_and_complete():
    from limap_b200.dist import partition_by_cost, slice_tracks
    rng = np.random.default_rng(0)
    cost = rng.integers(2, 60, 1001)
    for world in (1, 2, 3, 8):
        parts = partition_by_cost(cost, world)
        allidx = np.concatenate(parts)
        assert len(parts) == world and np.array_equal(np.sort(allidx), np.arange(1001))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
        tot = [cost[p].sum() for p in parts]
        assert max(tot) - min(tot) <= 60
        assert all(np.all(np.diff(p) > 0) for p in parts)
    sup_off = np.concatenate([[0], np.cumsum(cost)])
    flat = np.arange(sup_off[-1])
    off, (sub,) = slice_tracks(np.array([3, 10]), sup_off, flat)
    assert off.tolist() == [0, cost[3], cost[3] + cost[10]]
    assert np.array_equal(sub, np.concatenate([flat[sup_off[3]:sup_off[4]], flat[sup_off[10]:sup_off[11]]]))


def test_partition_ranges_by_cost_are_contiguous_and_balanced():
    from limap_b200.dist import partition_ranges_by_cost
    rng = np.random.default_rng(1)
    cost = rng.integers(2, 60, 1001)
    for world in (1, 2, 3, 8):
        r = partition_ranges_by_cost(cost, world)
        assert len(r) == world and r[0][0] == 0 and r[-1][1] == 1001
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        tot = [cost[b:e].sum() for b, e in r]
        assert max(tot) - min(tot) <= 2 * 60
    assert partition_ranges_by_cost(np.zeros(0), 3) == [(0, 0)] * 3
    assert partition_ranges_by_cost([5.0], 2) in ([(0, 0), (0, 1)], [(0, 1), (1, 1)])


def _ba_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from limap_b200.dist import solve_line_ba_sharded
    rng = np.random.default_rng(5)
    T = 37
    cnt = rng.integers(2, 9, T)
    sup_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    S = int(sup_off[-1])
    sup_view = rng.integers(0, 5, S).astype(np.int32)
    segs = rng.uniform(0, 100, (S, 4))
    line3d = rng.normal(size=(S, 6))
    init = rng.normal(size=(T, 6))
    seen = {}

    def fake_solve(kvec, qvec, tvec, off, sv, sg, l3, li, **kw):  # stands in for BAEngine.solve (needs a GPU)
        n = len(off) - 1
        seen["n"] = n
        seen["s"] = int(off[-1])
        # a function of the track's own data only: line = init + sum of its segments' first coordinate
        s = np.array([sg[off[t]:off[t + 1], 0].sum() for t in range(n)])
        return dict(line=li + s[:, None], iters=np.stack([off[1:] - off[:-1], np.zeros(n, np.int64)], 1).astype(np.int32),
                    cost=np.stack([s, 0.5 * s], 1))
    out = solve_line_ba_sharded(fake_solve, None, None, None, sup_off, sup_view, segs, line3d, init, rank, world,
                                device="cpu")
    s_all = np.array([segs[sup_off[t]:sup_off[t + 1], 0].sum() for t in range(T)])
    ok = np.allclose(out["line"], init + s_all[:, None]) and np.array_equal(out["iters"][:, 0], cnt)
    ok &= np.allclose(out["cost"][:, 1], 0.5 * s_all) and abs(seen["s"] - S / world) <= 9  # shares are contiguous ranges with equal support counts
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_line_ba_gather_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_ba_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def _vp_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from limap_b200.dist import detect_vps_sharded
    from limap_b200.synth import make_vp_images
    from oracle import oracle as orc
    imgs = make_vp_images(5, 60, seed=3) + [np.zeros((0, 4))]

    class R:  # VPResult-like
        def __init__(self, labels, vps):
            self.labels, self.vps = labels, vps

    def cpu_detect(segs_list, image_index=None):  # stands in for JLinkageDetector.detect_batch (needs a GPU)
        off = np.concatenate([[0], np.cumsum([len(s) for s in segs_list])]).astype(np.int64)
        segs = np.concatenate(segs_list, 0) if len(segs_list) else np.zeros((0, 4))
        lab, vo, vps = orc.detect_vps(off, segs, min_length=40, inlier_threshold=1.0, min_num_supports=5, seed=11,
                                      n_models=500, threads=1, image_index=image_index)
        return [R(lab[off[i]:off[i + 1]], vps[vo[i]:vo[i + 1]]) for i in range(len(segs_list))]
    labels, vps = detect_vps_sharded(cpu_detect, imgs, rank, world, device="cpu")
    ref = cpu_detect(imgs)  # the single call on all images
    ok = all(np.array_equal(labels[i], np.asarray(ref[i].labels, np.int32)) for i in range(len(imgs)))
    ok &= all(np.allclose(vps[i], np.asarray(ref[i].vps).reshape(-1, 3)) for i in range(len(imgs)))
    ok &= any(len(v) > 0 for v in vps)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_jlinkage_identical_to_single_call_gloo_world2():
    from oracle import oracle as orc
    orc.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_vp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
