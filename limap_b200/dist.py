"""Multi-GPU plumbing of the triangulation path: one process per GPU, the scene is
replicated, source images are sharded in contiguous blocks balanced by match rows, and the per-node
results (best candidate record + valid connections) are exchanged with one all-gather per stage.
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the transport; no data-path collective runs
inside the kernels because the path has no reduction, only this exchange."""
import ctypes as C

import numpy as np


def partition_views(weights, world):
    """Contiguous blocks of views, balanced by weight (match rows per source image).
    Returns [(begin, end)] * world covering [0, len(weights))."""
    w = np.asarray(weights, dtype=np.float64)
    n = len(w)
    if world <= 1:
        return [(0, n)]
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        k = int(np.searchsorted(cum, target, side="left"))
        # pick the closer boundary
        if k > 0 and abs(cum[k - 1] - target) <= abs(cum[min(k, n)] - target):
            k -= 1
        k = min(max(k, cuts[-1]), n)
        cuts.append(k)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def all_gather_padded(t, group=None):
    """Variable-length all-gather of 1-D tensors: a counts exchange followed by one all_gather of padded
    buffers. Returns the list of per-rank tensors (views into one buffer)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(max(counts), 1)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = torch.empty(world * m, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return [out[r * m: r * m + counts[r]] for r in range(world)]


class NodeGather:
    """All-gather of the per-node records (96 B each) and of the valid-connection pairs of every rank's
    shard, imported into the local engine so that ComputeLineTracks can run replicated (or on rank 0)."""

    def __init__(self, eng, world, rank, shards=None):
        import torch
        from ._cabi import NODE_RECORD_DTYPE, lib
        self.eng, self.world, self.rank = eng, world, rank
        self.lib = lib()
        V = len(eng.img_ids)
        if shards is None:
            per = V // world
            shards = [(r * per, (r + 1) * per if r < world - 1 else V) for r in range(world)]
        self.shards = shards
        self.node_rng = [(int(eng.line_off[b]), int(eng.line_off[e])) for b, e in shards]
        self.rec = NODE_RECORD_DTYPE.itemsize
        self.max_nodes = max(e - b for b, e in self.node_rng)
        self.send = torch.zeros(self.max_nodes * self.rec, dtype=torch.uint8, device="cuda")
        self.recv = torch.zeros(world * self.max_nodes * self.rec, dtype=torch.uint8, device="cuda")

    def all_gather(self, with_edges=True):
        import torch
        import torch.distributed as dist
        from ._cabi import check
        h = self.eng.ctx.handle
        b, e = self.node_rng[self.rank]
        check(self.lib.lm_tri_export_nodes(h, b, e, C.c_void_p(self.send.data_ptr())))
        dist.all_gather_into_tensor(self.recv, self.send)
        for r in range(self.world):
            if r == self.rank:
                continue
            rb, re_ = self.node_rng[r]
            off = r * self.max_nodes * self.rec
            check(self.lib.lm_tri_import_nodes(h, rb, re_, C.c_void_p(self.recv.data_ptr() + off)))
        if with_edges:
            n = check(self.lib.lm_tri_num_valid_edges(h))
            mine = torch.empty(max(2 * n, 1), dtype=torch.int64, device="cuda")
            if n:
                check(self.lib.lm_tri_export_edges(h, C.c_void_p(mine.data_ptr())))
            parts = all_gather_padded(mine[: 2 * n])
            for r, p in enumerate(parts):
                if r == self.rank or p.numel() == 0:
                    continue
                p = p.contiguous()
                check(self.lib.lm_tri_import_edges(h, p.numel() // 2, C.c_void_p(p.data_ptr()), 1))
        torch.cuda.current_stream().synchronize()


# ---- track-sharded stages (LM refinement with constant cameras, J-Linkage per image) -------------
def partition_by_cost(costs, world):
    """Units (tracks by #supports, images by #segments) dealt to `world` ranks: sorted by cost descending and dealt
    in snake order, so every rank gets the same count (+-1) and nearly the same total cost. Returns a list of
    ascending index arrays; deterministic on every rank."""
    c = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-c, kind="stable")
    out = [[] for _ in range(max(world, 1))]
    for k, idx in enumerate(order):
        r = k % world
        if (k // world) % 2 == 1:
            r = world - 1 - r
        out[r].append(int(idx))
    return [np.asarray(sorted(x), dtype=np.int64) for x in out]


def gather_rows(local_index, local_rows, total, group=None):
    """Every rank holds rows of a [total, k] table for its own indices; returns the full table on every rank (one
    padded all-gather of the indices and one of the rows)."""
    import torch
    idx = torch.as_tensor(local_index, dtype=torch.int64, device=local_rows.device).reshape(-1)
    rows = local_rows.reshape(idx.numel(), -1) if idx.numel() else local_rows.reshape(0, local_rows.shape[-1])
    k = rows.shape[1]
    parts_i = all_gather_padded(idx, group)
    parts_r = all_gather_padded(rows.reshape(-1).contiguous(), group)
    out = torch.zeros((int(total), k), dtype=local_rows.dtype, device=local_rows.device)
    for pi, pr in zip(parts_i, parts_r):
        if pi.numel():
            out[pi] = pr.reshape(pi.numel(), k)
    return out


def slice_tracks(index, sup_off, *per_support):
    """Sub-problem of the tracks in `index`: (sup_off', per-support arrays restricted and re-packed)."""
    sup_off = np.asarray(sup_off, dtype=np.int64)
    counts = (sup_off[1:] - sup_off[:-1])[index]
    new_off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    sel = np.concatenate([np.arange(sup_off[t], sup_off[t + 1]) for t in index]) if len(index) else np.zeros(0, np.int64)
    return new_off, [np.ascontiguousarray(np.asarray(a)[sel]) for a in per_support]


def solve_line_ba_sharded(solve, kvec, qvec, tvec, sup_off, sup_view, segs, line3d, line_init, rank, world, group=None,
                          device="cuda", **kw):
    """Line BA with constant cameras is block-separable per track (hybrid_bundle_adjustment.cc:106-123,156-197):
    tracks are dealt by #supports, every rank solves its share with `solve` (BAEngine.solve) and the refined lines
    (+ iteration counts and costs) are all-gathered. Returns dict(line[T,6], iters[T,2], cost[T,2]) on every rank."""
    import torch
    sup_off = np.asarray(sup_off, dtype=np.int64)
    T = len(sup_off) - 1
    mine = partition_by_cost(sup_off[1:] - sup_off[:-1], world)[rank]
    off, (sv, sg, l3) = slice_tracks(mine, sup_off, sup_view, segs, line3d)
    res = solve(kvec, qvec, tvec, off, sv, sg, l3, np.ascontiguousarray(np.asarray(line_init)[mine]), **kw)
    table = np.concatenate([res["line"], res["iters"].astype(np.float64), res["cost"]], axis=1) if len(mine) else \
        np.zeros((0, 10))
    full = gather_rows(mine, torch.as_tensor(table, dtype=torch.float64, device=device), T, group).cpu().numpy()
    return dict(line=full[:, :6], iters=full[:, 6:8].astype(np.int32), cost=full[:, 8:10])
