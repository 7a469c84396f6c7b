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
    """Exchange of every rank's per-node results (96-byte node records + valid connections) so that ComputeLineTracks
    can run replicated (or on rank 0): pack (one kernel) -> ONE all_gather_into_tensor of a fixed-size message ->
    unpack (one kernel). The message capacity for valid connections is agreed once (a counts all-reduce on the first
    call, with head-room) and reused; steady-state calls do not synchronise the host. `check()` (called by
    lm_tri_build_tracks too) reports an overflow, in which case the exchange is repeated with a larger message."""

    def __init__(self, eng, world, rank, shards=None, group=None):
        import torch
        from ._cabi import lib
        self.eng, self.world, self.rank, self.group = eng, world, rank, group
        self.lib = lib()
        # pack / unpack kernels are enqueued on the engine's stream and the collective on torch's current stream:
        # make them the same stream so that stream order is the only synchronisation needed
        eng.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        V = len(eng.img_ids)
        if shards is None:
            per = V // world
            shards = [(r * per, (r + 1) * per if r < world - 1 else V) for r in range(world)]
        self.shards = shards
        self.node_rng = [(int(eng.line_off[b]), int(eng.line_off[e])) for b, e in shards]
        self.node_begin = np.ascontiguousarray([b for b, _ in self.node_rng], np.int64)
        self.max_nodes = max(e - b for b, e in self.node_rng)
        self.cap_edges = 0
        self.send = self.recv = None

    def _size(self, cap_edges):
        import torch
        self.cap_edges = int(cap_edges)
        self.msg_bytes = int(self.lib.lm_tri_gather_message_bytes(self.max_nodes, self.cap_edges))
        self.send = torch.zeros(self.msg_bytes, dtype=torch.uint8, device="cuda")
        self.recv = torch.zeros(self.world * self.msg_bytes, dtype=torch.uint8, device="cuda")

    def _agree_capacity(self):
        """Largest valid-connection count over ranks (one tiny all-reduce + host read, first call only)."""
        import torch
        import torch.distributed as dist
        n = int(self.eng.stats()["n_valid_edges"])
        t = torch.tensor([n], dtype=torch.int64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return int(t.item())

    def all_gather(self):
        import torch.distributed as dist
        from ._cabi import check
        h = self.eng.ctx.handle
        if self.send is None:
            self._size(max(1024, int(self._agree_capacity() * 1.25) + 1024))
        check(self.lib.lm_tri_pack_message(h, self.max_nodes, self.cap_edges, C.c_void_p(self.send.data_ptr())))
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        check(self.lib.lm_tri_unpack_messages(h, self.world, self.node_begin.ctypes.data_as(C.c_void_p), self.max_nodes,
                                              self.cap_edges, C.c_void_p(self.recv.data_ptr())))

    def check(self):
        """Synchronise; if some rank held more valid connections than the message capacity, grow it and repeat the
        exchange. Returns the number of directed valid connections now held by the engine."""
        from ._cabi import check
        tot = C.c_int64(0)
        over = check(self.lib.lm_tri_gather_status(self.eng.ctx.handle, C.byref(tot)))
        if over:
            self._size(max(self.cap_edges * 2, int(self._agree_capacity() * 1.25) + 1024))
            self.all_gather()
            over = check(self.lib.lm_tri_gather_status(self.eng.ctx.handle, C.byref(tot)))
            assert not over
        return int(tot.value)


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


def partition_ranges_by_cost(costs, world):
    """Contiguous ranges [begin, end) of units for `world` ranks with nearly equal total cost (boundaries where the
    running cost crosses k / world of the total). Contiguous shares of flat per-support arrays are views, not copies."""
    c = np.asarray(costs, dtype=np.float64)
    n = len(c)
    cum = np.concatenate([[0.0], np.cumsum(c)])
    cuts = [0]
    for k in range(1, max(world, 1)):
        b = int(np.searchsorted(cum, cum[-1] * k / world, side="left"))
        cuts.append(min(max(b, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(max(world, 1))]


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
    # concatenated index ranges [sup_off[t], sup_off[t + 1]) of the chosen tracks, without a Python loop
    sel = (np.repeat(sup_off[index] - new_off[:-1], counts) + np.arange(int(new_off[-1]), dtype=np.int64)) if len(index) \
        else np.zeros(0, np.int64)
    return new_off, [np.ascontiguousarray(np.asarray(a)[sel]) for a in per_support]


def solve_line_ba_sharded(solve, kvec, qvec, tvec, sup_off, sup_view, segs, line3d, line_init, rank, world, group=None,
                          device="cuda", **kw):
    """Line BA with constant cameras is block-separable per track (hybrid_bundle_adjustment.cc:106-123,156-197):
    every rank solves a contiguous range of tracks with the same number of supports as the others (the share is a view
    of the caller's flat arrays: nothing is re-packed on the host; inside a rank the kernel's warps fetch tracks
    dynamically) with `solve` (BAEngine.solve), and the refined lines (+ iteration counts and costs) are all-gathered.
    Returns dict(line[T,6], iters[T,2], cost[T,2]) on every rank."""
    import torch
    sup_off = np.asarray(sup_off, dtype=np.int64)
    T = len(sup_off) - 1
    b, e = partition_ranges_by_cost(sup_off[1:] - sup_off[:-1], world)[rank]
    s0, s1 = int(sup_off[b]), int(sup_off[e])
    off = sup_off[b:e + 1] - s0
    res = solve(kvec, qvec, tvec, off, np.asarray(sup_view)[s0:s1], np.asarray(segs)[s0:s1], np.asarray(line3d)[s0:s1],
                np.asarray(line_init)[b:e], **kw)
    table = np.concatenate([res["line"], res["iters"].astype(np.float64), res["cost"]], axis=1) if e > b else \
        np.zeros((0, 10))
    full = gather_rows(np.arange(b, e, dtype=np.int64), torch.as_tensor(table, dtype=torch.float64, device=device), T,
                       group).cpu().numpy()
    return dict(line=full[:, :6], iters=full[:, 6:8].astype(np.int32), cost=full[:, 8:10])


def detect_vps_sharded(detect_batch, segs_list, rank, world, group=None, device="cuda"):
    """J-Linkage over the images of a scene, sharded per image (vplib/base_vp_detector.py:46-78 fans images out with
    joblib): images are dealt by segment count, every rank clusters its share with `detect_batch(images,
    image_index=...)` -- the index seeds each image's hypotheses, so the labels are those of the single-rank call --
    and labels + VPs are all-gathered. Returns (labels[n_images] int32 arrays, vps[n_images] (k,3) arrays) on every rank."""
    import torch
    n = len(segs_list)
    counts = np.asarray([len(s) for s in segs_list], np.int64)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    mine = partition_by_cost(counts, world)[rank]
    res = detect_batch([segs_list[i] for i in mine], image_index=mine) if len(mine) else []
    lab = np.concatenate([np.asarray(r.labels, np.int64) for r in res]) if len(res) else np.zeros(0, np.int64)
    nv = np.asarray([len(np.asarray(r.vps).reshape(-1, 3)) for r in res], np.int64)
    vps = np.concatenate([np.asarray(r.vps, np.float64).reshape(-1, 3) for r in res]) if len(res) else np.zeros((0, 3))
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    parts_i = all_gather_padded(t(mine, torch.int64), group)
    parts_l = all_gather_padded(t(lab, torch.int64), group)
    parts_n = all_gather_padded(t(nv, torch.int64), group)
    parts_v = all_gather_padded(t(vps.reshape(-1), torch.float64), group)
    labels, out_vps = [None] * n, [None] * n
    for pi, pl, pn, pv in zip(parts_i, parts_l, parts_n, parts_v):
        pi, pl, pn, pv = pi.cpu().numpy(), pl.cpu().numpy(), pn.cpu().numpy(), pv.cpu().numpy().reshape(-1, 3)
        lo = vo = 0
        for k, i in enumerate(pi):
            labels[i] = pl[lo:lo + counts[i]].astype(np.int32)
            out_vps[i] = pv[vo:vo + pn[k]].copy()
            lo += counts[i]
            vo += pn[k]
    return labels, out_vps
