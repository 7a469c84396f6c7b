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
