"""Wall-clock breakdown of one end-to-end step through the public API (host buffers -> results on host)."""
import os
import sys
import time

import numpy as np
import torch

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
from limap_b200._cabi import NODE_RECORD_DTYPE
from limap_b200.config import DEFAULT_YAML_TRIANGULATION
from limap_b200.engine import TriEngine
from limap_b200.synth import CONFIGS, make_scene

sc = make_scene(**CONFIGS["hypersim100"])
bsrc, bng, boff, bpairs = sc.bulk_matches()
tp = torch.empty(bpairs.shape, dtype=torch.int32, pin_memory=True)
tp.numpy()[...] = bpairs
pp = tp.numpy()
tsegs = torch.empty(sc.segs.shape, dtype=torch.float64, pin_memory=True)
tsegs.numpy()[...] = sc.segs
sc.segs = tsegs.numpy()
eng = TriEngine(dict(DEFAULT_YAML_TRIANGULATION))
nodes_out = torch.empty(int(sc.line_off[-1]) * NODE_RECORD_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True).numpy().view(NODE_RECORD_DTYPE)
off_out = torch.empty(int(sc.line_off[-1]) + 1, dtype=torch.int64, pin_memory=True).numpy()
edges_out = torch.empty((4000000, 2), dtype=torch.int32, pin_memory=True).numpy()
acc = {}
def tick(name, t0):
    torch.cuda.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)
for it in range(6):
    t = time.perf_counter(); eng.upload(sc); eng.set_ranges(*sc.ranges); tick("upload_scene", t)
    t = time.perf_counter(); eng.add_matches_bulk(bsrc, bng, boff, pp); eng.ctx.synchronize(); tick("add_matches(H2D 160MB)", t)
    t = time.perf_counter(); st = eng.run(); tick("run", t)
    t = time.perf_counter(); eng.get_nodes(nodes_out); tick("get_nodes", t)
    t = time.perf_counter(); eng.get_all_valid_edges(off_out, edges_out); tick("get_edges", t)
for it in range(5):
    eng.upload(sc); eng.set_ranges(*sc.ranges); torch.cuda.synchronize()
    t = time.perf_counter(); eng.add_matches_bulk(bsrc, bng, boff, pp); t1 = time.perf_counter(); st = eng.run(); tick("add+run (pipelined)", t)
    acc.setdefault("  add call returns after", []).append((t1 - t) * 1e3)
for k, v in acc.items():
    print(f"{k:28s} {np.median(v[2:]):8.3f} ms")
print("device run ms", st["last_run_ms"], "kernel", st["last_node_kernel_ms"])

# ---- full public-API step (scene up, matches up, run, results down) for several pipeline group counts ----
for groups in (1, 2, 4, 6, 8, 12):
    eng.set_pipeline_groups(groups)
    tot, dev, ker = [], [], []
    for it in range(7):
        torch.cuda.synchronize()
        t = time.perf_counter()
        eng.upload(sc); eng.set_ranges(*sc.ranges)
        eng.add_matches_bulk(bsrc, bng, boff, pp)
        st = eng.run(nodes_out=nodes_out)
        eng.get_all_valid_edges(off_out, edges_out)
        tot.append((time.perf_counter() - t) * 1e3)
        dev.append(st["last_run_ms"]); ker.append(st["last_node_kernel_ms"])
    print(f"groups={groups}: e2e step {np.median(tot[2:]):.3f} ms; device run {np.median(dev[2:]):.3f} ms; node kernel {np.median(ker[2:]):.3f} ms")
