"""Line-BA micro-benchmark (BASELINE.json configs[3]: 10k tracks x 30 supporting views): LM iterations/s of
the batched CUDA solver, and the CPU restatement beside it when --cpu is given."""
import argparse
import json
import os
import sys
import time

import numpy as np

R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)

ap = argparse.ArgumentParser()
ap.add_argument("--tracks", type=int, default=10000)
ap.add_argument("--supports", type=int, default=30)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cpu", action="store_true")
a = ap.parse_args()
from limap_b200.engine import BAEngine
from limap_b200.synth import make_tracks
ts = make_tracks(T=a.tracks, S=a.supports, V=300, seed=1237)
eng = BAEngine()
out = None
for _ in range(a.reps):
    t0 = time.perf_counter()
    out = eng.solve_trackset(ts, max_num_iterations=a.iters)
    wall = time.perf_counter() - t0
st = out["stats"]
res = {"tracks": a.tracks, "supports": a.supports, "max_iters": a.iters, "total_iterations": st["total_iterations"],
       "solve_ms": st["solve_ms"], "prepare_ms": st["prepare_ms"], "wall_ms": wall * 1e3,
       "lm_iters_per_s_kernel": st["total_iterations"] / (st["solve_ms"] * 1e-3),
       "lm_iters_per_s_e2e": st["total_iterations"] / wall}
if a.cpu:
    from oracle import oracle as orc
    n = min(a.tracks, 2000)
    sub = make_tracks(T=n, S=a.supports, V=300, seed=1237)
    t0 = time.perf_counter()
    o = orc.refine_tracks(sub, max_num_iterations=a.iters, threads=orc.usable_cpus())
    dt = time.perf_counter() - t0
    res["cpu_lm_iters_per_s"] = float(o["iters"][:, 0].sum() / dt)
    res["cpu_threads"] = orc.usable_cpus()
    res["cpu_sample_tracks"] = n
print(json.dumps(res))
