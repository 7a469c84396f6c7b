#!/usr/bin/env python
"""bench.py — headline benchmark of the line-triangulation hot path (BASELINE.json metric
"3D line candidates triangulated+scored/sec").

A step = one pass of candidate generation + scoring + selection (all TriangulateImage calls of a scene,
SURVEY.md §8d M1) over one synthetic scene of BASELINE.json configs[1] shape ("hypersim100": V=100 views,
L=1000 lines/view, N=20 neighbours, K=10 matches per line per neighbour = 2e7 match rows). The unit
counted is the match row tested by triangulateOneNode.

  value : whole-job rows/s with the scene and match tables already resident in HBM (CUDA-event timed,
          max over ranks).
  e2e   : the same metric through the public engine API with HOST (pinned) buffers: scene upload,
          match upload, run, and the per-node results + valid connections read back, every step.
  roofline : the fused generate+score kernel against the measured HBM copy bandwidth
          (MEASURED_PEAKS.json), algorithmic bytes per SURVEY.md §8(d).
  cpu_baseline : the fp64 oracle restatement of the reference's CPU path ("port"; the reference itself
          cannot be built here) on a bounded sample of the same scene, on this host's cores.

N > 1 (torchrun): weak scaling -- every rank owns 100 source images of a scene with 100*N views (scene
replicated, matches sharded by source image) and the per-node results are exchanged with
one NCCL all-gather inside the timed region.

--impl reference times the reference-faithful CPU restatement (oracle/) on bounded samples instead.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "3D line candidates triangulated+scored/sec"
UNIT = "match rows/s"
WORKLOAD = "hypersim100"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for k, n in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def get_scene(n_gpus, rank):
    from limap_b200.synth import CONFIGS, make_scene
    cfg = dict(CONFIGS[WORKLOAD])
    per = cfg["V"]
    cfg["V"] = per * n_gpus
    mv = range(per * rank, per * (rank + 1))
    return make_scene(match_views=mv, **cfg), per


def algorithmic_bytes(n_rows, n_nodes, n_views, n_cand, n_valid):
    """B_gen + B_score for the fused kernel (fp32-storage convention of the survey)."""
    b_gen = 24 * n_rows + 16 * n_nodes + 44 * n_views + 48 * n_cand
    b_score = n_cand * (48 + 16) + 4 * n_cand + n_nodes * 48 + 4 * n_valid
    return b_gen + b_score


def pick_cpu_threads(scene):
    """The reference parallelises with OpenMP inside one node (n ~ 10..200 iterations per region), which does
    not scale to every core of a large host; calibrate on one source image and keep the fastest count."""
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from oracle import oracle as orc
    n = orc.usable_cpus()
    cands = sorted({1, min(8, n), min(32, n), n})
    i0 = int(scene.img_ids[0]) if int(scene.img_ids[0]) in scene.matches else sorted(scene.matches)[0]
    f = scene.flat_matches(i0)
    best, best_t = 1, None
    for th in cands:
        o = orc.OracleTri(dict(DEFAULT_YAML_TRIANGULATION), threads=th)
        o.upload(scene)
        o.set_ranges(*scene.ranges)
        t0 = time.perf_counter()
        o.add_image_matches(i0, *f)
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = th, dt
        del o
    return best


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of the reference's own code path (oracle/, 'port': the reference
    needs Eigen/Ceres/COLMAP and cannot be built in this image), all host threads, bounded samples."""
    if rank != 0:
        return
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from oracle import oracle as orc
    orc.build()
    scene, per = get_scene(1, 0)
    cores = pick_cpu_threads(scene)
    sample_imgs = max(1, args.ref_images)
    ids = [int(i) for i in scene.img_ids]
    times, rows = [], []
    nsteps = args.warmup + args.steps
    for s in range(nsteps):
        o = orc.OracleTri(dict(DEFAULT_YAML_TRIANGULATION), threads=cores)
        o.upload(scene)
        o.set_ranges(*scene.ranges)
        pick = [ids[(s * sample_imgs + k) % len(ids)] for k in range(sample_imgs)]
        flat = [scene.flat_matches(i) for i in pick]
        t0 = time.perf_counter()
        for i, f in zip(pick, flat):
            o.add_image_matches(i, *f)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
            rows.append(o.rows_tested())
        del o
    value = float(sum(rows) / sum(times))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "V": 100, "L": 1000, "N": 20, "K": 10,
                       "sample": f"{sample_imgs} source image(s) per step of the {len(ids)}-image scene"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{sample_imgs} source images x {int(np.mean(rows))} rows per step, "
                                       f"{args.steps} steps"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--ref-images", type=int, default=10, help="source images per reference step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-lm", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: limap_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL's own log lines (version banner, NCCL_DEBUG output) go to stderr: stdout carries the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from limap_b200._cabi import NODE_RECORD_DTYPE
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from limap_b200.engine import TriEngine
    from limap_b200 import dist as lmdist

    scene, per = get_scene(world, rank)
    my_ids = [int(scene.img_ids[v]) for v in range(per * rank, per * (rank + 1))]
    n_rows_rank = scene.n_rows(my_ids)
    cfg = dict(DEFAULT_YAML_TRIANGULATION)
    eng = TriEngine(cfg, device=local_rank)
    eng.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.upload(scene)
    eng.set_ranges(*scene.ranges)
    flat = {i: scene.flat_matches(i) for i in my_ids}
    for i in my_ids:
        eng.add_image_matches(i, *flat[i])
    eng.set_shard(per * rank, per * (rank + 1))
    gather = lmdist.NodeGather(eng, world, rank) if world > 1 else None

    def step():
        st = eng.run()
        if gather is not None:
            gather.all_gather()
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        st = step()
    launches0 = eng.stats()["n_kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record()
    for _ in range(args.steps):
        st = step()
        kernel_ms.append(st["last_node_kernel_ms"])
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    launches = eng.stats()["n_kernel_launches"] - launches0
    t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
    r = torch.tensor([float(n_rows_rank)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.SUM)
    ms_total = float(t.item())
    rows_all = float(r.item())
    value = rows_all * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (fused generate+score), rank 0's launch ----------------
    peak, peak_src = load_peaks()
    n_nodes_shard = int(scene.line_off[per * (rank + 1)] - scene.line_off[per * rank])
    alg = algorithmic_bytes(st["n_rows"], n_nodes_shard, scene.n_views, st["n_candidates"], st["n_valid_edges"])
    k_ms = float(np.mean(kernel_ms))
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM bytes of one launch from the committed `ncu --set full` capture of this workload (world == 1)
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "tri_node_kernel_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("workload") == WORKLOAD and world == 1:
            traffic, traffic_src = float(tj["dram_bytes_per_launch"]), tj.get("source")
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "hbm", "kernel": "tri_node_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                "kernel_share_of_step": k_ms * args.steps / ms_total,
                "note": "ALU/SFU-bound fp64 geometry (C^2 pair scores per node); HBM fraction is low by "
                        "construction, see DESIGN.md"}

    # ---- e2e through the public API with host buffers (rank-local; N=1 headline) -------------------
    e2e = None
    if not args.no_e2e:
        bsrc, bng, boff, bpairs = scene.bulk_matches(my_ids)
        tp = torch.empty(bpairs.shape, dtype=torch.int32, pin_memory=True)
        tp.numpy()[...] = bpairs
        pinned_pairs = tp.numpy()
        eng2 = TriEngine(cfg, device=local_rank)
        eng2.ctx.set_stream(torch.cuda.current_stream().cuda_stream)
        h2d = (scene.segs.nbytes + scene.kvec.nbytes + scene.qvec.nbytes + scene.tvec.nbytes +
               scene.line_off.nbytes + pinned_pairs.nbytes)
        nodes_out = torch.empty(int(scene.line_off[-1]) * NODE_RECORD_DTYPE.itemsize, dtype=torch.uint8,
                                pin_memory=True).numpy().view(NODE_RECORD_DTYPE)
        off_out = torch.empty(int(scene.line_off[-1]) + 1, dtype=torch.int64, pin_memory=True).numpy()
        edges_out = torch.empty((max(int(st["n_valid_edges"]) * 2, 1), 2), dtype=torch.int32, pin_memory=True).numpy()

        def e2e_step():
            # public API, bulk form: Init + SetRanges + TriangulateImage(all images) + run + results to host
            eng2.upload(scene)
            eng2.set_ranges(*scene.ranges)
            eng2.add_matches_bulk(bsrc, bng, boff, pinned_pairs)
            eng2.set_shard(per * rank, per * (rank + 1))
            s2 = eng2.run()
            nodes = eng2.get_nodes(nodes_out)
            off, edges = eng2.get_all_valid_edges(off_out, edges_out)
            return s2, nodes.nbytes + off.nbytes + edges.nbytes

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            s2, _ = e2e_step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h = n_nodes_shard * NODE_RECORD_DTYPE.itemsize + 4 * (n_nodes_shard + 1) + 4 * s2["n_valid_edges"]
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": rows_all * args.steps / float(tt.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h)}
        eng2.close()

    # ---- CPU baseline (rank 0, N=1 only): oracle restatement on a bounded sample -------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        orc.build()
        cores = pick_cpu_threads(scene)
        o = orc.OracleTri(cfg, threads=cores)
        o.upload(scene)
        o.set_ranges(*scene.ranges)
        t0 = time.perf_counter()
        n_img = 0
        for i in my_ids:
            o.add_image_matches(i, *flat[i])
            n_img += 1
            if time.perf_counter() - t0 > args.cpu_seconds:
                break
        dt = time.perf_counter() - t0
        cpu = {"value": o.rows_tested() / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "host_cpus": orc.usable_cpus(),
               "sample": f"first {n_img} source images of the scene ({o.rows_tested()} rows, {dt:.1f} s), "
                         "fp64 restatement with the reference's loop structure (OpenMP over connections / "
                         "candidates of one node)"}
        # the same arithmetic with the OpenMP loop moved out to the 2D lines of an image (identical results): what a
        # throughput-tuned CPU implementation would do; reported beside the reference's own schedule
        try:
            o2 = orc.OracleTri(cfg, threads=orc.usable_cpus(), node_parallel=True)
            o2.upload(scene)
            o2.set_ranges(*scene.ranges)
            t0 = time.perf_counter()
            n2 = 0
            for i in my_ids:
                o2.add_image_matches(i, *flat[i])
                n2 += 1
                if time.perf_counter() - t0 > 0.5 * args.cpu_seconds:
                    break
            dt2 = time.perf_counter() - t0
            cpu["node_parallel"] = {"value": o2.rows_tested() / dt2, "cores": orc.usable_cpus(),
                                    "sample": f"first {n2} source images ({o2.rows_tested()} rows, {dt2:.1f} s)"}
        except Exception as e:  # the extra figure must never cost the bench line
            cpu["node_parallel"] = {"error": str(e)}

    # ---- M2: line-BA LM iterations/s (BASELINE.json configs[3]: 10k tracks x 30 supporting views per rank) ----
    lm_ba = None
    if not args.no_lm:
        from limap_b200.engine import BAEngine
        from limap_b200.synth import make_tracks
        ts = make_tracks(T=10000, S=30, V=300, seed=1237 + rank)
        ba = BAEngine(ctx=eng.ctx)
        views, first_idx = np.unique(ts.img_ids, return_index=True)
        remap = np.zeros(int(views.max()) + 1, np.int32)
        remap[views] = np.arange(len(views), dtype=np.int32)
        a = (ts.kvec[first_idx], ts.qvec[first_idx], ts.tvec[first_idx], ts.sup_off, remap[ts.img_ids], ts.segs,
             ts.line3d, ts.line_init)
        for _ in range(2):
            out = ba.solve(*a, max_num_iterations=100)
        barrier()
        t0 = time.perf_counter()
        k_ms, iters = [], 0
        for _ in range(args.steps):
            out = ba.solve(*a, max_num_iterations=100)
            k_ms.append(out["stats"]["solve_ms"])
            iters += out["stats"]["total_iterations"]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, float(np.sum(k_ms)) * 1e-3], device="cuda", dtype=torch.float64)
        ii = torch.tensor([float(iters)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(ii, op=dist.ReduceOp.SUM)
        lm_ba = {"metric": "line-BA LM iters/sec", "unit": "track LM iterations/s",
                 "config": {"workload": "ba10k", "tracks_per_gpu": 10000, "supports": 30, "max_num_iterations": 100},
                 "value_kernel": float(ii.item() / tt[1].item()), "e2e": {"value": float(ii.item() / tt[0].item()),
                                                                         "note": "lm_ba_solve from host arrays: H2D, "
                                                                                 "solve, segment cut, D2H"},
                 "kernel_ms": float(np.mean(k_ms)), "iterations_per_solve": int(iters // args.steps),
                 "dtype": "f64"}
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            sub = make_tracks(T=1500, S=30, V=300, seed=1237)
            best = None
            for th in sorted({1, min(8, orc.usable_cpus()), orc.usable_cpus()}):
                t0 = time.perf_counter()
                o = orc.refine_tracks(sub, max_num_iterations=100, threads=th)
                v = float(o["iters"][:, 0].sum() / (time.perf_counter() - t0))
                if best is None or v > best[0]:
                    best = (v, th)
            lm_ba["cpu_baseline"] = {"value": best[0], "unit": "track LM iterations/s", "cores": best[1], "kind": "port",
                                     "sample": "1500 tracks x 30 supports, Ceres-style LM restatement, OpenMP over "
                                               "tracks"}

    # ---- M3: remerge pair test: all-pairs check_connection over 1e5 track lines, rank 0 ----
    remerge = None
    if not args.no_lm and rank == 0:
        from limap_b200.config import LINKER3D_DEFAULTS, make_linker
        from limap_b200.engine import MergeEngine
        from limap_b200.synth import make_track_lines
        lk = dict(score_th=0.5, th_angle=5.0, th_overlap=0.001, th_smartoverlap=0.1, th_smartangle=1.0, th_perp=1.0,
                  th_innerseg=1.0)  # cfgs/triangulation/default.yaml:99-108
        Tm = 100000
        TL = make_track_lines(Tm, dup_frac=0.3, seed=1, extent=60.0)
        me = MergeEngine(ctx=eng.ctx)
        act = np.ones(Tm, np.uint8)
        k_ms, w_ms = [], []
        for it in range(2 + args.steps):
            t0 = time.perf_counter()
            _, ng, ne = me.remerge_labels(TL, act, make_linker(LINKER3D_DEFAULTS, lk))
            if it >= 2:
                w_ms.append((time.perf_counter() - t0) * 1e3)
                k_ms.append(me.stats()["last_remerge_kernel_ms"])
        pairs = Tm * (Tm - 1) / 2
        remerge = {"metric": "remerge pair tests/sec", "unit": "track pairs/s",
                   "config": {"workload": "remerge100k", "tracks": Tm, "groups": ng, "edges": ne,
                              "pairs_past_fp32_gate": int(me.stats()["n_pairs_gated"])},
                   "value_kernel": pairs / (float(np.mean(k_ms)) * 1e-3), "kernel_ms": float(np.mean(k_ms)),
                   "e2e": {"value": pairs / (float(np.mean(w_ms)) * 1e-3),
                           "note": "lm_remerge_labels from host arrays: H2D, pair kernel, edge list D2H, host union-find"},
                   "dtype": "f32 gate + f64 check"}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as orc
            Ts = 20000
            sub = make_track_lines(Ts, dup_frac=0.3, seed=1, extent=60.0)
            t0 = time.perf_counter()
            orc.remerge_labels(sub, np.ones(Ts, np.uint8), lk, threads=orc.usable_cpus())
            dtc = time.perf_counter() - t0
            remerge["cpu_baseline"] = {"value": Ts * (Ts - 1) / 2 / dtc, "unit": "track pairs/s",
                                       "cores": orc.usable_cpus(), "kind": "port",
                                       "sample": f"{Ts} tracks, all pairs, OpenMP over tracks ({dtc:.2f} s)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "V": 100 * world, "L": 1000, "N": 20, "K": 10,
                           "rows_per_step": int(rows_all), "candidates_per_step_rank0": int(st["n_candidates"]),
                           "valid_connections_rank0": int(st["n_valid_edges"]),
                           "pairs": {"past_3d_gates": int(st["n_pairs_gated"]), "scored_exact_fp64": int(st["n_pairs_exact"])},
                           "parallelism": f"source-image shards x{world}",
                           "l2": "inputs larger than L2 (match rows + sort buffers > 126 MB per step)"},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
                "cpu_baseline": cpu, "lm_ba": lm_ba, "remerge": remerge}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
