"""One workload per hot kernel, for `ncu -k regex:<kernel> ...` captures (never a bench value):
  python scripts/ncu_target.py tri      # tri_node_kernel on hypersim100 (the bench workload), 3 runs
  python scripts/ncu_target.py lm       # lm_refine_kernel on 10k tracks x 30 supports (configs[3])
  python scripts/ncu_target.py remerge  # remerge_pairs_kernel on 1e5 track lines
  python scripts/ncu_target.py vp       # jlinkage_kernel on 296 images x 300 segments x 5000 hypotheses"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
what = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if what == "tri":
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from limap_b200.engine import TriEngine
    from limap_b200.synth import CONFIGS, make_scene
    cfg_scene = dict(CONFIGS[os.environ.get("LM_WORKLOAD", "hypersim100")])
    if os.environ.get("LM_K"):  # matches per (line, neighbour): rows per node = N * K (occupancy experiments)
        cfg_scene["K"] = int(os.environ["LM_K"])
    sc = make_scene(**cfg_scene)
    eng = TriEngine(dict(DEFAULT_YAML_TRIANGULATION))
    eng.upload(sc)
    eng.set_ranges(*sc.ranges)
    eng.add_matches_bulk(*sc.bulk_matches())
    for _ in range(n):
        st = eng.run()
    print("tri", st["last_node_kernel_ms"], st["n_candidates"], "max rows/node", st["max_rows_per_node"])
elif what == "lm":
    from limap_b200.engine import BAEngine
    from limap_b200.synth import make_tracks
    ts = make_tracks(T=10000, S=30, V=300, seed=1237)
    ba = BAEngine()
    for _ in range(n):
        o = ba.solve_trackset(ts, max_num_iterations=100)
    print("lm", o["stats"]["solve_ms"], o["stats"]["total_iterations"])
elif what == "remerge":
    from limap_b200.config import LINKER3D_DEFAULTS, make_linker
    from limap_b200.engine import MergeEngine
    from limap_b200.synth import make_track_lines
    lk = dict(score_th=0.5, th_angle=5.0, th_overlap=0.001, th_smartoverlap=0.1, th_smartangle=1.0, th_perp=1.0, th_innerseg=1.0)
    TL = make_track_lines(100000, dup_frac=0.3, seed=1, extent=60.0)
    me = MergeEngine()
    for _ in range(n):
        me.remerge_labels(TL, np.ones(len(TL), np.uint8), make_linker(LINKER3D_DEFAULTS, lk))
    print("remerge", me.stats()["last_remerge_kernel_ms"])
elif what == "vp":
    from limap_b200.synth import make_vp_images
    from limap_b200.vplib import JLinkageDetector
    imgs = make_vp_images(296, 300, seed=77)
    det = JLinkageDetector(dict(min_num_supports=10, min_length=40, inlier_threshold=1.0), seed=7)
    for _ in range(n):
        det.detect_batch(imgs)
    print("vp", det.stats()["kernel_ms"])
