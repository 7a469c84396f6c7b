"""Freezes seeded hot-path fixtures: inputs AND the oracle's outputs for a small triangulation
scene, a VP-proposal scene and a line-refinement problem, as tests/golden/hotpath/*.npz.

    python tests/golden/make_hotpath_golden.py

The reference has no golden vectors for this path and cannot be built here, so these are produced by
the restatement (oracle/) -- which oracle/_ref pins to the reference's own compiled functions where they compile --
and frozen: `-m "not gpu"` tests check that the oracle still reproduces them, `-m gpu` tests check the CUDA path
against them without the oracle in the loop."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
OUT = os.path.join(HERE, "hotpath")


def tri_fixture(name, cfg_over, scene_kw, vp_seed=None):
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from limap_b200.synth import make_scene
    from oracle.oracle import OracleTri
    sc = make_scene(**scene_kw)
    cfg = dict(DEFAULT_YAML_TRIANGULATION)
    cfg.update(cfg_over)
    o = OracleTri(cfg)
    o.upload(sc)
    o.set_ranges(*sc.ranges)
    vp = {}
    if vp_seed is not None:
        rng = np.random.default_rng(vp_seed)
        labs, vps = [], []
        for v, i in enumerate(sc.img_ids):
            L = int(sc.line_off[v + 1] - sc.line_off[v])
            q = rng.normal(size=(3, 3))
            q[:, :2] *= 1000.0
            q /= np.linalg.norm(q, axis=1, keepdims=True)
            lab = rng.integers(0, 3, L)
            lab[rng.random(L) < 0.4] = -1
            labs.append(lab.astype(np.int32))
            vps.append(q)

            class R:
                pass
            r = R()
            r.labels, r.vps = labs[-1], q
            vp[int(i)] = r
        o.set_vps(vp, sc.img_ids, sc.line_off)
    for i in sc.img_ids:
        o.add_image_matches(int(i), *sc.flat_matches(int(i)))
    best, ng, ncand, eoff, edges = [], [], [], [0], []
    for i in sc.img_ids:
        l, g, c = o.get_best(int(i))
        best.append(l); ng.append(g); ncand.append(c)
        off, e = o.get_valid_edges(int(i))
        for k in range(len(off) - 1):
            ee = sorted(map(tuple, e[off[k]:off[k + 1]]))
            edges.extend(ee)
            eoff.append(eoff[-1] + len(ee))
    tr = o.build_tracks()
    members = []
    for t in range(len(tr["track_off"]) - 1):
        a, b = tr["track_off"][t], tr["track_off"][t + 1]
        members.append(sorted(zip(tr["img_ids"][a:b].tolist(), tr["line_ids"][a:b].tolist())))
    order = sorted(range(len(members)), key=lambda k: members[k])
    src, ngb, boff, pairs = sc.bulk_matches()
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        cfg_keys=np.array(sorted(cfg_over.keys())), cfg_vals=np.array([float(cfg_over[k]) for k in sorted(cfg_over.keys())]),
        img_ids=sc.img_ids, model_ids=sc.model_ids, kvec=sc.kvec, qvec=sc.qvec, tvec=sc.tvec, line_off=sc.line_off,
        segs=sc.segs, range_lo=sc.ranges[0], range_hi=sc.ranges[1], blk_src=src, blk_ng=ngb, blk_off=boff, pairs=pairs,
        vp_labels=np.concatenate(labs) if vp else np.zeros(0, np.int32), vp_vps=np.stack(vps) if vp else np.zeros((0, 3, 3)),
        best_line=np.concatenate(best), best_ng=np.concatenate(ng), n_cand=np.concatenate(ncand),
        edge_off=np.asarray(eoff, np.int64), edges=np.asarray(edges, np.int32).reshape(-1, 2),
        track_off=np.concatenate([[0], np.cumsum([len(members[k]) for k in order])]).astype(np.int64),
        track_members=np.asarray([m for k in order for m in members[k]], np.int32).reshape(-1, 2),
        track_line=tr["track_line"][order])
    print(name, "nodes", len(np.concatenate(ncand)), "candidates", int(np.concatenate(ncand).sum()), "edges", eoff[-1],
          "tracks", len(members))


def lm_fixture(name, T, S, V, seed, max_iter):
    from limap_b200.synth import make_tracks
    from oracle import oracle as orc
    ts = make_tracks(T=T, S=S, V=V, seed=seed)
    o = orc.refine_tracks(ts, max_num_iterations=max_iter, threads=1)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), sup_off=ts.sup_off, segs=ts.segs, kvec=ts.kvec, qvec=ts.qvec,
                        tvec=ts.tvec, img_ids=ts.img_ids, line3d=ts.line3d, line_init=ts.line_init,
                        max_iter=np.int64(max_iter), line=o["line"], cost=o["cost"], iters=o["iters"])
    print(name, "tracks", T, "iterations", int(o["iters"][:, 0].sum()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    from oracle import oracle as orc
    orc.build()
    tri_fixture("tri_default_yaml", {}, dict(V=6, L=50, N=4, K=4, seed=101))
    tri_fixture("tri_mixed_cameras_halfpix", {"add_halfpix": 1}, dict(V=6, L=50, N=4, K=4, seed=102, camera_mix=True, scale=100.0))
    tri_fixture("tri_vp_proposals", {"use_vp": 1}, dict(V=5, L=40, N=3, K=3, seed=103), vp_seed=9)
    lm_fixture("lm_refine", T=60, S=10, V=40, seed=104, max_iter=100)
