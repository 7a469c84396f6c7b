"""Frozen hot-path fixtures (tests/golden/hotpath/*.npz, made by tests/golden/make_hotpath_golden.py).
CPU (`-m "not gpu"`): the oracle reproduces them -- a drift of the restatement shows up here, not as a silent change
of what the CUDA path is compared with. GPU (`-m gpu`): the CUDA path against the frozen outputs, oracle not in the
loop. Bars: ids / counts / valid connections / track membership bit-exact, endpoints 1e-4, scores 1e-6."""
import os

import numpy as np
import pytest

from limap_b200.config import DEFAULT_YAML_TRIANGULATION

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath")
TRI = ["tri_default_yaml", "tri_mixed_cameras_halfpix", "tri_vp_proposals"]


class _R:
    def __init__(self, labels, vps):
        self.labels, self.vps = labels, vps


def _load(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = dict(DEFAULT_YAML_TRIANGULATION)
    for k, v in zip(z["cfg_keys"], z["cfg_vals"]):
        cfg[str(k)] = int(v) if float(v).is_integer() else float(v)
    return z, cfg


def _feed(t, z):
    t.upload_scene(z["img_ids"], z["model_ids"], z["kvec"], z["qvec"], z["tvec"], z["line_off"], z["segs"])
    t.set_ranges(z["range_lo"], z["range_hi"])
    if len(z["vp_labels"]):
        vp, lo = {}, z["line_off"]
        for v, i in enumerate(z["img_ids"]):
            vp[int(i)] = _R(z["vp_labels"][lo[v]:lo[v + 1]], z["vp_vps"][v])
        t.set_vps(vp, z["img_ids"], z["line_off"])
    src, ng, off, pairs = z["blk_src"], z["blk_ng"], z["blk_off"], z["pairs"]
    for i in z["img_ids"]:
        sel = np.flatnonzero(src == i)
        row_off = np.concatenate([[0], np.cumsum(off[sel + 1] - off[sel])]).astype(np.int64)
        pr = np.concatenate([pairs[off[b]:off[b + 1]] for b in sel]) if len(sel) else np.zeros((0, 2), np.int32)
        t.add_image_matches(int(i), ng[sel], row_off, np.ascontiguousarray(pr, np.int32))


def _check_tri(t, z):
    best, ng, nc, eoff, edges = [], [], [], [0], []
    for i in z["img_ids"]:
        l, g, c = t.get_best(int(i))
        best.append(l); ng.append(g); nc.append(c)
        off, e = t.get_valid_edges(int(i))
        for k in range(len(off) - 1):
            ee = sorted(map(tuple, e[off[k]:off[k + 1]]))
            edges.extend(ee)
            eoff.append(eoff[-1] + len(ee))
    best, ng, nc = np.concatenate(best), np.concatenate(ng), np.concatenate(nc)
    assert np.array_equal(nc, z["n_cand"])
    has = nc > 0
    assert np.array_equal(ng[has], z["best_ng"][has])
    assert np.abs(best[has, :9] - z["best_line"][has, :9]).max() <= 1e-4
    assert np.abs(best[has, 9] - z["best_line"][has, 9]).max() <= 1e-6
    assert np.array_equal(np.asarray(eoff), z["edge_off"])
    assert np.array_equal(np.asarray(edges, np.int32).reshape(-1, 2), z["edges"])
    tr = t.build_tracks()
    members = []
    for k in range(len(tr["track_off"]) - 1):
        a, b = tr["track_off"][k], tr["track_off"][k + 1]
        members.append(sorted(zip(tr["img_ids"][a:b].tolist(), tr["line_ids"][a:b].tolist())))
    order = sorted(range(len(members)), key=lambda k: members[k])
    flat = np.asarray([m for k in order for m in members[k]], np.int32).reshape(-1, 2)
    assert np.array_equal(flat, z["track_members"])
    assert np.array_equal(np.concatenate([[0], np.cumsum([len(members[k]) for k in order])]), z["track_off"])
    tl, gl = tr["track_line"][order], z["track_line"]
    d = np.minimum(np.abs(tl[:, :6] - gl[:, :6]).max(1), np.abs(tl[:, :6] - gl[:, [3, 4, 5, 0, 1, 2]]).max(1))
    big = [k for k in range(len(order)) if len(members[order[k]]) >= 4]  # < 4 members: take-best ties
    assert d[big].max(initial=0) <= 1e-4


def _trackset(z):
    from limap_b200.synth import TrackSet
    return TrackSet(sup_off=z["sup_off"], segs=z["segs"], kvec=z["kvec"], qvec=z["qvec"], tvec=z["tvec"],
                    img_ids=z["img_ids"], line3d=z["line3d"], line_init=z["line_init"], gt=z["line_init"])


def _check_lm(out, z):
    assert np.abs(out["cost"][:, 0] - z["cost"][:, 0]).max() <= 1e-9 * (1 + z["cost"][:, 0].max())
    rel = np.abs(out["cost"][:, 1] - z["cost"][:, 1]) / (1e-12 + z["cost"][:, 1])
    assert rel.max() < 1e-6
    d = np.minimum(np.abs(out["line"] - z["line"]).max(1), np.abs(out["line"] - z["line"][:, [3, 4, 5, 0, 1, 2]]).max(1))
    assert d.max() <= 1e-4


@pytest.mark.parametrize("name", TRI)
def test_oracle_reproduces_golden_triangulation(name):
    from oracle.oracle import OracleTri
    z, cfg = _load(name)
    o = OracleTri(cfg)
    _feed(o, z)
    _check_tri(o, z)


def test_oracle_reproduces_golden_refinement():
    from oracle import oracle as orc
    z = np.load(os.path.join(GOLD, "lm_refine.npz"))
    o = orc.refine_tracks(_trackset(z), max_num_iterations=int(z["max_iter"]), threads=1)
    _check_lm(o, z)
    assert np.array_equal(o["iters"], z["iters"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", TRI)
def test_cuda_matches_golden_triangulation(name):
    from limap_b200.engine import TriEngine
    z, cfg = _load(name)
    e = TriEngine(cfg)
    _feed(e, z)
    _check_tri(e, z)


@pytest.mark.gpu
def test_cuda_matches_golden_refinement():
    from limap_b200.engine import BAEngine
    z = np.load(os.path.join(GOLD, "lm_refine.npz"))
    g = BAEngine().solve_trackset(_trackset(z), max_num_iterations=int(z["max_iter"]))
    _check_lm(g, z)
