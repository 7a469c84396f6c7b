"""Helpers of the runner-level drop-in tests: artefacts of a synthetic scene in the reference's on-disk layout
(segments_<id>.txt, matches_<id>.npy) and oracle-backed stand-ins for the three engine classes, so that the Python
operator surface can be driven on a machine without a GPU (test infrastructure: the product never uses them)."""
import os

import numpy as np


def write_artifacts(sc, cfg, root):
    """Detections and matches of scene `sc` where runners.compute_2d_segs / compute_matches (load branches) look."""
    import limap.runners as runners
    import limap.util.io as limapio
    segdir = runners.segments_folder(cfg, root)
    mdir = runners.matches_folder(cfg, root)
    os.makedirs(segdir, exist_ok=True)
    os.makedirs(mdir, exist_ok=True)
    for v, i in enumerate(sc.img_ids):
        limapio.save_txt_segments(segdir, int(i), sc.lines_of(v))
        limapio.save_npy(os.path.join(mdir, f"matches_{int(i)}.npy"), {int(g): m for g, m in sc.matches[int(i)].items()})


def imagecols_of(sc):
    import limap.base as base
    cams, imgs = {}, {}
    for v, i in enumerate(sc.img_ids):
        k = sc.kvec[v]
        if int(sc.model_ids[v]) == 0:
            cams[v] = base.Camera("SIMPLE_PINHOLE", [k[0], k[2], k[3]], v, (600, 800))
        else:
            cams[v] = base.Camera("PINHOLE", [k[0], k[1], k[2], k[3]], v, (600, 800))
        imgs[int(i)] = base.CameraImage(v, base.CameraPose(sc.qvec[v], sc.tvec[v]), f"img_{int(i)}.png")
    return base.ImageCollection(cams, imgs)


class OracleTriEngine:
    """TriEngine's array interface on top of oracle.OracleTri (CPU)."""

    def __init__(self, cfg=None, device=0):
        from oracle.oracle import OracleTri
        self._o = OracleTri(cfg, threads=1)
        self.cfg = self._o.cfg

    def __getattr__(self, name):
        return getattr(self._o, name)

    def unset_ranges(self):
        self._o._c("tri_unset_ranges")(self._o._h)

    def add_image_matches_dict(self, img_id, matches):
        ngs = sorted(matches.keys())
        row_off = np.zeros(len(ngs) + 1, np.int64)
        parts = []
        for k, g in enumerate(ngs):
            m = np.asarray(matches[g]).reshape(-1, 2).astype(np.int32)
            parts.append(m)
            row_off[k + 1] = row_off[k] + len(m)
        pairs = np.concatenate(parts, 0) if parts else np.zeros((0, 2), np.int32)
        self._o.add_image_matches(img_id, np.asarray(ngs, np.int32), row_off, pairs)

    def stats(self):
        return {"n_candidates": 0, "n_valid_edges": 0, "n_rows": self._o.rows_tested()}


class OracleBAEngine:
    def __init__(self, device=0, ctx=None):
        pass

    def solve(self, kvec, qvec, tvec, sup_off, sup_view, segs, line3d, line_init, max_num_iterations=100, min_num_images=4,
              num_outliers=2, geometric_alpha=10.0, cauchy_scale=0.25, max_num_consecutive_invalid_steps=10, sup_vp=None,
              vp_multiplier=1.0):
        from limap_b200.synth import TrackSet
        from oracle import oracle as orc
        sv = np.asarray(sup_view, np.int64)
        ts = TrackSet(sup_off=np.asarray(sup_off, np.int64), segs=np.asarray(segs, float), kvec=np.asarray(kvec, float)[sv],
                      qvec=np.asarray(qvec, float)[sv], tvec=np.asarray(tvec, float)[sv], img_ids=sv.astype(np.int32),
                      line3d=np.asarray(line3d, float), line_init=np.asarray(line_init, float), gt=np.asarray(line_init, float))
        o = orc.refine_tracks(ts, max_num_iterations=max_num_iterations, min_num_images=min_num_images,
                              num_outliers=num_outliers, geometric_alpha=geometric_alpha, cauchy_scale=cauchy_scale,
                              sup_vp=sup_vp, vp_multiplier=vp_multiplier, threads=1,
                              max_num_consecutive_invalid_steps=max_num_consecutive_invalid_steps)
        o["stats"] = {"total_iterations": int(o["iters"][:, 0].sum())}
        return o


class OracleMergeEngine:
    def __init__(self, device=0, ctx=None):
        pass

    def support_flags(self, model_ids, kvec, qvec, tvec, sup_off, sup_view, segs, track_line, **th):
        from oracle import oracle as orc
        return orc.track_support_flags(model_ids, kvec, qvec, tvec, sup_off, sup_view, segs, track_line, threads=1, **th)

    def remerge_labels(self, track_line, active, linker3d):
        from oracle import oracle as orc
        lk = {n: getattr(linker3d, n) for n, _ in linker3d._fields_}
        return orc.remerge_labels(track_line, active, lk, threads=1)

    @staticmethod
    def aggregate(off, lines, scores, num_outliers):
        from oracle import oracle as orc
        return orc.aggregate_lines(off, lines, scores, num_outliers)


def install_oracle_backend(monkeypatch):
    """Point the operator surface at the CPU oracle (tests only)."""
    import limap_b200.merging as merging
    import limap_b200.optimize as optimize
    import limap_b200.triangulation as triangulation
    monkeypatch.setattr(triangulation, "TriEngine", OracleTriEngine)
    monkeypatch.setattr(optimize, "BAEngine", OracleBAEngine)
    monkeypatch.setattr(merging, "MergeEngine", OracleMergeEngine)
    monkeypatch.setattr(merging, "_engine", None, raising=False)


def summarize(tracks):
    """Comparable digest of a runner result: per-track membership, supports, line."""
    mem = [tuple(sorted(zip(t.image_id_list, t.line_id_list))) for t in tracks]
    order = sorted(range(len(tracks)), key=lambda k: mem[k])
    lines = np.array([np.concatenate([tracks[k].line.start, tracks[k].line.end]) for k in order]) if tracks else np.zeros((0, 6))
    return [mem[k] for k in order], lines
