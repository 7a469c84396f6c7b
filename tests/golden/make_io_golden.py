"""Generates tests/golden/io/* by running the REFERENCE's own writers (src/limap/util/io.py, pure Python) in this
container, and cross-checks both directions (reference readers on files written by limap_b200.util.io).

The reference package itself cannot be imported here (its __init__ pulls the compiled _limap module and pycolmap), so
io.py is loaded from its path with `pycolmap` stubbed (only logging.info is used) and `limap.base` resolved to this
repository's value types (the writers only call count_lines / count_images / .line / id lists on the objects they get).
Run from the repository root:  python tests/golden/make_io_golden.py   (needs /root/reference; the GPU box only
reads the committed fixtures)."""
import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "io")
REF_IO = "/root/reference/src/limap/util/io.py"


def load_reference_io():
    pyc = types.ModuleType("pycolmap")
    pyc.logging = types.SimpleNamespace(info=lambda *a, **k: None, warning=lambda *a, **k: None)
    sys.modules["pycolmap"] = pyc
    import limap  # noqa: F401  (alias package of this repository: limap.base -> limap_b200.base)
    spec = importlib.util.spec_from_file_location("reference_limap_util_io", REF_IO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def fixtures():
    import limap.base as base
    rng = np.random.default_rng(2024)
    segs = {3: np.round(rng.uniform(0, 640, (5, 4)), 3), 11: rng.uniform(0, 640, (2, 4)), 12: np.zeros((0, 4))}
    neighbors = {3: [11, 12], 11: [3], 12: []}
    ranges = (np.array([-1.5, -2.25, 0.0]), np.array([3.0, 4.125, 10.0]))
    tracks = []
    for t in range(3):
        tr = base.LineTrack()
        tr.line = base.Line3d(rng.normal(size=3), rng.normal(size=3))
        n = 3 + 2 * t
        tr.image_id_list = [int(x) for x in rng.integers(0, 4 + t, n)]
        tr.line_id_list = [int(x) for x in rng.integers(0, 100, n)]
        tr.line2d_list = [base.Line2d(rng.uniform(0, 100, 2), rng.uniform(0, 100, 2)) for _ in range(n)]
        tracks.append(tr)
    matches = {11: rng.integers(0, 50, (7, 2)).astype(np.int64), 12: np.zeros((0, 2), np.int64)}
    return segs, neighbors, ranges, tracks, matches


def main():
    ref = load_reference_io()
    import limap.util.io as mine
    os.makedirs(OUT, exist_ok=True)
    segs, neighbors, ranges, tracks, matches = fixtures()
    for i, s in segs.items():
        ref.save_txt_segments(OUT, i, s)
    ref.save_txt_metainfos(os.path.join(OUT, "metainfos.txt"), neighbors, ranges)
    ref.save_txt_linetracks(os.path.join(OUT, "alltracks_nv1.txt"), tracks, n_visible_views=1)
    ref.save_txt_linetracks(os.path.join(OUT, "alltracks_nv3.txt"), tracks, n_visible_views=3)
    ref.save_npy(os.path.join(OUT, "matches_3.npy"), matches)
    np.savez(os.path.join(OUT, "inputs.npz"), seg_ids=np.array(sorted(segs)), **{f"segs_{i}": s for i, s in segs.items()},
             range_lo=ranges[0], range_hi=ranges[1],
             track_lines=np.array([np.concatenate([t.line.start, t.line.end]) for t in tracks]),
             track_img=np.array([t.image_id_list for t in tracks], dtype=object),
             track_lid=np.array([t.line_id_list for t in tracks], dtype=object), allow_pickle=True)
    # the other direction: the reference's readers on files written by this repository
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        for i, s in segs.items():
            mine.save_txt_segments(d, i, s)
            if len(s):  # the reference's reader returns a (0,) array for an empty file; nothing to compare
                assert np.array_equal(ref.read_txt_segments(d, i), s)
        mine.save_txt_metainfos(os.path.join(d, "m.txt"), neighbors, ranges)
        nb, rg = ref.read_txt_metainfos(os.path.join(d, "m.txt"))
        assert nb == neighbors and np.array_equal(rg[0], ranges[0]) and np.array_equal(rg[1], ranges[1])
        mine.save_match(d, 3, matches)
        back = ref.read_npy(os.path.join(d, "matches_3.npy")).item()
        assert sorted(back) == sorted(matches) and np.array_equal(back[11], matches[11])
    # camera pose math: the reference's pure-Python quaternion -> rotation (util/geometry.py:40-58) on random,
    # unnormalised quaternions; pins base.CameraPose.R() and the oracle's projection
    spec = importlib.util.spec_from_file_location("reference_limap_util_geometry",
                                                  "/root/reference/src/limap/util/geometry.py")
    geo = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(geo)
    rng = np.random.default_rng(7)
    q = rng.normal(size=(64, 4)) * rng.uniform(0.1, 3.0, (64, 1))
    R = np.stack([geo.rotation_from_quaternion(x) for x in q])
    np.savez(os.path.join(OUT, "quaternion_rotation.npz"), qvec=q, R=R)
    print("golden files written to", OUT, sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
