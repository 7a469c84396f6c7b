"""The limap-style operator surface ([D] triangulation and [E] line bundle adjustment of
src/limap/runners/line_triangulation.py:99-219) driven the way the runner drives it, checked against the
oracle on the same scene."""
import numpy as np
import pytest

from limap_b200.config import DEFAULT_YAML_TRIANGULATION
from limap_b200.synth import make_scene

pytestmark = pytest.mark.gpu


def _imagecols(sc):
    import limap.base as base
    cams = {0: base.Camera("SIMPLE_PINHOLE", [sc.kvec[0, 0], sc.kvec[0, 2], sc.kvec[0, 3]], 0, (600, 800))}
    imgs = {int(i): base.CameraImage(0, base.CameraPose(sc.qvec[v], sc.tvec[v]), f"img_{int(i)}.png")
            for v, i in enumerate(sc.img_ids)}
    return base.ImageCollection(cams, imgs)


def test_runner_style_triangulation_and_ba():
    import limap.base as base
    import limap.optimize as optimize
    import limap.triangulation as triangulation
    from oracle import oracle as orc
    sc = make_scene(V=10, L=150, N=6, K=6, seed=41)
    cfg = dict(DEFAULT_YAML_TRIANGULATION)
    imagecols = _imagecols(sc)
    all_2d_segs = {int(i): sc.lines_of(v) for v, i in enumerate(sc.img_ids)}
    # [D]
    Triangulator = triangulation.GlobalLineTriangulator(cfg)
    Triangulator.SetRanges(sc.ranges)
    all_2d_lines = base.get_all_lines_2d(all_2d_segs)
    Triangulator.Init(all_2d_lines, imagecols)
    for img_id in imagecols.get_img_ids():
        Triangulator.TriangulateImage(img_id, sc.matches[img_id])
    linetracks = Triangulator.ComputeLineTracks()
    assert Triangulator.CountImages() == 10 and Triangulator.CountLines(0) == 150
    # oracle on the same data
    o = orc.OracleTri(cfg)
    o.upload(sc)
    o.set_ranges(*sc.ranges)
    for i in sc.img_ids:
        o.add_image_matches(int(i), *sc.flat_matches(int(i)))
    ot = o.build_tracks()
    assert len(linetracks) == len(ot["track_off"]) - 1
    got = {frozenset(zip(t.image_id_list, t.line_id_list)) for t in linetracks}
    exp = {frozenset(zip(ot["img_ids"][a:b].tolist(), ot["line_ids"][a:b].tolist()))
           for a, b in zip(ot["track_off"][:-1], ot["track_off"][1:])}
    assert got == exp
    t0 = linetracks[0]
    assert len(t0.line2d_list) == len(t0.line3d_list) == len(t0.score_list) == t0.count_lines()
    assert isinstance(t0.line2d_list[0], base.Line2d) and isinstance(t0.line, base.Line3d)
    # [E] refinement exactly as the runner calls it
    cfg_ref = dict(constant_intrinsics=True, constant_principal_point=True, constant_pose=True, constant_line=False,
                   min_num_images=4, num_outliers_aggregator=2, use_geometric=True, geometric_alpha=10.0)
    ba_engine = optimize.solve_line_bundle_adjustment(cfg_ref, imagecols, linetracks, max_num_iterations=200)
    out = ba_engine.GetOutputLineTracks(num_outliers=2)
    assert sorted(out) == list(range(len(linetracks)))
    # oracle refinement on the same tracks
    from limap_b200.optimize import _tracks_to_arrays
    from limap_b200.synth import TrackSet
    view_of = {int(i): v for v, i in enumerate(sc.img_ids)}
    sup_off, sup_view, segs, l3d, init = _tracks_to_arrays(linetracks, view_of)
    ts = TrackSet(sup_off=sup_off, segs=segs, kvec=sc.kvec[sup_view], qvec=sc.qvec[sup_view], tvec=sc.tvec[sup_view],
                  img_ids=sup_view.astype(np.int32), line3d=l3d, line_init=init, gt=init)
    ref = orc.refine_tracks(ts, max_num_iterations=200, min_num_images=4, num_outliers=2)
    lines = np.array([np.concatenate([out[k].line.start, out[k].line.end]) for k in range(len(linetracks))])
    d = np.minimum(np.abs(lines - ref["line"]).max(1), np.abs(lines - ref["line"][:, [3, 4, 5, 0, 1, 2]]).max(1))
    assert d.max() <= 1e-4
    moved = np.abs(lines - init).max(1) > 1e-9
    n_img = np.array([t.count_images() for t in linetracks])
    assert moved[n_img >= 4].mean() > 0.9
    # GetOutputLineTracks with another num_outliers only re-cuts the refined infinite lines (no second solve): the result
    # is the one a solve configured with that trim gives
    out0 = ba_engine.GetOutputLineTracks(num_outliers=0)
    ref0 = orc.refine_tracks(ts, max_num_iterations=200, min_num_images=4, num_outliers=0)
    lines0 = np.array([np.concatenate([out0[k].line.start, out0[k].line.end]) for k in range(len(linetracks))])
    d0 = np.minimum(np.abs(lines0 - ref0["line"]).max(1), np.abs(lines0 - ref0["line"][:, [3, 4, 5, 0, 1, 2]]).max(1))
    ok0 = np.isfinite(ref0["line"]).all(1)
    assert ok0.sum() > 10 and d0[ok0].max() <= 1e-4
    assert np.abs(lines0 - lines).max() > 1e-6  # (the untrimmed cut is a different segment)


def test_triangulate_image_index_error_and_model_check():
    import limap.base as base
    import limap.triangulation as triangulation
    sc = make_scene(V=4, L=20, N=2, K=2, seed=42)
    imagecols = _imagecols(sc)
    tri = triangulation.GlobalLineTriangulator(dict(DEFAULT_YAML_TRIANGULATION))
    tri.Init(base.get_all_lines_2d({int(i): sc.lines_of(v) for v, i in enumerate(sc.img_ids)}), imagecols)
    bad = {k: v.copy() for k, v in sc.matches[0].items()}
    next(iter(bad.values()))[0, 0] = 999
    with pytest.raises(RuntimeError, match="IndexError"):
        tri.TriangulateImage(0, bad)
    with pytest.raises(RuntimeError):
        base.Camera("OPENCV", [1, 1, 0, 0, 0, 0, 0, 0], 0, (10, 10))


def test_disk_formats_and_device_tensor_ingress(tmp_path):
    """§8(f) rank 2: the runner's inputs read back from the reference's file formats, and matches handed over as CUDA
    tensors (GPU matcher output), give the tracks of the in-memory host path."""
    import torch
    import limap.base as base
    import limap.triangulation as triangulation
    import limap.util.io as limapio
    sc = make_scene(V=8, L=120, N=5, K=5, seed=43)
    cfg = dict(DEFAULT_YAML_TRIANGULATION)
    imagecols = _imagecols(sc)
    d = str(tmp_path)
    os_segs = {int(i): sc.lines_of(v) for v, i in enumerate(sc.img_ids)}
    for i, s in os_segs.items():
        limapio.save_txt_segments(d, i, s)
        limapio.save_match(d, i, sc.matches[i])
    neighbors = {int(i): sorted(sc.matches[int(i)].keys()) for i in sc.img_ids}
    limapio.save_txt_metainfos(d + "/metainfos.txt", neighbors, sc.ranges)
    limapio.save_npy(d + "/imagecols.npy", imagecols.as_dict())

    def run(segs, ranges, ic, matches_of):
        tri = triangulation.GlobalLineTriangulator(cfg)
        tri.SetRanges(ranges)
        tri.Init(base.get_all_lines_2d(segs), ic)
        for img_id in ic.get_img_ids():
            tri.TriangulateImage(img_id, matches_of(img_id))
        return tri.ComputeLineTracks()

    ref = run(os_segs, sc.ranges, imagecols, lambda i: sc.matches[i])
    # from disk
    nb2, ranges2 = limapio.read_txt_metainfos(d + "/metainfos.txt")
    ic2 = base.ImageCollection(limapio.read_npy(d + "/imagecols.npy").item())
    segs2 = {i: limapio.read_txt_segments(d, i) for i in ic2.get_img_ids()}
    assert nb2 == neighbors
    got = run(segs2, ranges2, ic2, lambda i: limapio.read_match(d, i))
    # from device tensors
    dev = run(os_segs, sc.ranges, imagecols,
              lambda i: {g: torch.as_tensor(np.asarray(m), device="cuda") for g, m in sc.matches[i].items()})
    for other in (got, dev):
        assert len(other) == len(ref) > 20
        for a, b in zip(ref, other):
            assert a.image_id_list == b.image_id_list and a.line_id_list == b.line_id_list
            assert np.abs(a.line.as_array() - b.line.as_array()).max() <= 1e-9
    # tracks written and read back
    limapio.save_folder_linetracks(d + "/finaltracks", ref)
    back = limapio.read_folder_linetracks(d + "/finaltracks")
    assert len(back) == len(ref)
    assert all(x.line_id_list == y.line_id_list and x.node_id_list == y.node_id_list for x, y in zip(ref, back))
    assert max(np.abs(x.line.as_array() - y.line.as_array()).max() for x, y in zip(ref, back)) <= 1e-9


def test_valid_tris_getters_match_oracle_debug_mode():
    """GetValidScoredTrisNode / ...NodeSet / GetValidTrisNode / GetValidTrisImage / GetAllValidTris
    (global_line_triangulator.cc:380-470): valid_tris_ holds the candidates with score >= fullscore_th in
    std::greater<(score, tri_id)> order, capped at max_valid_conns."""
    import limap.base as base
    import limap.triangulation as triangulation
    from oracle import oracle as orc
    sc = make_scene(V=6, L=60, N=4, K=6, seed=43)
    for max_conns in (1000, 2):
        cfg = dict(DEFAULT_YAML_TRIANGULATION, debug_mode=True, max_valid_conns=max_conns)
        tri = triangulation.GlobalLineTriangulator(cfg)
        tri.SetRanges(sc.ranges)
        tri.Init(base.get_all_lines_2d({int(i): sc.lines_of(v) for v, i in enumerate(sc.img_ids)}), _imagecols(sc))
        for img_id in sc.img_ids:
            tri.TriangulateImage(int(img_id), sc.matches[int(img_id)])
        o = orc.OracleTri(cfg)
        o.upload(sc)
        o.set_ranges(*sc.ranges)
        for i in sc.img_ids:
            o.add_image_matches(int(i), *sc.flat_matches(int(i)))
        n_valid_total = 0
        for i in sc.img_ids[:3]:
            i = int(i)
            ooff, oe = o.get_valid_edges(i)
            n_img = 0
            for l in range(tri.CountLines(i)):
                cl, cng = o.get_cands_node(i, l)
                order = sorted(range(len(cl)), key=lambda k: (-cl[k, 9], -k))[:max_conns]
                exp = [k for k in order if cl[k, 9] >= cfg["fullscore_th"]]
                got = tri.GetValidScoredTrisNode(i, l)
                assert [g[2] for g in got] == [tuple(int(x) for x in cng[k]) for k in exp]
                assert np.allclose([g[1] for g in got], cl[exp, 9], atol=1e-6)
                assert sorted(g[2] for g in got) == sorted(map(tuple, oe[ooff[l]:ooff[l + 1]].tolist()))
                assert len(tri.GetValidTrisNode(i, l)) == len(got)
                ns = tri.GetValidScoredTrisNodeSet(i, l)
                assert len(ns) == len({g[2][0] for g in got}) and all(isinstance(x[0], base.Line3d) for x in ns)
                assert len(tri.GetValidTrisNodeSet(i, l)) == len(ns)
                n_img += len(got)
            assert len(tri.GetValidTrisImage(i)) == n_img == len(oe)
            n_valid_total += n_img
        assert n_valid_total > 0
        assert len(tri.GetAllValidTris()) == tri.CountAllValidTris()
