"""GPU parity of the post-triangulation filters and the remerge: CUDA engine through
the C ABI vs the CPU oracle (oracle/orc_merging.cpp) on the same tracks. Boolean outputs and group labels are
compared bit-exactly, aggregated lines to 1e-9."""
import numpy as np
import pytest

from limap_b200.config import DEFAULT_YAML_TRIANGULATION, LINKER3D_DEFAULTS, make_linker
from limap_b200.synth import make_scene, make_track_lines

pytestmark = pytest.mark.gpu

REMERGE_LINKER = dict(score_th=0.5, th_angle=5.0, th_overlap=0.001, th_smartoverlap=0.1, th_smartangle=1.0,
                      th_perp=1.0, th_innerseg=1.0)  # cfgs/triangulation/default.yaml:99-108
FILTER = dict(th_angular_2d=8.0, th_perp_2d=5.0, th_sv_angular_3d=75.0, th_overlap=0.5)  # default.yaml:109-115


def _imagecols(sc):
    import limap.base as base
    cams = {0: base.Camera("SIMPLE_PINHOLE", [sc.kvec[0, 0], sc.kvec[0, 2], sc.kvec[0, 3]], 0, (600, 800))}
    imgs = {int(i): base.CameraImage(0, base.CameraPose(sc.qvec[v], sc.tvec[v]), f"img_{int(i)}.png")
            for v, i in enumerate(sc.img_ids)}
    return base.ImageCollection(cams, imgs)


def _tracks(sc):
    import limap.base as base
    import limap.triangulation as triangulation
    imagecols = _imagecols(sc)
    tri = triangulation.GlobalLineTriangulator(dict(DEFAULT_YAML_TRIANGULATION))
    tri.SetRanges(sc.ranges)
    tri.Init(base.get_all_lines_2d({int(i): sc.lines_of(v) for v, i in enumerate(sc.img_ids)}), imagecols)
    for img_id in imagecols.get_img_ids():
        tri.TriangulateImage(img_id, sc.matches[img_id])
    return imagecols, tri.ComputeLineTracks()


def test_support_flags_match_oracle():
    from limap_b200.engine import MergeEngine
    from limap_b200.merging import _flatten
    from oracle import oracle as orc
    sc = make_scene(V=12, L=250, N=8, K=8, seed=51)
    imagecols, tracks = _tracks(sc)
    assert len(tracks) > 100
    arr = _flatten(tracks, imagecols)
    got = MergeEngine().support_flags(*arr, **FILTER)
    exp = orc.track_support_flags(*arr, **FILTER)
    assert got.shape == exp.shape and len(got) > 500
    assert np.array_equal(got, exp)
    # tighter thresholds, PINHOLE default model ids; every bit is exercised in both states
    th = dict(th_angular_2d=1.0, th_perp_2d=0.8, th_sv_angular_3d=40.0, th_overlap=0.97)
    a2 = (None,) + tuple(arr[1:])
    got2 = MergeEngine().support_flags(*a2, **th)
    assert np.array_equal(got2, orc.track_support_flags(*a2, **th))
    for bit in (1, 2, 4):
        assert 0 < int((got2 & bit).astype(bool).sum()) < len(got2), bit


def test_aggregate_lines_match_oracle():
    from limap_b200.engine import MergeEngine
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    sizes = np.array([1, 2, 3, 4, 5, 9, 30, 0, 7])
    off = np.concatenate([[0], np.cumsum(sizes)])
    n = int(off[-1])
    base_dir = rng.normal(size=3)
    lines = np.zeros((n, 7))
    mid = rng.normal(size=(n, 3)) * 0.05
    lines[:, :3] = mid - base_dir * rng.uniform(0.5, 1.5, (n, 1))
    lines[:, 3:6] = mid + base_dir * rng.uniform(0.5, 1.5, (n, 1))
    lines[:, 6] = rng.uniform(0.01, 0.1, n)
    scores = rng.uniform(0, 5, n)
    for no in (0, 2):
        got = MergeEngine.aggregate(off, lines, scores, no)
        exp = orc.aggregate_lines(off, lines, scores, no)
        # the TLS direction is defined up to sign: endpoints may swap
        d = np.minimum(np.abs(got - exp).max(1), np.abs(got[:, [3, 4, 5, 0, 1, 2, 6]] - exp).max(1))
        assert d.max() <= 1e-9


@pytest.mark.parametrize("T,inactive", [(1, 0.0), (2, 0.0), (700, 0.0), (700, 0.4), (3000, 0.0), (3000, 0.1)])
def test_remerge_labels_match_oracle(T, inactive):
    from limap_b200.engine import MergeEngine
    from oracle import oracle as orc
    L = make_track_lines(T, dup_frac=0.35, seed=T + int(100 * inactive))
    rng = np.random.default_rng(T)
    active = (rng.uniform(size=T) >= inactive).astype(np.uint8)
    lab, ng, ne = MergeEngine().remerge_labels(L, active, make_linker(LINKER3D_DEFAULTS, REMERGE_LINKER))
    lab_o, ng_o, ne_o = orc.remerge_labels(L, active, REMERGE_LINKER)
    assert (ng, ne) == (ng_o, ne_o)
    assert np.array_equal(lab, lab_o)
    if T >= 700:
        assert ne > T // 10 and ng < T


def test_remerge_without_angle_gate_and_wide_thresholds():
    """th_angle >= 89 switches the fp32 gate off; every pair goes through the fp64 check."""
    from limap_b200.engine import MergeEngine
    from oracle import oracle as orc
    L = make_track_lines(400, dup_frac=0.4, seed=9, extent=3.0)
    lk = dict(REMERGE_LINKER, th_angle=89.5, th_smartangle=30.0, th_innerseg=4.0)
    lab, ng, ne = MergeEngine().remerge_labels(L, np.ones(400, np.uint8), make_linker(LINKER3D_DEFAULTS, lk))
    lab_o, ng_o, ne_o = orc.remerge_labels(L, np.ones(400, np.uint8), lk)
    assert (ng, ne) == (ng_o, ne_o) and np.array_equal(lab, lab_o)


def test_runner_sequence_filters_and_remerge():
    """runners/line_triangulation.py:171-200 through limap.merging, against the same sequence rebuilt from oracle
    flags / labels on flat arrays."""
    import limap.base as base
    import limap.merging as merging
    from limap_b200.merging import _flatten
    from oracle import oracle as orc
    sc = make_scene(V=12, L=250, N=8, K=8, seed=52)
    imagecols, tracks = _tracks(sc)
    n0 = len(tracks)
    t1 = merging.filter_tracks_by_reprojection(tracks, imagecols, 8.0, 5.0)
    # oracle: same selection
    arr = _flatten(tracks, imagecols)
    f = orc.track_support_flags(*arr, **FILTER)
    off = arr[4]
    keep = [[k for k in range(off[t], off[t + 1]) if f[k] & 1] for t in range(n0)]
    keep = [k for k in keep if k]
    assert len(t1) == len(keep)
    for tr, ks in zip(t1, keep):
        assert tr.count_lines() == len(ks)
    assert merging.check_track_by_reprojection(tracks[0], imagecols, 8.0, 5.0) == \
        [bool(x & 1) for x in f[off[0]:off[1]]]
    linker3d = base.LineLinker3d(REMERGE_LINKER)
    t2 = merging.remerge(linker3d, t1)
    assert 0 < len(t2) <= len(t1)
    assert sum(t.count_lines() for t in t2) == sum(t.count_lines() for t in t1)
    # first pass against the oracle labels
    TL = np.array([np.concatenate([t.line.start, t.line.end, [t.line.uncertainty]]) for t in t1])
    lab_o, ng_o, _ = orc.remerge_labels(TL, np.ones(len(t1), np.uint8), REMERGE_LINKER)
    one = merging._remerge_once(t1, make_linker(LINKER3D_DEFAULTS, REMERGE_LINKER), 2)
    assert len(one) == ng_o
    sizes = np.bincount(lab_o, minlength=ng_o)
    assert [t.active for t in one] == [bool(s > 1) for s in sizes]
    t3 = merging.filter_tracks_by_reprojection(t2, imagecols, 8.0, 5.0)
    t4 = merging.filter_tracks_by_sensitivity(t3, imagecols, 75.0, 3)
    t5 = merging.filter_tracks_by_overlap(t4, imagecols, 0.5, 3)
    assert len(t5) <= len(t4) <= len(t3) <= len(t2)
    arr5 = _flatten(t3, imagecols)
    f5 = orc.track_support_flags(*arr5, **FILTER)
    o5 = arr5[4]
    n_sens = sum(len({t3[t].image_id_list[k - o5[t]] for k in range(o5[t], o5[t + 1]) if f5[k] & 2}) >= 3
                 for t in range(len(t3)))
    assert len(t4) == n_sens
    assert merging.check_sensitivity(t3[0], imagecols, 75.0) == [bool(x & 2) for x in f5[o5[0]:o5[1]]]
