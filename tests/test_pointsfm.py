"""Visual-neighbour ranking and robust ranges: the oracle restatement on hand-checkable inputs (CPU)
and the CUDA path against it (GPU), through limap.pointsfm.SfmModel as pointsfm/functions.py drives the reference's."""
import numpy as np
import pytest

from limap_b200.synth import make_scene, make_sfm_points


def test_oracle_ranking_on_a_hand_made_model():
    from oracle import oracle as orc
    # three cameras on a line looking down -z at points 10 units away; image 0 and 1 share 3 points, 1 and 2 share 1
    centres = np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0]], float)
    xyz = np.array([[0.5, 0, 10], [0.6, 0.2, 10], [0.4, -0.1, 10], [1.5, 0, 10], [1.0, 1.0, 10]], float)
    off = np.array([0, 2, 4, 6, 8, 9])
    img = np.array([0, 1, 0, 1, 0, 1, 1, 2, 1], np.int32)  # the last point is seen by image 1 only
    out, cnt = orc.rank_neighbors(centres, xyz, off, img, 2, min_triangulation_angle=1.0, mode=0)
    assert cnt.tolist() == [1, 2, 1] and out[0, 0] == 1 and out[1].tolist() == [0, 2] and out[2, 0] == 1
    # a 45-degree threshold removes every pair (baseline 1 at depth 10 is about 5.7 degrees)
    out, cnt = orc.rank_neighbors(centres, xyz, off, img, 2, min_triangulation_angle=45.0, mode=0)
    assert cnt.tolist() == [0, 0, 0]
    # IoU of (0,1): 3 shared / (3 + 5 - 3); Dice = 6 / 8: both rank 0 before 2 for image 1
    out_d, _ = orc.rank_neighbors(centres, xyz, off, img, 2, min_triangulation_angle=1.0, mode=1)
    assert out_d[1].tolist() == [0, 2]
    lo, hi = orc.robust_ranges(np.column_stack([np.arange(100.0), np.zeros(100), -np.arange(100.0)]), 0.05, 0.95, 1.25)
    assert np.allclose(lo, [5 - 1.25 * 90, 0, -94 - 1.25 * 90]) and np.allclose(hi, [95 + 1.25 * 90, 0, -4 + 1.25 * 90])


def _model(sc, centres, xyz, off, img):
    import limap.pointsfm as pointsfm
    m = pointsfm.SfmModel()
    from limap_b200.base import CameraPose
    for v, i in enumerate(sc.img_ids):
        k = sc.kvec[v]
        K = [[k[0], 0, k[2]], [0, k[1], k[3]], [0, 0, 1]]
        m.addImage(pointsfm.CreateSfmImage(f"img_{int(i)}.png", 800, 600, K, CameraPose(sc.qvec[v], sc.tvec[v]).R(), sc.tvec[v]),
                   int(i))
    for p in range(len(xyz)):
        m.addPoint(*xyz[p], img[off[p]:off[p + 1]])
    return m


@pytest.mark.gpu
@pytest.mark.parametrize("ntype,mode", [("iou", 0), ("dice", 1), ("overlap", 2)])
def test_neighbor_ranking_matches_oracle(ntype, mode):
    import limap.pointsfm as pointsfm
    from oracle import oracle as orc
    sc = make_scene(V=40, L=10, N=3, K=1, seed=61, id_stride=3)
    centres, xyz, off, img = make_sfm_points(sc, n_points=6000, seed=5)
    m = _model(sc, centres, xyz, off, img)
    for n_nb, ang in ((20, 1.0), (5, 8.0)):
        got = pointsfm.compute_neighbors(m, n_nb, min_triangulation_angle=ang, neighbor_type=ntype)
        # the model keeps float poses and points like colmap::mvs::Model: the oracle gets what the model holds
        c_m = np.stack([im.centre() for im in m.images])
        xyz_m = m._arrays()[0]
        exp, cnt = orc.rank_neighbors(c_m, xyz_m, off, img, n_nb, min_triangulation_angle=ang, mode=mode)
        ids = sc.img_ids
        assert sorted(got) == [int(i) for i in ids]
        n_total = 0
        for v, i in enumerate(ids):
            assert got[int(i)] == [int(ids[j]) for j in exp[v, :cnt[v]]], (ntype, n_nb, int(i))
            n_total += cnt[v]
        assert n_total > 40 * min(n_nb, 5) // 2
    assert m.ComputeNumPoints() == np.bincount(img, minlength=40).tolist()


@pytest.mark.gpu
def test_robust_ranges_and_metainfos_match_oracle():
    import limap.pointsfm as pointsfm
    from oracle import oracle as orc
    sc = make_scene(V=20, L=10, N=3, K=1, seed=62, scale=100.0)
    centres, xyz, off, img = make_sfm_points(sc, n_points=5000, seed=6)
    m = _model(sc, centres, xyz, off, img)
    cfg = dict(min_triangulation_angle=1.0, neighbor_type="dice", ranges=dict(range_robust=[0.05, 0.95], k_stretch=1.25))
    neighbors, ranges = pointsfm.compute_metainfos(cfg, m, n_neighbors=10)
    lo, hi = orc.robust_ranges(xyz, 0.05, 0.95, 1.25)
    assert np.array_equal(ranges[0], lo) and np.array_equal(ranges[1], hi)  # float arithmetic, bit-exact
    assert all(len(v) <= 10 for v in neighbors.values()) and sum(len(v) for v in neighbors.values()) > 100
