"""Oracle vs the reference's own known-answer vector and vs ground-truth self-checks (CPU only).

The only test in the reference that touches a hot-path type is tests/base/test_linebase.py:8-17
(Line2d((0,0),(1,1)): length sqrt(2), direction (1,1)/sqrt(2)). Everything else is 'parity unpinned'
by the reference, so the remaining checks here are exact-geometry self-checks:
noise-free projections of a known 3D line must triangulate back to it.
"""
import numpy as np

from oracle import oracle as orc


def test_reference_kat_line2d(oracle_lib):
    seg = np.array([0.0, 0.0, 1.0, 1.0])
    assert abs(oracle_lib.orc_line2d_length(orc._p(seg)) - np.sqrt(2)) < 1e-15
    d = np.zeros(2)
    oracle_lib.orc_line2d_direction(orc._p(seg), orc._p(d))
    assert np.allclose(d, np.array([1.0, 1.0]) / np.sqrt(2), atol=1e-15)


def _cam(C, seed):
    from limap_b200.synth import _rot_to_quat
    rng = np.random.default_rng(seed)
    z = -C / np.linalg.norm(C)
    x = np.cross(z, np.array([0, 0, 1.0]) + rng.normal(scale=0.1, size=3))
    x /= np.linalg.norm(x)
    R = np.stack([x, np.cross(z, x), z])
    return orc.cam_array(0, [692.82, 692.82, 400, 300], _rot_to_quat(R), -R @ C)


def test_triangulate_exact_projections(oracle_lib):
    rng = np.random.default_rng(0)
    for k in range(20):
        X0, X1 = rng.uniform(-3, 3, 3), rng.uniform(-3, 3, 3)
        c1 = _cam(np.array([12.0, 1.0, 0.5]), k)
        c2 = _cam(np.array([10.0, 6.0, -1.0]), k + 100)
        l1, l2 = np.zeros(4), np.zeros(4)
        for cam, l in ((c1, l1), (c2, l2)):
            for e, X in enumerate((X0, X1)):
                p = np.zeros(2)
                oracle_lib.orc_project_point(orc._p(cam), orc._p(X), orc._p(p))
                l[2 * e:2 * e + 2] = p
        # endpoints of l2 slide along the same 2D line: the plane-pair intersection must not care
        d = l2[2:] - l2[:2]
        l2b = np.concatenate([l2[:2] - 0.3 * d, l2[2:] + 0.2 * d])
        out = np.zeros(9)
        oracle_lib.orc_triangulate_line(orc._p(l1), orc._p(c1), orc._p(l2b), orc._p(c2), 0, orc._p(out))
        assert out[8] == 1.0
        assert np.allclose(out[:3], X0, atol=1e-7) and np.allclose(out[3:6], X1, atol=1e-7)
        iou = oracle_lib.orc_compute_epipolar_IoU(orc._p(l1), orc._p(c1), orc._p(l2), orc._p(c2))
        assert abs(iou - 1.0) < 1e-4  # dehomogeneous() adds EPS to a normalised w
        oracle_lib.orc_triangulate_line(orc._p(l1), orc._p(c1), orc._p(l2), orc._p(c2), 1, orc._p(out))
        assert np.allclose(out[:3], X0, atol=1e-7) and np.allclose(out[3:6], X1, atol=1e-7)


def test_linker_scores_thresholds(oracle_lib):
    import ctypes as C
    from limap_b200.config import LINKER2D_DEFAULTS, make_linker
    lk = make_linker(LINKER2D_DEFAULTS, dict(score_th=0.5, th_angle=5.0, th_perp=2.0, th_overlap=0.05))
    a = np.array([0.0, 0.0, 100.0, 0.0])
    same = oracle_lib.orc_score_2d(C.byref(lk), orc._p(a), orc._p(a))
    assert same == 1.0
    # 2 px perpendicular offset is exactly the threshold: score == score_th (up to rounding) or 0
    b = np.array([0.0, 2.0, 100.0, 2.0])
    s = oracle_lib.orc_score_2d(C.byref(lk), orc._p(a), orc._p(b))
    assert s == 0.0 or abs(s - 0.5) < 1e-12
    b = np.array([0.0, 1.0, 100.0, 1.0])
    s = oracle_lib.orc_score_2d(C.byref(lk), orc._p(a), orc._p(b))
    assert abs(s - np.exp(-0.5 * (1.0 / (2.0 / np.sqrt(-2 * np.log(0.5)))) ** 2)) < 1e-12
    # 6 degrees apart fails the 5 degree angle test
    t = np.deg2rad(6.0)
    b = np.array([0.0, 0.0, 100 * np.cos(t), 100 * np.sin(t)])
    assert oracle_lib.orc_score_2d(C.byref(lk), orc._p(a), orc._p(b)) == 0.0
    # no overlap along the line
    b = np.array([200.0, 0.0, 300.0, 0.0])
    assert oracle_lib.orc_score_2d(C.byref(lk), orc._p(a), orc._p(b)) == 0.0


def test_oracle_recovers_ground_truth_tracks():
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from limap_b200.synth import make_scene
    sc = make_scene(V=8, L=60, N=5, K=3, seed=7, noise_px=0.2)
    o = orc.OracleTri(dict(DEFAULT_YAML_TRIANGULATION))
    o.upload(sc)
    o.set_ranges(*sc.ranges)
    for i in sc.img_ids:
        o.add_image_matches(int(i), *sc.flat_matches(int(i)))
    assert o.rows_tested() == sc.n_rows()
    tr = o.build_tracks()
    T = len(tr["track_off"]) - 1
    assert T > 20
    view = {int(i): v for v, i in enumerate(sc.img_ids)}
    pure, close = 0, 0
    for t in range(T):
        a, b = tr["track_off"][t], tr["track_off"][t + 1]
        gts = {int(sc.gt_id[sc.line_off[view[int(i)]] + l]) for i, l in zip(tr["img_ids"][a:b], tr["line_ids"][a:b])}
        if len(gts) == 1 and -1 not in gts and b - a >= 4:
            pure += 1
            g = sc.gt_lines[gts.pop()]
            L = tr["track_line"][t]
            d = g[3:] - g[:3]
            d /= np.linalg.norm(d)
            dist = max(np.linalg.norm(np.cross(L[:3] - g[:3], d)), np.linalg.norm(np.cross(L[3:6] - g[:3], d)))
            close += dist < 0.05
    assert pure > 0.5 * T
    assert close > 0.9 * pure


def test_pose_rotation_against_reference_python_golden(oracle_lib):
    """tests/golden/io/quaternion_rotation.npz holds rotation matrices computed by the reference's own pure-Python
    rotation_from_quaternion (util/geometry.py:40-58) on unnormalised quaternions (tests/golden/make_io_golden.py).
    Pins CameraPose::R() of the Python value types and the pose math inside the oracle's projection."""
    import os
    from limap_b200 import base
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io", "quaternion_rotation.npz"))
    rng = np.random.default_rng(1)
    K = np.array([[500.0, 0, 320.0], [0, 480.0, 240.0], [0, 0, 1.0]])
    for q, R in zip(z["qvec"], z["R"]):
        assert np.abs(base.CameraPose(q, np.zeros(3)).R() - R).max() <= 2e-15
        T = rng.normal(size=3)
        X = R.T @ (np.array([0.3, -0.2, 4.0]) - T)  # in front of the camera
        h = K @ (R @ X + T)
        exp = h[:2] / (h[2] + 1e-12)
        cam = orc.cam_array(1, [500.0, 480.0, 320.0, 240.0], q, T)
        p = np.zeros(2)
        oracle_lib.orc_project_point(orc._p(cam), orc._p(X), orc._p(p))
        assert np.abs(p - exp).max() <= 1e-9


def test_node_parallel_schedule_gives_identical_results():
    """The oracle's throughput schedule (OpenMP over the 2D lines of an image) is the reference-structured restatement
    with the loop moved outwards: every output is bit-identical for any thread count."""
    from limap_b200.config import DEFAULT_YAML_TRIANGULATION
    from limap_b200.synth import make_scene
    sc = make_scene(V=8, L=120, N=5, K=6, seed=77)
    outs = []
    for mode, th in ((False, 1), (True, 1), (True, 4)):
        o = orc.OracleTri(dict(DEFAULT_YAML_TRIANGULATION), threads=th, node_parallel=mode)
        o.upload(sc)
        o.set_ranges(*sc.ranges)
        for i in sc.img_ids:
            o.add_image_matches(int(i), *sc.flat_matches(int(i)))
        tr = o.build_tracks()
        outs.append(([o.get_best(int(i)) for i in sc.img_ids], [o.get_valid_edges(int(i)) for i in sc.img_ids], tr,
                     o.rows_tested()))
    ref = outs[0]
    for other in outs[1:]:
        assert other[3] == ref[3]
        for a, b in zip(ref[0], other[0]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        for a, b in zip(ref[1], other[1]):
            assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert np.array_equal(ref[2]["track_off"], other[2]["track_off"])
        assert np.array_equal(ref[2]["track_line"], other[2]["track_line"])
