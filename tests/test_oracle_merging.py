"""CPU checks of the merging oracle (oracle/orc_merging.cpp) against a second, pure-numpy restatement of the same
reference code (merging/merging_utils.cc:27-155, merging/merging.cc:513-598, merging/aggregator.cc:9-101) on
small inputs, plus exact-geometry cases with known answers."""
import numpy as np

from limap_b200.synth import make_track_lines, make_tracks
from oracle import oracle as orc

LK = dict(score_th=0.5, th_angle=5.0, th_overlap=0.001, th_smartoverlap=0.1, th_smartangle=1.0, th_perp=1.0,
          th_innerseg=1.0)


def _dir(l):
    d = l[3:6] - l[0:3]
    n = np.linalg.norm(d)
    return d / n if n > 0 else d


def _overlap(a, b):  # compute_overlap(l1=a, l2=b), line_dists.h:190-201
    v = _dir(b)
    ln = np.linalg.norm(b[3:6] - b[0:3])
    p1, p2 = np.dot(a[0:3] - b[0:3], v) / ln, np.dot(a[3:6] - b[0:3], v) / ln
    if p1 > p2:
        p1, p2 = p2, p1
    return min(p2, 1.0) - max(p1, 0.0)


def _innerseg(l1, l2):  # get_innerseg(l1, l2): the part of l2 inside l1's slab, line_dists.h:160-187
    d1 = _dir(l1)
    d2 = l2[3:6] - l2[0:3]
    den = np.dot(d2, d1) + 1e-12
    t1, t2 = np.dot(l1[0:3] - l2[0:3], d1) / den, np.dot(l1[3:6] - l2[0:3], d1) / den
    if t1 > t2:
        t1, t2 = t2, t1
    if t1 >= 1.0 or t2 <= 0.0:
        return None
    return np.concatenate([l2[0:3] + d2 * max(t1, 0.0), l2[0:3] + d2 * min(t2, 1.0)])


def _perp(a, b):  # dist_endpoints_perpendicular, line_dists.h:126-133
    def oneway(x, y):
        v = _dir(y)
        out = []
        for pt in (x[0:3], x[3:6]):
            d = pt - y[0:3]
            out.append(np.sqrt(max(np.dot(d, d) - np.dot(d, v) ** 2, 0.0)))
        return out
    return max(oneway(a, b) + oneway(b, a))


def _check(l1, l2, u1, u2, c):  # LineLinker3d::check_connection after set_to_spatial_merging
    mult = 1.0 / np.sqrt(-2.0 * np.log(c["score_th"]))
    ang = np.degrees(np.arccos(min(abs(np.dot(_dir(l1), _dir(l2))), 1.0)))
    if not ang <= c["th_angle"]:
        return False
    bio = max(_overlap(l1, l2), _overlap(l2, l1))
    if not bio > c["th_overlap"]:
        return False
    th = c["th_angle"]
    if bio < c["th_smartoverlap"]:
        r = min((c["th_smartoverlap"] - bio) / (c["th_smartoverlap"] - c["th_overlap"]), 1.0)
        th = c["th_angle"] - r * (c["th_angle"] - c["th_smartangle"])
    if np.exp(-(ang / (th * mult)) ** 2 / 2) < c["score_th"]:
        return False
    a, b = _innerseg(l2, l1), _innerseg(l1, l2)
    if a is None or b is None:
        return False
    d = _perp(a, b)
    return np.exp(-(d / (c["th_innerseg"] * min(u1, u2) * mult)) ** 2 / 2) >= c["score_th"]


def _remerge_py(L, active, c):
    T = len(L)
    edges = set()
    act = [i for i in range(T) if active[i]]
    for i in act:
        for j in range(T):
            if i == j:
                continue
            if len(act) == T:
                if i < j and (i + j) % 2 == 0:
                    continue
                if i > j and (i + j) % 2 == 1:
                    continue
            if _check(L[i, :6], L[j, :6], L[i, 6], L[j, 6], c):
                edges.add((min(i, j), max(i, j)))
    parent = [-1] * T
    size = [1] * T

    def root(x):
        while parent[x] != -1:
            x = parent[x]
        return x
    for a, b in sorted(edges):
        ra, rb = root(a), root(b)
        if ra != rb:
            if size[ra] < size[rb]:
                parent[ra] = rb
                size[rb] += size[ra]
            else:
                parent[rb] = ra
                size[ra] += size[rb]
    lab = [-1] * T
    n = 0
    for t in range(T):
        if parent[t] == -1:
            lab[t] = n
            n += 1
    for t in range(T):
        if lab[t] == -1:
            lab[t] = lab[root(t)]
    return np.array(lab, np.int32), n, len(edges)


def test_remerge_labels_against_numpy_restatement():
    for T, inactive, seed in ((60, 0.0, 1), (61, 0.0, 2), (80, 0.3, 3)):
        L = make_track_lines(T, dup_frac=0.5, seed=seed, extent=2.0)
        active = (np.random.default_rng(seed).uniform(size=T) >= inactive).astype(np.uint8)
        lab, ng, ne = orc.remerge_labels(L, active, LK)
        lab2, ng2, ne2 = _remerge_py(L, active, LK)
        assert (ng, ne) == (ng2, ne2) and np.array_equal(lab, lab2)
        assert ne > 5 and ng < T


def test_remerge_known_answers():
    # two collinear overlapping fragments merge; a parallel line 1 unit away and a crossing line do not
    L = np.array([[0, 0, 0, 1, 0, 0, 0.05], [0.5, 0.001, 0, 1.5, 0.001, 0, 0.05], [0, 1, 0, 1, 1, 0, 0.05],
                  [0.5, -0.5, 0, 0.5, 0.5, 0, 0.05], [10, 0, 0, 11, 0, 0, 0.05]], float)
    lab, ng, ne = orc.remerge_labels(L, np.ones(5, np.uint8), LK)
    assert ne == 1 and ng == 4 and lab[0] == lab[1] and len({lab[0], lab[2], lab[3], lab[4]}) == 4
    # inactive tracks are still merge targets of active ones, but two inactive tracks never merge
    lab, ng, ne = orc.remerge_labels(L, np.array([1, 0, 0, 0, 0], np.uint8), LK)
    assert ne == 1 and lab[0] == lab[1]
    lab, ng, ne = orc.remerge_labels(L, np.zeros(5, np.uint8), LK)
    assert ne == 0 and ng == 5


def test_aggregate_against_numpy_svd():
    rng = np.random.default_rng(5)
    d = np.array([1.0, 2.0, -0.5])
    d /= np.linalg.norm(d)
    n = 9
    mid = rng.normal(size=(n, 3)) * 0.02 + np.array([3.0, 1.0, 2.0])
    lines = np.zeros((n, 7))
    lines[:, :3] = mid - d * rng.uniform(0.5, 1.5, (n, 1))
    lines[:, 3:6] = mid + d * rng.uniform(0.5, 1.5, (n, 1))
    lines[:, 6] = rng.uniform(0.01, 0.1, n)
    sc = rng.uniform(0.1, 3.0, n)
    for no in (0, 2):
        out = orc.aggregate_lines(np.array([0, n]), lines, sc, no)[0]
        pts = np.concatenate([lines[:, :3], lines[:, 3:6]])
        c = pts.mean(0)
        _, _, vt = np.linalg.svd(pts - c)
        dirv = vt[0]
        proj = np.sort((pts - c) @ dirv)
        exp = np.concatenate([c + dirv * proj[no], c + dirv * proj[2 * n - 1 - no]])
        got = out[:6]
        assert min(np.abs(got - exp).max(), np.abs(got - exp[[3, 4, 5, 0, 1, 2]]).max()) <= 1e-9
        assert out[6] == lines[:, 6].min()
    # fewer than 4 lines: the best-scored line with the smallest uncertainty (aggregator.cc:9-29)
    out = orc.aggregate_lines(np.array([0, 3]), lines[:3], np.array([0.2, 0.9, 0.5]), 2)[0]
    assert np.array_equal(out[:6], lines[1, :6]) and out[6] == lines[:3, 6].min()


def test_support_flags_exact_geometry():
    ts = make_tracks(T=30, S=12, V=40, seed=7, noise_px=0.0, perturb=0.0)
    views, first = np.unique(ts.img_ids, return_index=True)
    remap = np.zeros(int(views.max()) + 1, np.int32)
    remap[views] = np.arange(len(views))
    a = (None, ts.kvec[first], ts.qvec[first], ts.tvec[first], ts.sup_off, remap[ts.img_ids], ts.segs, ts.gt)
    f = orc.track_support_flags(*a, th_angular_2d=0.01, th_perp_2d=1e-3, th_sv_angular_3d=89.9, th_overlap=0.05)
    # noise-free supports are sub-segments of the projection: zero angle and distance, full overlap of the support
    assert np.all(f & 1) and np.all(f & 4)
    # shifting every support by 3 px across the line breaks the distance test only
    segs = ts.segs.copy()
    d = segs[:, 2:] - segs[:, :2]
    nrm = np.stack([-d[:, 1], d[:, 0]], 1) / np.linalg.norm(d, axis=1)[:, None]
    segs[:, :2] += 3 * nrm
    segs[:, 2:] += 3 * nrm
    a2 = a[:6] + (segs, ts.gt)
    f2 = orc.track_support_flags(*a2, th_angular_2d=0.5, th_perp_2d=2.9, th_sv_angular_3d=89.9, th_overlap=0.05)
    assert not np.any(f2 & 1)
    f3 = orc.track_support_flags(*a2, th_angular_2d=0.5, th_perp_2d=3.1, th_sv_angular_3d=89.9, th_overlap=0.05)
    assert np.all(f3 & 1)
    # sensitivity bit against Line3d.sensitivity of the Python value types
    import limap.base as base
    t = 0
    line = base.Line3d(ts.gt[t, :3], ts.gt[t, 3:])
    for k in range(ts.sup_off[t], ts.sup_off[t + 1]):
        v = remap[ts.img_ids[k]]
        view = base.CameraView(base.Camera("PINHOLE", ts.kvec[first][v], 0, (600, 800)),
                               base.CameraPose(ts.qvec[first][v], ts.tvec[first][v]))
        s = line.sensitivity(view)
        fk = orc.track_support_flags(*a, th_sv_angular_3d=s + 1e-6)[k]
        fk2 = orc.track_support_flags(*a, th_sv_angular_3d=s - 1e-6)[k]
        assert (fk & 2) and not (fk2 & 2)
