"""CPU known-answer checks of the LM and J-Linkage oracles (oracle/orc_lm.*, oracle/orc_vp.*): neither Ceres nor
the JLinkage library can be run here, so the restatements are anchored on geometry with known answers."""
import numpy as np

from limap_b200.synth import make_tracks
from oracle import oracle as orc


def _line_dist(line, gt):
    d = gt[3:] - gt[:3]
    d = d / np.linalg.norm(d)
    return max(np.linalg.norm(np.cross(line[:3] - gt[:3], d)), np.linalg.norm(np.cross(line[3:] - gt[:3], d)))


def test_lm_converges_to_ground_truth_on_noise_free_supports():
    ts = make_tracks(T=40, S=12, V=60, seed=11, noise_px=0.0, perturb=0.05)
    out = orc.refine_tracks(ts, max_num_iterations=100, min_num_images=4)
    after = np.array([_line_dist(out["line"][t], ts.gt[t]) for t in range(ts.n_tracks)])
    before = np.array([_line_dist(ts.line_init[t], ts.gt[t]) for t in range(ts.n_tracks)])
    assert np.median(before) > 1e-2 and after.max() < 1e-5  # infinite line recovered from exact observations
    assert (out["cost"][:, 1] <= 1e-12).all() and (out["cost"][:, 0] > 1e-3).all()
    assert (out["iters"][:, 0] > 0).all()
    # minimal Pluecker form: unit quaternion (direction frame) and unit 2-vector (moment split)
    m = out["minimal"]
    assert np.abs(np.linalg.norm(m[:, :4], axis=1) - 1).max() < 1e-12
    assert np.abs(np.linalg.norm(m[:, 4:], axis=1) - 1).max() < 1e-12


def test_lm_noise_lowers_cost_and_small_tracks_stay_constant():
    ts = make_tracks(T=30, S=10, V=50, seed=12, noise_px=0.5, perturb=0.05)
    out = orc.refine_tracks(ts, max_num_iterations=100, min_num_images=4)
    assert (out["cost"][:, 1] <= out["cost"][:, 0] + 1e-12).all()
    hi = orc.refine_tracks(ts, max_num_iterations=100, min_num_images=11)  # every track has 10 images: all constant
    assert (hi["iters"] == 0).all() and np.array_equal(hi["cost"][:, 0], hi["cost"][:, 1])
    # determinism, and independence of the OpenMP schedule
    a = orc.refine_tracks(ts, max_num_iterations=100, threads=1)
    b = orc.refine_tracks(ts, max_num_iterations=100, threads=4)
    assert np.array_equal(a["line"], b["line"]) and np.array_equal(a["iters"], b["iters"])


def _vp_image(rng, vps, counts, clutter, noise=0.2):
    segs, truth = [], []
    for c, (vp, n) in enumerate(zip(vps, counts)):
        for _ in range(n):
            p = rng.uniform([0, 0], [800, 600])
            d = vp[:2] / vp[2] - p
            d = d / np.linalg.norm(d)
            L = rng.uniform(45, 160)
            segs.append([*(p + rng.normal(scale=noise, size=2)), *(p + d * L + rng.normal(scale=noise, size=2))])
            truth.append(c)
    for _ in range(clutter):
        a = rng.uniform([0, 0], [800, 600])
        ang = rng.uniform(0, np.pi)
        segs.append([*a, *(a + rng.uniform(45, 150) * np.array([np.cos(ang), np.sin(ang)]))])
        truth.append(-1)
    return np.asarray(segs), np.asarray(truth)


def test_jlinkage_recovers_planted_vanishing_points():
    rng = np.random.default_rng(4)
    vps = [np.array([3000.0, 320.0, 1.0]), np.array([410.0, -3500.0, 1.0]), np.array([395.0, 290.0, 1.0])]
    segs, truth = _vp_image(rng, vps, [40, 35, 30], 20)
    short = np.array([[10.0, 10.0, 20.0, 12.0]])  # below min_length: label -1, not part of the clustering
    allsegs = np.concatenate([segs, short])
    off = np.array([0, len(allsegs), len(allsegs)], np.int64)  # second image is empty
    lab, vp_off, out = orc.detect_vps(off, allsegs, min_length=40, inlier_threshold=1.0, min_num_supports=10, seed=3)
    assert lab[-1] == -1 and vp_off[1] == vp_off[2] and vp_off[1] >= 3
    # every planted family is dominated by one label, and that label's VP is the planted point (as a direction)
    used = set()
    for c, vp in enumerate(vps):
        fam = lab[:len(segs)][truth == c]
        vals, cnt = np.unique(fam[fam >= 0], return_counts=True)
        best = int(vals[np.argmax(cnt)])
        assert cnt.max() >= 0.8 * (truth == c).sum() and best not in used
        used.add(best)
        est = out[vp_off[0] + best]
        a, b = est / np.linalg.norm(est), vp / np.linalg.norm(vp)
        assert min(np.linalg.norm(a - b), np.linalg.norm(a + b)) < 1.5e-2  # < 1 deg as a homogeneous direction
    # same seed, same answer; thread count does not matter
    lab2, _, out2 = orc.detect_vps(off, allsegs, min_length=40, inlier_threshold=1.0, min_num_supports=10, seed=3,
                                   threads=1)
    assert np.array_equal(lab, lab2) and np.array_equal(out, out2)
