"""limap.triangulation operator surface over the CUDA engine.

GlobalLineTriangulator mirrors the pybind class of src/limap/triangulation/bindings.cc:78-119 (same method
names and argument meaning): Init / InitVPResults / SetRanges / TriangulateImage /
TriangulateImageExhaustiveMatch / ComputeLineTracks / GetTracks / Count* and the debug getters.
TriangulateImage enqueues; the batched kernels run at the first getter or at ComputeLineTracks (results are
only observable through getters in the reference as well). The ten free functions of bindings.cc:19-32 are
small numpy restatements for interactive use; they are not on the hot path.
"""
import numpy as np

from . import base
from .config import TRI_DEFAULTS, make_tri_config
from .engine import TriEngine


class GlobalLineTriangulatorConfig:
    """global_line_triangulator.h:11-25 (+ base_line_triangulator.h:22-43)."""

    def __init__(self, d=None):
        vals = dict(TRI_DEFAULTS)
        vals.update({k: v for k, v in (d or {}).items() if k in vals})
        for k, v in vals.items():
            setattr(self, k, v)
        self.linker2d_config = base.LineLinker2dConfig((d or {}).get("linker2d_config"))
        self.linker3d_config = base.LineLinker3dConfig((d or {}).get("linker3d_config"))

    def as_dict(self):
        d = {k: getattr(self, k) for k in TRI_DEFAULTS}
        d["linker2d_config"] = self.linker2d_config.as_dict()
        d["linker3d_config"] = self.linker3d_config.as_dict()
        return d


def _line3d(rec):
    return base.Line3d(rec[0:3], rec[3:6], 1.0, rec[6], rec[7], rec[8])


class GlobalLineTriangulator:
    def __init__(self, cfg=None, device=0):
        if isinstance(cfg, GlobalLineTriangulatorConfig):
            cfg = cfg.as_dict()
        self._cfg_dict = dict(cfg or {})
        self.config_ = GlobalLineTriangulatorConfig(self._cfg_dict)
        self._eng = TriEngine(make_tri_config(self._cfg_dict), device=device)
        self._lines2d = None
        self._imagecols = None
        self._tracks = []
        self._vpresults = {}
        self._neighbors = {}

    # ---- interfaces (base_line_triangulator.h:52-84) ------------------------------------------
    def Init(self, all_2d_segs, imagecols):
        if not imagecols.IsUndistorted():
            raise RuntimeError("THROW_CHECK_EQ(imagecols->IsUndistorted(), true)")
        ids, model, kvec, qvec, tvec = imagecols.arrays()
        segs, off = [], [0]
        for i in ids:
            lines = all_2d_segs[int(i)]
            arr = getattr(lines, "array", None)
            if arr is None or len(arr) != len(lines):
                arr = (np.array([[l.start[0], l.start[1], l.end[0], l.end[1]] for l in lines], dtype=np.float64)
                       if len(lines) else np.zeros((0, 4)))
            segs.append(arr)
            off.append(off[-1] + len(arr))
        self._segs = np.concatenate(segs, 0) if segs else np.zeros((0, 4))
        self._off = np.asarray(off, np.int64)
        self._ids = ids
        self._view = {int(i): v for v, i in enumerate(ids)}
        self._eng.upload_scene(ids, model, kvec, qvec, tvec, self._off, self._segs)
        self._lines2d = all_2d_segs
        self._imagecols = imagecols  # the reference keeps a raw pointer to the caller's object as well

    def InitVPResults(self, vpresults):
        self._vpresults = dict(vpresults)
        self._eng.set_vps(vpresults, self._ids, self._off)

    def SetRanges(self, ranges):
        self._eng.set_ranges(np.asarray(ranges[0], np.float64), np.asarray(ranges[1], np.float64))

    def UnsetRanges(self):
        self._eng.unset_ranges()

    def TriangulateImage(self, img_id, matches):
        n_lines = self.CountLines(img_id)
        if matches and all(type(m).__module__.startswith("torch") and m.is_cuda for m in matches.values()):
            # device tensors straight from a GPU matcher: out-of-range ids are reported by the engine at run time
            self._neighbors[int(img_id)] = sorted(int(k) for k in matches)
            self._eng.add_image_matches_torch(int(img_id), {int(k): v for k, v in matches.items()})
            return
        for ng, m in matches.items():
            m = np.asarray(m)
            if m.size:
                if m.ndim != 2 or m.shape[1] != 2:
                    raise RuntimeError("THROW_CHECK_EQ(match_info.cols(), 2)")
                if int(m[:, 0].max()) >= n_lines:  # base_line_triangulator.cc:87-94
                    raise RuntimeError(
                        f"IndexError! Out-of-index matches exist between image (img_id = {img_id}) and neighbor "
                        f"image (img_id = {ng}). Please make sure you are reusing the correct descriptors and "
                        "matches when using the --skip_exists option.")
        self._neighbors[int(img_id)] = sorted(int(k) for k in matches)
        self._eng.add_image_matches_dict(int(img_id), {int(k): v for k, v in matches.items()})

    def TriangulateImageExhaustiveMatch(self, img_id, neighbors):
        self._neighbors[int(img_id)] = [int(n) for n in neighbors]
        self._eng.add_image_exhaustive(int(img_id), neighbors)

    def SetBipartites2d(self, all_bpt2ds):
        raise NotImplementedError("point-based proposals (use_pointsfm) are outside the hot path (SURVEY.md §8f-3)")

    def SetSfMPoints(self, points):
        raise NotImplementedError("point-based proposals (use_pointsfm) are outside the hot path (SURVEY.md §8f-3)")

    def ComputeLineTracks(self):
        tr = self._eng.build_tracks()
        tracks = []
        off = tr["track_off"]
        for t in range(len(off) - 1):
            lt = base.LineTrack()
            L = tr["track_line"][t]
            lt.line = base.Line3d(L[0:3], L[3:6])
            lt.line.uncertainty = float(L[6])
            for k in range(off[t], off[t + 1]):
                img, line = int(tr["img_ids"][k]), int(tr["line_ids"][k])
                lt.node_id_list.append(int(tr["node_ids"][k]))
                lt.image_id_list.append(img)
                lt.line_id_list.append(line)
                lt.line2d_list.append(self._lines2d[img][line])
                rec = tr["line3d"][k]
                lt.line3d_list.append(_line3d(rec))
                lt.score_list.append(float(rec[9]))
            tracks.append(lt)
        self._tracks = tracks
        return self.GetTracks()

    def GetTracks(self):
        return list(self._tracks)

    def GetVPResult(self, image_id):
        return self._vpresults[image_id]

    def GetVPResults(self):
        return dict(self._vpresults)

    def CountImages(self):
        return len(self._ids)

    def CountLines(self, img_id):
        v = self._view[int(img_id)]
        return int(self._off[v + 1] - self._off[v])

    def GetLinker(self):
        return base.LineLinker(self.config_.linker2d_config, self.config_.linker3d_config)

    # ---- interface for visualisation (global_line_triangulator.h:43-66) ----------------------
    def _tris(self, img_id, line_id, valid_only=False):
        line, ng = self._eng.get_cands_node(int(img_id), int(line_id))  # needs debug_mode, like the reference
        out = [(_line3d(r), float(r[9]), (int(g[0]), int(g[1]))) for r, g in zip(line, ng)]
        if valid_only:
            # valid_tris_ is filled in std::greater<pair<score, tri_id>> order (global_line_triangulator.cc:124-142)
            order = sorted(range(len(out)), key=lambda i: (-out[i][1], -i))
            out = [out[i] for i in order[: self.config_.max_valid_conns] if out[i][1] >= self.config_.fullscore_th]
        return out

    def CountAllTris(self):
        return int(self._eng.stats()["n_candidates"]) if self.config_.debug_mode else 0

    def GetScoredTrisNode(self, image_id, line_id):
        return self._tris(image_id, line_id)

    def GetValidScoredTrisNode(self, image_id, line_id):
        return self._tris(image_id, line_id, valid_only=True)

    def GetValidScoredTrisNodeSet(self, image_id, line_id):
        best = {}
        for t in self._tris(image_id, line_id, valid_only=True):
            if t[2][0] not in best or t[1] > best[t[2][0]][1]:
                best[t[2][0]] = t
        return [best[k] for k in sorted(best)]

    def CountAllValidTris(self):
        return int(self._eng.stats()["n_valid_edges"]) if self.config_.debug_mode else 0

    def GetValidTrisNode(self, image_id, line_id):
        return [t[0] for t in self.GetValidScoredTrisNode(image_id, line_id)]

    def GetValidTrisNodeSet(self, image_id, line_id):
        return [t[0] for t in self.GetValidScoredTrisNodeSet(image_id, line_id)]

    def GetValidTrisImage(self, image_id):
        return [l for k in range(self.CountLines(image_id)) for l in self.GetValidTrisNode(image_id, k)]

    def GetAllValidTris(self):
        return [l for i in self._ids for l in self.GetValidTrisImage(int(i))]

    def GetBestTrisImage(self, image_id):
        line, _, _ = self._eng.get_best(int(image_id))
        return [_line3d(r) for r in line]

    def GetAllBestTris(self):
        return [l for i in self._ids for l in self.GetBestTrisImage(int(i))]

    def GetAllValidBestTris(self):
        return self.GetAllBestTris()  # min_num_outer_edges filtering only affects the clustering graph

    def GetBestTriNode(self, image_id, line_id):
        return self.GetBestTrisImage(image_id)[line_id]

    def GetBestScoredTriNode(self, image_id, line_id):
        line, ng, _ = self._eng.get_best(int(image_id))
        r = line[line_id]
        return (_line3d(r), float(r[9]), (int(ng[line_id][0]), int(ng[line_id][1])))

    def GetSurvivedLinesImage(self, image_id, n_visible_views):
        out = []
        for t in self._tracks:
            if t.count_images() < n_visible_views:
                continue
            out += [l for i, l in zip(t.image_id_list, t.line_id_list) if i == image_id]
        return out


# ---- free functions (triangulation/functions.cc), numpy restatements --------------------------------
def get_normal_direction(l, view):
    Rt_Kinv = view.R().T @ view.K_inv()
    n = np.cross(Rt_Kinv @ np.append(l.start, 1.0), Rt_Kinv @ np.append(l.end, 1.0))
    return n / np.linalg.norm(n)


def get_direction_from_VP(vp, view):
    d = (view.R().T @ view.K_inv()) @ np.asarray(vp, float)
    return d / np.linalg.norm(d)


def compute_essential_matrix(view1, view2):
    relR = view2.R() @ view1.R().T
    t = view2.T() - relR @ view1.T()
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return tx @ relR


def compute_fundamental_matrix(view1, view2):
    return view2.K_inv().T @ compute_essential_matrix(view1, view2) @ view1.K_inv()


def compute_epipolar_IoU(l1, view1, l2, view2):
    F = compute_fundamental_matrix(view1, view2)
    coor_l2 = l2.coords()

    def hit(p):
        ep = F @ np.append(p, 1.0)
        ep = ep / np.linalg.norm(ep)
        h = np.cross(coor_l2, ep)
        return h[:2] / (h[2] + base.EPS)
    c1 = np.dot(hit(l1.start) - l2.start, l2.direction()) / l2.length()
    c2 = np.dot(hit(l1.end) - l2.start, l2.direction()) / l2.length()
    if c1 > c2:
        c1, c2 = c2, c1
    return (min(c2, 1.0) - max(c1, 0.0)) / (max(c2, 1.0) - min(c1, 0.0))


def triangulate_point(p1, view1, p2, view2):
    C1, C2 = view1.pose.center(), view2.pose.center()
    n1, n2 = view1.ray_direction(p1), view2.ray_direction(p2)
    A = np.array([[n1 @ n1, -n1 @ n2], [-n2 @ n1, n2 @ n2]])
    b = np.array([n1 @ (C2 - C1), n2 @ (C1 - C2)])
    res = np.linalg.solve(A, b)
    point = 0.5 * (n1 * res[0] + C1 + n2 * res[1] + C2)
    if view1.pose.projdepth(point) < base.EPS or view2.pose.projdepth(point) < base.EPS:
        return np.zeros(3), False
    return point, True


def triangulate_line_by_endpoints(l1, view1, l2, view2):
    ps, oks = triangulate_point(l1.start, view1, l2.start, view2)
    pe, oke = triangulate_point(l1.end, view1, l2.end, view2)
    if not (oks and oke):
        return base.Line3d(np.zeros(3), np.ones(3), -1.0)
    return base.Line3d(ps, pe, 1.0, view1.pose.projdepth(ps), view1.pose.projdepth(pe))


def triangulate_line(l1, view1, l2, view2):
    c1s, c1e = view1.ray_direction(l1.start), view1.ray_direction(l1.end)
    c2s, c2e = view2.ray_direction(l2.start), view2.ray_direction(l2.end)
    B = view2.pose.center() - view1.pose.center()
    fail = base.Line3d(np.zeros(3), np.ones(3), -1.0)
    try:
        ls = np.linalg.inv(np.stack([c1s, -c2s, -c2e], 1)) @ B
        le = np.linalg.inv(np.stack([c1e, -c2s, -c2e], 1)) @ B
    except np.linalg.LinAlgError:
        return fail
    Xs, Xe = c1s * ls[0] + view1.pose.center(), c1e * le[0] + view1.pose.center()
    zs, ze = view1.pose.projdepth(Xs), view1.pose.projdepth(Xe)
    if zs < base.EPS or ze < base.EPS or view2.pose.projdepth(Xs) < base.EPS or view2.pose.projdepth(Xe) < base.EPS:
        return fail
    if np.isnan(Xs[0]) or np.isnan(Xe[0]):
        return fail
    return base.Line3d(Xs, Xe, 1.0, zs, ze)


def triangulate_line_with_direction(l1, view1, l2, view2, direction):
    fail = base.Line3d(np.zeros(3), np.ones(3), -1.0)
    direction = np.asarray(direction, float)
    n1 = get_normal_direction(l1, view1)
    direc = direction - n1.dot(direction) * n1
    if np.linalg.norm(direc) < base.EPS:
        return fail
    direc /= np.linalg.norm(direc)
    perp = np.cross(n1, direc)
    v1s, v1e = view1.ray_direction(l1.start), view1.ray_direction(l1.end)
    a1s, a1e = v1s.dot(perp), v1e.dot(perp)
    if a1s < 0:
        a1s, a1e = -a1s, -a1e
    if a1s < 0.001 or a1e < 0.001:
        return fail
    C1, C2 = view1.pose.center(), view2.pose.center()
    n2 = get_normal_direction(l2, view2)
    c1, c2 = n2.dot(v1s), n2.dot(v1e) * a1s / a1e
    d1s = (c1 + c2) * n2.dot(C2 - C1) / (c1 * c1 + c2 * c2)
    d1e = d1s * a1s / a1e
    Xs, Xe = d1s * v1s + C1, d1e * v1e + C1
    zs, ze = view1.pose.projdepth(Xs), view1.pose.projdepth(Xe)
    if zs < base.EPS or ze < base.EPS or view2.pose.projdepth(Xs) < base.EPS or view2.pose.projdepth(Xe) < base.EPS:
        return fail
    if np.isnan(Xs[0]) or np.isnan(Xe[0]):
        return fail
    return base.Line3d(Xs, Xe, 1.0, zs, ze)


def triangulate_line_with_one_point(l1, view1, l2, view2, point):
    raise NotImplementedError("one-point proposals (PoseLib quartic) are outside the hot path (SURVEY.md §8f-3)")
