"""limap.optimize operator surface for line refinement / line bundle adjustment over the CUDA LM solver.

Mirrors src/limap/optimize/hybrid_bundle_adjustment/{bindings.cc:17-69, solve.py} and
src/limap/optimize/line_refinement/{bindings.cc:17-79, solve.py, line_refinement.py} for the line-only,
constant-camera case that runners/line_triangulation.py:210-219 and runners/refinement.py use. Free cameras,
point tracks and the heatmap / feature residuals are outside the hot path.
A Ceres-like options object (SolverOptions / LoggingType) is provided because the reference's Python glue
writes to `config.solver_options` (solve.py:6-7).
"""
import enum

import numpy as np

from . import base
from .engine import BAEngine


class LoggingType(enum.IntEnum):  # ceresbase/bindings.cc
    SILENT = 0
    PER_MINIMIZER_ITERATION = 1


class SolverOptions:
    """The fields of ceres::Solver::Options the path honours (refinement_config.h:26-36)."""

    def __init__(self):
        self.function_tolerance = 0.0
        self.gradient_tolerance = 0.0
        self.parameter_tolerance = 0.0
        self.minimizer_progress_to_stdout = True
        self.max_num_iterations = 100
        self.max_linear_solver_iterations = 200
        self.max_num_consecutive_invalid_steps = 10
        self.max_consecutive_nonmonotonic_steps = 10
        self.num_threads = -1
        self.logging_type = LoggingType.SILENT


class RefinementConfig:
    """optimize/line_refinement/refinement_config.h:18-92"""
    _fields = dict(use_geometric=True, min_num_images=4, num_outliers_aggregate=2, print_summary=True,
                   geometric_alpha=10.0, vp_multiplier=1.0)

    def __init__(self, d=None):
        for k, v in self._fields.items():
            setattr(self, k, v)
        for k, v in (d or {}).items():
            if k in self._fields:
                setattr(self, k, v)
        self.solver_options = SolverOptions()
        self.line_geometric_loss_scale = 0.25  # ceres::CauchyLoss(0.25), refinement_config.h:21


class HybridBAConfig(RefinementConfig):
    """optimize/hybrid_bundle_adjustment/hybrid_bundle_adjustment_config.h:17-49"""
    _ba_fields = dict(constant_intrinsics=False, constant_principal_point=True, constant_pose=False,
                      constant_point=False, constant_line=False, lw_point=0.1)

    def __init__(self, d=None):
        super().__init__(d)
        for k, v in self._ba_fields.items():
            setattr(self, k, v)
        for k, v in (d or {}).items():
            if k in self._ba_fields:
                setattr(self, k, v)

    def set_constant_camera(self):
        self.constant_intrinsics = True
        self.constant_pose = True


def _tracks_to_arrays(tracks, view_of):
    sup_off, sup_view, segs, l3d, init = [0], [], [], [], []
    for t in tracks:
        for img, l2 in zip(t.image_id_list, t.line2d_list):
            sup_view.append(view_of[img])
            segs.append([l2.start[0], l2.start[1], l2.end[0], l2.end[1]])
        for l3 in t.line3d_list:
            l3d.append(np.concatenate([l3.start, l3.end]))
        if len(t.line3d_list) != len(t.line2d_list):
            raise RuntimeError("track.line3d_list must hold one 3D line per supporting 2D line")
        sup_off.append(len(sup_view))
        init.append(np.concatenate([t.line.start, t.line.end]))
    return (np.asarray(sup_off, np.int64), np.asarray(sup_view, np.int32),
            np.asarray(segs, np.float64).reshape(-1, 4), np.asarray(l3d, np.float64).reshape(-1, 6),
            np.asarray(init, np.float64).reshape(-1, 6))


def _support_vps(tracks, vpresult_of):
    """[n_support, 3] VP of every supporting 2D line (NaN where the line has none)."""
    out = []
    for t in tracks:
        for img, line_id in zip(t.image_id_list, t.line_id_list):
            r = vpresult_of(img)
            out.append(r.GetVP(line_id) if (r is not None and r.HasVP(line_id)) else [np.nan] * 3)
    return np.asarray(out, np.float64).reshape(-1, 3)


class HybridBAEngine:
    """HybridBAEngine (hybrid_bundle_adjustment.h) for line tracks with constant cameras."""

    def __init__(self, cfg=None, device=0):
        self.config_ = cfg if isinstance(cfg, HybridBAConfig) else HybridBAConfig(cfg)
        self._imagecols = None
        self._tracks = {}
        self._res = None
        self._ba = BAEngine(device=device)

    def InitImagecols(self, imagecols):
        self._imagecols = imagecols

    def InitLineTracks(self, line_tracks):
        self._tracks = dict(line_tracks) if isinstance(line_tracks, dict) else dict(enumerate(line_tracks))

    def InitPointTracks(self, point_tracks):
        if point_tracks:
            raise NotImplementedError("point tracks are outside the hot path (SURVEY.md §8f-3)")

    def SetUp(self):
        c = self.config_
        if not (c.constant_intrinsics and c.constant_pose):
            # yaml default (cfgs/triangulation/default.yaml:138-140) and the runner keep cameras constant
            raise NotImplementedError("free-camera bundle adjustment is outside the hot path (SURVEY.md §8f-3)")

    def Solve(self):
        if not self._tracks:
            return False
        ids, _, kvec, qvec, tvec = self._imagecols.arrays()
        view_of = {int(i): v for v, i in enumerate(ids)}
        keys = list(self._tracks)
        arr = _tracks_to_arrays([self._tracks[k] for k in keys], view_of)
        c = self.config_
        if c.constant_line:
            min_img = 1 << 30
        else:
            min_img = c.min_num_images
        self._res = self._ba.solve(kvec, qvec, tvec, *arr, max_num_iterations=c.solver_options.max_num_iterations,
                                   min_num_images=min_img, num_outliers=c.num_outliers_aggregate,
                                   geometric_alpha=c.geometric_alpha, cauchy_scale=c.line_geometric_loss_scale,
                                   max_num_consecutive_invalid_steps=c.solver_options.max_num_consecutive_invalid_steps)
        self._keys, self._arr = keys, arr
        return True

    def GetOutputLineTracks(self, num_outliers=2):
        out = {}
        lines = self._res["line"]
        if num_outliers != self.config_.num_outliers_aggregate:
            lines = self._resegment(num_outliers)
        for k, key in enumerate(self._keys):
            t = base.LineTrack(self._tracks[key])
            t.line = base.Line3d(lines[k, 0:3], lines[k, 3:6])
            out[key] = t
        return out

    def _resegment(self, num_outliers):
        """GetOutputLineTracks(num_outliers != num_outliers_aggregate): only the segment cut
        (GetLineSegmentFromInfiniteLine3d, hybrid_bundle_adjustment.cc:288-301) depends on it, so the refined infinite
        lines are cut again -- a zero-iteration pass over the stored arrays -- instead of solving the problem twice."""
        ids, _, kvec, qvec, tvec = self._imagecols.arrays()
        sup_off, sup_view, segs, l3d, _ = self._arr
        c = self.config_
        res = self._ba.solve(kvec, qvec, tvec, sup_off, sup_view, segs, l3d, np.ascontiguousarray(self._res["line"]),
                             max_num_iterations=0, min_num_images=1 << 30, num_outliers=num_outliers,
                             geometric_alpha=c.geometric_alpha, cauchy_scale=c.line_geometric_loss_scale)
        return res["line"]

    def GetOutputLines(self, num_outliers=2):
        return {k: t.line for k, t in self.GetOutputLineTracks(num_outliers).items()}

    def GetOutputImagecols(self):
        return self._imagecols

    def summary(self):
        return dict(self._res["stats"]) if self._res else {}


def _init_bundle_adjustment_engine(cfg, imagecols, max_num_iterations=100):  # solve.py:4-11
    ba_config = HybridBAConfig(cfg) if isinstance(cfg, dict) else cfg
    ba_config.solver_options.logging_type = LoggingType.SILENT
    ba_config.solver_options.max_num_iterations = max_num_iterations
    ba_engine = HybridBAEngine(ba_config)
    ba_engine.InitImagecols(imagecols)
    return ba_engine


def solve_line_bundle_adjustment(cfg, imagecols, linetracks, max_num_iterations=100):  # solve.py:31-39
    ba_engine = _init_bundle_adjustment_engine(cfg, imagecols, max_num_iterations=max_num_iterations)
    ba_engine.InitLineTracks(linetracks)
    ba_engine.SetUp()
    ba_engine.Solve()
    return ba_engine


class RefinementEngine:
    """RefinementEngine<DTYPE, CHANNELS> (optimize/line_refinement/refine.h) with geometric (+VP) residuals."""

    def __init__(self, cfg=None, device=0):
        self.config_ = cfg if isinstance(cfg, RefinementConfig) else RefinementConfig(cfg)
        self._track, self._views, self._res, self._vpresults = None, None, None, None
        self._ba = BAEngine(device=device)

    def Initialize(self, track, p_camviews):
        self._track, self._views = track, list(p_camviews)

    def InitializeVPs(self, p_vpresults):
        """p_vpresults[i] = VPResult of the i-th image of track.GetSortedImageIds() (refine.cc:30-35)."""
        self._vpresults = list(p_vpresults)

    def SetUp(self):
        pass

    def Solve(self):
        t = self._track
        sorted_ids = t.GetSortedImageIds()
        view_of = {i: k for k, i in enumerate(sorted_ids)}
        kvec = np.array([v.cam.kvec() for v in self._views])
        qvec = np.array([v.pose.qvec for v in self._views])
        tvec = np.array([v.pose.tvec for v in self._views])
        arr = _tracks_to_arrays([t], view_of)
        c = self.config_
        sup_vp = None
        if self._vpresults is not None:  # AddVPResiduals (refine.cc:86-127)
            sup_vp = _support_vps([t], lambda img: self._vpresults[view_of[img]])
        self._res = self._ba.solve(kvec, qvec, tvec, *arr, max_num_iterations=c.solver_options.max_num_iterations,
                                   min_num_images=0, num_outliers=c.num_outliers_aggregate,
                                   geometric_alpha=c.geometric_alpha, cauchy_scale=c.line_geometric_loss_scale,
                                   sup_vp=sup_vp, vp_multiplier=c.vp_multiplier)
        return True

    def GetLine3d(self):
        L = self._res["line"][0]
        return base.Line3d(L[0:3], L[3:6])


RefinementEngine_f16_c128 = RefinementEngine  # optimize/line_refinement/bindings.cc name used by solve.py:29-30


def solve_line_refinement(cfg, track, p_camviews, p_vpresults=None, p_heatmaps=None, p_patches=None,
                          p_features=None, dtype="float16"):  # line_refinement/solve.py:4-51
    rf_config = RefinementConfig(cfg)
    rf_config.solver_options.logging_type = LoggingType.SILENT
    if track.count_images() < rf_config.min_num_images:
        return None
    if p_heatmaps is not None or p_patches is not None or p_features is not None:
        raise NotImplementedError("pixel-wise residuals need INTERPOLATION_ENABLED (off by default, CMakeLists.txt:25)")
    rf_engine = RefinementEngine(rf_config)
    rf_engine.Initialize(track, p_camviews)
    if p_vpresults is not None:
        rf_engine.InitializeVPs(p_vpresults)
    rf_engine.SetUp()
    rf_engine.Solve()
    return rf_engine


def line_refinement(cfg, tracks, imagecols, heatmap_dir=None, patch_dir=None, featuremap_dir=None, vpresults=None,
                    n_visible_views=4):
    """line_refinement.py:15-147: refine each track (>= n_visible_views images) with fixed cameras. The
    reference loops over tracks with one Ceres problem each; here all selected tracks go to the GPU in one
    batched solve."""
    if cfg.get("use_heatmap") or cfg.get("use_feature"):
        raise NotImplementedError("pixel-wise residuals need INTERPOLATION_ENABLED (off by default)")
    use_vp = bool(cfg.get("use_vp"))
    rf_config = RefinementConfig(cfg)
    ids = [k for k in range(len(tracks)) if tracks[k].count_images() >= n_visible_views]
    sel = [k for k in ids if tracks[k].count_images() >= rf_config.min_num_images]
    newtracks = list(tracks)
    if sel:
        img_ids, _, kvec, qvec, tvec = imagecols.arrays()
        view_of = {int(i): v for v, i in enumerate(img_ids)}
        arr = _tracks_to_arrays([tracks[k] for k in sel], view_of)
        sup_vp = _support_vps([tracks[k] for k in sel], lambda img: vpresults[img]) if use_vp else None
        res = BAEngine().solve(kvec, qvec, tvec, *arr, max_num_iterations=rf_config.solver_options.max_num_iterations,
                               min_num_images=0, num_outliers=rf_config.num_outliers_aggregate,
                               geometric_alpha=rf_config.geometric_alpha,
                               cauchy_scale=rf_config.line_geometric_loss_scale, sup_vp=sup_vp,
                               vp_multiplier=rf_config.vp_multiplier)
        for n, k in enumerate(sel):
            t = base.LineTrack(tracks[k])
            t.line = base.Line3d(res["line"][n, 0:3], res["line"][n, 3:6])
            newtracks[k] = t
    for k in ids:
        if k not in sel:
            newtracks[k] = base.LineTrack(tracks[k])
    return newtracks
