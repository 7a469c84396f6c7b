"""limap.runners as far as the triangulation path needs it: `setup`, the LOAD branches of `compute_2d_segs` /
`compute_matches` (detections and matches come from disk or from a GPU matcher; the learned front end is out of scope)
and `line_triangulation` -- the steps of src/limap/runners/line_triangulation.py:18-245 ([A] metainfos, [B] segments,
[C] matches, [D] triangulation + track filters + remerge, [E] line bundle adjustment, [F] output + track report) driven
through this package's operator surface. tests/test_runner_dropin.py additionally executes the reference's own runner
file, unmodified and loaded by path, against the same surface."""
import logging
import os

import numpy as np

from . import base, merging, optimize, triangulation, visualize, vplib
from .util import io as limapio

_log = logging.getLogger("limap_b200")


def setup(cfg):  # runners/functions.py:11-28
    folder_save = cfg["output_dir"] or "tmp"
    limapio.check_makedirs(folder_save)
    folder_load = "tmp" if cfg.get("use_tmp") else cfg.get("load_dir")
    cfg["dir_save"] = folder_save
    cfg["dir_load"] = folder_load if folder_load is not None else folder_save
    return cfg


def _out_of_scope(what):
    def f(*a, **k):
        raise NotImplementedError(f"{what} is outside the hot path (SURVEY.md §2); provide the artefacts and set "
                                  "load_det / load_match, or pass neighbors and ranges")
    return f


undistort_images = _out_of_scope("image undistortion")
compute_sfminfos = _out_of_scope("COLMAP neighbours / ranges")
compute_2d_bipartites_from_colmap = _out_of_scope("point-line bipartites from a COLMAP model")
compute_exhaustive_matches = _out_of_scope("descriptor matching")


def segments_folder(cfg, root):
    """<root>/line_detections/<detector>/segments (runners/functions.py:232-235, line2d/base_detector.py:155-164)."""
    return os.path.join(root, "line_detections", cfg["line2d"]["detector"]["method"], "segments")


def matches_folder(cfg, root):
    """<root>/line_matchings/<detector>/feats_<extractor>/<matcher>_n<neighbours>_top<k>
    (runners/functions.py:316-320, line2d/base_matcher.py:60-72)."""
    m = cfg["line2d"]["matcher"]
    return os.path.join(root, "line_matchings", cfg["line2d"]["detector"]["method"],
                        "feats_{}".format(cfg["line2d"]["extractor"]["method"]),
                        "{}_n{}_top{}".format(m["method"], cfg["n_neighbors"], m.get("topk", 10)))


def compute_2d_segs(cfg, imagecols, compute_descinfo=True):  # runners/functions.py:197-290, load branch
    if not cfg["load_det"]:
        return _out_of_scope("2D line detection")()
    segs = limapio.read_all_segments_from_folder(segments_folder(cfg, cfg["dir_load"]))
    return {i: segs[i] for i in imagecols.get_img_ids()}, None


def compute_matches(cfg, descinfo_folder, image_ids, neighbors):  # runners/functions.py:293-345, load branch
    if not cfg["load_match"]:
        return _out_of_scope("2D line matching")()
    return matches_folder(cfg, cfg["dir_load"])


def line_triangulation(cfg, imagecols, neighbors=None, ranges=None):
    """Main interface of line triangulation over multi-view images (runners/line_triangulation.py:18-245)."""
    cfg = setup(cfg)
    tcfg = cfg["triangulation"]
    if tcfg["var2d"] == -1:
        tcfg["var2d"] = cfg["var2d"][cfg["line2d"]["detector"]["method"]]
    if not imagecols.IsUndistorted():
        imagecols = undistort_images(imagecols)
    if cfg.get("max_image_dim") not in (-1, None):
        imagecols.set_max_image_dim(cfg["max_image_dim"])
    save = cfg["dir_save"]
    limapio.save_txt_imname_dict(os.path.join(save, "image_list.txt"), imagecols.get_image_name_dict())
    limapio.save_npy(os.path.join(save, "imagecols.npy"), imagecols.as_dict())
    # [A] neighbours and ranges
    if neighbors is None:
        _, neighbors, ranges = compute_sfminfos(cfg, imagecols)
    else:
        neighbors = imagecols.update_neighbors(neighbors)
        neighbors = {i: list(n)[: cfg["n_neighbors"]] for i, n in neighbors.items()}
    limapio.save_txt_metainfos(os.path.join(save, "metainfos.txt"), neighbors, ranges)
    # [B] segments, [C] matches
    all_2d_segs, descinfo_folder = compute_2d_segs(cfg, imagecols, compute_descinfo=False)
    exhaustive = tcfg["use_exhaustive_matcher"]
    matches_dir = None if exhaustive else compute_matches(cfg, descinfo_folder, imagecols.get_img_ids(), neighbors)
    # [D] triangulation
    tri = triangulation.GlobalLineTriangulator(tcfg)
    tri.SetRanges(ranges)
    all_2d_lines = base.get_all_lines_2d(all_2d_segs)
    tri.Init(all_2d_lines, imagecols)
    if tcfg["use_vp"]:
        det = vplib.get_vp_detector(tcfg["vpdet_config"], n_jobs=tcfg["vpdet_config"]["n_jobs"])
        tri.InitVPResults(det.detect_vp_all_images(all_2d_lines, imagecols.get_map_camviews()))
    if tcfg["use_pointsfm"]["enable"]:
        compute_2d_bipartites_from_colmap()
    for img_id in imagecols.get_img_ids():
        if exhaustive:
            tri.TriangulateImageExhaustiveMatch(img_id, neighbors[img_id])
        else:
            tri.TriangulateImage(img_id, limapio.read_npy(os.path.join(matches_dir, f"matches_{img_id}.npy")).item())
    linetracks = tri.ComputeLineTracks()
    f2d = tcfg["filtering2d"]
    linetracks = merging.filter_tracks_by_reprojection(linetracks, imagecols, f2d["th_angular_2d"], f2d["th_perp_2d"])
    if not tcfg["remerging"]["disable"]:
        linetracks = merging.remerge(base.LineLinker3d(tcfg["remerging"]["linker3d"]), linetracks)
        linetracks = merging.filter_tracks_by_reprojection(linetracks, imagecols, f2d["th_angular_2d"], f2d["th_perp_2d"])
    linetracks = merging.filter_tracks_by_sensitivity(linetracks, imagecols, f2d["th_sv_angular_3d"], f2d["th_sv_num_supports"])
    linetracks = merging.filter_tracks_by_overlap(linetracks, imagecols, f2d["th_overlap"], f2d["th_overlap_num_supports"])
    # [E] line bundle adjustment (cameras constant)
    if not cfg["refinement"]["disable"]:
        ba = optimize.solve_line_bundle_adjustment(cfg["refinement"], imagecols, linetracks, max_num_iterations=200)
        tracks_map = ba.GetOutputLineTracks(num_outliers=cfg["refinement"]["num_outliers_aggregator"])
        linetracks = [t for _, t in tracks_map.items()]
    # [F] output + track report
    limapio.save_txt_linetracks(os.path.join(save, "alltracks.txt"), linetracks, n_visible_views=4)
    limapio.save_folder_linetracks_with_info(os.path.join(save, cfg["output_folder"]), linetracks, config=cfg,
                                             imagecols=imagecols, all_2d_segs=all_2d_segs)
    vis = visualize.Open3DTrackVisualizer(linetracks)
    vis.report()
    limapio.save_obj(os.path.join(save, "triangulated_lines_nv{}.obj".format(cfg["n_visible_views"])),
                     vis.get_lines_np(n_visible_views=cfg["n_visible_views"]))
    return linetracks
