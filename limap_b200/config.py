"""Configuration mirrors: Python dict -> C structs with the reference's C++ defaults.

The reference builds its config objects with ASSIGN_PYDICT_ITEM (src/limap/internal/helpers.h:25-27):
keys that are present overwrite the C++ default, missing keys keep it, unknown keys are ignored.
Defaults: src/limap/triangulation/base_line_triangulator.h:22-43,
src/limap/triangulation/global_line_triangulator.h:11-25, src/limap/base/line_linker.h:18-46,80-143.
Layouts match include/limap_b200.h (lm_linker_config, lm_tri_config).
"""
import ctypes as C


class LinkerConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "score_th", "th_angle", "th_overlap", "th_smartoverlap", "th_smartangle", "th_perp",
        "th_innerseg", "th_scaleinv")] + [(n, C.c_int32) for n in (
            "use_angle", "use_overlap", "use_smartangle", "use_perp", "use_innerseg", "use_scaleinv")]


LINKER2D_DEFAULTS = dict(score_th=0.5, th_angle=8.0, th_overlap=0.1, th_smartoverlap=0.2,
                         th_smartangle=1.0, th_perp=5.0, th_innerseg=5.0, th_scaleinv=0.0,
                         use_angle=True, use_overlap=True, use_smartangle=True, use_perp=True,
                         use_innerseg=False, use_scaleinv=False)
LINKER3D_DEFAULTS = dict(score_th=0.5, th_angle=10.0, th_overlap=0.01, th_smartoverlap=0.1,
                         th_smartangle=1.0, th_perp=0.02, th_innerseg=0.02, th_scaleinv=0.01,
                         use_angle=True, use_overlap=True, use_smartangle=True, use_perp=False,
                         use_innerseg=True, use_scaleinv=False)


def make_linker(defaults, d=None):
    vals = dict(defaults)
    for k, v in (d or {}).items():
        if k in vals:
            vals[k] = v
    out = LinkerConfig()
    for name, ctype in LinkerConfig._fields_:
        setattr(out, name, float(vals[name]) if ctype is C.c_double else int(bool(vals[name])))
    return out


class TriConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "min_length_2d", "line_tri_angle_threshold", "IoU_threshold", "sensitivity_threshold",
        "var2d", "fullscore_th")] + [(n, C.c_int32) for n in (
            "debug_mode", "add_halfpix", "use_vp", "use_endpoints_triangulation",
            "disable_many_points_triangulation", "disable_one_point_triangulation",
            "disable_algebraic_triangulation", "disable_vp_triangulation",
            "max_valid_conns", "min_num_outer_edges", "num_outliers_aggregator",
            "merging_strategy")] + [("linker2d", LinkerConfig), ("linker3d", LinkerConfig)]


TRI_DEFAULTS = dict(min_length_2d=20.0, line_tri_angle_threshold=5.0, IoU_threshold=0.1,
                    sensitivity_threshold=70.0, var2d=2.0, fullscore_th=1.0, debug_mode=False,
                    add_halfpix=False, use_vp=False, use_endpoints_triangulation=False,
                    disable_many_points_triangulation=False, disable_one_point_triangulation=False,
                    disable_algebraic_triangulation=False, disable_vp_triangulation=False,
                    max_valid_conns=1000, min_num_outer_edges=1, num_outliers_aggregator=2,
                    merging_strategy="greedy")
MERGING = {"greedy": 0, "exhaustive": 1, "avg": 2}
_BOOLS = {"debug_mode", "add_halfpix", "use_vp", "use_endpoints_triangulation",
          "disable_many_points_triangulation", "disable_one_point_triangulation",
          "disable_algebraic_triangulation", "disable_vp_triangulation"}


def make_tri_config(d=None):
    """GlobalLineTriangulatorConfig(py::dict) (global_line_triangulator.cc:18-30)."""
    d = d or {}
    vals = dict(TRI_DEFAULTS)
    for k, v in d.items():
        if k in vals:
            vals[k] = v
    out = TriConfig()
    for name, ctype in TriConfig._fields_:
        if name in ("linker2d", "linker3d"):
            continue
        v = vals[name]
        if name == "merging_strategy":
            if v not in MERGING:
                raise RuntimeError("Error!The given merging strategy is not implemented")
            v = MERGING[v]
        elif name in _BOOLS:
            v = int(bool(v))
        setattr(out, name, float(v) if ctype is C.c_double else int(v))
    out.linker2d = make_linker(LINKER2D_DEFAULTS, d.get("linker2d_config"))
    out.linker3d = make_linker(LINKER3D_DEFAULTS, d.get("linker3d_config"))
    return out


# cfgs/triangulation/default.yaml:70-98 (the values a default quickstart run passes in), with var2d
# resolved for the LSD detector (default.yaml:61-66, runners/line_triangulation.py:39-40).
DEFAULT_YAML_TRIANGULATION = dict(
    use_exhaustive_matcher=False, use_endpoints_triangulation=False, add_halfpix=False,
    min_length_2d=0.0, var2d=2.0, line_tri_angle_threshold=1.0, IoU_threshold=0.1,
    sensitivity_threshold=70.0, fullscore_th=1.0, max_valid_conns=1000, min_num_outer_edges=0,
    merging_strategy="greedy", num_outliers_aggregator=2, debug_mode=False,
    linker2d_config=dict(score_th=0.5, th_angle=5.0, th_perp=2.0, th_overlap=0.05),
    linker3d_config=dict(score_th=0.5, th_angle=10.0, th_overlap=0.05, th_smartoverlap=0.1,
                         th_smartangle=2.0, th_perp=1.0, th_innerseg=1.0, th_scaleinv=0.015),
    use_vp=False,
)


def default_runner_config(**over):
    """cfgs/triangulation/default.yaml as the runner reads it (the keys line_triangulation touches), with the
    load_det / load_match switches on: detections and matches are artefacts on disk here, not computed."""
    cfg = dict(
        cfg_type="triangulation", weight_path=None, load_meta=False, load_det=True, load_match=True, load_undistort=False,
        use_tmp=False, n_visible_views=4, n_neighbors=20, use_cuda=True, visualize=False, max_image_dim=1600,
        skip_exists=False, output_dir=None, output_folder="finaltracks", load_dir=None, n_jobs=-1,
        undistortion_output_dir="undistorted_images",
        line2d=dict(max_num_2d_segs=3000, do_merge_lines=False, visualize=False, save_l3dpp=False, compute_descinfo=False,
                    detector=dict(method="lsd", skip_exists=False), extractor=dict(method="superpoint_endpoints", skip_exists=False),
                    matcher=dict(method="nn_endpoints", n_jobs=1, topk=10, skip_exists=False)),
        var2d=dict(sold2=5.0, lsd=2.0, hawpv3=5.0, tp_lsd=5.0, deeplsd=4.0),
        triangulation=dict(DEFAULT_YAML_TRIANGULATION, var2d=-1.0, use_exhaustive_matcher=False, debug_mode=False,
                           remerging=dict(disable=False, linker3d=dict(score_th=0.5, th_angle=5.0, th_overlap=0.001,
                                                                       th_smartoverlap=0.1, th_smartangle=1.0, th_perp=1.0,
                                                                       th_innerseg=1.0)),
                           filtering2d=dict(th_angular_2d=8.0, th_perp_2d=5.0, th_sv_angular_3d=75.0, th_sv_num_supports=3,
                                            th_overlap=0.5, th_overlap_num_supports=3),
                           use_vp=False, vpdet_config=dict(method="jlinkage", n_jobs=8, min_length=40, inlier_threshold=1.0,
                                                           min_num_supports=10),
                           use_pointsfm=dict(enable=False, colmap_folder=None, reuse_sfminfos_colmap=True,
                                             use_triangulated_points=True, use_neighbors=True)),
        refinement=dict(disable=False, constant_intrinsics=True, constant_principal_point=True, constant_pose=True,
                        constant_line=False, min_num_images=4, num_outliers_aggregator=2, use_geometric=True,
                        geometric_alpha=10.0, use_vp=False, vp_multiplier=0.1, use_heatmap=False, use_feature=False),
        structures=dict(bpt2d=dict(threshold_keypoints=2.0, threshold_intersection=2.0, threshold_merge_junctions=2.0)),
    )
    for k, v in over.items():
        cfg[k] = v
    return cfg
