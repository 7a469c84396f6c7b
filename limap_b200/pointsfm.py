"""limap.pointsfm placeholder: point SfM (COLMAP / pycolmap) is outside the hot path; the
triangulation runner only reaches it when neighbours / ranges are not given or `use_pointsfm` is enabled."""


def _out_of_scope(*a, **k):
    raise NotImplementedError("limap.pointsfm needs COLMAP; pass `neighbors` and `ranges` to line_triangulation "
                              "(SURVEY.md §8f-4)")


check_exists_colmap_model = run_colmap_sfm_with_known_poses = compute_neighbors = compute_ranges = _out_of_scope
