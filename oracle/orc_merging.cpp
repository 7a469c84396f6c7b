// orc_merging.cpp — CPU restatement of the post-triangulation track filters and the iterative remerge.
//
// TEST INFRASTRUCTURE (oracle). PARITY PINNED to the reference's compiled merging_utils.cc / RemergeLineTracks
// (oracle/_ref, tests/test_ref_pinning.py::test_track_filters_and_remerge). This file restates, in plain fp64 C++ and with the reference's loop structure,
//   merging::CheckReprojection / FilterSupportingLines     merging/merging_utils.cc:27-87
//   merging::CheckSensitivity / FilterTracksBySensitivity  merging/merging_utils.cc:89-131
//   merging::FilterTracksByOverlap                         merging/merging_utils.cc:133-155
//   merging::RemergeLineTracks                             merging/merging.cc:513-645
//   Aggregator::aggregate_line3d_list                      merging/aggregator.cc:9-101
// Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it.
#include "orc_geom.h"
#include "orc_triangulation.h"

#include <cstdint>
#include <cstring>
#include <set>
#include <vector>

using namespace orc;

namespace {

// LineLinker3d::check_connection (base/line_linker.cc:212-306)
bool check_connection3d(const LineLinker3d &lk, const Line3d &l1, const Line3d &l2) {
  const LinkerConfig &c = lk.config;
  if (c.use_angle)
    if (!(compute_angle(l1, l2) <= c.th_angle)) return false;
  if (c.use_overlap)
    if (!(lk.score_overlap(l1, l2) == 1.0)) return false;
  if (c.use_angle && c.use_overlap && c.use_smartangle)
    if (!(lk.score_smartangle(l1, l2) >= c.score_th)) return false;
  if (c.use_perp)
    if (!(lk.score_perp(l1, l2) >= c.score_th)) return false;
  if (c.use_innerseg)
    if (!(lk.score_innerseg(l1, l2) >= c.score_th)) return false;
  if (c.use_scaleinv)
    if (!(lk.score_scaleinv(l1, l2) >= c.score_th)) return false;
  return true;
}

std::vector<CameraView> make_views(int n_views, const int32_t *model_ids, const double *kvec, const double *qvec,
                                   const double *tvec) {
  std::vector<CameraView> views(n_views);
  for (int v = 0; v < n_views; ++v) {
    views[v].cam.model_id = model_ids ? model_ids[v] : 1;
    for (int k = 0; k < 4; ++k) views[v].cam.kvec[k] = kvec[4 * v + k];
    views[v].pose.set(qvec + 4 * v, tvec + 3 * v);
  }
  return views;
}

} // namespace

extern "C" {

// Per supporting line s of track t (sup_off[T+1]; sup_view = view index; segs[S][4]; track_line[T][6]):
// bit0 CheckReprojection true, bit1 CheckSensitivity true, bit2 overlap(proj, seg) >= th_overlap.
int orc_track_support_flags(int32_t n_views, const int32_t *model_ids, const double *kvec, const double *qvec,
                            const double *tvec, int64_t T, const int64_t *sup_off, const int32_t *sup_view,
                            const double *segs, const double *track_line, double th_angular2d, double th_perp2d,
                            double th_sv_angular3d, double th_overlap, uint8_t *out_flags) {
  const std::vector<CameraView> views = make_views(n_views, model_ids, kvec, qvec, tvec);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t t = 0; t < T; ++t) {
    const double *tl = track_line + 6 * t;
    Line3d line(V3(tl[0], tl[1], tl[2]), V3(tl[3], tl[4], tl[5]));
    for (int64_t s = sup_off[t]; s < sup_off[t + 1]; ++s) {
      const CameraView &view = views[sup_view[s]];
      Line2d line2d(V2(segs[4 * s], segs[4 * s + 1]), V2(segs[4 * s + 2], segs[4 * s + 3]));
      Line2d proj = line.projection(view);
      uint8_t f = 0;
      { // merging_utils.cc:33-48
        bool ok = true;
        double angle = compute_angle(line2d, proj);
        if (angle > th_angular2d) ok = false;
        if (ok) {
          auto d = dists_endpoints_perpendicular_oneway(line2d, proj);
          if (smax(d.first, d.second) > th_perp2d) ok = false; // std::max(first, second)
        }
        if (ok) f |= 1;
      }
      if (!(line.sensitivity(view) > th_sv_angular3d)) f |= 2; // merging_utils.cc:101-107
      if (compute_overlap(proj, line2d) >= th_overlap) f |= 4;  // merging_utils.cc:147-149
      out_flags[s] = f;
    }
  }
  return 0;
}

// Aggregator::aggregate_line3d_list for T groups: lines[n][7] = start, end, uncertainty; out[T][7].
int orc_aggregate_lines(int64_t T, const int64_t *off, const double *lines, const double *scores, int32_t num_outliers,
                        double *out) {
  for (int64_t t = 0; t < T; ++t) {
    std::vector<Line3d> ls;
    std::vector<double> sc;
    for (int64_t s = off[t]; s < off[t + 1]; ++s) {
      const double *l = lines + 7 * s;
      Line3d x(V3(l[0], l[1], l[2]), V3(l[3], l[4], l[5]));
      x.uncertainty = l[6];
      ls.push_back(x);
      sc.push_back(scores[s]);
    }
    double *o = out + 7 * t;
    if (ls.empty()) { memset(o, 0, 7 * sizeof(double)); continue; }
    Line3d r = aggregate_line3d_list(ls, sc, num_outliers);
    o[0] = r.start.x; o[1] = r.start.y; o[2] = r.start.z;
    o[3] = r.end.x; o[4] = r.end.y; o[5] = r.end.z;
    o[6] = r.uncertainty;
  }
  return 0;
}

// One pass of RemergeLineTracks up to the group labels (merging.cc:515-598): track_line[T][7] (start, end,
// uncertainty), active[T]; out_labels[T]; returns the number of groups; *n_edges = |edges|.
int64_t orc_remerge_labels(int64_t T, const double *track_line, const uint8_t *active, const LinkerConfig *linker_cfg,
                           int32_t *out_labels, int64_t *n_edges) {
  LineLinker3d linker3d;
  linker3d.config = *linker_cfg;
  linker3d.config.set_to_spatial_merging();
  const size_t n_tracks = (size_t)T;
  std::vector<Line3d> lines(n_tracks);
  for (size_t i = 0; i < n_tracks; ++i) {
    const double *l = track_line + 7 * i;
    lines[i] = Line3d(V3(l[0], l[1], l[2]), V3(l[3], l[4], l[5]));
    lines[i].uncertainty = l[6];
  }
  std::set<std::pair<size_t, size_t>> edges;
  std::vector<std::set<std::pair<size_t, size_t>>> edges_per_track(n_tracks);
  std::vector<int> active_ids;
  for (size_t i = 0; i < n_tracks; ++i)
    if (active[i]) active_ids.push_back((int)i);
  const size_t n_active_ids = active_ids.size();
#pragma omp parallel for schedule(dynamic, 16)
  for (size_t k = 0; k < n_active_ids; ++k) {
    const size_t i = (size_t)active_ids[k];
    const Line3d &l1 = lines[i];
    for (size_t j = 0; j < n_tracks; ++j) {
      if (i == j) continue;
      if (n_active_ids == n_tracks) {
        if (i < j && (i + j) % 2 == 0) continue;
        if (i > j && (i + j) % 2 == 1) continue;
      }
      if (!check_connection3d(linker3d, l1, lines[j])) continue;
      if (i < j) edges_per_track[i].insert(std::make_pair(i, j));
      else edges_per_track[i].insert(std::make_pair(j, i));
    }
  }
  for (size_t i = 0; i < n_tracks; ++i) edges.insert(edges_per_track[i].begin(), edges_per_track[i].end());
  if (n_edges) *n_edges = (int64_t)edges.size();
  std::vector<int> parent_tracks(n_tracks, -1);
  std::vector<std::set<int>> tracks_in_group(n_tracks);
  for (size_t i = 0; i < n_tracks; ++i) tracks_in_group[i].insert((int)i);
  for (auto it = edges.begin(); it != edges.end(); ++it) {
    size_t root1 = union_find_get_root(it->first, parent_tracks);
    size_t root2 = union_find_get_root(it->second, parent_tracks);
    if (root1 != root2) {
      if (tracks_in_group[root1].size() < tracks_in_group[root2].size()) {
        parent_tracks[root1] = (int)root2;
        tracks_in_group[root2].insert(tracks_in_group[root1].begin(), tracks_in_group[root1].end());
        tracks_in_group[root1].clear();
      } else {
        parent_tracks[root2] = (int)root1;
        tracks_in_group[root1].insert(tracks_in_group[root2].begin(), tracks_in_group[root2].end());
        tracks_in_group[root2].clear();
      }
    }
  }
  std::vector<int64_t> group_labels(n_tracks, -1);
  int64_t n_groups = 0;
  for (size_t t = 0; t < n_tracks; ++t)
    if (parent_tracks[t] == -1) group_labels[t] = n_groups++;
  for (size_t t = 0; t < n_tracks; ++t) {
    if (group_labels[t] != -1) continue;
    group_labels[t] = group_labels[union_find_get_root(t, parent_tracks)];
  }
  for (size_t t = 0; t < n_tracks; ++t) out_labels[t] = (int32_t)group_labels[t];
  return n_groups;
}

} // extern "C"
