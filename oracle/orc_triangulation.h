// oracle/orc_triangulation.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// PARITY PINNED to the reference's compiled GlobalLineTriangulator (oracle/_ref, tests/test_ref_pinning.py; see orc_geom.h).
//
// fp64 CPU restatement of the triangulation / scoring / track-building path of
// cvg/limap. Loop structure, container choices
// and per-call recomputation of R()/K_inv() follow the reference so that this
// is also the "reference CPU path" for timing (BASELINE.md §3 mode 1).
//   triangulation/functions.cc
//   triangulation/base_line_triangulator.cc
//   triangulation/global_line_triangulator.cc
//   merging/merging.cc:18-103, merging/aggregator.cc, base/graph.cc
#pragma once
#include "orc_geom.h"
#include <queue>
#include <stdexcept>

namespace orc {

// ---------------------------------------------------------------------------
// triangulation/functions.cc
inline bool test_line_inside_ranges(const Line3d &line, const V3 &lo, const V3 &hi) { // :8-26
  for (int k = 0; k < 3; ++k)
    if (line.start[k] < lo[k] || line.start[k] > hi[k]) return false;
  for (int k = 0; k < 3; ++k)
    if (line.end[k] < lo[k] || line.end[k] > hi[k]) return false;
  return true;
}
inline V3 getNormalDirection(const Line2d &l, const CameraView &view) { // :28-35
  const M3 K_inv = view.K_inv();
  const M3 R = view.R();
  V3 c_start = (R.transpose() * K_inv) * V3(l.start.x, l.start.y, 1);
  V3 c_end = (R.transpose() * K_inv) * V3(l.end.x, l.end.y, 1);
  return c_start.cross(c_end).normalized();
}
inline V3 getDirectionFromVP(const V3 &vp, const CameraView &view) { // :37-42
  return ((view.R().transpose() * view.K_inv()) * vp).normalized();
}
inline M3 compute_essential_matrix(const CameraView &view1, const CameraView &view2) { // :44-66
  const M3 R1 = view1.R();
  const V3 T1 = view1.T();
  const M3 R2 = view2.R();
  const V3 T2 = view2.T();
  M3 relR = R2 * R1.transpose();
  V3 relT = T2 - relR * T1;
  M3 tskew;
  tskew.m[0][1] = -relT.z; tskew.m[0][2] = relT.y;
  tskew.m[1][0] = relT.z;  tskew.m[1][2] = -relT.x;
  tskew.m[2][0] = -relT.y; tskew.m[2][1] = relT.x;
  return tskew * relR;
}
inline M3 compute_fundamental_matrix(const CameraView &view1, const CameraView &view2) { // :68-74
  M3 E = compute_essential_matrix(view1, view2);
  return (view2.K_inv().transpose() * E) * view1.K_inv();
}
inline double compute_epipolar_IoU(const Line2d &l1, const CameraView &view1,
                                   const Line2d &l2, const CameraView &view2) { // :76-98
  M3 F = compute_fundamental_matrix(view1, view2);
  V3 coor_l2 = l2.coords();
  V3 coor_epline_start = (F * V3(l1.start.x, l1.start.y, 1)).normalized();
  V2 c_start = dehomogeneous(coor_l2.cross(coor_epline_start));
  V3 coor_epline_end = (F * V3(l1.end.x, l1.end.y, 1)).normalized();
  V2 c_end = dehomogeneous(coor_l2.cross(coor_epline_end));
  double c1 = (c_start - l2.start).dot(l2.direction()) / l2.length();
  double c2 = (c_end - l2.start).dot(l2.direction()) / l2.length();
  if (c1 > c2) std::swap(c1, c2);
  return (smin(c2, 1.0) - smax(c1, 0.0)) / (smax(c2, 1.0) - smin(c1, 0.0));
}
inline std::pair<V3, bool> triangulate_point(const V2 &p1, const CameraView &view1,
                                             const V2 &p2, const CameraView &view2) { // :100-117
  V3 C1 = view1.pose.center();
  V3 C2 = view2.pose.center();
  V3 n1e = view1.ray_direction(p1);
  V3 n2e = view2.ray_direction(p2);
  double a00 = n1e.dot(n1e), a01 = -n1e.dot(n2e), a10 = -n2e.dot(n1e), a11 = n2e.dot(n2e);
  double b0 = n1e.dot(C2 - C1), b1 = n2e.dot(C1 - C2);
  // A.ldlt().solve(b) for a 2x2 SPD system (no pivoting needed: a00 >= a11
  // up to rounding for unit rays; Eigen pivots on the larger diagonal).
  double r0, r1;
  if (a00 >= a11) {
    double l10 = a10 / a00;
    double d1 = a11 - l10 * a01;
    double y1 = b1 - l10 * b0;
    r1 = y1 / d1;
    r0 = (b0 - a01 * r1) / a00;
  } else {
    double l01 = a01 / a11;
    double d0 = a00 - l01 * a10;
    double y0 = b0 - l01 * b1;
    r0 = y0 / d0;
    r1 = (b1 - a10 * r0) / a11;
  }
  V3 point = (n1e * r0 + C1 + n2e * r1 + C2) * 0.5;
  if (view1.pose.projdepth(point) < EPS || view2.pose.projdepth(point) < EPS)
    return std::make_pair(V3(0, 0, 0), false);
  return std::make_pair(point, true);
}
inline Line3d triangulate_line_by_endpoints(const Line2d &l1, const CameraView &view1,
                                            const Line2d &l2, const CameraView &view2) { // :172-190
  auto rs = triangulate_point(l1.start, view1, l2.start, view2);
  if (!rs.second) return Line3d(V3(0, 0, 0), V3(1, 1, 1), -1.0);
  auto re = triangulate_point(l1.end, view1, l2.end, view2);
  if (!re.second) return Line3d(V3(0, 0, 0), V3(1, 1, 1), -1.0);
  double z_start = view1.pose.projdepth(rs.first);
  double z_end = view1.pose.projdepth(re.first);
  return Line3d(rs.first, re.first, 1.0, z_start, z_end);
}
inline std::pair<Line3d, bool> line_triangulation(const Line2d &l1, const CameraView &view1,
                                                  const Line2d &l2, const CameraView &view2) { // :194-233
  V3 c1_start = view1.ray_direction(l1.start);
  V3 c1_end = view1.ray_direction(l1.end);
  V3 c2_start = view2.ray_direction(l2.start);
  V3 c2_end = view2.ray_direction(l2.end);
  V3 B = view2.pose.center() - view1.pose.center();
  M3 A_start = M3::fromCols(c1_start, -c2_start, -c2_end);
  V3 res_start = A_start.inverse() * B;
  V3 l3d_start = c1_start * res_start.x + view1.pose.center();
  double z_start = view1.pose.projdepth(l3d_start);
  M3 A_end = M3::fromCols(c1_end, -c2_start, -c2_end);
  V3 res_end = A_end.inverse() * B;
  V3 l3d_end = c1_end * res_end.x + view1.pose.center();
  double z_end = view1.pose.projdepth(l3d_end);
  if (z_start < EPS || z_end < EPS) return std::make_pair(Line3d(), false);
  double d21 = view2.pose.projdepth(l3d_start);
  double d22 = view2.pose.projdepth(l3d_end);
  if (d21 < EPS || d22 < EPS) return std::make_pair(Line3d(), false);
  if (std::isnan(l3d_start.x) || std::isnan(l3d_end.x)) return std::make_pair(Line3d(), false);
  return std::make_pair(Line3d(l3d_start, l3d_end, 1.0, z_start, z_end), true);
}
inline Line3d triangulate_line(const Line2d &l1, const CameraView &view1,
                               const Line2d &l2, const CameraView &view2) { // :295-303
  auto res = line_triangulation(l1, view1, l2, view2);
  if (!res.second) return Line3d(V3(0, 0, 0), V3(1, 1, 1), -1.0);
  return res.first;
}
inline Line3d triangulate_line_with_direction(const Line2d &l1, const CameraView &view1,
                                              const Line2d &l2, const CameraView &view2,
                                              const V3 &direction) { // :389-446
  const Line3d FAIL(V3(0, 0, 0), V3(1, 1, 1), -1.0);
  V3 n1 = getNormalDirection(l1, view1);
  V3 direc = direction - n1 * (n1.dot(direction));
  if (direc.norm() < EPS) return FAIL;
  direc = direc.normalized();
  V3 perp_direc = n1.cross(direc);
  V3 v1s = view1.ray_direction(l1.start);
  double a1s = v1s.dot(perp_direc);
  V3 v1e = view1.ray_direction(l1.end);
  double a1e = v1e.dot(perp_direc);
  const double MIN_VALUE = 0.001;
  if (a1s < 0) { a1s *= -1; a1e *= -1; }
  if (a1s < MIN_VALUE || a1e < MIN_VALUE) return FAIL;
  V3 C1 = view1.pose.center();
  V3 C2 = view2.pose.center();
  V3 n2 = getNormalDirection(l2, view2);
  double c1s = n2.dot(v1s);
  double c1e = n2.dot(v1e);
  double b = n2.dot(C2 - C1);
  double c1 = c1s;
  double c2 = c1e * a1s / a1e;
  double d1s_num = (c1 + c2) * b;
  double d1s_denom = (c1 * c1 + c2 * c2);
  double d1s = d1s_num / d1s_denom;
  double d1e = d1s * a1s / a1e;
  V3 lstart = v1s * d1s + C1;
  V3 lend = v1e * d1e + C1;
  double z_start = view1.pose.projdepth(lstart);
  double z_end = view1.pose.projdepth(lend);
  if (z_start < EPS || z_end < EPS) return FAIL;
  double d21 = view2.pose.projdepth(lstart);
  double d22 = view2.pose.projdepth(lend);
  if (d21 < EPS || d22 < EPS) return FAIL;
  if (std::isnan(lstart.x) || std::isnan(lend.x)) return FAIL;
  return Line3d(lstart, lend, 1.0, z_start, z_end);
}

// ---------------------------------------------------------------------------
// vplib/vpbase.h:18-47
struct VPResult {
  std::vector<int> labels;
  std::vector<V3> vps;
  bool HasVP(int line_id) const { return line_id < (int)labels.size() && labels[line_id] >= 0; }
  V3 GetVP(int line_id) const { return vps[labels[line_id]]; }
};

// ---------------------------------------------------------------------------
// base_line_triangulator.h:22-43, global_line_triangulator.h:11-25
struct TriConfig {
  bool debug_mode = false;
  bool add_halfpix = false;
  bool use_vp = false;
  bool use_endpoints_triangulation = false;
  bool disable_many_points_triangulation = false;
  bool disable_one_point_triangulation = false;
  bool disable_algebraic_triangulation = false;
  bool disable_vp_triangulation = false;
  double min_length_2d = 20.0;
  double line_tri_angle_threshold = 5.0;
  double IoU_threshold = 0.1;
  double sensitivity_threshold = 70.0;
  double var2d = 2.0;
  double fullscore_th = 1.0;
  int max_valid_conns = 1000;
  int min_num_outer_edges = 1;
  int num_outliers_aggregator = 2;
  LinkerConfig linker2d = default_linker2d();
  LinkerConfig linker3d = default_linker3d();
};

struct TriTuple { // base_line_triangulator.h:17-18
  Line3d line;
  double score = 0; // default-constructed std::tuple value-initialises the double
  int ng_img = 0, ng_line = 0;
};

struct LineTrack { // base/linetrack.h:19-57 (fields used on the path)
  Line3d line;
  std::vector<int> image_id_list, line_id_list, node_id_list;
  std::vector<Line2d> line2d_list;
  std::vector<Line3d> line3d_list;
  std::vector<double> score_list;
};

// base/graph.cc:157-166
inline size_t union_find_get_root(size_t node_idx, std::vector<int> &parent_nodes) {
  if (parent_nodes[node_idx] == -1) return node_idx;
  parent_nodes[node_idx] = (int)union_find_get_root(parent_nodes[node_idx], parent_nodes);
  return parent_nodes[node_idx];
}

// Symmetric 3x3 Jacobi eigen-solver; returns the unit eigenvector of the
// largest eigenvalue. Stands in for JacobiSVD(endpoints).matrixV().col(0)
// (merging/aggregator.cc:76-78): the right singular vector of the largest
// singular value of X equals the dominant eigenvector of X^T X up to sign.
inline V3 dominant_eigenvector_sym3(const double Ain[3][3]) {
  double A[3][3], Vv[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = Ain[i][j];
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off <= 1e-32 * diag || off == 0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0) continue;
        double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) { // A = A * G
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) { // A = G^T * A
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = Vv[k][p], vkq = Vv[k][q];
          Vv[k][p] = c * vkp - s * vkq; Vv[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  if (A[1][1] > A[best][best]) best = 1;
  if (A[2][2] > A[best][best]) best = 2;
  return V3(Vv[0][best], Vv[1][best], Vv[2][best]);
}

// merging/aggregator.cc
inline Line3d aggregate_line3d_list_takebest(const std::vector<Line3d> &lines,
                                             const std::vector<double> &scores) { // :9-29
  int n_lines = (int)lines.size();
  double best_score = 0.0;
  int best_idx = -1;
  double min_uncertainty = std::numeric_limits<double>::max();
  for (int i = 0; i < n_lines; ++i) {
    if (scores[i] > best_score) { best_score = scores[i]; best_idx = i; }
    if (lines[i].uncertainty < min_uncertainty) min_uncertainty = lines[i].uncertainty;
  }
  // The reference indexes lines[-1] when no score is > 0 (undefined
  // behaviour); this restatement takes index 0 there (documented divergence,
  // SURVEY.md §7 "Degenerate defaults").
  if (best_idx < 0) best_idx = 0;
  Line3d best_line = lines[best_idx];
  best_line.uncertainty = min_uncertainty;
  return best_line;
}
inline Line3d aggregate_line3d_list(const std::vector<Line3d> &lines,
                                    const std::vector<double> &scores, int num_outliers) { // :53-101
  int n_lines = (int)lines.size();
  if (n_lines < 4) return aggregate_line3d_list_takebest(lines, scores);
  V3 center(0, 0, 0);
  for (int i = 0; i < n_lines; ++i) { center = center + lines[i].start; center = center + lines[i].end; }
  center = center / (2 * n_lines);
  double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < n_lines; ++i) {
    V3 p[2] = {lines[i].start - center, lines[i].end - center};
    for (int e = 0; e < 2; ++e)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) S[a][b] += p[e][a] * p[e][b];
  }
  V3 direc = dominant_eigenvector_sym3(S);
  direc = direc / direc.norm();
  std::vector<double> projections;
  for (int i = 0; i < n_lines; ++i) {
    projections.push_back((lines[i].start - center).dot(direc));
    projections.push_back((lines[i].end - center).dot(direc));
  }
  std::sort(projections.begin(), projections.end());
  double min_uncertainty = std::numeric_limits<double>::max();
  for (int i = 0; i < n_lines; ++i)
    if (lines[i].uncertainty < min_uncertainty) min_uncertainty = lines[i].uncertainty;
  Line3d final_line;
  final_line.start = center + direc * projections[num_outliers];
  final_line.end = center + direc * projections[n_lines * 2 - 1 - num_outliers];
  final_line.uncertainty = min_uncertainty;
  return final_line;
}

// merging/merging.cc:18-103. Graph = (nodes[(img,line)], edges[(sim,n1,n2)]).
typedef std::tuple<double, size_t, size_t> edge_tuple;
inline std::vector<int> ComputeLineTrackLabelsGreedy(const std::vector<std::pair<int, int>> &nodes,
                                                     const std::vector<edge_tuple> &edges_in) {
  const size_t n_nodes = nodes.size();
  std::vector<edge_tuple> edges = edges_in;
  std::sort(edges.begin(), edges.end());
  std::reverse(edges.begin(), edges.end());
  std::vector<int> parent_nodes(n_nodes, -1);
  std::vector<std::set<int>> images_in_track(n_nodes);
  for (size_t i = 0; i < n_nodes; ++i) images_in_track[i].insert(nodes[i].first);
  for (size_t e = 0; e < edges.size(); ++e) {
    size_t node_idx1 = std::get<1>(edges[e]);
    size_t node_idx2 = std::get<2>(edges[e]);
    size_t root1 = union_find_get_root(node_idx1, parent_nodes);
    size_t root2 = union_find_get_root(node_idx2, parent_nodes);
    if (root1 != root2) {
      if (images_in_track[root1].size() < images_in_track[root2].size()) {
        parent_nodes[root1] = (int)root2;
        images_in_track[root2].insert(images_in_track[root1].begin(), images_in_track[root1].end());
        images_in_track[root1].clear();
      } else {
        parent_nodes[root2] = (int)root1;
        images_in_track[root1].insert(images_in_track[root2].begin(), images_in_track[root2].end());
        images_in_track[root2].clear();
      }
    }
  }
  std::vector<int> track_labels(n_nodes, -1);
  size_t n_tracks = 0;
  for (size_t node_idx = 0; node_idx < n_nodes; ++node_idx) {
    if (parent_nodes[node_idx] == -1) continue;
    size_t parent_idx = parent_nodes[node_idx];
    if (parent_nodes[parent_idx] == -1 && track_labels[parent_idx] == -1)
      track_labels[parent_idx] = (int)n_tracks++;
  }
  for (size_t node_idx = 0; node_idx < n_nodes; ++node_idx) {
    if (parent_nodes[node_idx] == -1) continue;
    track_labels[node_idx] = track_labels[union_find_get_root(node_idx, parent_nodes)];
  }
  return track_labels;
}

// ---------------------------------------------------------------------------
// BaseLineTriangulator + GlobalLineTriangulator
class GlobalLineTriangulator {
public:
  TriConfig config_;
  bool faithful_ = true; // true: per-call camview() copies, R()/K_inv() recomputation as in the reference
  // false: the reference's schedule (OpenMP inside one node: over the connections / candidates of a 2D line).
  // true: the same arithmetic with the OpenMP loop moved out to the 2D lines of the image (nodes are independent), the
  // schedule a CPU implementation tuned for throughput would use; results are identical.
  bool node_parallel_ = false;

  explicit GlobalLineTriangulator(const TriConfig &cfg) : config_(cfg) {
    linker2d_.config = cfg.linker2d;
    linker3d_.config = cfg.linker3d;
  }

  // base_line_triangulator.cc:45-63 + global_line_triangulator.cc:32-57
  void Init(const std::map<int, std::vector<Line2d>> &all_2d_segs,
            const std::map<int, CameraView> &views) {
    all_lines_2d_ = all_2d_segs;
    views_ = views;
    if (config_.add_halfpix) // offsetHalfPixel :32-43
      for (auto &kv : all_lines_2d_)
        for (auto &line : kv.second) {
          line.start = line.start + V2(0.5, 0.5);
          line.end = line.end + V2(0.5, 0.5);
        }
    for (auto &kv : views_) {
      int img_id = kv.first;
      size_t n_lines = all_lines_2d_.at(img_id).size();
      neighbors_[img_id];
      edges_[img_id].assign(n_lines, {});
      tris_[img_id].assign(n_lines, {});
      valid_edges_[img_id].assign(n_lines, {});
      valid_tris_[img_id].assign(n_lines, {});
      tris_best_[img_id].assign(n_lines, TriTuple());
      already_scored_[img_id].assign(n_lines, 0);
    }
  }
  void InitVPResults(const std::map<int, VPResult> &v) { vpresults_ = v; }
  void SetRanges(const V3 &lo, const V3 &hi) { ranges_flag_ = true; ranges_lo_ = lo; ranges_hi_ = hi; }
  void UnsetRanges() { ranges_flag_ = false; }
  size_t CountLines(int img_id) const { return all_lines_2d_.at(img_id).size(); }

  // base_line_triangulator.cc:71-109. matches: ng_img_id -> rows (line_id, ng_line_id).
  void TriangulateImage(int img_id, const std::map<int, std::vector<std::pair<int, int>>> &matches) {
    neighbors_[img_id].clear();
    for (auto it = matches.begin(); it != matches.end(); ++it) {
      int ng_img_id = it->first;
      neighbors_[img_id].push_back(ng_img_id);
      for (const auto &row : it->second) {
        int line_id = row.first, ng_line_id = row.second;
        if (line_id >= (int)edges_[img_id].size())
          throw std::runtime_error("IndexError! Out-of-index matches exist between image (img_id = " +
                                   std::to_string(img_id) + ") and neighbor image (img_id = " +
                                   std::to_string(ng_img_id) + ").");
        edges_[img_id][line_id].push_back(std::make_pair(ng_img_id, ng_line_id));
      }
      const long long n_lines = (long long)CountLines(img_id);
#pragma omp parallel for schedule(dynamic, 8) if (node_parallel_)
      for (long long line_id = 0; line_id < n_lines; ++line_id) {
        triangulateOneNode(img_id, (int)line_id);
        edges_[img_id][line_id].clear();
      }
    }
    ScoringCallback(img_id);
    if (!config_.debug_mode) tris_[img_id].assign(CountLines(img_id), {});
  }
  // base_line_triangulator.cc:111-136
  void TriangulateImageExhaustiveMatch(int img_id, const std::vector<int> &neighbors) {
    neighbors_[img_id] = neighbors;
    for (size_t nb = 0; nb < neighbors.size(); ++nb) {
      int ng_img_id = neighbors[nb];
      int n_lines_ng = (int)all_lines_2d_[ng_img_id].size();
      for (size_t line_id = 0; line_id < CountLines(img_id); ++line_id) {
        for (int ng_line_id = 0; ng_line_id < n_lines_ng; ++ng_line_id)
          edges_[img_id][line_id].push_back(std::make_pair(ng_img_id, ng_line_id));
        triangulateOneNode(img_id, (int)line_id);
        edges_[img_id][line_id].clear();
      }
    }
    ScoringCallback(img_id);
    if (!config_.debug_mode) tris_[img_id].assign(CountLines(img_id), {});
  }

  // global_line_triangulator.cc:353-359
  const std::vector<LineTrack> &ComputeLineTracks() {
    nodes_.clear(); graph_edges_.clear();
    run_clustering();
    build_tracks_from_clusters();
    return tracks_;
  }

  // results
  std::map<int, std::vector<std::vector<TriTuple>>> tris_;        // debug_mode only after scoring
  std::map<int, std::vector<std::vector<TriTuple>>> valid_tris_;  // debug_mode only
  std::map<int, std::vector<TriTuple>> tris_best_;
  std::map<int, std::vector<std::vector<std::pair<int, int>>>> valid_edges_; // (neighbor slot, line)
  std::map<int, std::vector<int>> neighbors_;
  std::map<int, std::vector<int>> n_tris_; // #candidates per node (kept for the metric)
  std::vector<LineTrack> tracks_;
  std::vector<std::pair<int, int>> nodes_;   // graph nodes in creation order
  std::vector<edge_tuple> graph_edges_;      // (sim, node1, node2)
  long long n_match_rows_tested_ = 0;

private:
  // imagecols_->camview(id) returns by value (base/image_collection.cc:366-371).
  CameraView camview(int img_id) const { return views_.at(img_id); }

  // base_line_triangulator.cc:161-337
  void triangulateOneNode(int img_id, int line_id) {
    auto &connections = edges_[img_id][line_id];
    const Line2d &l1 = all_lines_2d_[img_id][line_id];
#pragma omp atomic
    n_match_rows_tested_ += (long long)connections.size();
    if (l1.length() <= config_.min_length_2d) return;
    const CameraView view1 = camview(img_id);
    size_t n_conns = connections.size();
    std::vector<std::vector<TriTuple>> results(n_conns);
#pragma omp parallel for
    for (size_t conn_id = 0; conn_id < n_conns; ++conn_id) {
      int ng_img_id = connections[conn_id].first;
      int ng_line_id = connections[conn_id].second;
      const Line2d &l2 = all_lines_2d_[ng_img_id][ng_line_id];
      if (l2.length() <= config_.min_length_2d) continue;
      const CameraView view2 = camview(ng_img_id);
      auto push = [&](Line3d line) {
        double u1 = line.computeUncertainty(view1, config_.var2d);
        double u2 = line.computeUncertainty(view2, config_.var2d);
        line.uncertainty = smin(u1, u2);
        TriTuple t; t.line = line; t.score = -1.0; t.ng_img = ng_img_id; t.ng_line = ng_line_id;
        results[conn_id].push_back(t);
      };
      // Step 1 (pointsfm proposals) out of scope: use_pointsfm_ is never set here.
      // Step 2: triangulation with VPs (:258-288). NB view1 is used for both VPs.
      if (config_.use_vp && !config_.disable_vp_triangulation) {
        auto it1 = vpresults_.find(img_id);
        if (it1 != vpresults_.end() && it1->second.HasVP(line_id)) {
          V3 direc = getDirectionFromVP(it1->second.GetVP(line_id), view1);
          Line3d line = triangulate_line_with_direction(l1, view1, l2, view2, direc);
          if (line.score > 0) push(line);
        }
        auto it2 = vpresults_.find(ng_img_id);
        if (it2 != vpresults_.end() && it2->second.HasVP(ng_line_id)) {
          V3 direc = getDirectionFromVP(it2->second.GetVP(ng_line_id), view1);
          Line3d line = triangulate_line_with_direction(l1, view1, l2, view2, direc);
          if (line.score > 0) push(line);
        }
      }
      // Step 3: line triangulation (:290-326)
      if (!config_.disable_algebraic_triangulation) {
        V3 n2 = getNormalDirection(l2, view2);
        V3 ray1_start = view1.ray_direction(l1.start);
        double angle_start = 90 - std::acos(std::abs(n2.dot(ray1_start))) * 180.0 / M_PI;
        if (angle_start < config_.line_tri_angle_threshold) continue;
        V3 ray1_end = view1.ray_direction(l1.end);
        double angle_end = 90 - std::acos(std::abs(n2.dot(ray1_end))) * 180.0 / M_PI;
        if (angle_end < config_.line_tri_angle_threshold) continue;
        double IoU = compute_epipolar_IoU(l1, view1, l2, view2);
        if (IoU < config_.IoU_threshold) continue;
        Line3d line;
        if (!config_.use_endpoints_triangulation)
          line = triangulate_line(l1, view1, l2, view2);
        else
          line = triangulate_line_by_endpoints(l1, view1, l2, view2);
        if (line.sensitivity(view1) > config_.sensitivity_threshold &&
            line.sensitivity(view2) > config_.sensitivity_threshold)
          line.score = -1;
        if (line.score > 0) push(line);
      }
    }
    for (size_t conn_id = 0; conn_id < n_conns; ++conn_id)
      for (auto &t : results[conn_id]) {
        if (ranges_flag_ && !test_line_inside_ranges(t.line, ranges_lo_, ranges_hi_)) continue;
        tris_[img_id][line_id].push_back(t);
      }
  }

  // global_line_triangulator.cc:59-69
  void ScoringCallback(int img_id) {
    LineLinker3d linker3d_scoring = linker3d_;
    linker3d_scoring.config.set_to_shared_parent_scoring();
    n_tris_[img_id].assign(CountLines(img_id), 0);
    const long long n_lines = (long long)CountLines(img_id);
#pragma omp parallel for schedule(dynamic, 8) if (node_parallel_)
    for (long long line_id = 0; line_id < n_lines; ++line_id)
      scoreOneNode(img_id, (int)line_id, linker2d_, linker3d_scoring);
    if (!config_.debug_mode) valid_tris_[img_id].assign(CountLines(img_id), {});
  }

  // global_line_triangulator.cc:71-161
  void scoreOneNode(int img_id, int line_id, const LineLinker2d &linker2d, const LineLinker3d &linker3d) {
    if (already_scored_[img_id][line_id]) return;
    auto &tris = tris_[img_id][line_id];
    size_t n_tris = tris.size();
    n_tris_[img_id][line_id] = (int)n_tris;
    std::vector<double> scores(n_tris, 0);
#pragma omp parallel for
    for (size_t i = 0; i < n_tris; ++i) {
      std::map<int, std::vector<double>> score_table;
      const Line3d &l1 = tris[i].line;
      int img_id_i = tris[i].ng_img;
      const CameraView view1 = camview(img_id_i); // copied but unused, as in the reference (:89)
      (void)view1;
      for (size_t j = 0; j < n_tris; ++j) {
        if (i == j) continue;
        const Line3d &l2 = tris[j].line;
        int ng_img_id = tris[j].ng_img;
        int ng_line_id = tris[j].ng_line;
        if (ng_img_id == img_id_i) continue;
        const CameraView view2 = camview(ng_img_id);
        double score3d = linker3d.compute_score(l1, l2);
        if (score3d == 0) continue;
        double score2d = linker2d.compute_score(l1.projection(view2), all_lines_2d_[ng_img_id][ng_line_id]);
        if (score2d == 0) continue;
        double score = smin(score3d, score2d);
        score_table[ng_img_id].push_back(score);
      }
      for (auto it = score_table.begin(); it != score_table.end(); ++it)
        scores[i] += *std::max_element(it->second.begin(), it->second.end());
    }
    for (size_t i = 0; i < n_tris; ++i) tris[i].score = scores[i];

    std::map<int, int> reverse_mapper;
    int n_neighbors = (int)neighbors_[img_id].size();
    for (int i = 0; i < n_neighbors; ++i) reverse_mapper.insert(std::make_pair(neighbors_[img_id][i], i));
    std::vector<std::pair<double, int>> scores_to_sort;
    for (size_t tri_id = 0; tri_id < tris.size(); ++tri_id)
      scores_to_sort.push_back(std::make_pair(tris[tri_id].score, (int)tri_id));
    std::sort(scores_to_sort.begin(), scores_to_sort.end(), std::greater<std::pair<double, int>>());
    int n_valid_conns = std::min(int(scores_to_sort.size()), config_.max_valid_conns);
    for (int i = 0; i < n_valid_conns; ++i) {
      int tri_id = scores_to_sort[i].second;
      auto &tri = tris[tri_id];
      if (tri.score < config_.fullscore_th) continue;
      valid_tris_[img_id][line_id].push_back(tri);
      valid_edges_[img_id][line_id].push_back(std::make_pair(reverse_mapper.at(tri.ng_img), tri.ng_line));
    }
    double max_score = -1;
    for (size_t tri_id = 0; tri_id < n_tris; ++tri_id) {
      if (tris[tri_id].score > max_score) {
        tris_best_[img_id][line_id] = tris[tri_id];
        max_score = tris[tri_id].score;
      }
    }
    if (!config_.debug_mode) {
      tris_[img_id][line_id].clear();
      valid_tris_[img_id][line_id].clear();
    }
    already_scored_[img_id][line_id] = 1;
  }

  // global_line_triangulator.cc:168-232
  void filterNodeByNumOuterEdges(std::map<int, std::vector<bool>> &flags) {
    std::map<int, std::vector<std::vector<std::pair<int, int>>>> parent_neighbors;
    std::map<int, std::vector<int>> counters;
    for (auto &kv : views_) {
      int img_id = kv.first;
      size_t n_lines = CountLines(img_id);
      flags[img_id].assign(n_lines, true);
      parent_neighbors[img_id].assign(n_lines, {});
      counters[img_id].assign(n_lines, 0);
      for (size_t l = 0; l < n_lines; ++l) counters[img_id][l] = (int)valid_edges_.at(img_id)[l].size();
    }
    for (auto &kv : views_) {
      int img_id = kv.first;
      for (size_t line_id = 0; line_id < CountLines(img_id); ++line_id) {
        for (auto &e : valid_edges_.at(img_id)[line_id]) {
          int ng_img_id = neighbors_[img_id][e.first];
          parent_neighbors[ng_img_id][e.second].push_back(std::make_pair(img_id, (int)line_id));
        }
        if (counters[img_id][line_id] < config_.min_num_outer_edges) flags[img_id][line_id] = false;
      }
    }
    std::queue<std::pair<int, int>> q;
    for (auto &kv : views_)
      for (size_t l = 0; l < CountLines(kv.first); ++l)
        if (!flags[kv.first][l]) q.push(std::make_pair(kv.first, (int)l));
    while (!q.empty()) {
      auto node = q.front(); q.pop();
      for (auto &p : parent_neighbors[node.first][node.second]) {
        if (!flags[p.first][p.second]) continue;
        counters[p.first][p.second]--;
        if (counters[p.first][p.second] < config_.min_num_outer_edges) {
          flags[p.first][p.second] = false;
          q.push(p);
        }
      }
    }
  }

  // global_line_triangulator.cc:234-291
  void run_clustering() {
    LineLinker3d linker3d_clustering = linker3d_;
    linker3d_clustering.config.set_to_spatial_merging();
    valid_flags_.clear();
    filterNodeByNumOuterEdges(valid_flags_);
    typedef std::pair<int, int> LineNode;
    std::set<std::pair<LineNode, LineNode>> edges;
    for (auto &kv : views_) {
      int img_id = kv.first;
      for (size_t line_id = 0; line_id < CountLines(img_id); ++line_id) {
        for (auto &e : valid_edges_[img_id][line_id]) {
          LineNode node1 = std::make_pair(img_id, (int)line_id);
          if (!valid_flags_[node1.first][node1.second]) continue;
          LineNode node2 = std::make_pair(neighbors_[img_id][e.first], e.second);
          if (!valid_flags_[node2.first][node2.second]) continue;
          if (node1.first > node2.first || (node1.first == node2.first && node1.second > node2.second))
            std::swap(node1, node2);
          edges.insert(std::make_pair(node1, node2));
        }
      }
    }
    std::map<LineNode, size_t> node_map;
    for (auto it = edges.begin(); it != edges.end(); ++it) {
      const Line3d &line1 = tris_best_.at(it->first.first)[it->first.second].line;
      const Line3d &line2 = tris_best_.at(it->second.first)[it->second.second].line;
      // The reference also evaluates both 2d scores and then discards them
      // (score = score_3d, :277-283); only the 3d score decides.
      double score = linker3d_clustering.compute_score(line1, line2);
      if (score == 0) continue;
      size_t idx[2];
      const LineNode *ns[2] = {&it->first, &it->second};
      for (int k = 0; k < 2; ++k) {
        auto f = node_map.find(*ns[k]);
        if (f == node_map.end()) {
          nodes_.push_back(*ns[k]);
          idx[k] = nodes_.size() - 1;
          node_map.insert(std::make_pair(*ns[k], idx[k]));
        } else
          idx[k] = f->second;
      }
      graph_edges_.push_back(std::make_tuple(score, idx[0], idx[1]));
    }
  }

  // global_line_triangulator.cc:293-351
  void build_tracks_from_clusters() {
    tracks_.clear();
    std::vector<int> track_labels = ComputeLineTrackLabelsGreedy(nodes_, graph_edges_);
    if (track_labels.empty()) return;
    int n_tracks = *std::max_element(track_labels.begin(), track_labels.end()) + 1;
    tracks_.resize(n_tracks);
    for (size_t node_id = 0; node_id < nodes_.size(); ++node_id) {
      int img_id = nodes_[node_id].first, line_id = nodes_[node_id].second;
      int track_id = track_labels[node_id];
      if (track_id == -1) continue;
      const TriTuple &best = tris_best_.at(img_id)[line_id];
      tracks_[track_id].node_id_list.push_back((int)node_id);
      tracks_[track_id].image_id_list.push_back(img_id);
      tracks_[track_id].line_id_list.push_back(line_id);
      tracks_[track_id].line2d_list.push_back(all_lines_2d_[img_id][line_id]);
      tracks_[track_id].line3d_list.push_back(best.line);
      tracks_[track_id].score_list.push_back(best.score);
    }
    for (auto &t : tracks_)
      t.line = aggregate_line3d_list(t.line3d_list, t.score_list, config_.num_outliers_aggregator);
  }

  std::map<int, std::vector<Line2d>> all_lines_2d_;
  std::map<int, CameraView> views_;
  std::map<int, VPResult> vpresults_;
  std::map<int, std::vector<std::vector<std::pair<int, int>>>> edges_;
  std::map<int, std::vector<char>> already_scored_; // bytes, not bits: lines are scored concurrently in node-parallel mode
  std::map<int, std::vector<bool>> valid_flags_;
  bool ranges_flag_ = false;
  V3 ranges_lo_, ranges_hi_;
  LineLinker2d linker2d_;
  LineLinker3d linker3d_;
};

} // namespace orc
