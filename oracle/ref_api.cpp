// oracle/ref_api.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points (ctypes) over the REFERENCE'S OWN hot-path sources, compiled unchanged from /root/reference by
// oracle/Makefile (target _ref) against the header shims in oracle/ref_shim/ (Eigen subset, COLMAP camera struct,
// glog macros, PoseLib quartic: none of them is in this image). The entry points mirror limap_oracle.cpp's
// (prefix ref_ instead of orc_) so that tests pin the restatement to the reference's compiled code:
//   GlobalLineTriangulator::{Init, TriangulateImage(ExhaustiveMatch), ComputeLineTracks} and its result tables,
//   triangulation/functions.cc, LineLinker2d/3d::compute_score, Aggregator::aggregate_line3d_list,
//   MinimalInfiniteLine3d, GetLineSegmentFromInfiniteLine3d, CheckReprojection / CheckSensitivity / overlap,
//   RemergeLineTracks,
//   optimize/line_refinement/cost_functions.h: GeometricRefinementFunctor / VPConstraintsFunctor evaluated on
//   ceres::Jet<double, 6> (oracle/ref_shim/ceres: Jet arithmetic restated; the functors are the reference's),
//   vplib/JLinkage/JLinkage.cc + vplib/base_vp_detector.cc: limap's J-Linkage wrapper over oracle/ref_shim/JLinkage (the
//   library's sampling / clustering forward to the restatement in orc_vp.h),
//   pointsfm/sfm_model.cc: SfmModel::{GetMaxIoUImages, GetMaxDiceCoeffImages, GetMaxOverlapImages, ComputeRanges} on top
//   of oracle/ref_shim/colmap/mvs (COLMAP's shared-point / triangulation-angle statistics restated; the ranking loops,
//   score formulas, sorts and the float range arithmetic are the reference's).
// Nothing of the product links or loads this library.
// every standard / third-party header first, with its own access specifiers intact ...
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <pybind11/eigen.h>
#include <pybind11/embed.h>
#include <pybind11/eval.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
#include <Eigen/Dense>
#include <colmap/scene/camera.h>
#include <colmap/util/logging.h>
#include <third-party/half.h>
// ... then the reference's class definitions with their result tables readable (layout is unaffected)
#define private public
#define protected public
#include "limap/base/infinite_line.h"
#include "limap/base/line_linker.h"
#include "limap/merging/aggregator.h"
#include "limap/merging/merging.h"
#include "limap/merging/merging_utils.h"
#include "limap/optimize/line_refinement/cost_functions.h"
#include "limap/pointsfm/sfm_model.h"
#include "limap/vplib/JLinkage/JLinkage.h"
#include "limap/triangulation/functions.h"
#include "limap/triangulation/global_line_triangulator.h"
#undef private
#undef protected
#include "orc_vp.h" // the J-Linkage core the JLinkage-library shim forwards to (namespace orc)
#include <cstring>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace limap;
using limap::triangulation::GlobalLineTriangulator;
using limap::triangulation::GlobalLineTriangulatorConfig;
using limap::triangulation::TriTuple;

extern "C" {

struct ref_linker_cfg {
  double score_th, th_angle, th_overlap, th_smartoverlap, th_smartangle, th_perp, th_innerseg, th_scaleinv;
  int32_t use_angle, use_overlap, use_smartangle, use_perp, use_innerseg, use_scaleinv;
};
struct ref_tri_cfg {
  double min_length_2d, line_tri_angle_threshold, IoU_threshold, sensitivity_threshold, var2d, fullscore_th;
  int32_t debug_mode, add_halfpix, use_vp, use_endpoints_triangulation;
  int32_t disable_many_points_triangulation, disable_one_point_triangulation;
  int32_t disable_algebraic_triangulation, disable_vp_triangulation;
  int32_t max_valid_conns, min_num_outer_edges, num_outliers_aggregator, merging_strategy;
  ref_linker_cfg linker2d, linker3d;
};

} // extern "C"
template <typename L> static void fill_linker(L &l, const ref_linker_cfg &c) {
  l.score_th = c.score_th; l.th_angle = c.th_angle; l.th_overlap = c.th_overlap; l.th_smartoverlap = c.th_smartoverlap;
  l.th_smartangle = c.th_smartangle; l.th_perp = c.th_perp; l.th_innerseg = c.th_innerseg;
  l.use_angle = c.use_angle; l.use_overlap = c.use_overlap; l.use_smartangle = c.use_smartangle; l.use_perp = c.use_perp;
  l.use_innerseg = c.use_innerseg;
}
extern "C" {
static LineLinker2dConfig to_linker2d(const ref_linker_cfg &c) { LineLinker2dConfig l; fill_linker(l, c); return l; }
static LineLinker3dConfig to_linker3d(const ref_linker_cfg &c) {
  LineLinker3dConfig l;
  fill_linker(l, c);
  l.th_scaleinv = c.th_scaleinv;
  l.use_scaleinv = c.use_scaleinv;
  return l;
}

struct RefTri {
  std::unique_ptr<GlobalLineTriangulator> tri;
  std::map<int, std::vector<Line2d>> segs;
  std::unique_ptr<ImageCollection> imagecols;
  long long rows = 0;
};

static thread_local std::string g_err;
const char *ref_last_error() { return g_err.c_str(); }
void ref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void *ref_tri_create(const ref_tri_cfg *c) {
  GlobalLineTriangulatorConfig cfg;
  // debug_mode only decides whether the per-node candidate lists survive scoring (global_line_triangulator.cc:155-159);
  // they are kept so that candidate counts and lists can be read back
  cfg.debug_mode = true;
  cfg.add_halfpix = c->add_halfpix; cfg.use_vp = c->use_vp;
  cfg.use_endpoints_triangulation = c->use_endpoints_triangulation;
  cfg.disable_many_points_triangulation = c->disable_many_points_triangulation;
  cfg.disable_one_point_triangulation = c->disable_one_point_triangulation;
  cfg.disable_algebraic_triangulation = c->disable_algebraic_triangulation;
  cfg.disable_vp_triangulation = c->disable_vp_triangulation;
  cfg.min_length_2d = c->min_length_2d; cfg.line_tri_angle_threshold = c->line_tri_angle_threshold;
  cfg.IoU_threshold = c->IoU_threshold; cfg.sensitivity_threshold = c->sensitivity_threshold;
  cfg.var2d = c->var2d; cfg.fullscore_th = c->fullscore_th;
  cfg.max_valid_conns = c->max_valid_conns; cfg.min_num_outer_edges = c->min_num_outer_edges;
  cfg.num_outliers_aggregator = c->num_outliers_aggregator;
  cfg.merging_strategy = (c->merging_strategy == 0) ? "greedy" : "avg";
  cfg.linker2d_config = to_linker2d(c->linker2d);
  cfg.linker3d_config = to_linker3d(c->linker3d);
  RefTri *h = new RefTri();
  h->tri.reset(new GlobalLineTriangulator(cfg));
  return h;
}
void ref_tri_destroy(void *hp) { delete (RefTri *)hp; }

int ref_tri_init(void *hp, int n_views, const int32_t *img_ids, const int32_t *model_ids, const double *kvec,
                 const double *qvec, const double *tvec, const int64_t *line_off, const double *segs) {
  RefTri *h = (RefTri *)hp;
  try {
    std::map<int, Camera> cams;
    std::map<int, CameraImage> imgs;
    for (int v = 0; v < n_views; ++v) {
      const double *k = kvec + 4 * v;
      std::vector<double> params;
      if (model_ids[v] == 0) params = {k[0], k[2], k[3]};
      else params = {k[0], k[1], k[2], k[3]};
      cams[v] = Camera(model_ids[v], params, v);
      CameraPose pose(V4D(qvec[4 * v], qvec[4 * v + 1], qvec[4 * v + 2], qvec[4 * v + 3]),
                      V3D(tvec[3 * v], tvec[3 * v + 1], tvec[3 * v + 2]));
      imgs[img_ids[v]] = CameraImage(v, pose);
      std::vector<Line2d> lines;
      for (int64_t l = line_off[v]; l < line_off[v + 1]; ++l)
        lines.push_back(Line2d(V2D(segs[4 * l], segs[4 * l + 1]), V2D(segs[4 * l + 2], segs[4 * l + 3])));
      h->segs[img_ids[v]] = lines;
    }
    h->imagecols.reset(new ImageCollection(cams, imgs));
    h->tri->Init(h->segs, *h->imagecols);
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
void ref_tri_set_ranges(void *hp, const double *lo, const double *hi) {
  ((RefTri *)hp)->tri->SetRanges(std::make_pair(V3D(lo[0], lo[1], lo[2]), V3D(hi[0], hi[1], hi[2])));
}
void ref_tri_unset_ranges(void *hp) { ((RefTri *)hp)->tri->UnsetRanges(); }

int ref_tri_set_vps(void *hp, int n_images, const int32_t *img_ids, const int64_t *label_off, const int32_t *labels,
                    const int64_t *vp_off, const double *vps) {
  std::map<int, vplib::VPResult> res;
  for (int i = 0; i < n_images; ++i) {
    vplib::VPResult r;
    r.labels.assign(labels + label_off[i], labels + label_off[i + 1]);
    for (int64_t k = vp_off[i]; k < vp_off[i + 1]; ++k) r.vps.push_back(V3D(vps[3 * k], vps[3 * k + 1], vps[3 * k + 2]));
    res[img_ids[i]] = r;
  }
  ((RefTri *)hp)->tri->InitVPResults(res);
  return 0;
}

int ref_tri_triangulate_image(void *hp, int img_id, int n_ng, const int32_t *ng_ids, const int64_t *row_off,
                              const int32_t *pairs) {
  RefTri *h = (RefTri *)hp;
  std::map<int, Eigen::MatrixXi> matches;
  for (int g = 0; g < n_ng; ++g) {
    const int64_t n = row_off[g + 1] - row_off[g];
    Eigen::MatrixXi m(n, 2);
    for (int64_t r = 0; r < n; ++r) { m(r, 0) = pairs[2 * (row_off[g] + r)]; m(r, 1) = pairs[2 * (row_off[g] + r) + 1]; }
    matches[ng_ids[g]] = m;
    h->rows += n;
  }
  try {
    h->tri->TriangulateImage(img_id, matches);
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
int ref_tri_triangulate_image_exhaustive(void *hp, int img_id, int n_ng, const int32_t *ng_ids) {
  RefTri *h = (RefTri *)hp;
  std::vector<int> ngs(ng_ids, ng_ids + n_ng);
  try {
    h->tri->TriangulateImageExhaustiveMatch(img_id, ngs);
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
long long ref_tri_rows_tested(void *hp) { return ((RefTri *)hp)->rows; }

static void put_tri(const Line3d &l, double score, double *o) {
  o[0] = l.start[0]; o[1] = l.start[1]; o[2] = l.start[2]; o[3] = l.end[0]; o[4] = l.end[1]; o[5] = l.end[2];
  o[6] = l.depths[0]; o[7] = l.depths[1]; o[8] = l.uncertainty; o[9] = score;
}
int ref_tri_get_best(void *hp, int img_id, double *out_line, int32_t *out_ng, int32_t *out_ntris) {
  RefTri *h = (RefTri *)hp;
  auto &best = h->tri->tris_best_.at(img_id);
  auto &tris = h->tri->tris_.at(img_id);
  for (size_t l = 0; l < best.size(); ++l) {
    const TriTuple &t = best[l];
    const int n = (int)tris[l].size();
    if (n > 0) {
      put_tri(std::get<0>(t), std::get<1>(t), out_line + 10 * l);
      out_ng[2 * l] = std::get<2>(t).first; out_ng[2 * l + 1] = std::get<2>(t).second;
    } else { // the reference leaves tris_best_ default-constructed for nodes without candidates
      for (int k = 0; k < 10; ++k) out_line[10 * l + k] = 0.0;
      out_line[10 * l + 8] = -1.0;
      out_ng[2 * l] = out_ng[2 * l + 1] = 0;
    }
    if (out_ntris) out_ntris[l] = n;
  }
  return (int)best.size();
}
long long ref_tri_get_valid_edges(void *hp, int img_id, int64_t *off, int32_t *edges) {
  RefTri *h = (RefTri *)hp;
  auto &ve = h->tri->valid_edges_.at(img_id);
  auto &ngs = h->tri->neighbors_.at(img_id);
  long long n = 0;
  for (size_t l = 0; l < ve.size(); ++l) {
    if (off) off[l] = n;
    for (auto &e : ve[l]) {
      if (edges) { edges[2 * n] = ngs[e.first]; edges[2 * n + 1] = e.second; }
      ++n;
    }
  }
  if (off) off[ve.size()] = n;
  return n;
}
int ref_tri_get_tris_node(void *hp, int img_id, int line_id, int cap, double *out_line, int32_t *out_ng) {
  RefTri *h = (RefTri *)hp;
  auto &tris = h->tri->tris_.at(img_id)[line_id];
  const int n = (int)tris.size();
  for (int i = 0; i < n && i < cap; ++i) {
    put_tri(std::get<0>(tris[i]), std::get<1>(tris[i]), out_line + 10 * i);
    out_ng[2 * i] = std::get<2>(tris[i]).first; out_ng[2 * i + 1] = std::get<2>(tris[i]).second;
  }
  return n;
}
int ref_tri_compute_tracks(void *hp, int64_t *n_nodes_total) {
  RefTri *h = (RefTri *)hp;
  try {
    h->tri->ComputeLineTracks();
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  int64_t n = 0;
  for (auto &t : h->tri->tracks_) n += (int64_t)t.image_id_list.size();
  if (n_nodes_total) *n_nodes_total = n;
  return (int)h->tri->tracks_.size();
}
int ref_tri_get_tracks(void *hp, int64_t *track_off, int32_t *img_ids, int32_t *line_ids, int32_t *node_ids,
                       double *node_line3d, double *track_line) {
  RefTri *h = (RefTri *)hp;
  auto &tr = h->tri->tracks_;
  int64_t n = 0;
  for (size_t t = 0; t < tr.size(); ++t) {
    track_off[t] = n;
    for (size_t k = 0; k < tr[t].image_id_list.size(); ++k, ++n) {
      img_ids[n] = tr[t].image_id_list[k];
      line_ids[n] = tr[t].line_id_list[k];
      node_ids[n] = tr[t].node_id_list[k];
      put_tri(tr[t].line3d_list[k], tr[t].score_list[k], node_line3d + 10 * n);
    }
    const Line3d &L = tr[t].line;
    double *o = track_line + 7 * t;
    o[0] = L.start[0]; o[1] = L.start[1]; o[2] = L.start[2]; o[3] = L.end[0]; o[4] = L.end[1]; o[5] = L.end[2];
    o[6] = L.uncertainty;
  }
  track_off[tr.size()] = n;
  return (int)tr.size();
}

// ---- free functions ---------------------------------------------------------------------------------------------------
static CameraView mkview(const double *cam /*model,fx,fy,cx,cy,qw,qx,qy,qz,tx,ty,tz*/) {
  const int model = (int)cam[0];
  std::vector<double> params;
  if (model == 0) params = {cam[1], cam[3], cam[4]};
  else params = {cam[1], cam[2], cam[3], cam[4]};
  return CameraView(Camera(model, params, 0), CameraPose(V4D(cam[5], cam[6], cam[7], cam[8]), V3D(cam[9], cam[10], cam[11])));
}
static Line2d mk2(const double *s) { return Line2d(V2D(s[0], s[1]), V2D(s[2], s[3])); }
double ref_line2d_length(const double *seg) { return mk2(seg).length(); }
void ref_line2d_direction(const double *seg, double *out) { V2D d = mk2(seg).direction(); out[0] = d[0]; out[1] = d[1]; }
double ref_compute_epipolar_IoU(const double *l1, const double *cam1, const double *l2, const double *cam2) {
  return triangulation::compute_epipolar_IoU(mk2(l1), mkview(cam1), mk2(l2), mkview(cam2));
}
void ref_triangulate_line(const double *l1, const double *cam1, const double *l2, const double *cam2, int by_endpoints,
                          double *out) {
  Line3d L = by_endpoints ? triangulation::triangulate_line_by_endpoints(mk2(l1), mkview(cam1), mk2(l2), mkview(cam2))
                          : triangulation::triangulate_line(mk2(l1), mkview(cam1), mk2(l2), mkview(cam2));
  out[0] = L.start[0]; out[1] = L.start[1]; out[2] = L.start[2]; out[3] = L.end[0]; out[4] = L.end[1]; out[5] = L.end[2];
  out[6] = L.depths[0]; out[7] = L.depths[1]; out[8] = L.score;
}
// triangulate_line_with_direction (functions.cc:389-446): out as above
void ref_triangulate_line_with_direction(const double *l1, const double *cam1, const double *l2, const double *cam2,
                                         const double *direction, double *out) {
  Line3d L = triangulation::triangulate_line_with_direction(mk2(l1), mkview(cam1), mk2(l2), mkview(cam2),
                                                            V3D(direction[0], direction[1], direction[2]));
  out[0] = L.start[0]; out[1] = L.start[1]; out[2] = L.start[2]; out[3] = L.end[0]; out[4] = L.end[1]; out[5] = L.end[2];
  out[6] = L.depths[0]; out[7] = L.depths[1]; out[8] = L.score;
}
void ref_project_point(const double *cam, const double *X, double *out) {
  V2D p = mkview(cam).projection(V3D(X[0], X[1], X[2]));
  out[0] = p[0]; out[1] = p[1];
}
void ref_ray_direction(const double *cam, const double *p, double *out) {
  V3D r = mkview(cam).ray_direction(V2D(p[0], p[1]));
  out[0] = r[0]; out[1] = r[1]; out[2] = r[2];
}
static Line3d mkline3(const double *l) { return Line3d(V3D(l[0], l[1], l[2]), V3D(l[3], l[4], l[5]), 1.0, l[6], l[7], l[8]); }
double ref_line3d_sensitivity(const double *l, const double *cam) { return mkline3(l).sensitivity(mkview(cam)); }
double ref_line3d_uncertainty(const double *l, const double *cam, double var2d) { return mkline3(l).computeUncertainty(mkview(cam), var2d); }
double ref_score_3d(const ref_linker_cfg *c, const double *l1, const double *l2) {
  LineLinker3d lk(to_linker3d(*c));
  return lk.compute_score(mkline3(l1), mkline3(l2));
}
double ref_score_2d(const ref_linker_cfg *c, const double *l1, const double *l2) {
  LineLinker2d lk(to_linker2d(*c));
  return lk.compute_score(mk2(l1), mk2(l2));
}
int ref_check_connection_3d(const ref_linker_cfg *c, const double *l1, const double *l2) {
  LineLinker3d lk(to_linker3d(*c));
  return lk.check_connection(mkline3(l1), mkline3(l2)) ? 1 : 0;
}

// Aggregator::aggregate_line3d_list for T groups: lines[n][7] = start, end, uncertainty; out[T][7]
int ref_aggregate_lines(int64_t T, const int64_t *off, const double *lines, const double *scores, int32_t num_outliers,
                        double *out) {
  for (int64_t t = 0; t < T; ++t) {
    std::vector<Line3d> ls;
    std::vector<double> sc;
    for (int64_t s = off[t]; s < off[t + 1]; ++s) {
      const double *l = lines + 7 * s;
      Line3d x(V3D(l[0], l[1], l[2]), V3D(l[3], l[4], l[5]));
      x.uncertainty = l[6];
      ls.push_back(x);
      sc.push_back(scores[s]);
    }
    double *o = out + 7 * t;
    if (ls.empty()) { memset(o, 0, 7 * sizeof(double)); continue; }
    Line3d r = merging::Aggregator::aggregate_line3d_list(ls, sc, num_outliers);
    o[0] = r.start[0]; o[1] = r.start[1]; o[2] = r.start[2]; o[3] = r.end[0]; o[4] = r.end[1]; o[5] = r.end[2];
    o[6] = r.uncertainty;
  }
  return 0;
}

void ref_minimal_from_line(const double *line, double *out6) {
  MinimalInfiniteLine3d ml(Line3d(V3D(line[0], line[1], line[2]), V3D(line[3], line[4], line[5])));
  for (int i = 0; i < 4; ++i) out6[i] = ml.uvec[i];
  out6[4] = ml.wvec[0]; out6[5] = ml.wvec[1];
}
void ref_infinite_from_minimal(const double *x6, double *d3, double *m3) {
  MinimalInfiniteLine3d ml(std::vector<double>(x6, x6 + 6));
  InfiniteLine3d inf = ml.GetInfiniteLine();
  for (int i = 0; i < 3; ++i) { d3[i] = inf.d[i]; m3[i] = inf.m[i]; }
}
// GetLineSegmentFromInfiniteLine3d(inf_line(x6), line3ds, num_outliers) (base/infinite_line.cc:265-287)
void ref_segment_from_minimal(const double *x6, const double *line3d, int64_t n, int num_outliers, double *out6) {
  MinimalInfiniteLine3d ml(std::vector<double>(x6, x6 + 6));
  std::vector<Line3d> ls;
  for (int64_t k = 0; k < n; ++k) ls.push_back(Line3d(V3D(line3d[6 * k], line3d[6 * k + 1], line3d[6 * k + 2]),
                                                     V3D(line3d[6 * k + 3], line3d[6 * k + 4], line3d[6 * k + 5])));
  Line3d r = GetLineSegmentFromInfiniteLine3d(ml.GetInfiniteLine(), ls, num_outliers);
  for (int i = 0; i < 3; ++i) { out6[i] = r.start[i]; out6[3 + i] = r.end[i]; }
}

// CheckReprojection / CheckSensitivity / overlap bits per supporting line (merging_utils.cc:27-155), layout as in
// orc_track_support_flags.
int ref_track_support_flags(int32_t n_views, const int32_t *model_ids, const double *kvec, const double *qvec,
                            const double *tvec, int64_t T, const int64_t *sup_off, const int32_t *sup_view,
                            const double *segs, const double *track_line, double th_angular2d, double th_perp2d,
                            double th_sv_angular3d, double th_overlap, uint8_t *out_flags) {
  std::map<int, Camera> cams;
  std::map<int, CameraImage> imgs;
  for (int v = 0; v < n_views; ++v) {
    const double *k = kvec + 4 * v;
    const int model = model_ids ? model_ids[v] : 1;
    std::vector<double> params;
    if (model == 0) params = {k[0], k[2], k[3]};
    else params = {k[0], k[1], k[2], k[3]};
    cams[v] = Camera(model, params, v);
    imgs[v] = CameraImage(v, CameraPose(V4D(qvec[4 * v], qvec[4 * v + 1], qvec[4 * v + 2], qvec[4 * v + 3]),
                                        V3D(tvec[3 * v], tvec[3 * v + 1], tvec[3 * v + 2])));
  }
  ImageCollection imagecols(cams, imgs);
  for (int64_t t = 0; t < T; ++t) {
    LineTrack tr;
    const double *tl = track_line + 6 * t;
    tr.line = Line3d(V3D(tl[0], tl[1], tl[2]), V3D(tl[3], tl[4], tl[5]));
    for (int64_t s = sup_off[t]; s < sup_off[t + 1]; ++s) {
      tr.image_id_list.push_back(sup_view[s]);
      tr.line_id_list.push_back((int)(s - sup_off[t]));
      tr.line2d_list.push_back(mk2(segs + 4 * s));
    }
    std::vector<bool> rep, sens;
    merging::CheckReprojection(rep, tr, imagecols, th_angular2d, th_perp2d);
    merging::CheckSensitivity(sens, tr, imagecols, th_sv_angular3d);
    for (int64_t s = sup_off[t]; s < sup_off[t + 1]; ++s) {
      const size_t k = (size_t)(s - sup_off[t]);
      uint8_t f = 0;
      if (rep[k]) f |= 1;
      if (sens[k]) f |= 2;
      // FilterTracksByOverlap (merging_utils.cc:143-150)
      Line2d proj = tr.line.projection(imagecols.camview(sup_view[s]));
      if (compute_overlap<Line2d>(proj, tr.line2d_list[k]) >= th_overlap) f |= 4;
      out_flags[s] = f;
    }
  }
  return 0;
}

// RemergeLineTracks (merging/merging.cc:513-645) on tracks that carry only their 3D line (+uncertainty) and active flag:
// out_labels[t] = index of the output track that absorbed input track t is not recoverable from the reference's return
// value, so the comparison is on the PARTITION: out_group[t] = smallest input index merged with t.
int64_t ref_remerge_groups(int64_t T, const double *track_line, const uint8_t *active, const ref_linker_cfg *c,
                           int32_t *out_group) {
  std::vector<LineTrack> tracks((size_t)T);
  for (int64_t t = 0; t < T; ++t) {
    const double *l = track_line + 7 * t;
    tracks[t].line = Line3d(V3D(l[0], l[1], l[2]), V3D(l[3], l[4], l[5]));
    tracks[t].line.uncertainty = l[6];
    tracks[t].active = active[t] != 0;
    // one supporting line per track, tagged with the input index, so that the merged tracks tell the partition
    tracks[t].image_id_list.push_back((int)t);
    tracks[t].line_id_list.push_back(0);
    tracks[t].line2d_list.push_back(Line2d(V2D(0, 0), V2D(1, 1)));
    tracks[t].line3d_list.push_back(tracks[t].line);
    tracks[t].score_list.push_back(1.0);
    tracks[t].node_id_list.push_back((int)t);
  }
  LineLinker3d linker3d(to_linker3d(*c));
  std::vector<LineTrack> out = merging::RemergeLineTracks(tracks, linker3d, 0);
  for (const LineTrack &o : out) {
    int mn = o.image_id_list.empty() ? -1 : o.image_id_list[0];
    for (int i : o.image_id_list) mn = std::min(mn, i);
    for (int i : o.image_id_list) out_group[i] = mn;
  }
  return (int64_t)out.size();
}

} // extern "C"

// The reference's residual functors (optimize/line_refinement/cost_functions.h:35-190) evaluated with forward-mode
// jets over the 6 line parameters, camera constant (the refine.cc configuration): res[2] (+ jac[2x6] row-major) for the
// geometric functor, res[1] (+ jac[1x6]) for the VP functor. model: 0 = SIMPLE_PINHOLE (params f,cx,cy), 1 = PINHOLE.
template <typename CameraModel>
static void geometric_residual_impl(const double *x, const double *seg, const double *params, const double *qvec,
                                    const double *tvec, double alpha, double *res, double *jac) {
  typedef ceres::Jet<double, 6> J;
  Line2d l(V2D(seg[0], seg[1]), V2D(seg[2], seg[3]));
  optimize::line_refinement::GeometricRefinementFunctor<CameraModel> f(l, params, qvec, tvec, alpha);
  J u[4] = {J(x[0], 0), J(x[1], 1), J(x[2], 2), J(x[3], 3)}, w[2] = {J(x[4], 4), J(x[5], 5)}, rr[2];
  f(u, w, rr);
  for (int i = 0; i < 2; ++i) { res[i] = rr[i].a; if (jac) for (int j = 0; j < 6; ++j) jac[6 * i + j] = rr[i].v[j]; }
}
template <typename CameraModel>
static void vp_residual_impl(const double *x, const double *vp, const double *params, const double *qvec, double *res,
                             double *jac) {
  typedef ceres::Jet<double, 6> J;
  optimize::line_refinement::VPConstraintsFunctor<CameraModel> f(V3D(vp[0], vp[1], vp[2]), params, qvec);
  J u[4] = {J(x[0], 0), J(x[1], 1), J(x[2], 2), J(x[3], 3)}, w[2] = {J(x[4], 4), J(x[5], 5)}, rr[1];
  f(u, w, rr);
  res[0] = rr[0].a;
  if (jac) for (int j = 0; j < 6; ++j) jac[j] = rr[0].v[j];
}
extern "C" {
void ref_geometric_residual(int model, const double *x, const double *seg, const double *params, const double *qvec,
                            const double *tvec, double alpha, double *res, double *jac) {
  if (model == 0) geometric_residual_impl<colmap::SimplePinholeCameraModel>(x, seg, params, qvec, tvec, alpha, res, jac);
  else geometric_residual_impl<colmap::PinholeCameraModel>(x, seg, params, qvec, tvec, alpha, res, jac);
}
void ref_vp_residual(int model, const double *x, const double *vp, const double *params, const double *qvec,
                     double *res, double *jac) {
  if (model == 0) vp_residual_impl<colmap::SimplePinholeCameraModel>(x, vp, params, qvec, res, jac);
  else vp_residual_impl<colmap::PinholeCameraModel>(x, vp, params, qvec, res, jac);
}
}

// limap::pointsfm::SfmModel built the way pointsfm/colmap_reader.py builds it (CreateSfmImage + addImage(-1) + addPoint),
// then one of its neighbour rankings: mode 0 GetMaxIoUImages, 1 GetMaxDiceCoeffImages, 2 GetMaxOverlapImages.
// out[n_images][num_images] (-1 padded), out_count[n_images]. R row-major 3x3 per image, T 3 per image.
extern "C" {
int ref_sfm_rank_neighbors(int n_images, const double *R, const double *T, int64_t n_points, const double *xyz,
                           const int64_t *track_off, const int32_t *track_img, int num_images,
                           double min_triangulation_angle_deg, int mode, int32_t *out, int32_t *out_count) {
  pointsfm::SfmModel model;
  const std::vector<double> K = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < n_images; ++i)
    model.addImage(pointsfm::CreateSfmImage("img_" + std::to_string(i), 800, 600, K, std::vector<double>(R + 9 * i, R + 9 * i + 9),
                                            std::vector<double>(T + 3 * i, T + 3 * i + 3)),
                   -1);
  for (int64_t p = 0; p < n_points; ++p)
    model.addPoint(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2],
                   std::vector<int>(track_img + track_off[p], track_img + track_off[p + 1]));
  std::map<int, std::vector<int>> nb;
  if (mode == 0) nb = model.GetMaxIoUImages((size_t)num_images, min_triangulation_angle_deg);
  else if (mode == 1) nb = model.GetMaxDiceCoeffImages((size_t)num_images, min_triangulation_angle_deg);
  else nb = model.GetMaxOverlapImages((size_t)num_images, min_triangulation_angle_deg);
  for (int i = 0; i < n_images; ++i) {
    const std::vector<int> &v = nb.at(i);
    out_count[i] = (int32_t)v.size();
    for (int k = 0; k < num_images; ++k) out[(int64_t)i * num_images + k] = k < (int)v.size() ? v[k] : -1;
  }
  return 0;
}
void ref_sfm_robust_ranges(int64_t n_points, const double *xyz, double q_lo, double q_hi, double kstretch, double *out6) {
  pointsfm::SfmModel model;
  for (int64_t p = 0; p < n_points; ++p) model.addPoint(xyz[3 * p], xyz[3 * p + 1], xyz[3 * p + 2], std::vector<int>());
  const std::pair<V3D, V3D> r = model.ComputeRanges(std::make_pair(q_lo, q_hi), kstretch);
  for (int k = 0; k < 3; ++k) { out6[k] = r.first[k]; out6[3 + k] = r.second[k]; }
}
}

// ---- the JLinkage-library shim (oracle/ref_shim/JLinkage/include): VPSample::run only records the model count, VPCluster::run
// runs the restated J-Linkage (orc::jlinkage_cluster) with the seed / image index set by ref_vp_set_context -------------------
namespace {
thread_local uint64_t g_vp_seed = 0, g_vp_image_index = 0;
thread_local int g_vp_models = 5000;
} // namespace
namespace VPSample {
std::vector<std::vector<float> *> *run(std::vector<std::vector<float> *> *, int num_models, int, int, int) {
  g_vp_models = num_models;
  return new std::vector<std::vector<float> *>();
}
} // namespace VPSample
namespace VPCluster {
int run(std::vector<unsigned int> &labels, std::vector<unsigned int> &label_count, std::vector<std::vector<float> *> *pts,
        std::vector<std::vector<float> *> *, float inlier_threshold, int) {
  std::vector<std::array<float, 4>> p(pts->size());
  for (size_t i = 0; i < pts->size(); ++i)
    for (int k = 0; k < 4; ++k) p[i][k] = (*(*pts)[i])[k];
  orc::VPConfig cfg;
  cfg.inlier_threshold = inlier_threshold;
  cfg.n_models = g_vp_models;
  cfg.seed = g_vp_seed;
  std::vector<int> lab;
  const int nc = orc::jlinkage_cluster(p, cfg, g_vp_image_index, lab);
  labels.assign(lab.begin(), lab.end());
  label_count.assign((size_t)nc, 0u);
  for (int l : lab)
    if (l >= 0) label_count[(size_t)l] += 1;
  return nc;
}
} // namespace VPCluster

extern "C" {
void ref_vp_set_context(uint64_t seed, uint64_t image_index) { g_vp_seed = seed; g_vp_image_index = image_index; }
// vplib::JLinkage::JLinkage(config).AssociateVPs(lines) of the reference's compiled wrapper. Returns the number of VPs
// (may exceed cap); labels[n_lines], vps[cap][3].
long long ref_vp_associate(int64_t n_lines, const double *segs, double min_length, double inlier_threshold, int min_num_supports,
                           double th_perp_supports, int32_t *labels, double *vps, long long cap) {
  vplib::JLinkage::JLinkageConfig cfg;
  cfg.min_length = min_length;
  cfg.inlier_threshold = inlier_threshold;
  cfg.min_num_supports = min_num_supports;
  cfg.th_perp_supports = th_perp_supports;
  vplib::JLinkage::JLinkage det(cfg);
  std::vector<Line2d> lines;
  for (int64_t l = 0; l < n_lines; ++l)
    lines.push_back(Line2d(V2D(segs[4 * l], segs[4 * l + 1]), V2D(segs[4 * l + 2], segs[4 * l + 3])));
  const vplib::VPResult r = det.AssociateVPs(lines);
  for (int64_t l = 0; l < n_lines && l < (int64_t)r.labels.size(); ++l) labels[l] = r.labels[l];
  long long n = 0;
  for (const V3D &v : r.vps) {
    if (n < cap) { vps[3 * n] = v[0]; vps[3 * n + 1] = v[1]; vps[3 * n + 2] = v[2]; }
    ++n;
  }
  return n;
}
}

// LineTrack::Write / LineTrack::Read (base/linetrack.cc:133-270) and ComputeLineWeights (:315-322) of the reference's
// compiled code, on flat arrays: line[6]; per supporting line image id, line id, node id, score, line2d[4], line3d[6].
extern "C" {
int ref_linetrack_write(const char *path, const double *line, int64_t n, const int32_t *img, const int32_t *lid,
                        const int32_t *node, const double *score, const double *l2d, const double *l3d) {
  LineTrack t;
  t.line = Line3d(V3D(line[0], line[1], line[2]), V3D(line[3], line[4], line[5]));
  for (int64_t i = 0; i < n; ++i) {
    t.image_id_list.push_back(img[i]);
    t.line_id_list.push_back(lid[i]);
    if (node) t.node_id_list.push_back(node[i]);
    if (score) t.score_list.push_back(score[i]);
    t.line2d_list.push_back(Line2d(V2D(l2d[4 * i], l2d[4 * i + 1]), V2D(l2d[4 * i + 2], l2d[4 * i + 3])));
    if (l3d) t.line3d_list.push_back(Line3d(V3D(l3d[6 * i], l3d[6 * i + 1], l3d[6 * i + 2]), V3D(l3d[6 * i + 3], l3d[6 * i + 4], l3d[6 * i + 5])));
  }
  t.Write(std::string(path));
  return 0;
}
// returns the number of supporting lines (arrays sized by the caller: cap entries), -1 when cap is too small
int64_t ref_linetrack_read(const char *path, int64_t cap, double *line, int32_t *img, int32_t *lid, int32_t *node, double *score,
                           double *l2d, double *l3d, int32_t *n_images) {
  LineTrack t;
  t.Read(std::string(path));
  const int64_t n = (int64_t)t.count_lines();
  if (n > cap) return -1;
  for (int k = 0; k < 3; ++k) { line[k] = t.line.start[k]; line[3 + k] = t.line.end[k]; }
  for (int64_t i = 0; i < n; ++i) {
    img[i] = t.image_id_list[i];
    lid[i] = t.line_id_list[i];
    node[i] = i < (int64_t)t.node_id_list.size() ? t.node_id_list[i] : -1;
    score[i] = i < (int64_t)t.score_list.size() ? t.score_list[i] : 0.0;
    for (int k = 0; k < 2; ++k) { l2d[4 * i + k] = t.line2d_list[i].start[k]; l2d[4 * i + 2 + k] = t.line2d_list[i].end[k]; }
    if (i < (int64_t)t.line3d_list.size())
      for (int k = 0; k < 3; ++k) { l3d[6 * i + k] = t.line3d_list[i].start[k]; l3d[6 * i + 3 + k] = t.line3d_list[i].end[k]; }
  }
  *n_images = (int32_t)t.count_images();
  return n;
}
void ref_line_weights(int64_t n, const double *l2d, double *out) {
  LineTrack t;
  for (int64_t i = 0; i < n; ++i)
    t.line2d_list.push_back(Line2d(V2D(l2d[4 * i], l2d[4 * i + 1]), V2D(l2d[4 * i + 2], l2d[4 * i + 3])));
  t.image_id_list.assign((size_t)n, 0);
  t.line_id_list.assign((size_t)n, 0);
  std::vector<double> w;
  ComputeLineWeights(t, w);
  for (int64_t i = 0; i < n; ++i) out[i] = w[(size_t)i];
}
}

// Track-level post-triangulation operators of the reference's compiled merging code on flat track arrays:
//   op 0 FilterSupportingLines(th_angular2d = a, th_perp2d = b, num_outliers = n)        merging_utils.cc:51-83
//   op 1 FilterTracksBySensitivity(th_angular3d = a, min_support_ns = n)                 :105-128
//   op 2 FilterTracksByOverlap(th_overlap = a, min_support_ns = n)                       :130-155
//   op 3 RemergeLineTracks(linker3d, num_outliers = n), iterated until the track count is stable (merging.py:24-42)
// Tracks: T, sup_off[T+1]; track_line[T][9] = start, end, depth_start, depth_end, uncertainty; active[T];
// per supporting line: view index (= image id), line id, node id, score, line2d[4], line3d[9]. Outputs in the same layout
// (capacities = the inputs' sizes: no operator creates supporting lines). Returns the number of output tracks.
extern "C" {
int64_t ref_track_filter(int op, double a, double b, int n, const ref_linker_cfg *lk, int32_t n_views, const int32_t *model_ids,
                         const double *kvec, const double *qvec, const double *tvec, int64_t T, const int64_t *sup_off,
                         const double *track_line, const uint8_t *active, const int32_t *sup_view, const int32_t *sup_lid,
                         const int32_t *sup_node, const double *sup_score, const double *l2d, const double *l3d,
                         int64_t *o_off, double *o_line, uint8_t *o_active, int32_t *o_view, int32_t *o_lid, int32_t *o_node,
                         double *o_score, double *o_l2d, double *o_l3d) {
  std::map<int, Camera> cams;
  std::map<int, CameraImage> imgs;
  for (int v = 0; v < n_views; ++v) {
    const double *k = kvec + 4 * v;
    const int model = model_ids ? model_ids[v] : 1;
    std::vector<double> params;
    if (model == 0) params = {k[0], k[2], k[3]};
    else params = {k[0], k[1], k[2], k[3]};
    cams[v] = Camera(model, params, v);
    imgs[v] = CameraImage(v, CameraPose(V4D(qvec[4 * v], qvec[4 * v + 1], qvec[4 * v + 2], qvec[4 * v + 3]),
                                        V3D(tvec[3 * v], tvec[3 * v + 1], tvec[3 * v + 2])));
  }
  ImageCollection imagecols(cams, imgs);
  auto mk9 = [](const double *l) { return Line3d(V3D(l[0], l[1], l[2]), V3D(l[3], l[4], l[5]), 1.0, l[6], l[7], l[8]); };
  std::vector<LineTrack> tracks((size_t)T);
  for (int64_t t = 0; t < T; ++t) {
    LineTrack &tr = tracks[(size_t)t];
    tr.line = mk9(track_line + 9 * t);
    tr.active = active[t] != 0;
    for (int64_t s = sup_off[t]; s < sup_off[t + 1]; ++s) {
      tr.image_id_list.push_back(sup_view[s]);
      tr.line_id_list.push_back(sup_lid[s]);
      tr.node_id_list.push_back(sup_node[s]);
      tr.score_list.push_back(sup_score[s]);
      tr.line2d_list.push_back(mk2(l2d + 4 * s));
      tr.line3d_list.push_back(mk9(l3d + 9 * s));
    }
  }
  std::vector<LineTrack> out;
  if (op == 0) merging::FilterSupportingLines(out, tracks, imagecols, a, b, n);
  else if (op == 1) merging::FilterTracksBySensitivity(out, tracks, imagecols, a, n);
  else if (op == 2) merging::FilterTracksByOverlap(out, tracks, imagecols, a, n);
  else {
    out = tracks;
    size_t count = out.size();
    while (true) {
      out = merging::RemergeLineTracks(out, LineLinker3d(to_linker3d(*lk)), n);
      if (out.size() == count) break;
      count = out.size();
    }
  }
  int64_t k = 0;
  for (size_t t = 0; t < out.size(); ++t) {
    const LineTrack &tr = out[t];
    o_off[t] = k;
    double *ol = o_line + 9 * t;
    for (int q = 0; q < 3; ++q) { ol[q] = tr.line.start[q]; ol[3 + q] = tr.line.end[q]; }
    ol[6] = tr.line.depths[0]; ol[7] = tr.line.depths[1]; ol[8] = tr.line.uncertainty;
    o_active[t] = tr.active ? 1 : 0;
    for (size_t s = 0; s < tr.count_lines(); ++s, ++k) {
      o_view[k] = tr.image_id_list[s];
      o_lid[k] = tr.line_id_list[s];
      o_node[k] = tr.node_id_list[s];
      o_score[k] = tr.score_list[s];
      for (int q = 0; q < 2; ++q) { o_l2d[4 * k + q] = tr.line2d_list[s].start[q]; o_l2d[4 * k + 2 + q] = tr.line2d_list[s].end[q]; }
      const Line3d &l = tr.line3d_list[s];
      for (int q = 0; q < 3; ++q) { o_l3d[9 * k + q] = l.start[q]; o_l3d[9 * k + 3 + q] = l.end[q]; }
      o_l3d[9 * k + 6] = l.depths[0]; o_l3d[9 * k + 7] = l.depths[1]; o_l3d[9 * k + 8] = l.uncertainty;
    }
  }
  o_off[out.size()] = k;
  return (int64_t)out.size();
}
}

// Camera::set_max_image_dim (base/camera.cc:216-226; colmap::Camera::Rescale restated in the shim): params in / out, hw in / out
extern "C" void ref_camera_set_max_image_dim(int model, double *params, int32_t *hw, int val) {
  std::vector<double> p(params, params + (model == 0 ? 3 : 4));
  Camera cam(model, p, 0, std::make_pair(hw[0], hw[1]));
  cam.set_max_image_dim(val);
  for (size_t i = 0; i < cam.params.size(); ++i) params[i] = cam.params[i];
  hw[0] = (int32_t)cam.h();
  hw[1] = (int32_t)cam.w();
}
