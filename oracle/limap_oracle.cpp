// oracle/limap_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// PARITY PINNED to oracle/_ref (see orc_geom.h header).
//
// C entry points (ctypes) over the fp64 CPU restatement of the limap
// triangulation / scoring / track-building path. Built by oracle/Makefile into
// oracle/_build/liblimap_oracle.so. Used only by tests/, smoke() and bench.py's
// cpu_baseline / --impl reference legs.
#include "orc_triangulation.h"
#include <cstring>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

extern "C" {

struct orc_linker_cfg {
  double score_th, th_angle, th_overlap, th_smartoverlap, th_smartangle, th_perp, th_innerseg, th_scaleinv;
  int32_t use_angle, use_overlap, use_smartangle, use_perp, use_innerseg, use_scaleinv;
};
struct orc_tri_cfg {
  double min_length_2d, line_tri_angle_threshold, IoU_threshold, sensitivity_threshold, var2d, fullscore_th;
  int32_t debug_mode, add_halfpix, use_vp, use_endpoints_triangulation;
  int32_t disable_many_points_triangulation, disable_one_point_triangulation;
  int32_t disable_algebraic_triangulation, disable_vp_triangulation;
  int32_t max_valid_conns, min_num_outer_edges, num_outliers_aggregator, merging_strategy;
  orc_linker_cfg linker2d, linker3d;
};

static LinkerConfig to_linker(const orc_linker_cfg &c) {
  LinkerConfig l{c.score_th, c.th_angle, c.th_overlap, c.th_smartoverlap, c.th_smartangle,
                 c.th_perp, c.th_innerseg, c.th_scaleinv, c.use_angle, c.use_overlap,
                 c.use_smartangle, c.use_perp, c.use_innerseg, c.use_scaleinv};
  return l;
}

struct OrcTri {
  std::unique_ptr<GlobalLineTriangulator> tri;
  std::map<int, std::vector<Line2d>> segs;
  std::map<int, CameraView> views;
  std::string err;
};

static thread_local std::string g_err;
const char *orc_last_error() { return g_err.c_str(); }

int orc_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

// 0: the reference's OpenMP schedule (inside one node); 1: OpenMP over the 2D lines of the image (same results).
void orc_tri_set_node_parallel(void *h, int flag);

void *orc_tri_create(const orc_tri_cfg *c) {
  TriConfig cfg;
  cfg.debug_mode = c->debug_mode; cfg.add_halfpix = c->add_halfpix; cfg.use_vp = c->use_vp;
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
  cfg.linker2d = to_linker(c->linker2d); cfg.linker3d = to_linker(c->linker3d);
  if (c->merging_strategy != 0) { g_err = "only the greedy merging strategy is restated"; return nullptr; }
  OrcTri *h = new OrcTri();
  h->tri.reset(new GlobalLineTriangulator(cfg));
  return h;
}
void orc_tri_destroy(void *hp) { delete (OrcTri *)hp; }
void orc_tri_set_node_parallel(void *hp, int flag) { ((OrcTri *)hp)->tri->node_parallel_ = flag != 0; }

int orc_tri_init(void *hp, int n_views, const int32_t *img_ids, const int32_t *model_ids,
                 const double *kvec, const double *qvec, const double *tvec,
                 const int64_t *line_off, const double *segs) {
  OrcTri *h = (OrcTri *)hp;
  for (int v = 0; v < n_views; ++v) {
    CameraView view;
    view.cam.model_id = model_ids[v];
    for (int k = 0; k < 4; ++k) view.cam.kvec[k] = kvec[4 * v + k];
    view.pose.set(qvec + 4 * v, tvec + 3 * v);
    h->views[img_ids[v]] = view;
    std::vector<Line2d> lines;
    for (int64_t l = line_off[v]; l < line_off[v + 1]; ++l)
      lines.push_back(Line2d(V2(segs[4 * l], segs[4 * l + 1]), V2(segs[4 * l + 2], segs[4 * l + 3])));
    h->segs[img_ids[v]] = lines;
  }
  h->tri->Init(h->segs, h->views);
  return 0;
}
void orc_tri_set_ranges(void *hp, const double *lo, const double *hi) {
  ((OrcTri *)hp)->tri->SetRanges(V3(lo[0], lo[1], lo[2]), V3(hi[0], hi[1], hi[2]));
}
void orc_tri_unset_ranges(void *hp) { ((OrcTri *)hp)->tri->UnsetRanges(); }

static std::map<int, VPResult> g_dummy;
int orc_tri_set_vps(void *hp, int n_images, const int32_t *img_ids, const int64_t *label_off,
                    const int32_t *labels, const int64_t *vp_off, const double *vps) {
  std::map<int, VPResult> res;
  for (int i = 0; i < n_images; ++i) {
    VPResult r;
    r.labels.assign(labels + label_off[i], labels + label_off[i + 1]);
    for (int64_t k = vp_off[i]; k < vp_off[i + 1]; ++k) r.vps.push_back(V3(vps[3 * k], vps[3 * k + 1], vps[3 * k + 2]));
    res[img_ids[i]] = r;
  }
  ((OrcTri *)hp)->tri->InitVPResults(res);
  return 0;
}

int orc_tri_triangulate_image(void *hp, int img_id, int n_ng, const int32_t *ng_ids,
                              const int64_t *row_off, const int32_t *pairs) {
  OrcTri *h = (OrcTri *)hp;
  std::map<int, std::vector<std::pair<int, int>>> matches;
  for (int g = 0; g < n_ng; ++g) {
    auto &rows = matches[ng_ids[g]];
    for (int64_t r = row_off[g]; r < row_off[g + 1]; ++r)
      rows.push_back(std::make_pair((int)pairs[2 * r], (int)pairs[2 * r + 1]));
  }
  try {
    h->tri->TriangulateImage(img_id, matches);
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
int orc_tri_triangulate_image_exhaustive(void *hp, int img_id, int n_ng, const int32_t *ng_ids) {
  OrcTri *h = (OrcTri *)hp;
  std::vector<int> ngs(ng_ids, ng_ids + n_ng);
  try {
    h->tri->TriangulateImageExhaustiveMatch(img_id, ngs);
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  }
  return 0;
}
long long orc_tri_rows_tested(void *hp) { return ((OrcTri *)hp)->tri->n_match_rows_tested_; }

// best tri per line of an image: out_line[L][10] = start3,end3,depth2,unc,score ; out_ng[L][2]
int orc_tri_get_best(void *hp, int img_id, double *out_line, int32_t *out_ng, int32_t *out_ntris) {
  OrcTri *h = (OrcTri *)hp;
  auto &best = h->tri->tris_best_.at(img_id);
  auto nt = h->tri->n_tris_.find(img_id);
  for (size_t l = 0; l < best.size(); ++l) {
    const TriTuple &t = best[l];
    double *o = out_line + 10 * l;
    o[0] = t.line.start.x; o[1] = t.line.start.y; o[2] = t.line.start.z;
    o[3] = t.line.end.x; o[4] = t.line.end.y; o[5] = t.line.end.z;
    o[6] = t.line.depths[0]; o[7] = t.line.depths[1]; o[8] = t.line.uncertainty; o[9] = t.score;
    out_ng[2 * l] = t.ng_img; out_ng[2 * l + 1] = t.ng_line;
    if (out_ntris) out_ntris[l] = (nt == h->tri->n_tris_.end()) ? 0 : nt->second[l];
  }
  return (int)best.size();
}
// valid edges of an image. Call with edges == NULL to get the count.
long long orc_tri_get_valid_edges(void *hp, int img_id, int64_t *off, int32_t *edges) {
  OrcTri *h = (OrcTri *)hp;
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
// debug_mode: all scored candidates of a node, in reference order.
int orc_tri_get_tris_node(void *hp, int img_id, int line_id, int cap, double *out_line, int32_t *out_ng) {
  OrcTri *h = (OrcTri *)hp;
  auto &tris = h->tri->tris_.at(img_id)[line_id];
  int n = (int)tris.size();
  for (int i = 0; i < n && i < cap; ++i) {
    const TriTuple &t = tris[i];
    double *o = out_line + 10 * i;
    o[0] = t.line.start.x; o[1] = t.line.start.y; o[2] = t.line.start.z;
    o[3] = t.line.end.x; o[4] = t.line.end.y; o[5] = t.line.end.z;
    o[6] = t.line.depths[0]; o[7] = t.line.depths[1]; o[8] = t.line.uncertainty; o[9] = t.score;
    out_ng[2 * i] = t.ng_img; out_ng[2 * i + 1] = t.ng_line;
  }
  return n;
}

int orc_tri_compute_tracks(void *hp, int64_t *n_nodes_total) {
  OrcTri *h = (OrcTri *)hp;
  auto &tr = h->tri->ComputeLineTracks();
  int64_t n = 0;
  for (auto &t : tr) n += (int64_t)t.image_id_list.size();
  if (n_nodes_total) *n_nodes_total = n;
  return (int)tr.size();
}
// track_off[T+1]; per supporting node: img, line, node_id, line3d[10]; per track: line[7] = start3,end3,unc
int orc_tri_get_tracks(void *hp, int64_t *track_off, int32_t *img_ids, int32_t *line_ids,
                       int32_t *node_ids, double *node_line3d, double *track_line) {
  OrcTri *h = (OrcTri *)hp;
  auto &tr = h->tri->tracks_;
  int64_t n = 0;
  for (size_t t = 0; t < tr.size(); ++t) {
    track_off[t] = n;
    for (size_t k = 0; k < tr[t].image_id_list.size(); ++k, ++n) {
      img_ids[n] = tr[t].image_id_list[k];
      line_ids[n] = tr[t].line_id_list[k];
      node_ids[n] = tr[t].node_id_list[k];
      const Line3d &l = tr[t].line3d_list[k];
      double *o = node_line3d + 10 * n;
      o[0] = l.start.x; o[1] = l.start.y; o[2] = l.start.z; o[3] = l.end.x; o[4] = l.end.y; o[5] = l.end.z;
      o[6] = l.depths[0]; o[7] = l.depths[1]; o[8] = l.uncertainty; o[9] = tr[t].score_list[k];
    }
    const Line3d &L = tr[t].line;
    double *o = track_line + 7 * t;
    o[0] = L.start.x; o[1] = L.start.y; o[2] = L.start.z; o[3] = L.end.x; o[4] = L.end.y; o[5] = L.end.z;
    o[6] = L.uncertainty;
  }
  track_off[tr.size()] = n;
  return (int)tr.size();
}
// graph produced by run_clustering: nodes (img,line) in creation order and weighted edges.
long long orc_tri_get_graph(void *hp, int32_t *nodes, double *edge_w, int32_t *edge_nodes, long long *n_edges) {
  OrcTri *h = (OrcTri *)hp;
  auto &nd = h->tri->nodes_;
  auto &ed = h->tri->graph_edges_;
  if (nodes) for (size_t i = 0; i < nd.size(); ++i) { nodes[2 * i] = nd[i].first; nodes[2 * i + 1] = nd[i].second; }
  if (edge_w) for (size_t i = 0; i < ed.size(); ++i) {
    edge_w[i] = std::get<0>(ed[i]); edge_nodes[2 * i] = (int)std::get<1>(ed[i]); edge_nodes[2 * i + 1] = (int)std::get<2>(ed[i]);
  }
  if (n_edges) *n_edges = (long long)ed.size();
  return (long long)nd.size();
}

// ---- free functions for unit-level parity tests -----------------------------
static CameraView mkview(const double *cam /*model,fx,fy,cx,cy,qw,qx,qy,qz,tx,ty,tz*/) {
  CameraView v;
  v.cam.model_id = (int)cam[0];
  for (int k = 0; k < 4; ++k) v.cam.kvec[k] = cam[1 + k];
  v.pose.set(cam + 5, cam + 9);
  return v;
}
double orc_line2d_length(const double *seg) { return Line2d(V2(seg[0], seg[1]), V2(seg[2], seg[3])).length(); }
void orc_line2d_direction(const double *seg, double *out) {
  V2 d = Line2d(V2(seg[0], seg[1]), V2(seg[2], seg[3])).direction();
  out[0] = d.x; out[1] = d.y;
}
double orc_compute_epipolar_IoU(const double *l1, const double *cam1, const double *l2, const double *cam2) {
  return compute_epipolar_IoU(Line2d(V2(l1[0], l1[1]), V2(l1[2], l1[3])), mkview(cam1),
                              Line2d(V2(l2[0], l2[1]), V2(l2[2], l2[3])), mkview(cam2));
}
// out[9] = start3,end3,depth2,score
void orc_triangulate_line(const double *l1, const double *cam1, const double *l2, const double *cam2,
                          int by_endpoints, double *out) {
  Line2d a(V2(l1[0], l1[1]), V2(l1[2], l1[3])), b(V2(l2[0], l2[1]), V2(l2[2], l2[3]));
  Line3d L = by_endpoints ? triangulate_line_by_endpoints(a, mkview(cam1), b, mkview(cam2))
                          : triangulate_line(a, mkview(cam1), b, mkview(cam2));
  out[0] = L.start.x; out[1] = L.start.y; out[2] = L.start.z; out[3] = L.end.x; out[4] = L.end.y; out[5] = L.end.z;
  out[6] = L.depths[0]; out[7] = L.depths[1]; out[8] = L.score;
}
void orc_project_point(const double *cam, const double *X, double *out) {
  V2 p = mkview(cam).projection(V3(X[0], X[1], X[2]));
  out[0] = p.x; out[1] = p.y;
}
static Line3d mkline3(const double *l) { // start3,end3,depth2,unc
  Line3d L(V3(l[0], l[1], l[2]), V3(l[3], l[4], l[5]), 1.0, l[6], l[7], l[8]);
  return L;
}
double orc_score_3d(const orc_linker_cfg *c, const double *l1, const double *l2) {
  LineLinker3d lk; lk.config = to_linker(*c);
  return lk.compute_score(mkline3(l1), mkline3(l2));
}
void orc_triangulate_line_with_direction(const double *l1, const double *cam1, const double *l2, const double *cam2,
                                         const double *direction, double *out) {
  Line2d a(V2(l1[0], l1[1]), V2(l1[2], l1[3])), b(V2(l2[0], l2[1]), V2(l2[2], l2[3]));
  Line3d L = triangulate_line_with_direction(a, mkview(cam1), b, mkview(cam2), V3(direction[0], direction[1], direction[2]));
  out[0] = L.start.x; out[1] = L.start.y; out[2] = L.start.z; out[3] = L.end.x; out[4] = L.end.y; out[5] = L.end.z;
  out[6] = L.depths[0]; out[7] = L.depths[1]; out[8] = L.score;
}
void orc_ray_direction(const double *cam, const double *p, double *out) {
  V3 r = mkview(cam).ray_direction(V2(p[0], p[1]));
  out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
double orc_line3d_sensitivity(const double *l, const double *cam) { return mkline3(l).sensitivity(mkview(cam)); }
double orc_line3d_uncertainty(const double *l, const double *cam, double var2d) { return mkline3(l).computeUncertainty(mkview(cam), var2d); }
double orc_score_2d(const orc_linker_cfg *c, const double *l1, const double *l2) {
  LineLinker2d lk; lk.config = to_linker(*c);
  return lk.compute_score(Line2d(V2(l1[0], l1[1]), V2(l1[2], l1[3])), Line2d(V2(l2[0], l2[1]), V2(l2[2], l2[3])));
}

} // extern "C"
