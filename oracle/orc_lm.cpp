// oracle/orc_lm.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE. Pinning status: see orc_lm.h (functors pinned, solver loop unpinned).
// C entry points of the CPU restatement of the per-track line refinement / line bundle adjustment.
#include "orc_lm.h"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

extern "C" {

struct orc_lm_cfg {
  double geometric_alpha, cauchy_scale;
  int32_t max_num_iterations, min_num_images, num_outliers, mode, parallel_tracks, pad;
  double vp_multiplier;
};

// Batched refinement of T tracks with constant cameras.
//   mode 0: HybridBAEngine semantics (hybrid_bundle_adjustment.cc): every track is re-segmented from its
//           (possibly unchanged) minimal line; tracks with < min_num_images distinct images stay constant
//           (:106-123). The block-separable joint problem is solved track by track.
//   mode 1: RefinementEngine semantics (refine.cc) for every track given (the Python glue skips small ones).
// sup_off[T+1]; per support k: seg[4], kvec[4], qvec[4], tvec[3], img_id, l3d[6] (track.line3d_list[k]).
// out_line[T][6], out_min[T][6] = uvec,wvec; out_iters[T][2] = iterations, successful; out_cost[T][2].
int orc_refine_tracks(int T, const int64_t *sup_off, const double *segs, const double *kvec, const double *qvec,
                      const double *tvec, const int32_t *img_ids, const double *l3d, const double *line_init,
                      const double *sup_vp /* [n][3] or NULL; NaN = no VP */, const orc_lm_cfg *cfg, double *out_line, double *out_min, int32_t *out_iters, double *out_cost) {
#pragma omp parallel for schedule(dynamic, 8) if (cfg->parallel_tracks)
  for (int t = 0; t < T; ++t) {
    const int64_t a = sup_off[t], b = sup_off[t + 1];
    Line3d init(V3(line_init[6 * t], line_init[6 * t + 1], line_init[6 * t + 2]),
                V3(line_init[6 * t + 3], line_init[6 * t + 4], line_init[6 * t + 5]));
    MinimalLine ml = MinimalFromLine3d(init);
    double x[6] = {ml.uvec[0], ml.uvec[1], ml.uvec[2], ml.uvec[3], ml.wvec[0], ml.wvec[1]};
    std::set<int> imgs(img_ids + a, img_ids + b);
    LMSummary sum;
    LMProblem prob;
    prob.opt.max_num_iterations = cfg->max_num_iterations;
    prob.opt.geometric_alpha = cfg->geometric_alpha;
    prob.opt.cauchy_scale = cfg->cauchy_scale;
    for (int64_t k = a; k < b; ++k) {
      LMBlock blk;
      blk.p1[0] = segs[4 * k]; blk.p1[1] = segs[4 * k + 1]; blk.p2[0] = segs[4 * k + 2]; blk.p2[1] = segs[4 * k + 3];
      for (int i = 0; i < 4; ++i) { blk.kvec[i] = kvec[4 * k + i]; blk.qvec[i] = qvec[4 * k + i]; }
      for (int i = 0; i < 3; ++i) blk.tvec[i] = tvec[3 * k + i];
      // ComputeLineWeights (base/linetrack.cc:315-322): length / 30
      const double dx = blk.p1[0] - blk.p2[0], dy = blk.p1[1] - blk.p2[1];
      blk.w = std::sqrt(dx * dx + dy * dy) / 30.0;
      if (sup_vp && !std::isnan(sup_vp[3 * k])) {
        blk.has_vp = true;
        for (int i = 0; i < 3; ++i) blk.vp[i] = sup_vp[3 * k + i];
        blk.wvp = blk.w * cfg->vp_multiplier;
      }
      prob.blocks.push_back(blk);
    }
    const bool active = (int)imgs.size() >= cfg->min_num_images;
    if (active) sum = prob.solve(x);
    else { sum.initial_cost = sum.final_cost = prob.evaluate(x, nullptr, nullptr); }
    for (int i = 0; i < 6; ++i) out_min[6 * t + i] = x[i];
    MinimalLine res;
    for (int i = 0; i < 4; ++i) res.uvec[i] = x[i];
    res.wvec[0] = x[4]; res.wvec[1] = x[5];
    V3 d, m;
    InfiniteFromMinimal(res, d, m);
    std::vector<Line3d> l3;
    for (int64_t k = a; k < b; ++k)
      l3.push_back(Line3d(V3(l3d[6 * k], l3d[6 * k + 1], l3d[6 * k + 2]), V3(l3d[6 * k + 3], l3d[6 * k + 4], l3d[6 * k + 5])));
    Line3d seg = SegmentFromInfinite(d, m, l3, cfg->num_outliers);
    double *o = out_line + 6 * t;
    o[0] = seg.start.x; o[1] = seg.start.y; o[2] = seg.start.z; o[3] = seg.end.x; o[4] = seg.end.y; o[5] = seg.end.z;
    out_iters[2 * t] = sum.iterations; out_iters[2 * t + 1] = sum.successful;
    out_cost[2 * t] = sum.initial_cost; out_cost[2 * t + 1] = sum.final_cost;
  }
  return 0;
}

// residual (2) and ambient Jacobian (2x6) of one block at x = (uvec, wvec): unit-level parity hook
void orc_geometric_residual(const double *x, const double *seg, const double *kvec, const double *qvec,
                            const double *tvec, double alpha, double *res, double *jac) {
  LMBlock b;
  b.p1[0] = seg[0]; b.p1[1] = seg[1]; b.p2[0] = seg[2]; b.p2[1] = seg[3];
  for (int i = 0; i < 4; ++i) { b.kvec[i] = kvec[i]; b.qvec[i] = qvec[i]; }
  for (int i = 0; i < 3; ++i) b.tvec[i] = tvec[i];
  b.w = 1;
  typedef Jet<6> J;
  J u[4] = {J(x[0], 0), J(x[1], 1), J(x[2], 2), J(x[3], 3)}, w[2] = {J(x[4], 4), J(x[5], 5)}, rr[2];
  GeometricResidual<J>(b, u, w, alpha, rr);
  for (int i = 0; i < 2; ++i) { res[i] = rr[i].a; if (jac) for (int j = 0; j < 6; ++j) jac[6 * i + j] = rr[i].v[j]; }
}
// residual (1) and ambient Jacobian (1x6) of one VP block at x: unit-level parity hook
void orc_vp_residual(const double *x, const double *vp, const double *kvec, const double *qvec, double *res, double *jac) {
  LMBlock b;
  for (int i = 0; i < 4; ++i) { b.kvec[i] = kvec[i]; b.qvec[i] = qvec[i]; }
  for (int i = 0; i < 3; ++i) b.vp[i] = vp[i];
  b.has_vp = true;
  typedef Jet<6> J;
  J u[4] = {J(x[0], 0), J(x[1], 1), J(x[2], 2), J(x[3], 3)}, w[2] = {J(x[4], 4), J(x[5], 5)};
  J r = VPResidual<J>(b, u, w);
  res[0] = r.a;
  if (jac) for (int j = 0; j < 6; ++j) jac[j] = r.v[j];
}
void orc_minimal_from_line(const double *line, double *out6) {
  MinimalLine ml = MinimalFromLine3d(Line3d(V3(line[0], line[1], line[2]), V3(line[3], line[4], line[5])));
  for (int i = 0; i < 4; ++i) out6[i] = ml.uvec[i];
  out6[4] = ml.wvec[0]; out6[5] = ml.wvec[1];
}
// GetLineSegmentFromInfiniteLine3d(inf_line(x6), line3ds, num_outliers) (base/infinite_line.cc:265-287)
void orc_segment_from_minimal(const double *x6, const double *line3d, int64_t n, int num_outliers, double *out6) {
  MinimalLine ml;
  for (int i = 0; i < 4; ++i) ml.uvec[i] = x6[i];
  ml.wvec[0] = x6[4]; ml.wvec[1] = x6[5];
  V3 d, m;
  InfiniteFromMinimal(ml, d, m);
  std::vector<Line3d> ls;
  for (int64_t k = 0; k < n; ++k) ls.push_back(Line3d(V3(line3d[6 * k], line3d[6 * k + 1], line3d[6 * k + 2]),
                                                     V3(line3d[6 * k + 3], line3d[6 * k + 4], line3d[6 * k + 5])));
  Line3d r = SegmentFromInfinite(d, m, ls, num_outliers);
  out6[0] = r.start.x; out6[1] = r.start.y; out6[2] = r.start.z; out6[3] = r.end.x; out6[4] = r.end.y; out6[5] = r.end.z;
}
void orc_infinite_from_minimal(const double *x6, double *d3, double *m3) {
  MinimalLine ml;
  for (int i = 0; i < 4; ++i) ml.uvec[i] = x6[i];
  ml.wvec[0] = x6[4]; ml.wvec[1] = x6[5];
  V3 d, m;
  InfiniteFromMinimal(ml, d, m);
  d3[0] = d.x; d3[1] = d.y; d3[2] = d.z; m3[0] = m.x; m3[1] = m.y; m3[2] = m.z;
}

} // extern "C"
