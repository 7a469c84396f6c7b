// oracle/orc_vp.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE. Pinning status: see orc_vp.h (wrapper pinned, library core unpinned).
// Built with -ffp-contract=off so the float arithmetic of the consensus test matches the GPU kernel bit for bit.
#include <array>
#include "orc_vp.h"

using namespace orc;

extern "C" {

struct orc_vp_cfg {
  double min_length, inlier_threshold, th_perp_supports;
  int32_t min_num_supports, n_models;
  uint64_t seed;
};

// labels[sum L] out; vps_out[cap][3]; vp_off[n_images+1] out. Returns total number of VPs (may exceed cap).
// image_index[im] (may be NULL = im): the index that seeds the hypotheses of image im, so that a call on a subset of
// the images (one rank's share) draws what the call on all images draws for them.
long long orc_vp_detect_indexed(int n_images, const int64_t *line_off, const double *segs, const orc_vp_cfg *c,
                                const int64_t *image_index, int32_t *labels, int64_t *vp_off, double *vps_out, long long cap) {
  VPConfig cfg;
  cfg.min_length = c->min_length; cfg.inlier_threshold = c->inlier_threshold; cfg.th_perp_supports = c->th_perp_supports;
  cfg.min_num_supports = c->min_num_supports; cfg.n_models = c->n_models; cfg.seed = c->seed;
  std::vector<std::vector<int>> all_labels(n_images);
  std::vector<std::vector<V3>> all_vps(n_images);
#pragma omp parallel for schedule(dynamic, 1)
  for (int im = 0; im < n_images; ++im) {
    std::vector<Line2d> lines;
    for (int64_t l = line_off[im]; l < line_off[im + 1]; ++l)
      lines.push_back(Line2d(V2(segs[4 * l], segs[4 * l + 1]), V2(segs[4 * l + 2], segs[4 * l + 3])));
    detect_vp_image(lines, cfg, image_index ? (uint64_t)image_index[im] : (uint64_t)im, all_labels[im], all_vps[im]);
  }
  long long n = 0;
  for (int im = 0; im < n_images; ++im) {
    vp_off[im] = n;
    for (int64_t l = line_off[im]; l < line_off[im + 1]; ++l) labels[l] = all_labels[im][l - line_off[im]];
    for (const V3 &v : all_vps[im]) {
      if (n < cap) { vps_out[3 * n] = v.x; vps_out[3 * n + 1] = v.y; vps_out[3 * n + 2] = v.z; }
      ++n;
    }
  }
  vp_off[n_images] = n;
  return n;
}

long long orc_vp_detect(int n_images, const int64_t *line_off, const double *segs, const orc_vp_cfg *c,
                        int32_t *labels, int64_t *vp_off, double *vps_out, long long cap) {
  return orc_vp_detect_indexed(n_images, line_off, segs, c, nullptr, labels, vp_off, vps_out, cap);
}

} // extern "C"
