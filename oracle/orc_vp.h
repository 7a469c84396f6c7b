// oracle/orc_vp.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE. The reference-side wrapper (length filter, cluster filtering,
// label renumbering, VP fit) is PINNED to the reference's compiled vplib/JLinkage/JLinkage.cc + base_vp_detector.cc
// (oracle/_ref over oracle/ref_shim/JLinkage, tests/test_ref_pinning.py). The library core is PARITY UNPINNED:
// the J-Linkage arithmetic lives in a third-party library that is NOT under /root/reference
// (B1ueber2y/JLinkage @ 75dadd555f81b1cf1b0f016d8cac76f3b554ba9b, cmake/FindDependencies.cmake:73-77) and
// draws its 5000 minimal samples from an unseeded RNG, so not even the reference reproduces itself.
// This file restates the published algorithm (Toldo & Fusiello, "Robust multiple structures estimation with
// J-Linkage", ECCV 2008; Tardif, "Non-iterative approach for fast and accurate vanishing point detection",
// ICCV 2009) with the constants of the reference's call site, vplib/JLinkage/JLinkage.cc:44-46
// (VPSample::run(&pts, 5000, 2, 0, 3): 5000 models from minimal sets of 2 segments, uniform sampling;
// VPCluster::run(..., inlier_threshold, 2)), float arithmetic as in the library (JLinkage.cc:27-36), and a
// counter-based RNG so that oracle and GPU kernel agree bit for bit. The reference-side post-processing
// (JLinkage.cc:14-127, base_vp_detector.cc:41-73) is restated line by line.
//
// Specification shared with limap_b200/csrc/vp_kernels.cu (all float ops un-contracted, IEEE rn):
//   valid lines: length >= min_length (fp64), coordinates cast to float.
//   fewer than 2*max(min_num_supports,10) valid lines -> all labels -1.
//   model m: z = splitmix64(seed, image_index*M + m); i = lo32(z) % n; j = hi32(z) % (n-1); j += (j >= i);
//            l = (y1-y2, x2-x1, x1*y2 - x2*y1); vp = l_i x l_j, normalised if its norm is > 0.
//   consensus(p, m): mid = (p1+p2)*0.5; l = (mid,1) x vp; d = |l.(p1,1)| / sqrt(lx^2+ly^2); inlier iff d < th.
//   J-Linkage: singleton clusters; merge the pair (a<b) with the smallest Jaccard distance
//            1 - |PSa & PSb| / |PSa | PSb| (compared as integer fractions, ties: smallest a then b) while
//            some pair has a non-empty intersection; PS of a cluster = intersection of its members' PS.
//   labels: clusters numbered by their smallest member index, ascending.
#pragma once
#include "orc_geom.h"

namespace orc {

inline uint64_t splitmix64(uint64_t seed, uint64_t counter) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (counter + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

struct VPConfig {
  double min_length = 40, inlier_threshold = 1.0, th_perp_supports = 3.0;
  int min_num_supports = 5, n_models = 5000;
  uint64_t seed = 0;
};

struct F3 { float x, y, z; };
inline F3 line_coords_f(const float p[4]) {
  F3 l;
  l.x = p[1] - p[3];
  l.y = p[2] - p[0];
  const float a = p[0] * p[3], b = p[2] * p[1];
  l.z = a - b;
  return l;
}
inline F3 cross_f(const F3 &a, const F3 &b) {
  F3 c;
  { const float u = a.y * b.z, v = a.z * b.y; c.x = u - v; }
  { const float u = a.z * b.x, v = a.x * b.z; c.y = u - v; }
  { const float u = a.x * b.y, v = a.y * b.x; c.z = u - v; }
  return c;
}
inline bool vp_consensus(const float p[4], const F3 &vp, float th) {
  const float mx = (p[0] + p[2]) * 0.5f, my = (p[1] + p[3]) * 0.5f;
  F3 m{mx, my, 1.0f};
  const F3 l = cross_f(m, vp);
  const float xx = l.x * l.x, yy = l.y * l.y;
  const float den = std::sqrt(xx + yy);
  const float t0 = l.x * p[0], t1 = l.y * p[1];
  const float num = std::fabs((t0 + t1) + l.z);
  const float d = num / den;
  return d < th;
}

// J-Linkage labels of the valid lines of one image; returns the number of clusters.
inline int jlinkage_cluster(const std::vector<std::array<float, 4>> &pts, const VPConfig &cfg, uint64_t image_index,
                            std::vector<int> &labels) {
  const int n = (int)pts.size(), M = cfg.n_models, W = (M + 31) / 32;
  labels.assign(n, -1);
  std::vector<F3> models(M);
  for (int m = 0; m < M; ++m) {
    const uint64_t z = splitmix64(cfg.seed, image_index * (uint64_t)M + (uint64_t)m);
    int i = (int)((uint32_t)(z & 0xffffffffu) % (uint32_t)n);
    int j = (int)((uint32_t)(z >> 32) % (uint32_t)(n - 1));
    if (j >= i) ++j;
    F3 vp = cross_f(line_coords_f(pts[i].data()), line_coords_f(pts[j].data()));
    const float xx = vp.x * vp.x, yy = vp.y * vp.y, zz = vp.z * vp.z;
    const float nr = std::sqrt((xx + yy) + zz);
    if (nr > 0.0f) { vp.x = vp.x / nr; vp.y = vp.y / nr; vp.z = vp.z / nr; }
    models[m] = vp;
  }
  std::vector<uint32_t> ps((size_t)n * W, 0u);
  const float th = (float)cfg.inlier_threshold;
  for (int p = 0; p < n; ++p)
    for (int m = 0; m < M; ++m)
      if (vp_consensus(pts[p].data(), models[m], th)) ps[(size_t)p * W + (m >> 5)] |= 1u << (m & 31);
  std::vector<char> active(n, 1);
  std::vector<int> rep(n);
  for (int i = 0; i < n; ++i) rep[i] = i; // cluster representative (smallest member) of every line
  auto frac = [&](int a, int b, long long &inter, long long &uni) {
    inter = uni = 0;
    for (int w = 0; w < W; ++w) {
      const uint32_t x = ps[(size_t)a * W + w], y = ps[(size_t)b * W + w];
      inter += __builtin_popcount(x & y);
      uni += __builtin_popcount(x | y);
    }
  };
  while (true) {
    int ba = -1, bb = -1;
    long long bi = 0, bu = 1;
    for (int a = 0; a < n; ++a) {
      if (!active[a]) continue;
      for (int b = a + 1; b < n; ++b) {
        if (!active[b]) continue;
        long long in_, un_;
        frac(a, b, in_, un_);
        if (in_ == 0) continue;
        // smaller distance <=> larger in/un ; strict > keeps the first (smallest a, then b) on ties
        if (ba < 0 || in_ * bu > bi * un_) { ba = a; bb = b; bi = in_; bu = un_; }
      }
    }
    if (ba < 0) break;
    for (int w = 0; w < W; ++w) ps[(size_t)ba * W + w] &= ps[(size_t)bb * W + w];
    active[bb] = 0;
    for (int i = 0; i < n; ++i) if (rep[i] == bb) rep[i] = ba;
  }
  std::vector<int> cid(n, -1);
  int nc = 0;
  for (int i = 0; i < n; ++i) if (active[i]) cid[i] = nc++;
  for (int i = 0; i < n; ++i) labels[i] = cid[rep[i]];
  return nc;
}

// BaseVPDetector::count_valid_supports_2d (vplib/base_vp_detector.cc:41-73)
inline int count_valid_supports_2d(const std::vector<Line2d> &lines, double th_perp) {
  const size_t n = lines.size();
  std::vector<int> parent(n, -1);
  auto root = [&](size_t i) { while (parent[i] != -1) i = parent[i]; return i; };
  auto dist = [&](const Line2d &l, const V2 &q) { // InfiniteLine2d(l).point_distance(q) (base/infinite_line.cc:19-34)
    V3 c = l.coords();
    return std::abs(c.x * q.x + c.y * q.y + c.z) / std::sqrt(c.x * c.x + c.y * c.y);
  };
  for (size_t i = 0; i + 1 < n; ++i) {
    size_t ri = root(i);
    for (size_t j = i + 1; j < n; ++j) {
      size_t rj = root(j);
      if (rj == ri) continue;
      size_t k1 = i, k2 = j;
      if (lines[i].length() > lines[j].length()) { k1 = j; k2 = i; }
      double d = smax(dist(lines[k2], lines[k1].start), dist(lines[k2], lines[k1].end));
      if (d > th_perp) continue;
      parent[rj] = (int)ri;
    }
  }
  int cnt = 0;
  for (size_t i = 0; i < n; ++i) if (parent[i] == -1) ++cnt;
  return cnt;
}

// smallest-eigenvalue eigenvector of a symmetric 3x3 (JLinkage::fitVP, JLinkage.cc:86-100: V.col(2) of the SVD)
inline V3 smallest_eigvec_sym3(const double Ain[3][3]) {
  double A[3][3], Vv[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = Ain[i][j];
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
    if (off == 0 || off <= 1e-32 * diag) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (A[p][q] == 0) continue;
        double theta = (A[q][q] - A[p][p]) / (2 * A[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) { double a = A[k][p], b = A[k][q]; A[k][p] = c * a - s * b; A[k][q] = s * a + c * b; }
        for (int k = 0; k < 3; ++k) { double a = A[p][k], b = A[q][k]; A[p][k] = c * a - s * b; A[q][k] = s * a + c * b; }
        for (int k = 0; k < 3; ++k) { double a = Vv[k][p], b = Vv[k][q]; Vv[k][p] = c * a - s * b; Vv[k][q] = s * a + c * b; }
      }
  }
  int best = 0;
  if (A[1][1] < A[best][best]) best = 1;
  if (A[2][2] < A[best][best]) best = 2;
  return V3(Vv[0][best], Vv[1][best], Vv[2][best]).normalized();
}

// JLinkage::AssociateVPs (JLinkage.cc:102-127) given the raw cluster labels of the valid lines.
// `raw` may come from jlinkage_cluster (oracle) -- the post-processing is the reference's.
inline void associate_vps(const std::vector<Line2d> &lines, const std::vector<int> &valid_ids,
                          const std::vector<int> &raw, int n_clusters, const VPConfig &cfg,
                          std::vector<int> &final_labels, std::vector<V3> &vps) {
  final_labels.assign(lines.size(), -1);
  vps.clear();
  std::vector<std::vector<Line2d>> all_supports(n_clusters);
  for (size_t i = 0; i < valid_ids.size(); ++i) if (raw[i] >= 0) all_supports[raw[i]].push_back(lines[valid_ids[i]]);
  std::vector<int> vp_ids(n_clusters, -1);
  int counter = 0;
  for (int c = 0; c < n_clusters; ++c) {
    const auto &sup = all_supports[c];
    if ((int)sup.size() < cfg.min_num_supports) continue;
    if (count_valid_supports_2d(sup, cfg.th_perp_supports) < cfg.min_num_supports) continue;
    vp_ids[c] = counter++;
  }
  for (size_t i = 0; i < valid_ids.size(); ++i)
    if (raw[i] >= 0 && vp_ids[raw[i]] >= 0) final_labels[valid_ids[i]] = vp_ids[raw[i]];
  if (counter == 0) return;
  vps.resize(counter);
  std::vector<std::vector<Line2d>> supports(counter);
  for (size_t i = 0; i < lines.size(); ++i) if (final_labels[i] >= 0) supports[final_labels[i]].push_back(lines[i]);
  for (int v = 0; v < counter; ++v) { // fitVP
    double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (const Line2d &l : supports[v]) {
      V3 c = l.coords();
      for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] += c[a] * c[b];
    }
    vps[v] = smallest_eigvec_sym3(S);
  }
}

// JLinkage::ComputeVPLabels + AssociateVPs for one image
inline void detect_vp_image(const std::vector<Line2d> &lines, const VPConfig &cfg, uint64_t image_index,
                            std::vector<int> &labels, std::vector<V3> &vps) {
  labels.assign(lines.size(), -1);
  vps.clear();
  if (lines.empty()) return;
  std::vector<int> valid_ids;
  std::vector<std::array<float, 4>> pts;
  for (size_t i = 0; i < lines.size(); ++i) {
    if (lines[i].length() < cfg.min_length) continue;
    valid_ids.push_back((int)i);
    pts.push_back({(float)lines[i].start.x, (float)lines[i].start.y, (float)lines[i].end.x, (float)lines[i].end.y});
  }
  if ((int)pts.size() < 2 * std::max(cfg.min_num_supports, 10)) return;
  std::vector<int> raw;
  const int nc = jlinkage_cluster(pts, cfg, image_index, raw);
  associate_vps(lines, valid_ids, raw, nc, cfg, labels, vps);
}

} // namespace orc
