// oracle/orc_sfm.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE. The limap part (ranking loops, score formulas, ComputeRanges)
// is PINNED to the reference's compiled pointsfm/sfm_model.cc (oracle/_ref, tests/test_ref_pinning.py::
// test_sfm_model_neighbour_ranking_and_ranges); PARITY UNPINNED for the COLMAP part underneath it. The reference sorts with
// std::sort / std::partial_sort (unstable): among EQUAL scores its order is unspecified; this restatement keeps ascending
// image index (a stable sort over the std::map order).
//
// Visual-neighbour ranking and robust ranges of limap::pointsfm::SfmModel:
//   GetMaxIoUImages / GetMaxDiceCoeffImages   pointsfm/sfm_model.cc:101-162, 164-226
//   ComputeRanges / get_robust_range          pointsfm/sfm_model.cc:228-261
// The two inputs of the ranking come from COLMAP (pinned 1443d52 in cmake/FindDependencies.cmake:59-62, not in
// /root/reference) and are restated from its published source, colmap/mvs/model.cc:
//   ComputeSharedPoints: for every point, every unordered pair of images in its track counts one shared point
//     (both directions);
//   ComputeTriangulationAngles(percentile): for every point and pair of its images the triangulation angle between the
//     two viewing rays (law of cosines on baseline / ray lengths, folded to [0, pi/2]), as float; per image pair the
//     percentile = element round(p/100 (n-1)) of the sorted angles (colmap/math/math.h Percentile).
// Ties of the similarity score have no defined order in the reference (std::sort / std::partial_sort are not stable);
// here ties are broken by ascending image index, and the GPU path does the same.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

extern "C" {

// images: centres[n_images][3] (projection centres); points: xyz[n_points][3]; tracks: track_off[n_points+1],
// track_img[...] = image INDEX (0-based). mode 0 = IoU, 1 = Dice. out_neighbors[n_images][num_images] padded with -1,
// out_count[n_images].
int orc_rank_neighbors(int n_images, const double *centres, int64_t n_points, const double *xyz, const int64_t *track_off,
                       const int32_t *track_img, int num_images, double min_triangulation_angle_deg, int mode,
                       int32_t *out_neighbors, int32_t *out_count) {
  std::vector<int> num_points(n_images, 0);
  std::vector<std::map<int, int>> shared(n_images);
  std::vector<std::map<int, std::vector<float>>> angles(n_images);
  for (int64_t p = 0; p < n_points; ++p) {
    const double *X = xyz + 3 * p;
    for (int64_t a = track_off[p]; a < track_off[p + 1]; ++a) {
      const int i = track_img[a];
      num_points[i] += 1; // SfmModel::ComputeNumPoints (sfm_model.cc:59-68)
      for (int64_t b = track_off[p]; b < a; ++b) {
        const int j = track_img[b];
        if (i == j) continue;
        shared[i][j] += 1;
        shared[j][i] += 1;
        const double *c1 = centres + 3 * i, *c2 = centres + 3 * j;
        double bl2 = 0, r1 = 0, r2 = 0;
        for (int k = 0; k < 3; ++k) {
          bl2 += (c1[k] - c2[k]) * (c1[k] - c2[k]);
          r1 += (X[k] - c1[k]) * (X[k] - c1[k]);
          r2 += (X[k] - c2[k]) * (X[k] - c2[k]);
        }
        // CalculateTriangulationAngle (colmap/geometry/triangulation.cc)
        const double denom = 2.0 * std::sqrt(r1 * r2);
        double angle = 0.0;
        if (denom != 0.0) {
          const double nom = r1 + r2 - bl2;
          angle = std::abs(std::acos(nom / denom));
          angle = std::min(angle, M_PI - angle);
        }
        angles[i][j].push_back((float)angle);
        angles[j][i].push_back((float)angle);
      }
    }
  }
  const float min_angle = (float)(min_triangulation_angle_deg * M_PI / 180.0);
  for (int i = 0; i < n_images; ++i) {
    std::vector<std::pair<int, double>> ordered;
    for (const auto &kv : shared[i]) {
      std::vector<float> a = angles[i][kv.first];
      std::sort(a.begin(), a.end());
      const float perc = a[(size_t)std::lround(75.0 / 100.0 * (double)(a.size() - 1))];
      if (!(perc >= min_angle)) continue;
      const int inter = kv.second, uni = num_points[i] + num_points[kv.first] - inter;
      // mode 0: IoU (sfm_model.cc:130-133); 1: Dice (:194-196); 2: shared points, as colmap::mvs::Model::
      // GetMaxOverlappingImages, which GetMaxOverlapImages (:90-97) forwards to. COLMAP orders that list with an
      // UNSTABLE std::sort / partial_sort, so its tie order is unspecified; ties here keep ascending image index like
      // the two stable-sorted limap variants.
      const double score = (mode == 0) ? double(inter) / double(uni)
                           : (mode == 1) ? double(2 * inter) / double(uni + inter) : double(inter);
      ordered.emplace_back(kv.first, score);
    }
    std::stable_sort(ordered.begin(), ordered.end(),
                     [](const std::pair<int, double> &x, const std::pair<int, double> &y) { return x.second > y.second; });
    const int n = (int)std::min<size_t>(ordered.size(), (size_t)num_images);
    out_count[i] = n;
    for (int k = 0; k < num_images; ++k) out_neighbors[(int64_t)i * num_images + k] = (k < n) ? ordered[k].first : -1;
  }
  return 0;
}

// SfmModel::ComputeRanges (sfm_model.cc:245-261): per axis, float data sorted, data[size * q] at the two quantiles,
// stretched by kstretch * (hi - lo) on both sides. out[6] = lo3, hi3.
int orc_robust_ranges(int64_t n_points, const double *xyz, double q_lo, double q_hi, double kstretch, double *out) {
  for (int ax = 0; ax < 3; ++ax) {
    std::vector<float> d((size_t)n_points);
    for (int64_t p = 0; p < n_points; ++p) d[p] = (float)xyz[3 * p + ax];
    std::sort(d.begin(), d.end());
    const float kmin = (float)q_lo, kmax = (float)q_hi;
    float lo = d[(size_t)(d.size() * kmin)], hi = d[(size_t)(d.size() * kmax)];
    const float ks = (float)kstretch;
    const float diff = hi - lo;
    lo -= ks * diff;
    hi += ks * diff;
    out[ax] = lo;
    out[3 + ax] = hi;
  }
  return 0;
}

} // extern "C"
