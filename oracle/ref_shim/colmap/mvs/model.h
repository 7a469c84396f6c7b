// oracle/ref_shim: stands in for <colmap/mvs/model.h> (TEST INFRASTRUCTURE). The part of colmap::mvs::Model that
// limap::pointsfm::SfmModel builds on -- points with float coordinates and image-index tracks, ComputeSharedPoints,
// ComputeTriangulationAngles (projection centres in float, angles stored as float, percentile per image pair) and
// GetMaxOverlappingImages -- restated from COLMAP's published mvs/model.cc and geometry/triangulation.cc
// (CalculateTriangulationAngle). Reading COLMAP folders is not part of the path: ReadFromCOLMAP throws.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include <colmap/math/math.h>
#include <colmap/mvs/image.h>

namespace colmap {
namespace mvs {

struct Model {
  struct Point {
    float x = 0, y = 0, z = 0;
    std::vector<int> track;
  };
  std::vector<Image> images;
  std::vector<Point> points;

  void ReadFromCOLMAP(const std::string &, const std::string & = "sparse", const std::string & = "images") {
    throw std::runtime_error("ref_shim: reading COLMAP folders is outside the hot path");
  }
  std::string GetImageName(const int image_idx) const { return images.at(image_idx).GetPath(); }

  std::vector<std::map<int, int>> ComputeSharedPoints() const {
    std::vector<std::map<int, int>> shared_points(images.size());
    for (const auto &point : points)
      for (size_t i = 0; i < point.track.size(); ++i) {
        const int image_idx1 = point.track[i];
        for (size_t j = 0; j < i; ++j) {
          const int image_idx2 = point.track[j];
          if (image_idx1 != image_idx2) {
            shared_points.at(image_idx1)[image_idx2] += 1;
            shared_points.at(image_idx2)[image_idx1] += 1;
          }
        }
      }
    return shared_points;
  }

  static double CalculateTriangulationAngle(const double c1[3], const double c2[3], const double X[3]) {
    double bl2 = 0, r1 = 0, r2 = 0;
    for (int k = 0; k < 3; ++k) {
      bl2 += (c1[k] - c2[k]) * (c1[k] - c2[k]);
      r1 += (X[k] - c1[k]) * (X[k] - c1[k]);
      r2 += (X[k] - c2[k]) * (X[k] - c2[k]);
    }
    const double denominator = 2.0 * std::sqrt(r1 * r2);
    if (denominator == 0.0) return 0.0;
    const double nominator = r1 + r2 - bl2;
    const double angle = std::abs(std::acos(nominator / denominator));
    return std::min(angle, M_PI - angle);
  }

  std::vector<std::map<int, float>> ComputeTriangulationAngles(const float percentile = 50) const {
    // projection centres: C = -R^T T in float (mvs/image.cc ComputeProjectionCenter), then widened
    std::vector<double> proj_centers(3 * images.size());
    for (size_t image_idx = 0; image_idx < images.size(); ++image_idx) {
      const float *R = images[image_idx].GetR(), *T = images[image_idx].GetT();
      for (int i = 0; i < 3; ++i) {
        const float c = -(R[0 * 3 + i] * T[0] + R[1 * 3 + i] * T[1] + R[2 * 3 + i] * T[2]);
        proj_centers[3 * image_idx + i] = (double)c;
      }
    }
    std::vector<std::map<int, std::vector<float>>> all_triangulation_angles(images.size());
    for (const auto &point : points) {
      const double X[3] = {(double)point.x, (double)point.y, (double)point.z};
      for (size_t i = 0; i < point.track.size(); ++i) {
        const int image_idx1 = point.track[i];
        for (size_t j = 0; j < i; ++j) {
          const int image_idx2 = point.track[j];
          if (image_idx1 != image_idx2) {
            const float angle = (float)CalculateTriangulationAngle(&proj_centers[3 * image_idx1], &proj_centers[3 * image_idx2], X);
            all_triangulation_angles.at(image_idx1)[image_idx2].push_back(angle);
            all_triangulation_angles.at(image_idx2)[image_idx1].push_back(angle);
          }
        }
      }
    }
    std::vector<std::map<int, float>> triangulation_angles(images.size());
    for (size_t image_idx1 = 0; image_idx1 < all_triangulation_angles.size(); ++image_idx1)
      for (const auto &data : all_triangulation_angles[image_idx1])
        triangulation_angles.at(image_idx1).emplace(data.first, Percentile(data.second, percentile));
    return triangulation_angles;
  }

  std::vector<std::vector<int>> GetMaxOverlappingImages(const size_t num_images, const double min_triangulation_angle) const {
    std::vector<std::vector<int>> overlapping_images(images.size());
    const float min_triangulation_angle_rad = DegToRad(min_triangulation_angle);
    const auto shared_num_points = ComputeSharedPoints();
    const float kTriangulationAnglePercentile = 75;
    const auto triangulation_angles = ComputeTriangulationAngles(kTriangulationAnglePercentile);
    for (size_t image_idx = 0; image_idx < images.size(); ++image_idx) {
      const auto &shared_images = shared_num_points.at(image_idx);
      const auto &overlapping_triangulation_angles = triangulation_angles.at(image_idx);
      std::vector<std::pair<int, int>> ordered_images;
      ordered_images.reserve(shared_images.size());
      for (const auto &image : shared_images)
        if (overlapping_triangulation_angles.at(image.first) >= min_triangulation_angle_rad)
          ordered_images.emplace_back(image.first, image.second);
      const size_t eff_num_images = std::min(ordered_images.size(), num_images);
      auto cmp = [](const std::pair<int, int> image1, const std::pair<int, int> image2) { return image1.second > image2.second; };
      if (eff_num_images < shared_images.size())
        std::partial_sort(ordered_images.begin(), ordered_images.begin() + eff_num_images, ordered_images.end(), cmp);
      else
        std::sort(ordered_images.begin(), ordered_images.end(), cmp);
      overlapping_images[image_idx].reserve(eff_num_images);
      for (size_t i = 0; i < eff_num_images; ++i) overlapping_images[image_idx].push_back(ordered_images[i].first);
    }
    return overlapping_images;
  }
};

} // namespace mvs
} // namespace colmap
