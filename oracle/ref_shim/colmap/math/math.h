// oracle/ref_shim: stands in for <colmap/math/math.h> (TEST INFRASTRUCTURE): DegToRad and Percentile restated from
// COLMAP's header (Percentile: element round(p / 100 * (n - 1)) of the ordered elements).
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>

namespace colmap {

inline float DegToRad(const float deg) { return deg * 0.0174532925199432954743716805978692718781530857086181640625f; }
inline double DegToRad(const double deg) { return deg * 0.0174532925199432954743716805978692718781530857086181640625; }

template <typename T> T Percentile(const std::vector<T> &elems, const double p) {
  const int idx = static_cast<int>(std::round(p / 100 * (elems.size() - 1)));
  const size_t percentile_idx = std::max(0, std::min(static_cast<int>(elems.size() - 1), idx));
  std::vector<T> ordered_elems = elems;
  std::nth_element(ordered_elems.begin(), ordered_elems.begin() + percentile_idx, ordered_elems.end());
  return ordered_elems.at(percentile_idx);
}

} // namespace colmap
