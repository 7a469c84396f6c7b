// oracle/ref_shim: stands in for <colmap/scene/reconstruction.h> (TEST INFRASTRUCTURE): SfmModel::ReadFromCOLMAP names the
// type; reading COLMAP folders is outside the hot path, so Read throws.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
namespace colmap {
class Reconstruction {
public:
  void Read(const std::string &) { throw std::runtime_error("ref_shim: reading COLMAP folders is outside the hot path"); }
  std::vector<uint32_t> RegImageIds() const { return {}; }
};
} // namespace colmap
