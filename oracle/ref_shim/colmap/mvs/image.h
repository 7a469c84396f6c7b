// oracle/ref_shim: stands in for <colmap/mvs/image.h> (TEST INFRASTRUCTURE). COLMAP is not in this image; what
// limap::pointsfm::SfmModel needs of colmap::mvs::Image is restated from COLMAP's public header: the constructor that
// CreateSfmImage calls (K, R, T copied as float, row-major) and the getters.
#pragma once
#include <cstring>
#include <string>

namespace colmap {
namespace mvs {

class Image {
public:
  Image() {}
  Image(const std::string &path, const size_t width, const size_t height, const float *K, const float *R, const float *T)
      : path_(path), width_(width), height_(height) {
    std::memcpy(K_, K, 9 * sizeof(float));
    std::memcpy(R_, R, 9 * sizeof(float));
    std::memcpy(T_, T, 3 * sizeof(float));
  }
  size_t GetWidth() const { return width_; }
  size_t GetHeight() const { return height_; }
  const std::string &GetPath() const { return path_; }
  const float *GetR() const { return R_; }
  const float *GetT() const { return T_; }
  const float *GetK() const { return K_; }

private:
  std::string path_;
  size_t width_ = 0, height_ = 0;
  float K_[9] = {0}, R_[9] = {0}, T_[3] = {0};
};

} // namespace mvs
} // namespace colmap
