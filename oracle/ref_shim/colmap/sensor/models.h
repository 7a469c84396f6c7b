// oracle/ref_shim: stands in for <colmap/sensor/models.h> (TEST INFRASTRUCTURE). COLMAP's camera-model registry
// reduced to the facts limap's hot path reads: the model ids / names and the parameter layout of the two pinhole
// models (SIMPLE_PINHOLE: f, cx, cy; PINHOLE: fx, fy, cx, cy); other ids are known by name only.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
namespace colmap {
enum class CameraModelId {
  kInvalid = -1, kSimplePinhole = 0, kPinhole = 1, kSimpleRadial = 2, kRadial = 3, kOpenCV = 4, kOpenCVFisheye = 5,
  kFullOpenCV = 6, kFOV = 7, kSimpleRadialFisheye = 8, kRadialFisheye = 9, kThinPrismFisheye = 10
};
struct SimplePinholeCameraModel { static constexpr CameraModelId model_id = CameraModelId::kSimplePinhole; static constexpr size_t num_params = 3; };
struct PinholeCameraModel { static constexpr CameraModelId model_id = CameraModelId::kPinhole; static constexpr size_t num_params = 4; };
inline const char *const *shim_model_names() {
  static const char *const n[] = {"SIMPLE_PINHOLE", "PINHOLE", "SIMPLE_RADIAL", "RADIAL", "OPENCV", "OPENCV_FISHEYE",
                                  "FULL_OPENCV", "FOV", "SIMPLE_RADIAL_FISHEYE", "RADIAL_FISHEYE", "THIN_PRISM_FISHEYE"};
  return n;
}
inline CameraModelId CameraModelNameToId(const std::string &name) {
  for (int i = 0; i <= 10; ++i) if (name == shim_model_names()[i]) return static_cast<CameraModelId>(i);
  return CameraModelId::kInvalid;
}
inline std::string CameraModelIdToName(CameraModelId id) {
  const int i = static_cast<int>(id);
  return (i >= 0 && i <= 10) ? shim_model_names()[i] : "";
}
inline bool ExistsCameraModelWithName(const std::string &name) { return CameraModelNameToId(name) != CameraModelId::kInvalid; }
inline bool ExistsCameraModelWithId(CameraModelId id) { const int i = static_cast<int>(id); return i >= 0 && i <= 10; }
inline size_t CameraModelNumParams(CameraModelId id) {
  static const size_t n[] = {3, 4, 4, 5, 8, 8, 12, 5, 4, 5, 12};
  const int i = static_cast<int>(id);
  if (i < 0 || i > 10) throw std::domain_error("Camera model does not exist");
  return n[i];
}
inline std::vector<double> CameraModelInitializeParams(CameraModelId id, double f, size_t w, size_t h) {
  std::vector<double> p(CameraModelNumParams(id), 0.0);
  if (id == CameraModelId::kPinhole || id == CameraModelId::kOpenCV || id == CameraModelId::kOpenCVFisheye ||
      id == CameraModelId::kFullOpenCV || id == CameraModelId::kFOV || id == CameraModelId::kThinPrismFisheye) {
    p[0] = f; p[1] = f; p[2] = w / 2.0; p[3] = h / 2.0;
  } else { p[0] = f; p[1] = w / 2.0; p[2] = h / 2.0; }
  return p;
}
} // namespace colmap
