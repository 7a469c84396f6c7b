// oracle/ref_shim: stands in for <colmap/scene/camera.h> (TEST INFRASTRUCTURE): the struct limap::Camera derives
// from, with the members and accessors the hot path calls, for the two undistorted models.
#pragma once
#include <sstream>
#include <colmap/sensor/models.h>
#include <colmap/util/types.h>
namespace colmap {
struct Camera {
  camera_t camera_id = kInvalidCameraId;
  CameraModelId model_id = CameraModelId::kInvalid;
  size_t width = 0, height = 0;
  std::vector<double> params;
  bool has_prior_focal_length = false;
  bool two_focal() const {
    return model_id == CameraModelId::kPinhole || model_id == CameraModelId::kOpenCV || model_id == CameraModelId::kOpenCVFisheye ||
           model_id == CameraModelId::kFullOpenCV || model_id == CameraModelId::kFOV || model_id == CameraModelId::kThinPrismFisheye;
  }
  std::vector<size_t> FocalLengthIdxs() const { return two_focal() ? std::vector<size_t>{0, 1} : std::vector<size_t>{0}; }
  std::vector<size_t> PrincipalPointIdxs() const { return two_focal() ? std::vector<size_t>{2, 3} : std::vector<size_t>{1, 2}; }
  std::vector<size_t> ExtraParamsIdxs() const {
    std::vector<size_t> out;
    for (size_t i = two_focal() ? 4 : 3; i < params.size(); ++i) out.push_back(i);
    return out;
  }
  double FocalLength() const { return params[0]; }
  double FocalLengthX() const { return params[0]; }
  double FocalLengthY() const { return two_focal() ? params[1] : params[0]; }
  double PrincipalPointX() const { return two_focal() ? params[2] : params[1]; }
  double PrincipalPointY() const { return two_focal() ? params[3] : params[2]; }
  double MeanFocalLength() const { return (FocalLengthX() + FocalLengthY()) / 2.0; }
  Eigen::Matrix3d CalibrationMatrix() const {
    Eigen::Matrix3d K = Eigen::Matrix3d::Identity();
    K(0, 0) = FocalLengthX(); K(1, 1) = FocalLengthY(); K(0, 2) = PrincipalPointX(); K(1, 2) = PrincipalPointY();
    return K;
  }
  std::string ModelName() const { return CameraModelIdToName(model_id); }
  std::string ParamsToString() const { std::ostringstream os; for (size_t i = 0; i < params.size(); ++i) os << (i ? ", " : "") << params[i]; return os.str(); }
  bool VerifyParams() const { return ExistsCameraModelWithId(model_id) && params.size() == CameraModelNumParams(model_id); }
  bool IsUndistorted() const { return model_id == CameraModelId::kSimplePinhole || model_id == CameraModelId::kPinhole; }
  void Rescale(size_t new_width, size_t new_height) {
    const double sx = new_width / static_cast<double>(width), sy = new_height / static_cast<double>(height);
    width = new_width; height = new_height;
    if (two_focal()) { params[0] *= sx; params[1] *= sy; params[2] *= sx; params[3] *= sy; }
    else { params[0] *= (sx + sy) / 2.0; params[1] *= sx; params[2] *= sy; }
  }
};
} // namespace colmap
