// oracle/ref_shim: stands in for <colmap/geometry/sim3.h> (TEST INFRASTRUCTURE).
#pragma once
#include <Eigen/Geometry>
namespace colmap {
struct Sim3d {
  double scale = 1.0;
  Eigen::Quaterniond rotation = Eigen::Quaterniond::Identity();
  Eigen::Vector3d translation = Eigen::Vector3d::Zero();
};
} // namespace colmap
