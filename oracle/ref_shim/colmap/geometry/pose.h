// oracle/ref_shim: stands in for <colmap/geometry/pose.h> (TEST INFRASTRUCTURE).
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
