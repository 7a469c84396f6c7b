// oracle/ref_shim: stands in for <colmap/util/types.h> (TEST INFRASTRUCTURE). Only the id typedefs limap uses.
#pragma once
#include <cstdint>
#include <limits>
#include <Eigen/Core>
namespace colmap {
typedef uint32_t camera_t;
typedef uint32_t image_t;
typedef uint64_t image_pair_t;
typedef uint32_t point2D_t;
typedef uint64_t point3D_t;
const camera_t kInvalidCameraId = std::numeric_limits<camera_t>::max();
const image_t kInvalidImageId = std::numeric_limits<image_t>::max();
const point2D_t kInvalidPoint2DIdx = std::numeric_limits<point2D_t>::max();
const point3D_t kInvalidPoint3DId = std::numeric_limits<point3D_t>::max();
} // namespace colmap
