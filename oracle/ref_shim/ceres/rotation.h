// oracle/ref_shim: stands in for <ceres/rotation.h> (TEST INFRASTRUCTURE): the four helpers the path calls, restated
// from Ceres' header (QuaternionToRotation normalises by |q|^2; QuaternionRotatePoint normalises q first).
#pragma once
#include "ceres.h"
namespace ceres {
template <typename T> inline void QuaternionToRotation(const T q[4], T R[3 * 3]) { // row-major
  const T a = q[0], b = q[1], c = q[2], d = q[3];
  const T aa = a * a, ab = a * b, ac = a * c, ad = a * d, bb = b * b, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  R[0] = aa + bb - cc - dd; R[1] = T(2) * (bc - ad); R[2] = T(2) * (ac + bd);
  R[3] = T(2) * (ad + bc); R[4] = aa - bb + cc - dd; R[5] = T(2) * (cd - ab);
  R[6] = T(2) * (bd - ac); R[7] = T(2) * (ab + cd); R[8] = aa - bb - cc + dd;
  const T normalizer = T(1) / (aa + bb + cc + dd);
  for (int i = 0; i < 9; ++i) R[i] *= normalizer;
}
template <typename T> inline void UnitQuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  T uv0 = q[2] * pt[2] - q[3] * pt[1];
  T uv1 = q[3] * pt[0] - q[1] * pt[2];
  T uv2 = q[1] * pt[1] - q[2] * pt[0];
  uv0 += uv0; uv1 += uv1; uv2 += uv2;
  result[0] = pt[0] + q[0] * uv0;
  result[1] = pt[1] + q[0] * uv1;
  result[2] = pt[2] + q[0] * uv2;
  result[0] += q[2] * uv2 - q[3] * uv1;
  result[1] += q[3] * uv0 - q[1] * uv2;
  result[2] += q[1] * uv1 - q[2] * uv0;
}
template <typename T> inline void QuaternionRotatePoint(const T q[4], const T pt[3], T result[3]) {
  const T scale = T(1) / sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  const T unit[4] = {scale * q[0], scale * q[1], scale * q[2], scale * q[3]};
  UnitQuaternionRotatePoint(unit, pt, result);
}
template <typename T> inline void CrossProduct(const T x[3], const T y[3], T out[3]) {
  out[0] = x[1] * y[2] - x[2] * y[1];
  out[1] = x[2] * y[0] - x[0] * y[2];
  out[2] = x[0] * y[1] - x[1] * y[0];
}
template <typename T> inline T DotProduct(const T x[3], const T y[3]) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; }
} // namespace ceres
