// oracle/orc_geom.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// fp64 CPU restatement of the geometry primitives of cvg/limap that sit on the
// line-triangulation / line-refinement hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// PARITY PINNED to the reference's own compiled code: the reference's build system cannot run in this image (Eigen,
// Ceres, COLMAP, JLinkage absent), but its hot-path sources compile unchanged against the header shims in
// oracle/ref_shim/ (oracle/Makefile target `ref` -> oracle/_ref/liblimap_ref.so), and tests/test_ref_pinning.py holds
// every function of this restatement to that library on seeded inputs. The reference's own tests hold a
// single known-answer vector for this path (tests/base/test_linebase.py:8-17), checked in tests/test_oracle_kat.py.
// Written function-by-function from the files cited below (paths relative to /root/reference/src/limap/).
//
// No Eigen: the few Eigen operations the reference relies on are restated with
// the same evaluation structure (normalized(), 3x3 cofactor inverse,
// quaternion -> rotation matrix).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace orc {

// util/types.h:34
static const double EPS = 1e-12;

struct V2 {
  double x = 0, y = 0;
  V2() {}
  V2(double x_, double y_) : x(x_), y(y_) {}
  V2 operator+(const V2 &o) const { return V2(x + o.x, y + o.y); }
  V2 operator-(const V2 &o) const { return V2(x - o.x, y - o.y); }
  V2 operator*(double s) const { return V2(x * s, y * s); }
  V2 operator/(double s) const { return V2(x / s, y / s); }
  double dot(const V2 &o) const { return x * o.x + y * o.y; }
  double squaredNorm() const { return x * x + y * y; }
  double norm() const { return std::sqrt(squaredNorm()); }
  // Eigen normalized(): divide by the norm only when it is > 0.
  V2 normalized() const {
    double z = squaredNorm();
    if (z > 0)
      return *this / std::sqrt(z);
    return *this;
  }
};

struct V3 {
  double x = 0, y = 0, z = 0;
  V3() {}
  V3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
  V3 operator+(const V3 &o) const { return V3(x + o.x, y + o.y, z + o.z); }
  V3 operator-(const V3 &o) const { return V3(x - o.x, y - o.y, z - o.z); }
  V3 operator*(double s) const { return V3(x * s, y * s, z * s); }
  V3 operator/(double s) const { return V3(x / s, y / s, z / s); }
  V3 operator-() const { return V3(-x, -y, -z); }
  double dot(const V3 &o) const { return x * o.x + y * o.y + z * o.z; }
  V3 cross(const V3 &o) const {
    return V3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
  }
  double squaredNorm() const { return x * x + y * y + z * z; }
  double norm() const { return std::sqrt(squaredNorm()); }
  V3 normalized() const {
    double n = squaredNorm();
    if (n > 0)
      return *this / std::sqrt(n);
    return *this;
  }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

struct M3 {
  double m[3][3];
  M3() {
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        m[i][j] = 0;
  }
  static M3 fromCols(const V3 &a, const V3 &b, const V3 &c) {
    M3 r;
    r.m[0][0] = a.x; r.m[1][0] = a.y; r.m[2][0] = a.z;
    r.m[0][1] = b.x; r.m[1][1] = b.y; r.m[2][1] = b.z;
    r.m[0][2] = c.x; r.m[1][2] = c.y; r.m[2][2] = c.z;
    return r;
  }
  M3 transpose() const {
    M3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        r.m[i][j] = m[j][i];
    return r;
  }
  M3 operator*(const M3 &o) const {
    M3 r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        r.m[i][j] = m[i][0] * o.m[0][j] + m[i][1] * o.m[1][j] + m[i][2] * o.m[2][j];
    return r;
  }
  V3 operator*(const V3 &v) const {
    return V3(m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z,
              m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
              m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z);
  }
  V3 row(int i) const { return V3(m[i][0], m[i][1], m[i][2]); }
  V3 col(int j) const { return V3(m[0][j], m[1][j], m[2][j]); }
  // Eigen's fixed-size 3x3 inverse(): cofactors / determinant.
  M3 inverse() const {
    M3 r;
    double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
    double c10 = m[1][2] * m[2][0] - m[1][0] * m[2][2]; // cofactor(0,1) sign folded
    double c20 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
    double det = m[0][0] * c00 + m[0][1] * c10 + m[0][2] * c20;
    double invdet = 1.0 / det;
    r.m[0][0] = c00 * invdet;
    r.m[1][0] = c10 * invdet;
    r.m[2][0] = c20 * invdet;
    r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * invdet;
    r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * invdet;
    r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * invdet;
    r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * invdet;
    r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * invdet;
    r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * invdet;
    return r;
  }
};

// util/types.h:36-42
inline V3 homogeneous(const V2 &v) { return V3(v.x, v.y, 1.0); }
inline V2 dehomogeneous(const V3 &v) {
  return V2(v.x, v.y) / (v.z + EPS);
}

// std::min / std::max semantics (NaN behaviour matters: line_dists.h:53-66
// feeds acos() results that may be NaN into std::min).
inline double smin(double a, double b) { return (b < a) ? b : a; }
inline double smax(double a, double b) { return (a < b) ? b : a; }

// ---------------------------------------------------------------------------
// base/pose.cc:12-29 — COLMAP 3.8 quaternion helpers (wxyz).
inline void NormalizeQuaternion(const double q[4], double out[4]) {
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n == 0) {
    out[0] = 1.0; out[1] = q[1]; out[2] = q[2]; out[3] = q[3];
  } else {
    for (int i = 0; i < 4; ++i)
      out[i] = q[i] / n;
  }
}
// Eigen::Quaterniond::toRotationMatrix()
inline M3 QuaternionToRotationMatrix(const double qvec[4]) {
  double q[4];
  NormalizeQuaternion(qvec, q);
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M3 R;
  R.m[0][0] = 1 - (tyy + tzz); R.m[0][1] = txy - twz; R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz; R.m[1][1] = 1 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy; R.m[2][1] = tyz + twx; R.m[2][2] = 1 - (txx + tyy);
  return R;
}
// Eigen::Quaterniond(const Matrix3d&) (Shepperd), returns wxyz.
inline void RotationMatrixToQuaternion(const M3 &R, double q[4]) {
  double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (R.m[2][1] - R.m[1][2]) * t;
    q[2] = (R.m[0][2] - R.m[2][0]) * t;
    q[3] = (R.m[1][0] - R.m[0][1]) * t;
  } else {
    int i = 0;
    if (R.m[1][1] > R.m[0][0]) i = 1;
    if (R.m[2][2] > R.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R.m[k][j] - R.m[j][k]) * t;
    v[j] = (R.m[j][i] + R.m[i][j]) * t;
    v[k] = (R.m[k][i] + R.m[i][k]) * t;
    q[1] = v[0]; q[2] = v[1]; q[3] = v[2];
  }
}

// ---------------------------------------------------------------------------
// base/camera.h:33-95, base/camera.cc:228-242; only the two undistorted
// pinhole models are legal on this path (base/camera_models.h:29-44,
// triangulation/base_line_triangulator.cc:49). kvec = [fx, fy, cx, cy].
struct Camera {
  int model_id = 1; // 0 SIMPLE_PINHOLE, 1 PINHOLE
  double kvec[4] = {1, 1, 0, 0};
  M3 K() const {
    M3 k;
    k.m[0][0] = kvec[0]; k.m[0][2] = kvec[2];
    k.m[1][1] = kvec[1]; k.m[1][2] = kvec[3];
    k.m[2][2] = 1.0;
    return k;
  }
  M3 K_inv() const { return K().inverse(); } // camera.h:73 (recomputed per call)
  double uncertainty(double depth, double var2d) const { // camera.cc:228-242
    double f = (model_id == 0) ? kvec[0] : (kvec[0] + kvec[1]) / 2.0;
    return var2d * depth / f;
  }
};

// base/camera.h:89-112
struct CameraPose {
  double qvec[4] = {1, 0, 0, 0};
  V3 tvec;
  void set(const double q[4], const double t[3]) {
    // CameraPose(V4D qvec, V3D tvec): qvec(qvec.normalized())
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double n = std::sqrt(n2);
    for (int i = 0; i < 4; ++i)
      qvec[i] = (n2 > 0) ? q[i] / n : q[i];
    tvec = V3(t[0], t[1], t[2]);
  }
  M3 R() const { return QuaternionToRotationMatrix(qvec); } // camera.h:106
  V3 T() const { return tvec; }
  V3 center() const { return -(R().transpose() * T()); } // camera.h:109
  double projdepth(const V3 &p) const {                   // camera.cc:276-279
    V3 pc = R() * p + T();
    return pc.z;
  }
};

// base/camera_view.h / camera_view.cc:61-69
struct CameraView {
  Camera cam;
  CameraPose pose;
  M3 K() const { return cam.K(); }
  M3 K_inv() const { return cam.K_inv(); }
  M3 R() const { return pose.R(); }
  V3 T() const { return pose.T(); }
  V2 projection(const V3 &p3d) const {
    V3 p_homo = K() * (R() * p3d + T());
    return dehomogeneous(p_homo);
  }
  V3 ray_direction(const V2 &p2d) const {
    return ((R().transpose() * K_inv()) * homogeneous(p2d)).normalized();
  }
};

// ---------------------------------------------------------------------------
// base/linebase.h:17-37, linebase.cc:35-39
struct Line2d {
  V2 start, end;
  double score = -1;
  Line2d() {}
  Line2d(V2 s, V2 e, double sc = -1) : start(s), end(e), score(sc) {}
  double length() const { return (start - end).norm(); }
  V2 midpoint() const { return (start + end) * 0.5; }
  V2 direction() const { return (end - start).normalized(); }
  V3 coords() const {
    return homogeneous(start).cross(homogeneous(end)).normalized();
  }
};

// base/linebase.h:39-61, linebase.cc:93-116
struct Line3d {
  V3 start, end;
  double score = -1;
  double uncertainty = -1.0;
  double depths[2] = {0, 0};
  Line3d() {}
  Line3d(V3 s, V3 e, double sc = -1, double ds = -1, double de = -1,
         double unc = -1)
      : start(s), end(e), score(sc), uncertainty(unc) {
    depths[0] = ds;
    depths[1] = de;
  }
  double length() const { return (start - end).norm(); }
  V3 midpoint() const { return (start + end) * 0.5; }
  V3 direction() const { return (end - start).normalized(); }
  V3 point_projection(const V3 &p) const; // unused on the path
  Line2d projection(const CameraView &view) const {
    Line2d l;
    l.start = view.projection(start);
    l.end = view.projection(end);
    return l;
  }
  double sensitivity(const CameraView &view) const { // linebase.cc:100-107
    Line2d l2 = projection(view);
    V3 dir3d = view.ray_direction(l2.midpoint());
    double cos_val = std::abs(direction().dot(dir3d));
    double angle = std::acos(cos_val) * 180.0 / M_PI;
    return 90 - angle;
  }
  double computeUncertainty(const CameraView &view, double var2d) const {
    double d1 = view.pose.projdepth(start);
    double d2 = view.pose.projdepth(end);
    double d = (d1 + d2) / 2.0;
    return view.cam.uncertainty(d, var2d);
  }
};

// ---------------------------------------------------------------------------
// base/line_dists.h / line_dists.cc — only the distances the linkers call.
template <typename L> inline double cosine(const L &l1, const L &l2) {
  return std::abs(l1.direction().dot(l2.direction()));
}
template <typename L> inline double compute_angle(const L &l1, const L &l2) {
  // line_dists.h:62-66 — no clamp: |cos| > 1 gives NaN
  return std::acos(cosine(l1, l2)) * 180.0 / M_PI;
}
template <typename L>
inline std::pair<double, double>
dists_endpoints_perpendicular_oneway(const L &l1, const L &l2) {
  auto v2 = l2.direction();
  auto disps = l1.start - l2.start;
  double d12s_sq = disps.squaredNorm() - std::pow(disps.dot(v2), 2);
  double d12s = std::sqrt(smax(d12s_sq, 0.0));
  auto dispe = l1.end - l2.start;
  double d12e_sq = dispe.squaredNorm() - std::pow(dispe.dot(v2), 2);
  double d12e = std::sqrt(smax(d12e_sq, 0.0));
  return std::make_pair(d12s, d12e);
}
template <typename L>
inline double dist_endpoints_perpendicular(const L &l1, const L &l2) {
  auto a = dists_endpoints_perpendicular_oneway(l1, l2);
  auto b = dists_endpoints_perpendicular_oneway(l2, l1);
  double d[4] = {a.first, a.second, b.first, b.second};
  return *std::max_element(d, d + 4);
}
inline double dist_endpoints_scaleinv_oneway(const Line3d &l1, const Line3d &l2) {
  double ds = (l1.start - l2.start).norm();
  double de = (l1.end - l2.end).norm();
  return smax(ds / (l1.depths[0] + EPS), de / (l1.depths[1] + EPS));
}
template <typename L> inline bool get_innerseg(const L &l1, const L &l2, L &inner) {
  auto l1_dir = l1.direction();
  double denom = (l2.end - l2.start).dot(l1_dir);
  double nume_start = (l1.start - l2.start).dot(l1_dir);
  double t1 = nume_start / (denom + EPS);
  double nume_end = (l1.end - l2.start).dot(l1_dir);
  double t2 = nume_end / (denom + EPS);
  if (t1 > t2)
    std::swap(t1, t2);
  if (t1 >= 1.0 || t2 <= 0.0)
    return false;
  inner.start = l2.start + (l2.end - l2.start) * smax(t1, 0.0);
  inner.end = l2.start + (l2.end - l2.start) * smin(t2, 1.0);
  return true;
}
template <typename L> inline double dist_innerseg(const L &l1, const L &l2) {
  double MAX_DIST = std::numeric_limits<double>::max();
  L a, b;
  if (!get_innerseg(l2, l1, a))
    return MAX_DIST;
  if (!get_innerseg(l1, l2, b))
    return MAX_DIST;
  return dist_endpoints_perpendicular(a, b);
}
template <typename L> inline double compute_overlap(const L &l1, const L &l2) {
  double len = l2.length();
  auto v = l2.direction();
  double p1 = (l1.start - l2.start).dot(v) / len;
  double p2 = (l1.end - l2.start).dot(v) / len;
  if (p1 > p2)
    std::swap(p1, p2);
  return smin(p2, 1.0) - smax(p1, 0.0);
}
template <typename L> inline double compute_bioverlap(const L &l1, const L &l2) {
  double v1 = compute_overlap(l1, l2);
  double v2 = compute_overlap(l2, l1);
  return smax(v1, v2);
}

// ---------------------------------------------------------------------------
// base/line_linker.{h,cc}
struct LinkerConfig { // union of LineLinker2dConfig / LineLinker3dConfig
  double score_th, th_angle, th_overlap, th_smartoverlap, th_smartangle,
      th_perp, th_innerseg, th_scaleinv;
  int use_angle, use_overlap, use_smartangle, use_perp, use_innerseg,
      use_scaleinv;
  double multiplier() const { return 1.0 / std::sqrt(-std::log(score_th) * 2.0); }
  void set_to_shared_parent_scoring() { // line_linker.h:115-121
    use_angle = 1; use_overlap = 0; use_perp = 0; use_innerseg = 0; use_scaleinv = 1;
  }
  void set_to_spatial_merging() { // line_linker.h:123-129
    use_angle = 1; use_overlap = 1; use_perp = 0; use_innerseg = 1; use_scaleinv = 0;
  }
};
inline LinkerConfig default_linker2d() { // line_linker.h:23-45
  LinkerConfig c{0.5, 8.0, 0.1, 0.2, 1.0, 5.0, 5.0, 0.0, 1, 1, 1, 1, 0, 0};
  return c;
}
inline LinkerConfig default_linker3d() { // line_linker.h:85-111
  LinkerConfig c{0.5, 10.0, 0.01, 0.1, 1.0, 0.02, 0.02, 0.01, 1, 1, 1, 0, 1, 0};
  return c;
}
inline double expscore(double val, double sigma) { // line_linker.cc:15-17
  return std::exp(-std::pow(val / sigma, 2) / 2.0);
}

template <typename L, bool IS3D> struct Linker {
  LinkerConfig config;
  double score_angle(const L &l1, const L &l2) const {
    double angle = compute_angle(l1, l2);
    double s = expscore(angle, config.th_angle * config.multiplier());
    if (s < config.score_th) s = 0.0;
    return s;
  }
  double score_smartangle(const L &l1, const L &l2) const {
    double angle = compute_angle(l1, l2);
    double th_angle = config.th_angle;
    double overlap = compute_bioverlap(l1, l2);
    if (overlap < config.th_smartoverlap) {
      double ratio = (config.th_smartoverlap - overlap) /
                     (config.th_smartoverlap - config.th_overlap);
      ratio = smin(ratio, 1.0);
      th_angle = config.th_angle - ratio * (config.th_angle - config.th_smartangle);
    }
    double s = expscore(angle, th_angle * config.multiplier());
    if (s < config.score_th) s = 0.0;
    return s;
  }
  double score_overlap(const L &l1, const L &l2) const {
    return compute_bioverlap(l1, l2) > config.th_overlap ? 1.0 : 0.0;
  }
  double uncert(const L &l1, const L &l2) const;
  double score_perp(const L &l1, const L &l2) const {
    double dist = dist_endpoints_perpendicular(l1, l2);
    double s = expscore(dist, config.th_perp * uncert(l1, l2) * config.multiplier());
    if (s < config.score_th) s = 0.0;
    return s;
  }
  double score_innerseg(const L &l1, const L &l2) const {
    double dist = dist_innerseg(l1, l2);
    double s = expscore(dist, config.th_innerseg * uncert(l1, l2) * config.multiplier());
    if (s < config.score_th) s = 0.0;
    return s;
  }
  double score_scaleinv(const L &l1, const L &l2) const;
  // line_linker.cc:139-160 (2d) and :306-331 (3d)
  double compute_score(const L &l1, const L &l2) const {
    double score = 1.0;
    if (config.use_angle) score = smin(score, score_angle(l1, l2));
    if (score < config.score_th) return score;
    if (config.use_overlap) score = smin(score, score_overlap(l1, l2));
    if (score < config.score_th) return score;
    if (config.use_angle && config.use_overlap && config.use_smartangle)
      score = smin(score, score_smartangle(l1, l2));
    if (score < config.score_th) return score;
    if (config.use_perp) score = smin(score, score_perp(l1, l2));
    if (score < config.score_th) return score;
    if (config.use_innerseg) score = smin(score, score_innerseg(l1, l2));
    if (IS3D) {
      if (score < config.score_th) return score;
      if (config.use_scaleinv) score = smin(score, score_scaleinv(l1, l2));
    }
    return score;
  }
};
// 2d scores are in pixels (line_linker.cc:92-115); 3d perp / innerseg thresholds
// scale with min(uncertainty) (line_linker.cc:239-262).
template <> inline double Linker<Line2d, false>::uncert(const Line2d &, const Line2d &) const { return 1.0; }
template <> inline double Linker<Line3d, true>::uncert(const Line3d &a, const Line3d &b) const {
  return smin(a.uncertainty, b.uncertainty);
}
template <> inline double Linker<Line2d, false>::score_scaleinv(const Line2d &, const Line2d &) const { return 1.0; }
template <> inline double Linker<Line3d, true>::score_scaleinv(const Line3d &l1, const Line3d &l2) const {
  double dist = dist_endpoints_scaleinv_oneway(l1, l2);
  double s = expscore(dist, config.th_scaleinv * config.multiplier());
  if (s < config.score_th) s = 0.0;
  return s;
}
typedef Linker<Line2d, false> LineLinker2d;
typedef Linker<Line3d, true> LineLinker3d;

} // namespace orc
