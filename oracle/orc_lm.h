// oracle/orc_lm.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Residual functors, minimal line and segment cut: PINNED to the reference's compiled code (tests/test_ref_pinning.py:
// values and 6-column Jacobians of GeometricRefinementFunctor / VPConstraintsFunctor). The SOLVER LOOP is PARITY
// UNPINNED: the solver the reference calls (Ceres, version not pinned: cmake/FindDependencies.cmake:19) is absent from
// /root/reference. The trust-region
// Levenberg-Marquardt below restates the published Ceres algorithm (trust_region_minimizer.cc,
// levenberg_marquardt_strategy.cc, corrector.cc, manifold.cc of ceres-solver 2.1/2.2, Manifold API:
// QuaternionManifold + SphereManifold<2>) from the Ceres documentation; anchors are the reference's
// call sites:
//   optimize/line_refinement/cost_functions.h:96-194   GeometricRefinementFunctor
//   optimize/line_refinement/refine.cc:19-198          RefinementEngine
//   optimize/line_refinement/refinement_config.h:18-92 solver options (tolerances 0, 100 iterations)
//   optimize/hybrid_bundle_adjustment/hybrid_bundle_adjustment.cc:39-59,156-264,298-310
//   ceresbase/line_transforms.h:9-29, ceresbase/line_projection.h:15-80, ceresbase/line_dists.h:20-29
//   ceresbase/parameterization.h:8-35, base/infinite_line.cc:67-82,180-231,265-287
//   base/linetrack.cc:315-322 (ComputeLineWeights)
#pragma once
#include "orc_geom.h"

namespace orc {

// ---- forward-mode dual numbers (ceres::Jet) -----------------------------------------------------
template <int N> struct Jet {
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double x, int k) : a(x) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1; }
};
template <int N> inline Jet<N> operator+(const Jet<N> &f, const Jet<N> &g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N> &f, const Jet<N> &g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N> &f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N> &f, const Jet<N> &g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N> &f, const Jet<N> &g) {
  Jet<N> h; const double gi = 1.0 / g.a; const double q = f.a * gi; h.a = q;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi;
  return h;
}
template <int N> inline Jet<N> sqrt(const Jet<N> &f) { Jet<N> h; h.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <int N> inline Jet<N> exp(const Jet<N> &f) { Jet<N> h; h.a = std::exp(f.a); for (int i = 0; i < N; ++i) h.v[i] = h.a * f.v[i]; return h; }
template <int N> inline Jet<N> abs(const Jet<N> &f) { return f.a < 0 ? -f : f; }
template <int N> inline bool operator>(const Jet<N> &f, const Jet<N> &g) { return f.a > g.a; }
inline double sqrt(double x) { return std::sqrt(x); }
inline double exp(double x) { return std::exp(x); }
inline double abs(double x) { return std::abs(x); }

// ceres/rotation.h QuaternionToRotation (row-major R, normalises by q.q)
template <typename T> inline void CeresQuaternionToRotation(const T q[4], T R[9]) {
  const T a = q[0], b = q[1], c = q[2], d = q[3];
  const T aa = a * a, ab = a * b, ac = a * c, ad = a * d, bb = b * b, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  R[0] = aa + bb - cc - dd; R[1] = T(2) * (bc - ad); R[2] = T(2) * (ac + bd);
  R[3] = T(2) * (ad + bc); R[4] = aa - bb + cc - dd; R[5] = T(2) * (cd - ab);
  R[6] = T(2) * (bd - ac); R[7] = T(2) * (ab + cd); R[8] = aa - bb - cc + dd;
  const T normalizer = T(1) / (aa + bb + cc + dd);
  for (int i = 0; i < 9; ++i) R[i] = R[i] * normalizer;
}
// ceresbase/line_transforms.h:9-29
template <typename T> inline void MinimalPluckerToPlucker(const T uvec[4], const T wvec[2], T d[3], T m[3]) {
  T rotmat[9];
  CeresQuaternionToRotation(uvec, rotmat);
  T w1 = abs(wvec[0]), w2 = abs(wvec[1]);
  d[0] = rotmat[0]; d[1] = rotmat[3]; d[2] = rotmat[6];
  T b_norm = w2 / (w1 + T(EPS));
  m[0] = rotmat[1] * b_norm; m[1] = rotmat[4] * b_norm; m[2] = rotmat[7] * b_norm;
}
template <typename T> inline void mat3mul(const T A[9], const T B[9], T C[9]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
    C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
template <typename T> inline void mat3T(const T A[9], T B[9]) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) B[3 * i + j] = A[3 * j + i]; }
template <typename T> inline void skew(const T m[3], T S[9]) {
  S[0] = T(0.0); S[1] = -m[2]; S[2] = m[1]; S[3] = m[2]; S[4] = T(0.0); S[5] = -m[0]; S[6] = -m[1]; S[7] = m[0]; S[8] = T(0.0);
}
// ceresbase/line_projection.h:15-48
template <typename T> inline void Line_ImgFromCam(const T *kvec, const T *mvec, T *coor) {
  T mskew[9]; skew(mvec, mskew);
  T K[9] = {kvec[0], T(0.0), kvec[2], T(0.0), kvec[1], kvec[3], T(0.0), T(0.0), T(1.0)};
  T Kt[9], tmp[9], cs[9];
  mat3T(K, Kt); mat3mul(K, mskew, tmp); mat3mul(tmp, Kt, cs);
  coor[0] = cs[7]; coor[1] = cs[2]; coor[2] = cs[3];
  T n = sqrt(coor[0] * coor[0] + coor[1] * coor[1] + coor[2] * coor[2] + T(EPS));
  coor[0] = coor[0] / n; coor[1] = coor[1] / n; coor[2] = coor[2] / n;
}
// ceresbase/line_projection.h:51-80
template <typename T> inline void Line_WorldToPixel(const T *kvec, const T *qvec, const T *tvec, const T *dvec, const T *mvec, T *coor) {
  T R[9]; CeresQuaternionToRotation(qvec, R);
  T mskew[9]; skew(mvec, mskew);
  T Rt[9], tmp[9], A[9];
  mat3T(R, Rt); mat3mul(R, mskew, tmp); mat3mul(tmp, Rt, A);
  T Rd[3];
  for (int i = 0; i < 3; ++i) Rd[i] = R[3 * i] * dvec[0] + R[3 * i + 1] * dvec[1] + R[3 * i + 2] * dvec[2];
  // R [m]x R^T - t (R d)^T + (R d) t^T
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[3 * i + j] = A[3 * i + j] - tvec[i] * Rd[j] + Rd[i] * tvec[j];
  T mt[3] = {A[7], A[2], A[3]};
  Line_ImgFromCam(kvec, mt, coor);
}
// ceresbase/line_dists.h:20-29
template <typename T> inline T CeresComputeDist2D_cosine(const T dir1[2], const T dir2[2]) {
  T n1 = sqrt(dir1[0] * dir1[0] + dir1[1] * dir1[1] + T(EPS));
  T n2 = sqrt(dir2[0] * dir2[0] + dir2[1] * dir2[1] + T(EPS));
  T cosine = (dir1[0] * dir2[0] + dir1[1] * dir2[1]) / (n1 * n2);
  cosine = abs(cosine);
  if (cosine > T(1.0)) cosine = T(1.0);
  return cosine;
}
// optimize/line_refinement/cost_functions.h:96-127
template <typename T> inline void Ceres_CosineWeightedPerpendicularDist2D_1D(const T coor[3], const T p1[2], const T p2[2], T *res, double alpha) {
  T direc_norm = sqrt(coor[0] * coor[0] + coor[1] * coor[1] + T(EPS));
  T dir2d[2] = {-coor[1] / direc_norm, coor[0] / direc_norm};
  T direc[2] = {p2[0] - p1[0], p2[1] - p1[1]};
  T cosine = CeresComputeDist2D_cosine(dir2d, direc);
  T weight = exp(T(alpha) * (T(1.0) - cosine));
  T dn = sqrt(coor[0] * coor[0] + coor[1] * coor[1] + T(EPS));
  res[0] = (p1[0] * coor[0] + p1[1] * coor[1] + coor[2]) / dn;
  res[1] = (p2[0] * coor[0] + p2[1] * coor[1] + coor[2]) / dn;
  res[0] = res[0] * weight; res[1] = res[1] * weight;
}
// GeometricRefinementFunctor::operator() (cost_functions.h:170-186), cameras constant
struct LMBlock {
  double p1[2], p2[2], kvec[4], qvec[4], tvec[3], w;
  bool has_vp = false;  // VPConstraintsFunctor block attached to this support (refine.cc:86-127)
  double vp[3] = {0, 0, 0};
  double wvp = 0;       // weights[k] * vp_multiplier
};
// ceresbase/line_projection.h:125-135
template <typename T> inline void GetDirectionFromVP(const T vp[3], const T kvec[4], T direc[3]) {
  direc[0] = vp[0] / kvec[0] - kvec[2] / kvec[0] * vp[2];
  direc[1] = vp[1] / kvec[1] - kvec[3] / kvec[1] * vp[2];
  direc[2] = vp[2];
  T norm = sqrt(direc[0] * direc[0] + direc[1] * direc[1] + direc[2] * direc[2] + T(EPS));
  direc[0] = direc[0] / norm; direc[1] = direc[1] / norm; direc[2] = direc[2] / norm;
}
// ceresbase/line_dists.h:40-57
template <typename T> inline T CeresComputeDist3D_sine(const T dir1[3], const T dir2[3]) {
  T n1 = sqrt(dir1[0] * dir1[0] + dir1[1] * dir1[1] + dir1[2] * dir1[2] + T(EPS));
  T n2 = sqrt(dir2[0] * dir2[0] + dir2[1] * dir2[1] + dir2[2] * dir2[2] + T(EPS));
  T a[3] = {dir1[0] / n1, dir1[1] / n1, dir1[2] / n1}, b[3] = {dir2[0] / n2, dir2[1] / n2, dir2[2] / n2};
  T r[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  T sine = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + T(EPS));
  if (sine > T(1.0)) sine = T(1.0);
  return sine;
}
// VPConstraintsFunctor::operator() (optimize/line_refinement/cost_functions.h:60-85), camera constant
template <typename T> inline T VPResidual(const LMBlock &b, const T uvec[4], const T wvec[2]) {
  T kvec[4] = {T(b.kvec[0]), T(b.kvec[1]), T(b.kvec[2]), T(b.kvec[3])};
  T dir3d[3], m[3];
  MinimalPluckerToPlucker(uvec, wvec, dir3d, m);
  T q[4] = {T(b.qvec[0]), T(b.qvec[1]), T(b.qvec[2]), T(b.qvec[3])};
  T R[9];
  CeresQuaternionToRotation(q, R); // ceres::QuaternionRotatePoint normalises the quaternion: same rotation
  T rot[3];
  for (int i = 0; i < 3; ++i) rot[i] = R[3 * i] * dir3d[0] + R[3 * i + 1] * dir3d[1] + R[3 * i + 2] * dir3d[2];
  T vpvec[3] = {T(b.vp[0]), T(b.vp[1]), T(b.vp[2])}, direc[3];
  GetDirectionFromVP(vpvec, kvec, direc);
  return CeresComputeDist3D_sine(rot, direc);
}
template <typename T> inline void GeometricResidual(const LMBlock &b, const T uvec[4], const T wvec[2], double alpha, T res[2]) {
  T kvec[4] = {T(b.kvec[0]), T(b.kvec[1]), T(b.kvec[2]), T(b.kvec[3])};
  T qvec[4] = {T(b.qvec[0]), T(b.qvec[1]), T(b.qvec[2]), T(b.qvec[3])};
  T tvec[3] = {T(b.tvec[0]), T(b.tvec[1]), T(b.tvec[2])};
  T dvec[3], mvec[3];
  MinimalPluckerToPlucker(uvec, wvec, dvec, mvec);
  T coor[3];
  Line_WorldToPixel(kvec, qvec, tvec, dvec, mvec, coor);
  T p1[2] = {T(b.p1[0]), T(b.p1[1])}, p2[2] = {T(b.p2[0]), T(b.p2[1])};
  Ceres_CosineWeightedPerpendicularDist2D_1D(coor, p1, p2, res, alpha);
}

// ---- manifolds (ceres manifold.cc / sphere_manifold_functions.h / householder_vector.h) ---------
inline void QuaternionPlus(const double x[4], const double d[3], double out[4]) {
  const double sq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (sq == 0.0) { for (int i = 0; i < 4; ++i) out[i] = x[i]; return; }
  const double nd = std::sqrt(sq), sbd = std::sin(nd) / nd;
  const double z[4] = {std::cos(nd), sbd * d[0], sbd * d[1], sbd * d[2]};
  out[0] = z[0] * x[0] - z[1] * x[1] - z[2] * x[2] - z[3] * x[3];
  out[1] = z[0] * x[1] + z[1] * x[0] + z[2] * x[3] - z[3] * x[2];
  out[2] = z[0] * x[2] - z[1] * x[3] + z[2] * x[0] + z[3] * x[1];
  out[3] = z[0] * x[3] + z[1] * x[2] - z[2] * x[1] + z[3] * x[0];
}
inline void QuaternionPlusJacobian(const double x[4], double J[12]) { // 4x3 row-major
  J[0] = -x[1]; J[1] = -x[2]; J[2] = -x[3];
  J[3] = x[0];  J[4] = x[3];  J[5] = -x[2];
  J[6] = -x[3]; J[7] = x[0];  J[8] = x[1];
  J[9] = x[2];  J[10] = -x[1]; J[11] = x[0];
}
inline void Householder2(const double x[2], double v[2], double &beta) {
  const double sigma = x[0] * x[0];
  v[0] = x[0]; v[1] = 1.0; beta = 0.0;
  const double x_pivot = x[1];
  if (sigma <= std::numeric_limits<double>::epsilon()) { if (x_pivot < 0.0) beta = 2.0; return; }
  const double mu = std::sqrt(x_pivot * x_pivot + sigma);
  double v_pivot = 1.0;
  if (x_pivot <= 0.0) v_pivot = x_pivot - mu; else v_pivot = -sigma / (x_pivot + mu);
  beta = 2.0 * v_pivot * v_pivot / (sigma + v_pivot * v_pivot);
  v[0] /= v_pivot;
}
inline void Sphere2Plus(const double x[2], double delta, double out[2]) {
  const double nd = std::abs(delta);
  if (nd == 0.0) { out[0] = x[0]; out[1] = x[1]; return; }
  double v[2], beta;
  Householder2(x, v, beta);
  const double y[2] = {std::sin(nd) / nd * delta, std::cos(nd)};
  const double vty = v[0] * y[0] + v[1] * y[1];
  const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1]);
  out[0] = nx * (y[0] - v[0] * (beta * vty));
  out[1] = nx * (y[1] - v[1] * (beta * vty));
}
inline void Sphere2PlusJacobian(const double x[2], double J[2]) { // 2x1
  double v[2], beta;
  Householder2(x, v, beta);
  const double nx = std::sqrt(x[0] * x[0] + x[1] * x[1]);
  J[0] = (-beta * v[0] * v[0] + 1.0) * nx;
  J[1] = (-beta * v[0] * v[1]) * nx;
}

// ---- minimal line <-> Line3d (base/infinite_line.cc) ----------------------------------------------
struct MinimalLine { double uvec[4], wvec[2]; };
inline MinimalLine MinimalFromLine3d(const Line3d &line) { // :67-71 + :180-218
  V3 a = line.direction();
  V3 b = line.start.cross(a);
  MinimalLine out;
  double w1 = 1.0, w2 = b.norm();
  double denom = V2(w1, w2).norm();
  out.wvec[0] = w1 / denom; out.wvec[1] = w2 / denom;
  V3 q0 = a / a.norm(), q1, q2;
  if (b.norm() > EPS) {
    q1 = b / b.norm();
    V3 axb = a.cross(b);
    q2 = axb / axb.norm();
  } else {
    int best = 0;
    if (std::abs(a[1]) > std::abs(a[0])) best = 1;
    if (std::abs(a[2]) > std::abs(a[best])) best = 2;
    int i1 = (best + 1) % 3, i2 = (best + 2) % 3;
    double bp[3]; bp[i1] = 1.0; bp[i2] = 1.0; bp[best] = -(a[i1] * bp[i1] + a[i2] * bp[i2]) / a[best];
    V3 bprime(bp[0], bp[1], bp[2]);
    q1 = bprime / bprime.norm();
    V3 axb = a.cross(bprime);
    q2 = axb / axb.norm();
  }
  M3 Q = M3::fromCols(q0, q1, q2);
  RotationMatrixToQuaternion(Q, out.uvec);
  return out;
}
inline void InfiniteFromMinimal(const MinimalLine &ml, V3 &d, V3 &m) { // :220-231
  M3 Q = QuaternionToRotationMatrix(ml.uvec);
  d = Q.col(0);
  m = Q.col(1) * (std::abs(ml.wvec[1]) / std::abs(ml.wvec[0]));
}
// GetLineSegmentFromInfiniteLine3d(inf_line, line3ds, num_outliers) (:265-287)
inline Line3d SegmentFromInfinite(const V3 &d, const V3 &m, const std::vector<Line3d> &line3ds, int num_outliers) {
  auto point_projection = [&](const V3 &q) { V3 m_q = m + d.cross(q); return q + d.cross(m_q); }; // :73-78
  V3 p_ref = point_projection(line3ds[0].start);
  std::vector<double> values;
  for (const Line3d &l : line3ds) { values.push_back((l.start - p_ref).dot(d)); values.push_back((l.end - p_ref).dot(d)); }
  std::sort(values.begin(), values.end());
  int n = (int)line3ds.size();
  Line3d out;
  out.start = p_ref + d * values[num_outliers];
  out.end = p_ref + d * values[n * 2 - 1 - num_outliers];
  return out;
}

// ---- trust-region LM on one track ------------------------------------------------------------------
struct LMOptions {
  int max_num_iterations = 100;               // refinement_config.h:30 (runner passes 200)
  double geometric_alpha = 10.0;              // :59
  double cauchy_scale = 0.25;                 // :21
  double initial_trust_region_radius = 1e4;   // Ceres defaults below
  double max_trust_region_radius = 1e16;
  double min_trust_region_radius = 1e-32;
  double min_relative_decrease = 1e-3;
  double min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  int max_num_consecutive_invalid_steps = 10; // :32
  bool jacobi_scaling = true;
};
struct LMSummary { int iterations = 0, successful = 0; double initial_cost = 0, final_cost = 0; int termination = 0; };

struct LMProblem {
  std::vector<LMBlock> blocks;
  LMOptions opt;
  // cost (and optionally loss-corrected residuals + local Jacobian [2S x 4]) at x = (uvec, wvec)
  // Residual layout: 3 rows per support (2 geometric + 1 VP row, zero when the support has no VP).
  double evaluate(const double x[6], std::vector<double> *res, std::vector<double> *jac) const {
    const int S = (int)blocks.size();
    const double bq = opt.cauchy_scale * opt.cauchy_scale, cq = 1.0 / bq;
    double cost = 0;
    double Pq[12] = {0}, Ps[2] = {0, 0};
    if (jac) { QuaternionPlusJacobian(x, Pq); Sphere2PlusJacobian(x + 4, Ps); jac->assign(12 * S, 0.0); }
    if (res) res->assign(3 * S, 0.0);
    auto to_local = [&](const double J6[6], double Jl[4]) {
      for (int c = 0; c < 3; ++c) Jl[c] = J6[0] * Pq[c] + J6[1] * Pq[3 + c] + J6[2] * Pq[6 + c] + J6[3] * Pq[9 + c];
      Jl[3] = J6[4] * Ps[0] + J6[5] * Ps[1];
    };
    for (int k = 0; k < S; ++k) {
      double r[2], J6[12] = {0};
      typedef Jet<6> J;
      if (jac) {
        J u[4] = {J(x[0], 0), J(x[1], 1), J(x[2], 2), J(x[3], 3)}, w[2] = {J(x[4], 4), J(x[5], 5)}, rr[2];
        GeometricResidual<J>(blocks[k], u, w, opt.geometric_alpha, rr);
        for (int i = 0; i < 2; ++i) { r[i] = rr[i].a; for (int j = 0; j < 6; ++j) J6[6 * i + j] = rr[i].v[j]; }
      } else {
        double rr[2];
        GeometricResidual<double>(blocks[k], x, x + 4, opt.geometric_alpha, rr);
        r[0] = rr[0]; r[1] = rr[1];
      }
      // ScaledLoss(CauchyLoss(0.25), w) (refine.cc:77-78); ceres loss_function.cc
      const double s = r[0] * r[0] + r[1] * r[1];
      const double sum = 1.0 + s * cq, inv = 1.0 / sum;
      const double a = blocks[k].w;
      const double rho0 = a * bq * std::log(sum), rho1 = a * std::max(std::numeric_limits<double>::min(), inv), rho2 = a * (-cq * (inv * inv));
      cost += 0.5 * rho0;
      if (res || jac) {
        // Corrector (ceres corrector.cc)
        const double sqrt_rho1 = std::sqrt(rho1);
        double residual_scaling, alpha_sq_norm;
        if (s == 0.0 || rho2 <= 0.0) { residual_scaling = sqrt_rho1; alpha_sq_norm = 0.0; }
        else {
          const double D = 1.0 + 2.0 * s * rho2 / rho1;
          const double alpha = 1.0 - std::sqrt(D);
          residual_scaling = sqrt_rho1 / (1 - alpha);
          alpha_sq_norm = alpha / s;
        }
        if (jac) {
          double Jl[8];
          to_local(J6, Jl); to_local(J6 + 6, Jl + 4);
          if (alpha_sq_norm == 0.0) { for (int i = 0; i < 8; ++i) (*jac)[12 * k + i] = sqrt_rho1 * Jl[i]; }
          else {
            for (int c = 0; c < 4; ++c) {
              const double rtj = r[0] * Jl[c] + r[1] * Jl[4 + c];
              (*jac)[12 * k + c] = sqrt_rho1 * (Jl[c] - alpha_sq_norm * r[0] * rtj);
              (*jac)[12 * k + 4 + c] = sqrt_rho1 * (Jl[4 + c] - alpha_sq_norm * r[1] * rtj);
            }
          }
        }
        if (res) { (*res)[3 * k] = r[0] * residual_scaling; (*res)[3 * k + 1] = r[1] * residual_scaling; }
      }
      // VP block: ScaledLoss(TrivialLoss, weight * vp_multiplier) (refine.cc:100-123)
      if (blocks[k].has_vp) {
        double rv, Jv6[6] = {0, 0, 0, 0, 0, 0};
        if (jac) {
          J u[4] = {J(x[0], 0), J(x[1], 1), J(x[2], 2), J(x[3], 3)}, w[2] = {J(x[4], 4), J(x[5], 5)};
          J rj = VPResidual<J>(blocks[k], u, w);
          rv = rj.a;
          for (int j = 0; j < 6; ++j) Jv6[j] = rj.v[j];
        } else
          rv = VPResidual<double>(blocks[k], x, x + 4);
        const double wv = blocks[k].wvp, sq = std::sqrt(wv);
        cost += 0.5 * wv * rv * rv;
        if (res) (*res)[3 * k + 2] = sq * rv;
        if (jac) {
          double Jl[4];
          to_local(Jv6, Jl);
          for (int c = 0; c < 4; ++c) (*jac)[12 * k + 8 + c] = sq * Jl[c];
        }
      }
    }
    return cost;
  }
  static void plus(const double x[6], const double delta[4], double out[6]) {
    QuaternionPlus(x, delta, out);
    Sphere2Plus(x + 4, delta[3], out + 4);
  }
  // ceres TrustRegionMinimizer + LevenbergMarquardtStrategy, dense 4x4 normal equations (Cholesky)
  LMSummary solve(double x[6]) const {
    LMSummary sum;
    const int S = (int)blocks.size();
    if (S == 0) return sum;
    std::vector<double> r, J;
    double cost = evaluate(x, &r, &J);
    sum.initial_cost = cost;
    double scale[4] = {1, 1, 1, 1};
    if (opt.jacobi_scaling) for (int c = 0; c < 4; ++c) {
      double n2 = 0; for (int i = 0; i < 3 * S; ++i) n2 += J[4 * i + c] * J[4 * i + c];
      scale[c] = 1.0 / (1.0 + std::sqrt(n2));
    }
    auto scale_cols = [&](std::vector<double> &Jm) { for (int i = 0; i < 3 * S; ++i) for (int c = 0; c < 4; ++c) Jm[4 * i + c] *= scale[c]; };
    scale_cols(J);
    double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
    bool reuse_diagonal = false;
    double diag[4] = {0, 0, 0, 0};
    int invalid = 0;
    int it = 0;
    while (true) {
      // FinalizeIterationAndCheckIfMinimizerCanContinue
      if (it >= opt.max_num_iterations) { sum.termination = 1; break; }
      if (radius <= opt.min_trust_region_radius) { sum.termination = 2; break; }
      { double gmax = 0; for (int c = 0; c < 4; ++c) { double g = 0; for (int i = 0; i < 3 * S; ++i) g += J[4 * i + c] / scale[c] * r[i]; gmax = std::max(gmax, std::abs(g)); }
        if (gmax <= 0.0) { sum.termination = 3; break; } }
      ++it;
      // LevenbergMarquardtStrategy::ComputeStep
      double A[16], g[4];
      for (int a = 0; a < 4; ++a) { g[a] = 0; for (int b = 0; b < 4; ++b) A[4 * a + b] = 0; }
      for (int i = 0; i < 3 * S; ++i) for (int a = 0; a < 4; ++a) { g[a] += J[4 * i + a] * r[i]; for (int b = 0; b < 4; ++b) A[4 * a + b] += J[4 * i + a] * J[4 * i + b]; }
      if (!reuse_diagonal) for (int c = 0; c < 4; ++c) diag[c] = std::min(std::max(A[5 * c], opt.min_lm_diagonal), opt.max_lm_diagonal);
      for (int c = 0; c < 4; ++c) A[5 * c] += diag[c] / radius; // D^T D, D = sqrt(diag / radius)
      reuse_diagonal = true;
      double step[4];
      bool ok = chol_solve4(A, g, step);
      for (int c = 0; c < 4; ++c) { step[c] = -step[c]; if (!std::isfinite(step[c])) ok = false; }
      double model_cost_change = 0;
      if (ok) {
        // model_residuals = J step ; change = -model_residuals . (r + model_residuals / 2)
        for (int i = 0; i < 3 * S; ++i) {
          double mr = 0; for (int c = 0; c < 4; ++c) mr += J[4 * i + c] * step[c];
          model_cost_change -= mr * (r[i] + mr / 2.0);
        }
      }
      if (!ok || !(model_cost_change > 0.0)) {
        if (++invalid >= opt.max_num_consecutive_invalid_steps) { sum.termination = 4; break; }
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true; // StepIsInvalid -> StepRejected(0)
        continue;
      }
      invalid = 0;
      double delta[4], cand[6];
      for (int c = 0; c < 4; ++c) delta[c] = step[c] * scale[c];
      plus(x, delta, cand);
      const double cand_cost = evaluate(cand, nullptr, nullptr);
      double step_norm = 0; for (int c = 0; c < 6; ++c) step_norm += (x[c] - cand[c]) * (x[c] - cand[c]);
      if (!(std::sqrt(step_norm) > 0.0)) { sum.termination = 5; break; }          // parameter_tolerance = 0
      if (!(std::abs(cost - cand_cost) > 0.0)) { sum.termination = 6; break; }     // function_tolerance = 0
      const double relative_decrease = (cost - cand_cost) / model_cost_change;
      if (relative_decrease > opt.min_relative_decrease) {
        for (int c = 0; c < 6; ++c) x[c] = cand[c];
        cost = evaluate(x, &r, &J);
        scale_cols(J);
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
        radius = std::min(opt.max_trust_region_radius, radius);
        decrease_factor = 2.0; reuse_diagonal = false;
        ++sum.successful;
      } else {
        radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true;
      }
    }
    sum.iterations = it;
    sum.final_cost = cost;
    return sum;
  }
  static bool chol_solve4(const double A[16], const double b[4], double x[4]) {
    double L[16] = {0};
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j <= i; ++j) {
        double s = A[4 * i + j];
        for (int k = 0; k < j; ++k) s -= L[4 * i + k] * L[4 * j + k];
        if (i == j) { if (!(s > 0)) return false; L[4 * i + i] = std::sqrt(s); }
        else L[4 * i + j] = s / L[4 * j + j];
      }
    double y[4];
    for (int i = 0; i < 4; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[4 * i + k] * y[k]; y[i] = s / L[4 * i + i]; }
    for (int i = 3; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 4; ++k) s -= L[4 * k + i] * x[k]; x[i] = s / L[4 * i + i]; }
    return true;
  }
};

} // namespace orc
