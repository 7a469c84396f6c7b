// lm_block.cuh — one residual block of the line refinement and the line parameterisation it is evaluated on, as
// __host__ __device__ code: lm_kernels.cu runs it on the device, scripts/lm_block_check.cu on the host (derivative
// check of the analytic form against the dual-number form, no GPU needed).
#pragma once
#include "lm_kernels.cuh"
#include <cfloat>
#include <cmath>

namespace lm {

template <int N> struct Dual {
  double a;
  double v[N];
};
template <int N> LM_HD Dual<N> dconst(double x) { Dual<N> r; r.a = x;
#pragma unroll
  for (int i = 0; i < N; ++i) r.v[i] = 0; return r; }
template <int N> LM_HD Dual<N> operator+(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; h.a = f.a + g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> LM_HD Dual<N> operator-(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; h.a = f.a - g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> LM_HD Dual<N> operator-(const Dual<N> &f) { Dual<N> h; h.a = -f.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> LM_HD Dual<N> operator*(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; h.a = f.a * g.a;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> LM_HD Dual<N> operator*(const Dual<N> &f, double s) { Dual<N> h; h.a = f.a * s;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> LM_HD Dual<N> operator/(const Dual<N> &f, const Dual<N> &g) { Dual<N> h; const double gi = 1.0 / g.a, q = f.a * gi; h.a = q;
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h; }
template <int N> LM_HD Dual<N> dsqrt(const Dual<N> &f) { Dual<N> h; h.a = sqrt(f.a); const double t = 1.0 / (2.0 * h.a);
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <int N> LM_HD Dual<N> dexp(const Dual<N> &f) { Dual<N> h; h.a = exp(f.a);
#pragma unroll
  for (int i = 0; i < N; ++i) h.v[i] = h.a * f.v[i]; return h; }
template <int N> LM_HD Dual<N> dabs(const Dual<N> &f) { return f.a < 0 ? -f : f; }


// ceres QuaternionManifold / SphereManifold<2> (DESIGN.md "LM recipe")
LM_HD void quat_plus(const double x[4], const double d[3], double out[4]) {
  const double sq = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  if (sq == 0.0) { out[0] = x[0]; out[1] = x[1]; out[2] = x[2]; out[3] = x[3]; return; }
  const double nd = sqrt(sq);
  double sn_, cs_;
  sincos(nd, &sn_, &cs_);
  const double sbd = sn_ / nd;
  const double z0 = cs_, z1 = sbd * d[0], z2 = sbd * d[1], z3 = sbd * d[2];
  out[0] = z0 * x[0] - z1 * x[1] - z2 * x[2] - z3 * x[3];
  out[1] = z0 * x[1] + z1 * x[0] + z2 * x[3] - z3 * x[2];
  out[2] = z0 * x[2] - z1 * x[3] + z2 * x[0] + z3 * x[1];
  out[3] = z0 * x[3] + z1 * x[2] - z2 * x[1] + z3 * x[0];
}
LM_HD void householder2(const double x[2], double v[2], double &beta) {
  const double sigma = x[0] * x[0];
  v[0] = x[0]; v[1] = 1.0; beta = 0.0;
  const double xp = x[1];
  if (sigma <= DBL_EPSILON) { if (xp < 0.0) beta = 2.0; return; }
  const double mu = sqrt(xp * xp + sigma);
  const double vp = (xp <= 0.0) ? (xp - mu) : (-sigma / (xp + mu));
  beta = 2.0 * vp * vp / (sigma + vp * vp);
  v[0] /= vp;
}
// hh = {v[0], beta, |x|} of householder2(x) (v[1] is 1), or NULL to compute it here. The solver keeps the triple of the
// current point: it is a by-product of the line evaluation there (line_from_minimal_col).
LM_HD void sphere2_plus(const double x[2], double delta, double out[2], const double *hh = nullptr) {
  const double nd = fabs(delta);
  if (nd == 0.0) { out[0] = x[0]; out[1] = x[1]; return; }
  double v[2], beta, nx;
  if (hh) { v[0] = hh[0]; v[1] = 1.0; beta = hh[1]; nx = hh[2]; }
  else { householder2(x, v, beta); nx = sqrt(x[0] * x[0] + x[1] * x[1]); }
  double sn_, cs_;
  sincos(nd, &sn_, &cs_);
  const double y0 = sn_ / nd * delta, y1 = cs_;
  const double vty = v[0] * y0 + v[1] * y1;
  out[0] = nx * (y0 - v[0] * (beta * vty));
  out[1] = nx * (y1 - v[1] * (beta * vty));
}


// d and m (world-frame Pluecker line) with derivatives w.r.t. the 4 local (tangent) coordinates.
// MinimalPluckerToPlucker with ceres::QuaternionToRotation (normalised by |u|^2), composed with the plus
// Jacobians of the two manifolds.
struct LineLocal {
  Dual<4> d[3], m[3];
};
// The same, laid out for shared memory (one per warp): every lane evaluates a different residual block against the
// same line, so the line and its 24 tangent derivatives are warp-uniform -- kept here, not in 60 registers per lane.
struct LineShared {
  double d[3], m[3];
  double dv[3][4], mv[3][4];
};
LM_HD void line_from_minimal(const double x[6], bool want_jac, LineLocal &L) {
  // ambient duals (6-wide) would be wasteful: seed the 4 local directions directly through the plus Jacobians
  Dual<4> u[4], w[2];
  // QuaternionPlusJacobian (4x3)
  const double Pq[12] = {-x[1], -x[2], -x[3], x[0], x[3], -x[2], -x[3], x[0], x[1], x[2], -x[1], x[0]};
  double v2[2], beta;
  householder2(x + 4, v2, beta);
  const double nx = sqrt(x[4] * x[4] + x[5] * x[5]);
  const double Ps[2] = {(-beta * v2[0] * v2[0] + 1.0) * nx, (-beta * v2[0] * v2[1]) * nx};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    u[i].a = x[i];
    u[i].v[0] = want_jac ? Pq[3 * i] : 0; u[i].v[1] = want_jac ? Pq[3 * i + 1] : 0; u[i].v[2] = want_jac ? Pq[3 * i + 2] : 0;
    u[i].v[3] = 0;
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    w[i].a = x[4 + i];
    w[i].v[0] = w[i].v[1] = w[i].v[2] = 0;
    w[i].v[3] = want_jac ? Ps[i] : 0;
  }
  const Dual<4> a = u[0], b = u[1], c = u[2], dd = u[3];
  const Dual<4> aa = a * a, ab = a * b, ac = a * c, ad = a * dd, bb = b * b, bc = b * c, bd = b * dd, cc = c * c,
                cd = c * dd, d2 = dd * dd;
  const Dual<4> nrm = dconst<4>(1.0) / (aa + bb + cc + d2);
  // column 0 of R: (R00, R10, R20); column 1: (R01, R11, R21)
  L.d[0] = (aa + bb - cc - d2) * nrm;
  L.d[1] = ((ad + bc) * 2.0) * nrm;
  L.d[2] = ((bd - ac) * 2.0) * nrm;
  const Dual<4> w1 = dabs(w[0]), w2 = dabs(w[1]);
  const Dual<4> bn = w2 / (w1 + dconst<4>(consts<double>::eps()));
  L.m[0] = (((bc - ad) * 2.0) * nrm) * bn;
  L.m[1] = ((aa - bb + cc - d2) * nrm) * bn;
  L.m[2] = (((ab + cd) * 2.0) * nrm) * bn;
}

// One tangent column of the same computation (1-wide duals): lane c of a warp takes column c, so the 24 derivatives cost
// a quarter of the registers of the 4-wide form and a warp computes all of them in one pass. Bit-identical to
// line_from_minimal: every derivative component is the same expression.
LM_HD void line_from_minimal_col(const double x[6], int col, bool want_jac, Dual<1> d[3], Dual<1> m[3],
                                 double *hh_out = nullptr) {
  Dual<1> u[4], w[2];
  double v2[2], beta;
  householder2(x + 4, v2, beta);
  const double nx = sqrt(x[4] * x[4] + x[5] * x[5]);
  if (hh_out) { hh_out[0] = v2[0]; hh_out[1] = beta; hh_out[2] = nx; } // (see sphere2_plus)
  const double Ps[2] = {(-beta * v2[0] * v2[0] + 1.0) * nx, (-beta * v2[0] * v2[1]) * nx};
  // QuaternionPlusJacobian (4x3), column `col` (col 3 belongs to the sphere)
  double pq[4] = {0.0, 0.0, 0.0, 0.0};
  if (col == 0) { pq[0] = -x[1]; pq[1] = x[0]; pq[2] = -x[3]; pq[3] = x[2]; }
  else if (col == 1) { pq[0] = -x[2]; pq[1] = x[3]; pq[2] = x[0]; pq[3] = -x[1]; }
  else if (col == 2) { pq[0] = -x[3]; pq[1] = -x[2]; pq[2] = x[1]; pq[3] = x[0]; }
#pragma unroll
  for (int i = 0; i < 4; ++i) { u[i].a = x[i]; u[i].v[0] = want_jac ? pq[i] : 0.0; }
#pragma unroll
  for (int i = 0; i < 2; ++i) { w[i].a = x[4 + i]; w[i].v[0] = (want_jac && col == 3) ? Ps[i] : 0.0; }
  const Dual<1> a = u[0], b = u[1], c = u[2], dd = u[3];
  const Dual<1> aa = a * a, ab = a * b, ac = a * c, ad = a * dd, bb = b * b, bc = b * c, bd = b * dd, cc = c * c,
                cd = c * dd, d2 = dd * dd;
  const Dual<1> nrm = dconst<1>(1.0) / (aa + bb + cc + d2);
  d[0] = (aa + bb - cc - d2) * nrm;
  d[1] = ((ad + bc) * 2.0) * nrm;
  d[2] = ((bd - ac) * 2.0) * nrm;
  const Dual<1> w1 = dabs(w[0]), w2 = dabs(w[1]);
  const Dual<1> bn = w2 / (w1 + dconst<1>(consts<double>::eps()));
  m[0] = (((bc - ad) * 2.0) * nrm) * bn;
  m[1] = ((aa - bb + cc - d2) * nrm) * bn;
  m[2] = (((ab + cd) * 2.0) * nrm) * bn;
}

struct BlockEval {
  double r[2];     // raw residuals
  double J[8];     // 2x4 local Jacobian (raw)
  double rv;       // VP residual (VPConstraintsFunctor), only when B.wvp > 0
  double Jv[4];
};

// One residual block: from (d, m) to the two cosine-weighted point-line distances. Dual-number form (3-wide duals from
// the camera-frame moment to the residuals): the derivation reference of eval_block below; not used by the kernel.
LM_HD void eval_block_dual(const LMBlockDev &B, const LineShared &L, double alpha, bool want_jac, BlockEval &o) {
  // m_c = R m + t x (R d)   (Line_WorldToPixel, matrix form R [m]x R^T - t (Rd)^T + (Rd) t^T)
  double Rd[3], Rm[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Rd[i] = B.R[3 * i] * L.d[0] + B.R[3 * i + 1] * L.d[1] + B.R[3 * i + 2] * L.d[2];
    Rm[i] = B.R[3 * i] * L.m[0] + B.R[3 * i + 1] * L.m[1] + B.R[3 * i + 2] * L.m[2];
  }
  const double mc[3] = {Rm[0] + (B.t[1] * Rd[2] - B.t[2] * Rd[1]), Rm[1] + (B.t[2] * Rd[0] - B.t[0] * Rd[2]),
                        Rm[2] + (B.t[0] * Rd[1] - B.t[1] * Rd[0])};
  // 3-wide duals seeded on m_c
  Dual<3> q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) { q[i].a = mc[i]; q[i].v[0] = (i == 0); q[i].v[1] = (i == 1); q[i].v[2] = (i == 2); }
  // Line_ImgFromCam: coor = cof(K) m_c = (fy mx, fx my, fx fy mz - fy cx mx - fx cy my), then normalised (+EPS)
  const double fx = B.k[0], fy = B.k[1], cx = B.k[2], cy = B.k[3];
  Dual<3> c0 = q[0] * fy, c1 = q[1] * fx, c2 = q[2] * (fx * fy) - q[0] * (fy * cx) - q[1] * (fx * cy);
  const Dual<3> eps = dconst<3>(consts<double>::eps());
  const Dual<3> cn = dsqrt(c0 * c0 + c1 * c1 + c2 * c2 + eps);
  c0 = c0 / cn; c1 = c1 / cn; c2 = c2 / cn;
  // Ceres_CosineWeightedPerpendicularDist2D_1D
  const Dual<3> dn = dsqrt(c0 * c0 + c1 * c1 + eps);
  const Dual<3> dir0 = -c1 / dn, dir1 = c0 / dn;
  const double sx = B.p[2] - B.p[0], sy = B.p[3] - B.p[1];
  const Dual<3> n1 = dsqrt(dir0 * dir0 + dir1 * dir1 + eps);
  const double n2 = sqrt(sx * sx + sy * sy + consts<double>::eps());
  Dual<3> cosine = dabs((dir0 * sx + dir1 * sy) / (n1 * n2));
  if (cosine.a > 1.0) cosine = dconst<3>(1.0);
  const Dual<3> weight = dexp((dconst<3>(1.0) - cosine) * alpha);
  const Dual<3> r0 = ((c0 * B.p[0] + c1 * B.p[1] + c2) / dn) * weight;
  const Dual<3> r1 = ((c0 * B.p[2] + c1 * B.p[3] + c2) / dn) * weight;
  o.r[0] = r0.a;
  o.r[1] = r1.a;
  Dual<3> rvp = dconst<3>(0.0);
  if (B.wvp > 0.0) {
    // VPConstraintsFunctor (cost_functions.h:60-85): sine between R d and the VP direction
    // (CeresComputeDist3D_sine, ceresbase/line_dists.h:40-57); duals seeded on R d
    Dual<3> a[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { a[i].a = Rd[i]; a[i].v[0] = (i == 0); a[i].v[1] = (i == 1); a[i].v[2] = (i == 2); }
    const Dual<3> e3 = dconst<3>(consts<double>::eps());
    const Dual<3> n1 = dsqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + e3);
    const double n2 = sqrt(B.vdir[0] * B.vdir[0] + B.vdir[1] * B.vdir[1] + B.vdir[2] * B.vdir[2] + consts<double>::eps());
    const Dual<3> ax = a[0] / n1, ay = a[1] / n1, az = a[2] / n1;
    const double bx = B.vdir[0] / n2, by = B.vdir[1] / n2, bz = B.vdir[2] / n2;
    const Dual<3> cx_ = ay * bz - az * by, cy_ = az * bx - ax * bz, cz_ = ax * by - ay * bx;
    rvp = dsqrt(cx_ * cx_ + cy_ * cy_ + cz_ * cz_ + e3);
    if (rvp.a > 1.0) rvp = dconst<3>(1.0);
  }
  o.rv = rvp.a;
  if (!want_jac) return;
  // G = d m_c / d local (3x4) = R Dm + t x (R Dd)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double rd[3], rm[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      rd[i] = B.R[3 * i] * L.dv[0][c] + B.R[3 * i + 1] * L.dv[1][c] + B.R[3 * i + 2] * L.dv[2][c];
      rm[i] = B.R[3 * i] * L.mv[0][c] + B.R[3 * i + 1] * L.mv[1][c] + B.R[3 * i + 2] * L.mv[2][c];
    }
    const double g0 = rm[0] + (B.t[1] * rd[2] - B.t[2] * rd[1]);
    const double g1 = rm[1] + (B.t[2] * rd[0] - B.t[0] * rd[2]);
    const double g2 = rm[2] + (B.t[0] * rd[1] - B.t[1] * rd[0]);
    o.J[c] = r0.v[0] * g0 + r0.v[1] * g1 + r0.v[2] * g2;
    o.J[4 + c] = r1.v[0] * g0 + r1.v[1] * g1 + r1.v[2] * g2;
    o.Jv[c] = rvp.v[0] * rd[0] + rvp.v[1] * rd[1] + rvp.v[2] * rd[2];
  }
}


// The same block with the derivatives written out (the kernel's form). The VALUE path keeps the reference's operation
// order (Line_ImgFromCam normalisation -> Ceres_CosineWeightedPerpendicularDist2D_1D, EPS under every square root), so
// residuals and costs are those of the dual-number form to the last bit; the derivatives with respect to the camera-frame
// moment m_c follow from (all EPS terms kept):
//   q = l / cn, cn = sqrt(|l|^2 + EPS)                    =>  dq/dl = (I - q q^T) / cn
//   e_k = (q . (x_k, y_k, 1)) / dn, dn = sqrt(q0^2 + q1^2 + EPS)
//                                                         =>  de_k/dq = (x_k, y_k, 1)/dn - e_k (q0, q1, 0)/dn^2
//   dir = (-q1, q0)/dn, n1 = sqrt(|dir|^2 + EPS), T = dir . (sx, sy), cos = |T| / (n1 n2)
//                                                         =>  dcos/dq = sgn(T) [ (sx ddir0 + sy ddir1)/(n1 n2) - T dn1/(n1^2 n2) ]
//   w = exp(alpha (1 - cos)), r_k = e_k w                 =>  dr_k/dq = w de_k/dq - alpha w e_k dcos/dq
//   l = cof(K) m_c                                        =>  dr/dm_c = cof(K)^T dr/dl
// about 4x fewer operations than carrying 3-wide duals through the chain, and no dual registers.
LM_HD void eval_block(const LMBlockDev &B, const LineShared &L, double alpha, bool want_jac, BlockEval &o) {
  const double EPS = consts<double>::eps();
  double Rd[3], Rm[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Rd[i] = B.R[3 * i] * L.d[0] + B.R[3 * i + 1] * L.d[1] + B.R[3 * i + 2] * L.d[2];
    Rm[i] = B.R[3 * i] * L.m[0] + B.R[3 * i + 1] * L.m[1] + B.R[3 * i + 2] * L.m[2];
  }
  const double mc0 = Rm[0] + (B.t[1] * Rd[2] - B.t[2] * Rd[1]);
  const double mc1 = Rm[1] + (B.t[2] * Rd[0] - B.t[0] * Rd[2]);
  const double mc2 = Rm[2] + (B.t[0] * Rd[1] - B.t[1] * Rd[0]);
  const double fx = B.k[0], fy = B.k[1], cx = B.k[2], cy = B.k[3];
  // value path: operation for operation as in eval_block_dual
  const double l0 = mc0 * fy, l1 = mc1 * fx, l2 = mc2 * (fx * fy) - mc0 * (fy * cx) - mc1 * (fx * cy);
  const double cn = sqrt(l0 * l0 + l1 * l1 + l2 * l2 + EPS);
  const double icn = 1.0 / cn;
  const double q0 = l0 * icn, q1 = l1 * icn, q2 = l2 * icn; // (Dual operator/ multiplies by the reciprocal too)
  const double dn = sqrt(q0 * q0 + q1 * q1 + EPS);
  const double idn = 1.0 / dn;
  const double dir0 = (-q1) * idn, dir1 = q0 * idn;
  const double sx = B.p[2] - B.p[0], sy = B.p[3] - B.p[1];
  const double n1 = sqrt(dir0 * dir0 + dir1 * dir1 + EPS);
  const double n2 = sqrt(sx * sx + sy * sy + EPS);
  const double T = dir0 * sx + dir1 * sy;
  const double in12 = 1.0 / (n1 * n2);
  const double cs = T * in12;
  double cosine = cs < 0 ? -cs : cs;
  const bool clamped = cosine > 1.0;
  if (clamped) cosine = 1.0;
  const double w = exp((1.0 - cosine) * alpha);
  const double h0 = q0 * B.p[0] + q1 * B.p[1] + q2, h1 = q0 * B.p[2] + q1 * B.p[3] + q2;
  const double e0 = h0 * idn, e1 = h1 * idn;
  o.r[0] = e0 * w;
  o.r[1] = e1 * w;
  // VP residual (VPConstraintsFunctor): sine between R d and the VP direction, value path as in the dual form
  double rv = 0.0, gv[3] = {0.0, 0.0, 0.0}; // d rv / d (R d)
  if (B.wvp > 0.0) {
    const double nv1 = sqrt(Rd[0] * Rd[0] + Rd[1] * Rd[1] + Rd[2] * Rd[2] + EPS);
    const double nv2 = sqrt(B.vdir[0] * B.vdir[0] + B.vdir[1] * B.vdir[1] + B.vdir[2] * B.vdir[2] + EPS);
    const double inv1 = 1.0 / nv1;
    const double ax = Rd[0] * inv1, ay = Rd[1] * inv1, az = Rd[2] * inv1;
    const double bx = B.vdir[0] / nv2, by = B.vdir[1] / nv2, bz = B.vdir[2] / nv2;
    const double c0 = ay * bz - az * by, c1 = az * bx - ax * bz, c2 = ax * by - ay * bx;
    rv = sqrt(c0 * c0 + c1 * c1 + c2 * c2 + EPS);
    if (rv > 1.0) rv = 1.0; // (constant: zero derivative)
    else if (want_jac) {
      // rv^2 = |a x b|^2 + EPS; d|a x b|^2/da = 2 (b x (a x b))... written out: (c x b) with c = a x b gives -(b x c):
      // d(|c|^2)/da = 2 (b x c)^T-wise => grad_a = (b x c) * (-1)?  derive directly: c = a x b, dc/da applied to c:
      // d(|c|^2)/da_i = 2 c . (e_i x b) = 2 e_i . (b x c)
      const double ga0 = (by * c2 - bz * c1), ga1 = (bz * c0 - bx * c2), ga2 = (bx * c1 - by * c0); // b x c
      const double irv = 1.0 / rv;
      const double ha0 = ga0 * irv, ha1 = ga1 * irv, ha2 = ga2 * irv; // d rv / d a
      // a = x / sqrt(|x|^2 + EPS): da/dx = (I - a a^T) / nv1
      const double hd = ha0 * ax + ha1 * ay + ha2 * az;
      gv[0] = (ha0 - hd * ax) * inv1; gv[1] = (ha1 - hd * ay) * inv1; gv[2] = (ha2 - hd * az) * inv1;
    }
  }
  o.rv = rv;
  if (!want_jac) return;
  // gradients of r0, r1 with respect to q
  const double idn2 = idn * idn;
  double dc0 = 0.0, dc1 = 0.0; // d cosine / d (q0, q1); q2 does not enter
  if (!clamped) {
    // ddir0/dq = (0, -1/dn) - dir0 (q0, q1)/dn^2 ; ddir1/dq = (1/dn, 0) - dir1 (q0, q1)/dn^2
    const double d00 = -dir0 * q0 * idn2, d01 = -idn - dir0 * q1 * idn2;
    const double d10 = idn - dir1 * q0 * idn2, d11 = -dir1 * q1 * idn2;
    const double in1 = 1.0 / n1;
    const double dn1_0 = (dir0 * d00 + dir1 * d10) * in1, dn1_1 = (dir0 * d01 + dir1 * d11) * in1;
    const double sg = cs < 0 ? -1.0 : 1.0;
    dc0 = sg * ((sx * d00 + sy * d10) * in12 - T * dn1_0 * in12 * in1);
    dc1 = sg * ((sx * d01 + sy * d11) * in12 - T * dn1_1 * in12 * in1);
  }
  const double aw = alpha * w;
  double g0[3], g1[3]; // d r_k / d q
  g0[0] = w * (B.p[0] * idn - e0 * q0 * idn2) - aw * e0 * dc0;
  g0[1] = w * (B.p[1] * idn - e0 * q1 * idn2) - aw * e0 * dc1;
  g0[2] = w * idn;
  g1[0] = w * (B.p[2] * idn - e1 * q0 * idn2) - aw * e1 * dc0;
  g1[1] = w * (B.p[3] * idn - e1 * q1 * idn2) - aw * e1 * dc1;
  g1[2] = w * idn;
  // -> l: (g - (g.q) q)/cn ; -> m_c: cof(K)^T
  double m0[3], m1[3];
  {
    const double s0 = g0[0] * q0 + g0[1] * q1 + g0[2] * q2, s1 = g1[0] * q0 + g1[1] * q1 + g1[2] * q2;
    const double a0 = (g0[0] - s0 * q0) * icn, b0 = (g0[1] - s0 * q1) * icn, c0 = (g0[2] - s0 * q2) * icn;
    const double a1 = (g1[0] - s1 * q0) * icn, b1 = (g1[1] - s1 * q1) * icn, c1 = (g1[2] - s1 * q2) * icn;
    m0[0] = fy * a0 - (fy * cx) * c0; m0[1] = fx * b0 - (fx * cy) * c0; m0[2] = (fx * fy) * c0;
    m1[0] = fy * a1 - (fy * cx) * c1; m1[1] = fx * b1 - (fx * cy) * c1; m1[2] = (fx * fy) * c1;
  }
  // G = d m_c / d local (3x4) = R Dm + t x (R Dd)
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    double rd[3], rm[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      rd[i] = B.R[3 * i] * L.dv[0][c] + B.R[3 * i + 1] * L.dv[1][c] + B.R[3 * i + 2] * L.dv[2][c];
      rm[i] = B.R[3 * i] * L.mv[0][c] + B.R[3 * i + 1] * L.mv[1][c] + B.R[3 * i + 2] * L.mv[2][c];
    }
    const double gg0 = rm[0] + (B.t[1] * rd[2] - B.t[2] * rd[1]);
    const double gg1 = rm[1] + (B.t[2] * rd[0] - B.t[0] * rd[2]);
    const double gg2 = rm[2] + (B.t[0] * rd[1] - B.t[1] * rd[0]);
    o.J[c] = m0[0] * gg0 + m0[1] * gg1 + m0[2] * gg2;
    o.J[4 + c] = m1[0] * gg0 + m1[1] * gg1 + m1[2] * gg2;
    o.Jv[c] = gv[0] * rd[0] + gv[1] * rd[1] + gv[2] * rd[2];
  }
}

} // namespace lm
