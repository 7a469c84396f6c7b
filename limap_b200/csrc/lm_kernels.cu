// lm_kernels.cu — batched per-track Levenberg-Marquardt line refinement for sm_100a.
//
// Replaces the Ceres problems of
//   RefinementEngine::{SetUp,Solve}        (optimize/line_refinement/refine.cc:129-178)
//   HybridBAEngine::{SetUp,Solve}          (optimize/hybrid_bundle_adjustment/hybrid_bundle_adjustment.cc:199-264)
// for the line-only, constant-camera case: the cost is block-separable per track, so every track is an
// independent 4-dof problem (Quaternion manifold on uvec, Sphere<2> on wvec).
//
// One warp owns one track for the whole solve (all iterations in one launch): the track's supporting
// blocks are pre-digested once into 21 doubles each (2D endpoints, K, R, t, loss weight) that stay in
// L1/L2, lanes evaluate residual blocks, and the 4x4 normal equations are accumulated with warp-shuffle
// all-reductions so that every lane holds J^T J / J^T r and the trust-region logic stays warp-uniform.
//
// Residual: GeometricRefinementFunctor (optimize/line_refinement/cost_functions.h:129-194) =
//   MinimalPluckerToPlucker (ceresbase/line_transforms.h:9-29) -> Line_WorldToPixel
//   (ceresbase/line_projection.h:51-80) -> Ceres_CosineWeightedPerpendicularDist2D_1D (:107-127),
// loss ScaledLoss(CauchyLoss(0.25), |seg|/30) (refine.cc:77-78, base/linetrack.cc:315-322).
// Derivatives: the reference differentiates with 6-wide Ceres Jets through the whole chain; here the chain
// is split at the camera-frame Pluecker moment m_c = R m + t x (R d): d(m, d)/d(local) is computed once
// per evaluation per track with 4-wide duals, and each lane carries 3-wide duals from m_c to the residual.
// The solver restates Ceres' trust-region LM (DESIGN.md "LM recipe" lists the steps and their provenance).
#include "lm_kernels.cuh"
#include "lm_block.cuh"
#include <algorithm>
#include <cfloat>
#include <cstdlib>

namespace lm {

static constexpr int kSegMax = 128; // supports per track handled by the in-kernel segment cut

LM_D double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

struct Normal {
  double A[10]; // upper triangle of J^T J (scaled columns): 00 01 02 03 11 12 13 22 23 33
  double g[4];  // J^T r
  double cost;
};

// Per-warp solver state in shared memory. One warp owns one track; everything below is warp-uniform, so it lives here
// instead of in every lane's registers (the kernel used to need 255 registers, i.e. 8 warps per SM).
struct WarpState {
  LineShared L;
  double x[6], cand[6], scale[4], diag[4];
  Normal N;  // at x
  Normal Nc; // at the candidate point
  double cn2[4];
  // trust-region scalars (warp-uniform; every lane writes the same value, so no ordering is needed): kept here and
  // accessed through volatile references so that they are not live in registers across the track evaluations
  double cost0, cost, radius, decrease_factor, mcc, sn;
  double hh_e[3], hh_x[3]; // Householder triple of the sphere part at the last evaluated point / at x (lm_block.cuh)
};

// Evaluate the whole track at x: cost (always) and, if want_jac, the loss-corrected normal equations with the
// Jacobi column scaling applied (scale may be NULL -> unscaled, used to initialise the scaling). Results go to *out
// (and column norms to cn2_out) in shared memory; every lane may read them after the trailing __syncwarp.
// `acc` = this thread's column of the CTA's accumulator table ([19][kLmThreads] doubles in shared memory: the partial
// sums of J^T J (10), J^T r (4), cost (1), column norms (4) stay out of the registers while a block is evaluated).
static constexpr int kLmThreads = 128;
static constexpr int kAccStride = kLmThreads + 1; // doubles between two accumulator rows (bank spread for the row sums)
LM_D void eval_track(const LMBlockDev *blocks, int S, const double *x, double alpha, double bq, bool want_jac,
                     const double *scale, LineShared &Ls, Normal *out, double *cn2_out, double *acc, double *hh_out) {
  const int lane = threadIdx.x & 31;
  {
    double xr[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) xr[i] = x[i];
    Dual<1> dd[3], mm[3];
    double hh[3];
    line_from_minimal_col(xr, lane & 3, want_jac, dd, mm, hh); // lane c < 4 holds tangent column c (all lanes: the values)
    __syncwarp();
    if (lane == 4) { hh_out[0] = hh[0]; hh_out[1] = hh[1]; hh_out[2] = hh[2]; }
    if (lane < 4) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (lane == 0) { Ls.d[i] = dd[i].a; Ls.m[i] = mm[i].a; }
        Ls.dv[i][lane] = dd[i].v[0];
        Ls.mv[i][lane] = mm[i].v[0];
      }
    }
    __syncwarp();
  }
  // Accumulators: one column of the CTA's table per thread, written through a volatile pointer so that they really
  // stay in shared memory (promoted to registers they would be live across the whole block evaluation: 38 registers).
  volatile double *va = acc;
#pragma unroll
  for (int i = 0; i < 19; ++i) va[i * kAccStride] = 0.0;
  const double cq = 1.0 / bq;
  for (int k = lane; k < S; k += 32) {
    const LMBlockDev &B = blocks[k];
    BlockEval e;
    eval_block(B, Ls, alpha, want_jac, e);
    const double Bw = B.w, Bwvp = B.wvp;
    // ScaledLoss(CauchyLoss, w) + ceres Corrector
    const double s = e.r[0] * e.r[0] + e.r[1] * e.r[1];
    const double sum = 1.0 + s * cq, inv = 1.0 / sum;
    const double rho0 = Bw * bq * log(sum), rho1 = Bw * fmax(DBL_MIN, inv), rho2 = Bw * (-cq * (inv * inv));
    {
      double cost = 0.5 * rho0;
      if (Bwvp > 0.0) cost += 0.5 * Bwvp * e.rv * e.rv; // ScaledLoss(TrivialLoss, w * vp_multiplier)
      va[14 * kAccStride] += cost;
    }
    if (!want_jac) continue;
    const double sqrt_rho1 = sqrt(rho1);
    double residual_scaling = sqrt_rho1, alpha_sq_norm = 0.0;
    if (!(s == 0.0 || rho2 <= 0.0)) {
      const double D = 1.0 + 2.0 * s * rho2 / rho1;
      const double al = 1.0 - sqrt(D);
      residual_scaling = sqrt_rho1 / (1 - al);
      alpha_sq_norm = al / s;
    }
    double J0[4], J1[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (alpha_sq_norm == 0.0) { J0[c] = sqrt_rho1 * e.J[c]; J1[c] = sqrt_rho1 * e.J[4 + c]; }
      else {
        const double rtj = e.r[0] * e.J[c] + e.r[1] * e.J[4 + c];
        J0[c] = sqrt_rho1 * (e.J[c] - alpha_sq_norm * e.r[0] * rtj);
        J1[c] = sqrt_rho1 * (e.J[4 + c] - alpha_sq_norm * e.r[1] * rtj);
      }
    }
    const double r0 = e.r[0] * residual_scaling, r1 = e.r[1] * residual_scaling;
    double J2[4] = {0, 0, 0, 0}, r2 = 0.0;
    if (Bwvp > 0.0) {
      const double sq = sqrt(Bwvp);
      r2 = sq * e.rv;
#pragma unroll
      for (int c = 0; c < 4; ++c) J2[c] = sq * e.Jv[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      va[(15 + c) * kAccStride] += J0[c] * J0[c] + J1[c] * J1[c] + J2[c] * J2[c];
      if (scale) { const double sc = scale[c]; J0[c] *= sc; J1[c] *= sc; J2[c] *= sc; }
    }
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      va[(10 + a) * kAccStride] += J0[a] * r0 + J1[a] * r1 + J2[a] * r2;
#pragma unroll
      for (int b = a; b < 4; ++b) { va[idx * kAccStride] += J0[a] * J0[b] + J1[a] * J1[b] + J2[a] * J2[b]; ++idx; }
    }
  }
  __syncwarp();
  // Reduction across the warp's 32 columns: lane i adds up accumulator i (4 partial sums, fixed order) instead of 19
  // butterfly reductions of 5 shuffle rounds each. The row stride (kAccStride = 129 doubles) keeps the 19 lanes on
  // different banks.
  if (lane < 19) {
    const volatile double *row = acc - lane + lane * kAccStride; // this warp's 32 columns of row `lane`
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int j = 0; j < 32; j += 4) { s0 += row[j]; s1 += row[j + 1]; s2 += row[j + 2]; s3 += row[j + 3]; }
    const double v = (s0 + s1) + (s2 + s3);
    if (lane == 14) out->cost = v;
    if (want_jac) {
      if (lane < 10) out->A[lane] = v;
      else if (lane < 14) out->g[lane - 10] = v;
      else if (lane > 14 && cn2_out) cn2_out[lane - 15] = v;
    }
  }
  __syncwarp();
}

// Solve (A + diag(dg)) x = b for the symmetric positive definite 4x4 system of one LM step. Ceres' DENSE_NORMAL_CHOLESKY
// is an LL^T factorisation; this is the square-root-free L D L^T form of the same factorisation (x agrees to rounding;
// "not positive definite" <=> a pivot d_k <= 0 <=> LL^T meets a non-positive diagonal): 4 reciprocals instead of 4 square
// roots and 14 divisions in what is a strictly sequential, warp-uniform chain.
LM_D bool chol_solve4(const double Au[10], const double dg[4], const double b[4], double x[4]) {
  // upper-triangle storage 00 01 02 03 11 12 13 22 23 33
  const double a00 = Au[0] + dg[0], a10 = Au[1], a20 = Au[2], a30 = Au[3];
  const double a11 = Au[4] + dg[1], a21 = Au[5], a31 = Au[6];
  const double a22 = Au[7] + dg[2], a32 = Au[8];
  const double a33 = Au[9] + dg[3];
  const double d0 = a00;
  if (!(d0 > 0)) return false;
  const double i0 = 1.0 / d0;
  const double l10 = a10 * i0, l20 = a20 * i0, l30 = a30 * i0;
  const double d1 = a11 - l10 * a10;
  if (!(d1 > 0)) return false;
  const double i1 = 1.0 / d1;
  const double u21 = a21 - l20 * a10, u31 = a31 - l30 * a10; // (L D) entries of column 1
  const double l21 = u21 * i1, l31 = u31 * i1;
  const double d2 = a22 - l20 * a20 - l21 * u21;
  if (!(d2 > 0)) return false;
  const double i2 = 1.0 / d2;
  const double u32 = a32 - l30 * a20 - l31 * u21;
  const double l32 = u32 * i2;
  const double d3 = a33 - l30 * a30 - l31 * u31 - l32 * u32;
  if (!(d3 > 0)) return false;
  const double i3 = 1.0 / d3;
  // L y = b, D z = y, L^T x = z
  const double y0 = b[0];
  const double y1 = b[1] - l10 * y0;
  const double y2 = b[2] - l20 * y0 - l21 * y1;
  const double y3 = b[3] - l30 * y0 - l31 * y1 - l32 * y2;
  x[3] = y3 * i3;
  x[2] = y2 * i2 - l32 * x[3];
  x[1] = y1 * i1 - l21 * x[2] - l31 * x[3];
  x[0] = y0 * i0 - l10 * x[1] - l20 * x[2] - l30 * x[3];
  return true;
}

// MB = resident CTAs per SM the register allocation is bounded for (2: 255 registers, 3: 168, 4: 128).
template <int MB> __global__ void __launch_bounds__(128, MB) lm_refine_kernel(const __grid_constant__ LMParams p) {
  __shared__ WarpState s_ws[4];
  __shared__ double s_acc[19 * kAccStride];
  double *acc = s_acc + threadIdx.x;
  const int warp_in_block = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // Persistent warps: a warp fetches its next track when it is done. Solves take 10..max_num_iterations iterations, so a
  // static one-track-per-warp grid leaves every CTA slot waiting for its slowest track.
  for (;;) {
  long long t = 0;
  if (lane == 0) t = (long long)atomicAdd(p.next_track, 1ull);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t >= p.T) break;
  const int64_t s0 = p.sup_off[t];
  const int S = (int)(p.sup_off[t + 1] - s0);
  const LMBlockDev *blocks = p.blocks + s0;
  WarpState &ws = s_ws[warp_in_block];
  if (lane < 6) ws.x[lane] = p.x0[6 * t + lane];
  __syncwarp();
  const double bq = p.cauchy_scale * p.cauchy_scale;
  int it = 0, successful = 0, term = 0;
  volatile double &cost0 = ws.cost0, &cost = ws.cost, &radius = ws.radius, &decrease_factor = ws.decrease_factor;
  volatile double &mcc = ws.mcc, &sn = ws.sn;
  if (S == 0 || !p.active[t]) {
    eval_track(blocks, S, ws.x, p.geometric_alpha, bq, false, nullptr, ws.L, &ws.N, nullptr, acc, ws.hh_e);
    { const double c_ = ws.N.cost; cost0 = c_; cost = c_; }
  } else {
    // iteration 0: evaluate, fix the Jacobi scaling 1/(1+||J_col||)
    eval_track(blocks, S, ws.x, p.geometric_alpha, bq, true, nullptr, ws.L, &ws.N, ws.cn2, acc, ws.hh_e);
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) ws.scale[c] = 1.0 / (1.0 + sqrt(ws.cn2[c]));
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        ws.N.g[a] *= ws.scale[a];
#pragma unroll
        for (int b = a; b < 4; ++b) ws.N.A[idx++] *= ws.scale[a] * ws.scale[b];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) ws.diag[c] = 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) ws.hh_x[c] = ws.hh_e[c];
    }
    __syncwarp();
    { const double c_ = ws.N.cost; cost0 = c_; cost = c_; }
    radius = 1e4; decrease_factor = 2.0;
    bool reuse_diagonal = false;
    int invalid = 0;
    while (true) {
      if (it >= p.max_num_iterations) { term = 1; break; }
      if (radius <= 1e-32) { term = 2; break; }
      {
        // max |g_c / scale_c| <= 0 (the unscaled gradient is exactly zero): scale_c is in (0, 1], so the quotient is zero
        // exactly when g_c is
        double gmax = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) gmax = fmax(gmax, fabs(ws.N.g[c]));
        if (gmax <= 0.0) { term = 3; break; }
      }
      ++it;
      if (!reuse_diagonal) {
        __syncwarp();
        if (lane == 0) {
          const double dgi[4] = {ws.N.A[0], ws.N.A[4], ws.N.A[7], ws.N.A[9]};
#pragma unroll
          for (int c = 0; c < 4; ++c) ws.diag[c] = fmin(fmax(dgi[c], 1e-6), 1e32);
        }
        __syncwarp();
      }
      reuse_diagonal = true;
      // the 4x4 solve is warp-uniform: every lane computes it from the shared state (no divergence, no extra issue slots)
      double step[4];
      bool ok;
      double mcc_l = 0; // (kept in a register until the test below; the shared copy is written once, for after the evaluation)
      {
        double An[10], gn[4], dg[4];
#pragma unroll
        for (int c = 0; c < 10; ++c) An[c] = ws.N.A[c];
        const double rad_ = radius;
#pragma unroll
        for (int c = 0; c < 4; ++c) { gn[c] = ws.N.g[c]; dg[c] = ws.diag[c] / rad_; } // (4 independent divisions: one latency)
        ok = chol_solve4(An, dg, gn, step);
#pragma unroll
        for (int c = 0; c < 4; ++c) { step[c] = -step[c]; if (!isfinite(step[c])) ok = false; }
        if (ok) {
          // model_cost_change = -(step . g + step^T (J^T J) step / 2)
          double sg = 0, sAs = 0;
          int idx = 0;
#pragma unroll
          for (int a = 0; a < 4; ++a) {
            sg += step[a] * gn[a];
#pragma unroll
            for (int b = a; b < 4; ++b) { sAs += ((a == b) ? 1.0 : 2.0) * step[a] * step[b] * An[idx]; ++idx; }
          }
          mcc_l = -(sg + 0.5 * sAs);
        }
      }
      if (!ok || !(mcc_l > 0.0)) {
        if (++invalid >= p.max_invalid) { term = 4; break; }
        { const double r_ = radius, f_ = decrease_factor; __syncwarp(); radius = r_ / f_; decrease_factor = f_ * 2.0; } reuse_diagonal = true;
        continue;
      }
      invalid = 0;
      mcc = mcc_l; // every lane writes the same value, once per iteration; read after the evaluation's barriers
      {
        double snl = 0;
        double xr[6], delta[4], cand[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) xr[c] = ws.x[c];
#pragma unroll
        for (int c = 0; c < 4; ++c) delta[c] = step[c] * ws.scale[c];
        quat_plus(xr, delta, cand);
        sphere2_plus(xr + 4, delta[3], cand + 4, ws.hh_x);
#pragma unroll
        for (int c = 0; c < 6; ++c) snl += (xr[c] - cand[c]) * (xr[c] - cand[c]);
        sn = snl;
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 6; ++c) ws.cand[c] = cand[c];
        }
        __syncwarp();
      }
      // One evaluation per iteration: the candidate point is evaluated WITH its Jacobian; if the step is accepted the
      // normal equations are already there (a second evaluation at the same point would reproduce them bit for bit:
      // the value parts of the duals do not depend on want_jac).
      eval_track(blocks, S, ws.cand, p.geometric_alpha, bq, true, ws.scale, ws.L, &ws.Nc, nullptr, acc, ws.hh_e);
      const double cost_c = ws.Nc.cost;
      if (!(sqrt(sn) > 0.0)) { term = 5; break; }
      const double cost_x = cost;
      if (!(fabs(cost_x - cost_c) > 0.0)) { term = 6; break; }
      const double rel = (cost_x - cost_c) / mcc;
      if (rel > 1e-3) {
        __syncwarp();
        if (lane < 6) ws.x[lane] = ws.cand[lane];
        if (lane < 10) ws.N.A[lane] = ws.Nc.A[lane];
        if (lane < 4) ws.N.g[lane] = ws.Nc.g[lane];
        if (lane >= 4 && lane < 7) ws.hh_x[lane - 4] = ws.hh_e[lane - 4];
        __syncwarp();
        cost = cost_c;
        const double tq = 2.0 * rel - 1.0;
        { const double r_ = radius; __syncwarp(); radius = fmin(1e16, r_ / fmax(1.0 / 3.0, 1.0 - tq * tq * tq)); }
        decrease_factor = 2.0; reuse_diagonal = false;
        ++successful;
      } else {
        { const double r_ = radius, f_ = decrease_factor; __syncwarp(); radius = r_ / f_; decrease_factor = f_ * 2.0; } reuse_diagonal = true;
      }
    }
  }
  __syncwarp();
  double x[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) x[i] = ws.x[i];
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 6; ++i) p.x_out[6 * t + i] = x[i];
    p.iters[2 * t] = it; p.iters[2 * t + 1] = successful;
    p.cost[2 * t] = cost0; p.cost[2 * t + 1] = cost;
    p.term[t] = term;
  }
  // GetLineSegmentFromInfiniteLine3d(inf_line, track.line3d_list, num_outliers) (base/infinite_line.cc:265-287)
  // on the infinite line of MinimalInfiniteLine3d::GetInfiniteLine (:220-231)
  if (p.seg_out) {
    __shared__ double s_vals[4][2 * kSegMax];
    double *vals = s_vals[warp_in_block];
    const int n2 = 2 * S;
    const bool ok = p.line3d && S > 0 && S <= kSegMax && p.num_outliers >= 0 && p.num_outliers < n2 &&
                    n2 - 1 - p.num_outliers >= 0;
    if (!ok) {
      if (lane < 6) p.seg_out[6 * t + lane] = __longlong_as_double(0x7ff8000000000000ll);
      continue;
    }
    // limap QuaternionToRotationMatrix (base/pose.cc:12-18): normalised quaternion
    double q[4];
    {
      const double nq = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
      if (nq == 0) { q[0] = 1.0; q[1] = x[1]; q[2] = x[2]; q[3] = x[3]; }
      else { q[0] = x[0] / nq; q[1] = x[1] / nq; q[2] = x[2] / nq; q[3] = x[3] / nq; }
    }
    const double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz, twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    const double d0 = 1 - (tyy + tzz), d1 = txy + twz, d2 = txz - twy;       // Q.col(0)
    const double f = fabs(x[5]) / fabs(x[4]);
    const double m0 = (txy - twz) * f, m1 = (1 - (txx + tzz)) * f, m2 = (tyz + twx) * f; // Q.col(1) * |w1|/|w0|
    const double *l3 = p.line3d + 6 * s0;
    // p_ref = point_projection(line3ds[0].start): q + d x (m + d x q) (:73-78)
    double pr0, pr1, pr2;
    {
      const double a0 = l3[0], a1 = l3[1], a2 = l3[2];
      const double c0 = m0 + (d1 * a2 - d2 * a1), c1 = m1 + (d2 * a0 - d0 * a2), c2 = m2 + (d0 * a1 - d1 * a0);
      pr0 = a0 + (d1 * c2 - d2 * c1); pr1 = a1 + (d2 * c0 - d0 * c2); pr2 = a2 + (d0 * c1 - d1 * c0);
    }
    for (int k = lane; k < n2; k += 32) {
      const double *e = l3 + 3 * k; // endpoint k (start/end interleaved)
      vals[k] = (e[0] - pr0) * d0 + (e[1] - pr1) * d1 + (e[2] - pr2) * d2;
    }
    __syncwarp();
    // order statistics by rank counting (values[num_outliers] and values[2S-1-num_outliers] of the sorted list)
    const int ka = p.num_outliers, kb = n2 - 1 - p.num_outliers;
    for (int k = lane; k < n2; k += 32) {
      const double v = vals[k];
      int rank = 0;
      for (int j = 0; j < n2; ++j) { const double u = vals[j]; rank += (u < v) || (u == v && j < k); }
      if (rank == ka) { p.seg_out[6 * t] = pr0 + d0 * v; p.seg_out[6 * t + 1] = pr1 + d1 * v; p.seg_out[6 * t + 2] = pr2 + d2 * v; }
      if (rank == kb) { p.seg_out[6 * t + 3] = pr0 + d0 * v; p.seg_out[6 * t + 4] = pr1 + d1 * v; p.seg_out[6 * t + 5] = pr2 + d2 * v; }
    }
    __syncwarp(); // vals[] is reused by the warp's next track
  }
  } // track loop
}

// supports -> digested blocks: R from qvec via ceres::QuaternionToRotation (normalising), loss weight |seg|/30
__global__ void lm_prepare_blocks_kernel(const double *__restrict__ segs, const int32_t *__restrict__ sup_view,
                                         const double *__restrict__ kvec, const double *__restrict__ qvec,
                                         const double *__restrict__ tvec, const double *__restrict__ sup_vp,
                                         double vp_multiplier, int64_t n, LMBlockDev *__restrict__ out) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int v = sup_view[k];
  LMBlockDev B;
#pragma unroll
  for (int i = 0; i < 4; ++i) { B.p[i] = segs[4 * k + i]; B.k[i] = kvec[4 * v + i]; }
  const double a = qvec[4 * v], b = qvec[4 * v + 1], c = qvec[4 * v + 2], d = qvec[4 * v + 3];
  const double aa = a * a, ab = a * b, ac = a * c, ad = a * d, bb = b * b, bc = b * c, bd = b * d, cc = c * c, cd = c * d, dd = d * d;
  const double nr = 1.0 / (aa + bb + cc + dd);
  B.R[0] = (aa + bb - cc - dd) * nr; B.R[1] = 2 * (bc - ad) * nr; B.R[2] = 2 * (ac + bd) * nr;
  B.R[3] = 2 * (ad + bc) * nr; B.R[4] = (aa - bb + cc - dd) * nr; B.R[5] = 2 * (cd - ab) * nr;
  B.R[6] = 2 * (bd - ac) * nr; B.R[7] = 2 * (ab + cd) * nr; B.R[8] = (aa - bb - cc + dd) * nr;
#pragma unroll
  for (int i = 0; i < 3; ++i) B.t[i] = tvec[3 * v + i];
  const double dx = B.p[0] - B.p[2], dy = B.p[1] - B.p[3];
  B.w = sqrt(dx * dx + dy * dy) / 30.0; // ComputeLineWeights (base/linetrack.cc:315-322)
  B.vdir[0] = B.vdir[1] = B.vdir[2] = 0.0;
  B.wvp = 0.0;
  if (sup_vp && !isnan(sup_vp[3 * k])) {
    const double v0 = sup_vp[3 * k], v1 = sup_vp[3 * k + 1], v2 = sup_vp[3 * k + 2];
    const double e0 = v0 / B.k[0] - B.k[2] / B.k[0] * v2, e1 = v1 / B.k[1] - B.k[3] / B.k[1] * v2, e2 = v2;
    const double nn = sqrt(e0 * e0 + e1 * e1 + e2 * e2 + consts<double>::eps());
    B.vdir[0] = e0 / nn; B.vdir[1] = e1 / nn; B.vdir[2] = e2 / nn;
    B.wvp = B.w * vp_multiplier;
  }
  out[k] = B;
}

// Per-track prologue on the device (one thread per track): the minimal parameterisation of the start line
// (MinimalInfiniteLine3d(InfiniteLine3d(Line3d)), base/infinite_line.cc:67-71,180-218; Q -> quaternion as Eigen's
// Quaterniond(Matrix3d) does) and the constant-track flag of ParameterizeLines (hybrid_bundle_adjustment.cc:106-123:
// tracks seen in fewer than min_num_images distinct images keep their parameter blocks constant).
__global__ void lm_track_prologue_kernel(const double *__restrict__ line_init, const int64_t *__restrict__ sup_off,
                                         const int32_t *__restrict__ sup_view, int64_t T, int min_num_images,
                                         double *__restrict__ x0, uint8_t *__restrict__ active, int *__restrict__ err) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= T) return;
  const double *l = line_init + 6 * t;
  double out[6];
  {
    double sx = l[0], sy = l[1], sz = l[2];
    double ax = l[3] - sx, ay = l[4] - sy, az = l[5] - sz;
    const double an2 = ax * ax + ay * ay + az * az;
    if (!(an2 > 0) && sup_off[t + 1] > sup_off[t]) *err = 1; // CHECK_GT(line.length(), 0)
    if (an2 > 0) { const double an = sqrt(an2); ax /= an; ay /= an; az /= an; } // direction() = normalized()
    const double bx = sy * az - sz * ay, by = sz * ax - sx * az, bz = sx * ay - sy * ax; // m = p x d
    const double bn = sqrt(bx * bx + by * by + bz * bz);
    const double den = sqrt(1.0 + bn * bn);
    out[4] = 1.0 / den;
    out[5] = bn / den;
    const double na = sqrt(ax * ax + ay * ay + az * az);
    const double q0x = ax / na, q0y = ay / na, q0z = az / na;
    double q1x, q1y, q1z, q2x, q2y, q2z;
    if (bn > 1e-12) {
      q1x = bx / bn; q1y = by / bn; q1z = bz / bn;
      const double cx = ay * bz - az * by, cy = az * bx - ax * bz, cz = ax * by - ay * bx;
      const double cn = sqrt(cx * cx + cy * cy + cz * cz);
      q2x = cx / cn; q2y = cy / cn; q2z = cz / cn;
    } else {
      const double av[3] = {ax, ay, az};
      int best = 0;
      if (fabs(av[1]) > fabs(av[0])) best = 1;
      if (fabs(av[2]) > fabs(av[best])) best = 2;
      const int i1 = (best + 1) % 3, i2 = (best + 2) % 3;
      double bp[3];
      bp[i1] = 1.0; bp[i2] = 1.0; bp[best] = -(av[i1] * bp[i1] + av[i2] * bp[i2]) / av[best];
      const double pn = sqrt(bp[0] * bp[0] + bp[1] * bp[1] + bp[2] * bp[2]);
      q1x = bp[0] / pn; q1y = bp[1] / pn; q1z = bp[2] / pn;
      const double cx = ay * bp[2] - az * bp[1], cy = az * bp[0] - ax * bp[2], cz = ax * bp[1] - ay * bp[0];
      const double cn = sqrt(cx * cx + cy * cy + cz * cz);
      q2x = cx / cn; q2y = cy / cn; q2z = cz / cn;
    }
    const double R[3][3] = {{q0x, q1x, q2x}, {q0y, q1y, q2y}, {q0z, q1z, q2z}};
    double tr = R[0][0] + R[1][1] + R[2][2];
    if (tr > 0) {
      tr = sqrt(tr + 1.0);
      out[0] = 0.5 * tr;
      tr = 0.5 / tr;
      out[1] = (R[2][1] - R[1][2]) * tr; out[2] = (R[0][2] - R[2][0]) * tr; out[3] = (R[1][0] - R[0][1]) * tr;
    } else {
      int i = 0;
      if (R[1][1] > R[0][0]) i = 1;
      if (R[2][2] > R[i][i]) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      tr = sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
      double v[3];
      v[i] = 0.5 * tr;
      tr = 0.5 / tr;
      out[0] = (R[k][j] - R[j][k]) * tr;
      v[j] = (R[j][i] + R[i][j]) * tr;
      v[k] = (R[k][i] + R[i][k]) * tr;
      out[1] = v[0]; out[2] = v[1]; out[3] = v[2];
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) x0[6 * t + i] = out[i];
  // distinct images among the supports (tracks are short: quadratic scan)
  const int64_t a = sup_off[t], b = sup_off[t + 1];
  int n_img = 0;
  for (int64_t k = a; k < b && n_img < min_num_images; ++k) {
    const int v = sup_view[k];
    bool seen = false;
    for (int64_t q = a; q < k; ++q) if (sup_view[q] == v) { seen = true; break; }
    n_img += !seen;
  }
  active[t] = n_img >= min_num_images;
}
void launch_lm_prologue(const double *line_init, const int64_t *sup_off, const int32_t *sup_view, int64_t T,
                        int min_num_images, double *x0, uint8_t *active, int *err, cudaStream_t s) {
  if (T <= 0) return;
  lm_track_prologue_kernel<<<(int)((T + 127) / 128), 128, 0, s>>>(line_init, sup_off, sup_view, T, min_num_images, x0, active, err);
}

void launch_lm_prepare(const double *segs, const int32_t *sup_view, const double *kvec, const double *qvec,
                       const double *tvec, const double *sup_vp, double vp_multiplier, int64_t n, LMBlockDev *out,
                       cudaStream_t s) {
  if (n <= 0) return;
  lm_prepare_blocks_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(segs, sup_view, kvec, qvec, tvec, sup_vp,
                                                                  vp_multiplier, n, out);
}
void launch_lm_refine(const LMParams &p, cudaStream_t s) {
  if (p.T <= 0) return;
  const int warps = 4;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)std::min<int64_t>((p.T + warps - 1) / warps, (int64_t)sms * 4);
  static const int mb = [] { const char *e = getenv("LIMAP_B200_LM_OCC"); return e ? atoi(e) : 3; }();
  if (mb >= 4) lm_refine_kernel<4><<<grid, warps * 32, 0, s>>>(p);
  else if (mb == 3) lm_refine_kernel<3><<<grid, warps * 32, 0, s>>>(p);
  else lm_refine_kernel<2><<<grid, warps * 32, 0, s>>>(p);
}

} // namespace lm
