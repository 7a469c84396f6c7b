// oracle/ref_shim: stands in for <PoseLib/misc/univariate.h> (TEST INFRASTRUCTURE). PoseLib (pinned a84c545 in the
// reference's cmake/FindDependencies.cmake:44-47) is not in /root/reference; the one routine the path calls
// (solvers/triangulation/triangulate_line_with_one_point.cc:557) is restated: real roots of the monic quartic
// x^4 + b x^3 + c x^2 + d x + e by Ferrari's factorisation over a real root of the resolvent cubic, each root
// polished with two Newton steps on the original polynomial.
#pragma once
#include <algorithm>
#include <cmath>
namespace poselib {
namespace univariate {

// one real root of x^3 + a x^2 + b x + c (the largest one)
inline double shim_cubic_real_root(double a, double b, double c) {
  const double q = (a * a - 3.0 * b) / 9.0, r = (2.0 * a * a * a - 9.0 * a * b + 27.0 * c) / 54.0;
  const double r2 = r * r, q3 = q * q * q;
  double x;
  if (r2 < q3) {
    const double theta = std::acos(std::max(-1.0, std::min(1.0, r / std::sqrt(q3))));
    const double sq = -2.0 * std::sqrt(q);
    const double x0 = sq * std::cos(theta / 3.0) - a / 3.0;
    const double x1 = sq * std::cos((theta + 2.0 * M_PI) / 3.0) - a / 3.0;
    const double x2 = sq * std::cos((theta - 2.0 * M_PI) / 3.0) - a / 3.0;
    x = std::max(x0, std::max(x1, x2));
  } else {
    const double A = -std::copysign(std::cbrt(std::fabs(r) + std::sqrt(r2 - q3)), r);
    const double B = (A != 0.0) ? q / A : 0.0;
    x = A + B - a / 3.0;
  }
  for (int it = 0; it < 2; ++it) { // polish
    const double f = ((x + a) * x + b) * x + c, df = (3.0 * x + 2.0 * a) * x + b;
    if (df != 0.0) x -= f / df;
  }
  return x;
}

inline int solve_quartic_real(double b, double c, double d, double e, double roots[4]) {
  // depressed quartic y^4 + p y^2 + q y + r, x = y - b/4
  const double b2 = b * b;
  const double p = c - 3.0 * b2 / 8.0;
  const double q = d - b * c / 2.0 + b2 * b / 8.0;
  const double r = e - b * d / 4.0 + b2 * c / 16.0 - 3.0 * b2 * b2 / 256.0;
  int n = 0;
  if (std::fabs(q) < 1e-14 * (1.0 + std::fabs(p) + std::fabs(r))) { // biquadratic
    const double disc = p * p - 4.0 * r;
    if (disc >= 0) {
      const double s = std::sqrt(disc);
      for (double z : {(-p + s) / 2.0, (-p - s) / 2.0})
        if (z >= 0) { roots[n++] = std::sqrt(z) - b / 4.0; roots[n++] = -std::sqrt(z) - b / 4.0; }
    }
  } else {
    // resolvent cubic: m^3 + p m^2 + (p^2/4 - r) m - q^2/8 = 0, any root m > 0
    const double m = shim_cubic_real_root(p, p * p / 4.0 - r, -q * q / 8.0);
    if (m > 0) {
      const double s = std::sqrt(2.0 * m);
      // (y^2 + p/2 + m)^2 = (s y - q/(2 s))^2
      for (double sg : {1.0, -1.0}) {
        // y^2 - sg s y + p/2 + m + sg q/(2 s) = 0
        const double bb = -sg * s, cc = p / 2.0 + m + sg * q / (2.0 * s);
        const double disc = bb * bb - 4.0 * cc;
        if (disc >= 0) {
          const double sd = std::sqrt(disc);
          roots[n++] = (-bb + sd) / 2.0 - b / 4.0;
          roots[n++] = (-bb - sd) / 2.0 - b / 4.0;
        }
      }
    }
  }
  for (int i = 0; i < n; ++i) {
    double x = roots[i];
    for (int it = 0; it < 2; ++it) {
      const double f = (((x + b) * x + c) * x + d) * x + e, df = ((4.0 * x + 3.0 * b) * x + 2.0 * c) * x + d;
      if (df != 0.0 && std::isfinite(f / df)) x -= f / df;
    }
    roots[i] = x;
  }
  return n;
}

} // namespace univariate
} // namespace poselib
