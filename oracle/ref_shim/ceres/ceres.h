// oracle/ref_shim: stands in for <ceres/ceres.h> (TEST INFRASTRUCTURE). Ceres is not in this image; what the
// reference's residual functors need to COMPILE AND BE EVALUATED is restated from Ceres' public headers:
//   ceres::Jet<T, N> (jet.h): value + N partials, the arithmetic and sqrt / exp / abs / acos overloads the functors use;
//   ceres::CostFunction / AutoDiffCostFunction: declarations only (the functors' static Create() must compile; nothing
//     here builds or solves a ceres::Problem -- the LM restatement stays in oracle/orc_lm.h).
#pragma once
#include <cmath>
#include <limits>

namespace ceres {

template <typename T, int N> struct Jet {
  T a;
  T v[N];
  Jet() : a(T(0)) { for (int i = 0; i < N; ++i) v[i] = T(0); }
  Jet(const T &value) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(0); } // NOLINT (implicit, as in Ceres)
  Jet(const T &value, int k) : a(value) { for (int i = 0; i < N; ++i) v[i] = T(0); v[k] = T(1); }
  Jet &operator+=(const Jet &y) { a += y.a; for (int i = 0; i < N; ++i) v[i] += y.v[i]; return *this; }
  Jet &operator-=(const Jet &y) { a -= y.a; for (int i = 0; i < N; ++i) v[i] -= y.v[i]; return *this; }
  Jet &operator*=(const Jet &y) { *this = *this * y; return *this; }
  Jet &operator/=(const Jet &y) { *this = *this / y; return *this; }
  Jet &operator*=(const T &s) { a *= s; for (int i = 0; i < N; ++i) v[i] *= s; return *this; }
  Jet &operator/=(const T &s) { const T is = T(1) / s; a *= is; for (int i = 0; i < N; ++i) v[i] *= is; return *this; }
};
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N> &f) { return f; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N> &f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N> &f, const Jet<T, N> &g) { Jet<T, N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <typename T, int N> Jet<T, N> operator+(const Jet<T, N> &f, T s) { Jet<T, N> h(f); h.a += s; return h; }
template <typename T, int N> Jet<T, N> operator+(T s, const Jet<T, N> &f) { Jet<T, N> h(f); h.a += s; return h; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N> &f, const Jet<T, N> &g) { Jet<T, N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <typename T, int N> Jet<T, N> operator-(const Jet<T, N> &f, T s) { Jet<T, N> h(f); h.a -= s; return h; }
template <typename T, int N> Jet<T, N> operator-(T s, const Jet<T, N> &f) { Jet<T, N> h = -f; h.a += s; return h; }
template <typename T, int N> Jet<T, N> operator*(const Jet<T, N> &f, const Jet<T, N> &g) { Jet<T, N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <typename T, int N> Jet<T, N> operator*(const Jet<T, N> &f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> Jet<T, N> operator*(T s, const Jet<T, N> &f) { return f * s; }
template <typename T, int N> Jet<T, N> operator/(const Jet<T, N> &f, const Jet<T, N> &g) {
  // jet.h: g_a_inverse = 1 / g.a; f_a_by_g_a = f.a * g_a_inverse; v = (f.v - f_a_by_g_a * g.v) * g_a_inverse
  Jet<T, N> h; const T gi = T(1) / g.a; const T q = f.a * gi; h.a = q; for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - q * g.v[i]) * gi; return h;
}
template <typename T, int N> Jet<T, N> operator/(const Jet<T, N> &f, T s) { const T is = T(1) / s; return f * is; }
template <typename T, int N> Jet<T, N> operator/(T s, const Jet<T, N> &g) { const T minus_s_g_a_inverse2 = -s / (g.a * g.a); Jet<T, N> h; h.a = s / g.a; for (int i = 0; i < N; ++i) h.v[i] = g.v[i] * minus_s_g_a_inverse2; return h; }
#define SHIM_JET_CMP(op)                                                                                              \
  template <typename T, int N> bool operator op(const Jet<T, N> &f, const Jet<T, N> &g) { return f.a op g.a; }        \
  template <typename T, int N> bool operator op(const Jet<T, N> &f, T s) { return f.a op s; }                          \
  template <typename T, int N> bool operator op(T s, const Jet<T, N> &g) { return s op g.a; }
SHIM_JET_CMP(<) SHIM_JET_CMP(<=) SHIM_JET_CMP(>) SHIM_JET_CMP(>=) SHIM_JET_CMP(==) SHIM_JET_CMP(!=)
#undef SHIM_JET_CMP

inline double abs(double x) { return std::abs(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline double exp(double x) { return std::exp(x); }
inline double acos(double x) { return std::acos(x); }
inline bool IsNaN(double x) { return std::isnan(x); }
inline bool IsInfinite(double x) { return std::isinf(x); }
template <typename T, int N> Jet<T, N> abs(const Jet<T, N> &f) { return (f.a < T(0)) ? -f : f; }
template <typename T, int N> Jet<T, N> sqrt(const Jet<T, N> &f) { Jet<T, N> h; h.a = std::sqrt(f.a); const T t = T(1) / (T(2) * h.a); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * t; return h; }
template <typename T, int N> Jet<T, N> exp(const Jet<T, N> &f) { Jet<T, N> h; h.a = std::exp(f.a); for (int i = 0; i < N; ++i) h.v[i] = h.a * f.v[i]; return h; }
template <typename T, int N> Jet<T, N> acos(const Jet<T, N> &f) { Jet<T, N> h; h.a = std::acos(f.a); const T t = -T(1) / std::sqrt(T(1) - f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <typename T, int N> bool IsNaN(const Jet<T, N> &f) { if (std::isnan(f.a)) return true; for (int i = 0; i < N; ++i) if (std::isnan(f.v[i])) return true; return false; }
template <typename T, int N> bool IsInfinite(const Jet<T, N> &f) { if (std::isinf(f.a)) return true; for (int i = 0; i < N; ++i) if (std::isinf(f.v[i])) return true; return false; }

class CostFunction {
public:
  virtual ~CostFunction() {}
};
class LossFunction {
public:
  virtual ~LossFunction() {}
};
template <typename Functor, int kNumResiduals, int... Ns> class AutoDiffCostFunction : public CostFunction {
public:
  explicit AutoDiffCostFunction(Functor *f) : functor_(f) {}
  ~AutoDiffCostFunction() override { delete functor_; }
  Functor *functor_;
};

} // namespace ceres
