// oracle/ref_shim/mini_eigen.h — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A small, eager (no expression templates) stand-in for the subset of Eigen 3.4 that the hot-path sources of
// /root/reference use, so that those sources compile UNCHANGED from where they lie (oracle/Makefile, target _ref)
// although Eigen itself is absent from this image. It is written from Eigen's documented semantics:
//   * column-major storage, value semantics, operators evaluated left to right;
//   * Matrix3d::inverse() by cofactors / determinant (Eigen's fixed-size 3x3 path), 2x2 and 4x4 by cofactors too;
//   * Quaterniond(Matrix3d) by the trace / largest-diagonal (Shepperd) branches, toRotationMatrix() with the
//     tx, ty, tz / twx ... products, both as in Eigen/src/Geometry/Quaternion.h;
//   * ldlt().solve() by a symmetric LDL^T with diagonal pivoting; JacobiSVD by one-sided Jacobi (singular values
//     sorted descending; singular vectors are unique up to sign, callers on this path only use V up to sign).
// Only what the reference's files on the path need is here; anything else fails to compile, loudly.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <type_traits>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
using Index = std::ptrdiff_t;
enum { ComputeThinU = 1, ComputeThinV = 2, ComputeFullU = 4, ComputeFullV = 8 };
enum { ColMajor = 0, RowMajor = 1 };

template <typename T, int R, int C, int Opt = 0> class Matrix;

namespace detail {
template <typename T, int R, int C, bool Dyn = (R == Dynamic || C == Dynamic)> struct Storage {
  T d[R * C > 0 ? R * C : 1];
  Storage() { for (int i = 0; i < R * C; ++i) d[i] = T(0); }
  void resize(Index, Index) {}
  Index rows() const { return R; }
  Index cols() const { return C; }
  T *data() { return d; }
  const T *data() const { return d; }
};
template <typename T, int R, int C> struct Storage<T, R, C, true> {
  std::vector<T> d;
  Index r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
  void resize(Index rr, Index cc) { r = rr; c = cc; d.assign((size_t)(rr * cc), T(0)); }
  Index rows() const { return r; }
  Index cols() const { return c; }
  T *data() { return d.data(); }
  const T *data() const { return d.data(); }
};
} // namespace detail

template <typename T, int R, int C> class CommaInit;
template <typename M> class LDLT;
template <typename M> class JacobiSVD;

template <typename T, int R, int C, int Opt> class Matrix {
public:
  typedef T Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
  static constexpr bool IsVector = (R == 1 || C == 1);
  detail::Storage<T, R, C> s;

  Matrix() {}
  // dynamic sizes
  template <int RR = R, int CC = C, typename = std::enable_if_t<RR == Dynamic && CC == Dynamic>>
  Matrix(Index rows, Index cols) { s.resize(rows, cols); }
  template <int RR = R, int CC = C, typename = std::enable_if_t<(RR == Dynamic) != (CC == Dynamic)>>
  explicit Matrix(Index n) { s.resize(R == Dynamic ? n : R, C == Dynamic ? n : C); }
  // fixed-size vector constructors
  template <int N = R * C, typename = std::enable_if_t<N == 2 && IsVector>> Matrix(T a, T b) { s.d[0] = a; s.d[1] = b; }
  template <int N = R * C, typename = std::enable_if_t<N == 3 && IsVector>> Matrix(T a, T b, T c) { s.d[0] = a; s.d[1] = b; s.d[2] = c; }
  template <int N = R * C, typename = std::enable_if_t<N == 4 && IsVector>> Matrix(T a, T b, T c, T d) { s.d[0] = a; s.d[1] = b; s.d[2] = c; s.d[3] = d; }
  // conversion between scalar types / fixed <-> dynamic of the same shape
  template <typename U, int R2, int C2, typename = std::enable_if_t<!(std::is_same<U, T>::value && R2 == R && C2 == C)>>
  Matrix(const Matrix<U, R2, C2> &o) {
    constexpr bool same = (R == Dynamic || R2 == Dynamic || R == R2) && (C == Dynamic || C2 == Dynamic || C == C2);
    constexpr bool vecs = (R == 1 || C == 1) && (R2 == 1 || C2 == 1); // Eigen transposes vectors on assignment
    static_assert(same || vecs, "shape");
    if (same && !(vecs && ((R == 1 && R2 != 1 && C2 == 1) || (C == 1 && C2 != 1 && R2 == 1)))) {
      s.resize(o.rows(), o.cols());
      assert(rows() == o.rows() && cols() == o.cols());
      for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) (*this)(i, j) = (T)o(i, j);
    } else {
      s.resize(R == 1 ? 1 : o.size(), R == 1 ? o.size() : 1);
      assert(size() == o.size());
      for (Index i = 0; i < o.size(); ++i) (*this)[i] = (T)o[i];
    }
  }

  // a 1x1 product (row * column) is usable as a scalar, as in Eigen
  template <int RR = R, int CC = C, typename = std::enable_if_t<RR == 1 && CC == 1>> operator T() const { return s.d[0]; }

  Index rows() const { return s.rows(); }
  Index cols() const { return s.cols(); }
  Index size() const { return rows() * cols(); }
  void resize(Index r, Index c) { s.resize(r, c); }
  void resize(Index n) { s.resize(R == Dynamic ? n : R, C == Dynamic ? n : C); }
  T *data() { return s.data(); }
  const T *data() const { return s.data(); }

  T &operator()(Index i, Index j) { return s.data()[j * rows() + i]; }
  const T &operator()(Index i, Index j) const { return s.data()[j * rows() + i]; }
  T &operator()(Index i) { return s.data()[i]; }
  const T &operator()(Index i) const { return s.data()[i]; }
  T &operator[](Index i) { return s.data()[i]; }
  const T &operator[](Index i) const { return s.data()[i]; }
  T &x() { return s.data()[0]; }
  T &y() { return s.data()[1]; }
  T &z() { return s.data()[2]; }
  T &w() { return s.data()[3]; }
  const T &x() const { return s.data()[0]; }
  const T &y() const { return s.data()[1]; }
  const T &z() const { return s.data()[2]; }
  const T &w() const { return s.data()[3]; }

  static Matrix Zero() { Matrix m; return m; }
  static Matrix Zero(Index r, Index c) { Matrix m; m.s.resize(r, c); return m; }
  static Matrix Zero(Index n) { Matrix m; m.resize(n); return m; }
  static Matrix Ones() { Matrix m; for (Index i = 0; i < m.size(); ++i) m[i] = T(1); return m; }
  static Matrix Constant(T v) { Matrix m; for (Index i = 0; i < m.size(); ++i) m[i] = v; return m; }
  static Matrix Identity() { Matrix m; for (Index i = 0; i < std::min(m.rows(), m.cols()); ++i) m(i, i) = T(1); return m; }
  static Matrix Identity(Index r, Index c) { Matrix m; m.s.resize(r, c); for (Index i = 0; i < std::min(r, c); ++i) m(i, i) = T(1); return m; }
  void setZero() { for (Index i = 0; i < size(); ++i) (*this)[i] = T(0); }
  void setIdentity() { setZero(); for (Index i = 0; i < std::min(rows(), cols()); ++i) (*this)(i, i) = T(1); }
  void setConstant(T v) { for (Index i = 0; i < size(); ++i) (*this)[i] = v; }

  CommaInit<T, R, C> operator<<(T v);
  template <int R2, int C2> CommaInit<T, R, C> operator<<(const Matrix<T, R2, C2> &m);

  // arithmetic
  Matrix operator-() const { Matrix m(*this); for (Index i = 0; i < size(); ++i) m[i] = -m[i]; return m; }
  Matrix &operator+=(const Matrix &o) { assert(size() == o.size()); for (Index i = 0; i < size(); ++i) (*this)[i] += o[i]; return *this; }
  Matrix &operator-=(const Matrix &o) { assert(size() == o.size()); for (Index i = 0; i < size(); ++i) (*this)[i] -= o[i]; return *this; }
  Matrix &operator*=(T v) { for (Index i = 0; i < size(); ++i) (*this)[i] *= v; return *this; }
  Matrix &operator/=(T v) { for (Index i = 0; i < size(); ++i) (*this)[i] /= v; return *this; }
  friend Matrix operator+(Matrix a, const Matrix &b) { a += b; return a; }
  friend Matrix operator-(Matrix a, const Matrix &b) { a -= b; return a; }
  friend Matrix operator*(Matrix a, T v) { a *= v; return a; }
  friend Matrix operator*(T v, Matrix a) { a *= v; return a; }
  friend Matrix operator/(Matrix a, T v) { a /= v; return a; }
  template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>>
  friend Matrix operator*(Matrix a, U v) { a *= (T)v; return a; }
  template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>>
  friend Matrix operator*(U v, Matrix a) { a *= (T)v; return a; }
  template <typename U, typename = std::enable_if_t<std::is_arithmetic<U>::value && !std::is_same<U, T>::value>>
  friend Matrix operator/(Matrix a, U v) { a /= (T)v; return a; }
  bool operator==(const Matrix &o) const {
    if (rows() != o.rows() || cols() != o.cols()) return false;
    for (Index i = 0; i < size(); ++i) if (!((*this)[i] == o[i])) return false;
    return true;
  }
  bool operator!=(const Matrix &o) const { return !(*this == o); }

  template <int R2, int C2> Matrix<T, R, C2> operator*(const Matrix<T, R2, C2> &o) const {
    static_assert(C == R2 || C == Dynamic || R2 == Dynamic, "inner dimensions");
    assert(cols() == o.rows());
    Matrix<T, R, C2> m;
    m.s.resize(rows(), o.cols());
    for (Index j = 0; j < o.cols(); ++j)
      for (Index i = 0; i < rows(); ++i) {
        T acc = T(0);
        for (Index k = 0; k < cols(); ++k) acc += (*this)(i, k) * o(k, j);
        m(i, j) = acc;
      }
    return m;
  }
  LDLT<Matrix> ldlt() const;
  JacobiSVD<Matrix> jacobiSvd(unsigned opts = 0) const;

  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> m;
    m.s.resize(cols(), rows());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) m(j, i) = (*this)(i, j);
    return m;
  }
  void transposeInPlace() { *this = Matrix(transpose()); }

  T dot(const Matrix &o) const { assert(size() == o.size()); T a = T(0); for (Index i = 0; i < size(); ++i) a += (*this)[i] * o[i]; return a; }
  T squaredNorm() const { T a = T(0); for (Index i = 0; i < size(); ++i) a += (*this)[i] * (*this)[i]; return a; }
  T norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  Matrix normalized() const { T n = norm(); Matrix m(*this); if (n > T(0)) m /= n; return m; }
  void normalize() { T n = norm(); if (n > T(0)) *this /= n; }
  T sum() const { T a = T(0); for (Index i = 0; i < size(); ++i) a += (*this)[i]; return a; }
  T mean() const { return sum() / T(size()); }
  T trace() const { T a = T(0); for (Index i = 0; i < std::min(rows(), cols()); ++i) a += (*this)(i, i); return a; }
  T maxCoeff() const { T a = (*this)[0]; for (Index i = 1; i < size(); ++i) a = std::max(a, (*this)[i]); return a; }
  T minCoeff() const { T a = (*this)[0]; for (Index i = 1; i < size(); ++i) a = std::min(a, (*this)[i]); return a; }
  template <typename I> T maxCoeff(I *idx) const { Index b = 0; for (Index i = 1; i < size(); ++i) if ((*this)[i] > (*this)[b]) b = i; *idx = (I)b; return (*this)[b]; }
  template <typename I> T minCoeff(I *idx) const { Index b = 0; for (Index i = 1; i < size(); ++i) if ((*this)[i] < (*this)[b]) b = i; *idx = (I)b; return (*this)[b]; }
  bool hasNaN() const { for (Index i = 0; i < size(); ++i) if ((*this)[i] != (*this)[i]) return true; return false; }
  bool allFinite() const { for (Index i = 0; i < size(); ++i) if (!std::isfinite((double)(*this)[i])) return false; return true; }
  Matrix cwiseAbs() const { Matrix m(*this); for (Index i = 0; i < size(); ++i) m[i] = std::abs(m[i]); return m; }
  Matrix cwiseProduct(const Matrix &o) const { Matrix m(*this); for (Index i = 0; i < size(); ++i) m[i] *= o[i]; return m; }
  bool isApprox(const Matrix &o, T prec = T(1e-12)) const {
    return (*this - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
  }
  template <typename U> Matrix<U, R, C> cast() const { return Matrix<U, R, C>(*this); }
  const Matrix &eval() const { return *this; }
  const Matrix &array() const { return *this; }
  const Matrix &matrix() const { return *this; }

  Matrix cross(const Matrix &o) const {
    static_assert(R * C == 3, "cross() is for 3-vectors");
    Matrix m;
    m[0] = (*this)[1] * o[2] - (*this)[2] * o[1];
    m[1] = (*this)[2] * o[0] - (*this)[0] * o[2];
    m[2] = (*this)[0] * o[1] - (*this)[1] * o[0];
    return m;
  }

  // blocks (copies out; writable access through the helpers below)
  Matrix<T, R, 1> col(Index j) const { Matrix<T, R, 1> v; v.s.resize(rows(), 1); for (Index i = 0; i < rows(); ++i) v[i] = (*this)(i, j); return v; }
  Matrix<T, 1, C> row(Index i) const { Matrix<T, 1, C> v; v.s.resize(1, cols()); for (Index j = 0; j < cols(); ++j) v[j] = (*this)(i, j); return v; }
  template <int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v[i] = (*this)[i]; return v; }
  template <int N> Matrix<T, N, 1> tail() const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v[i] = (*this)[size() - N + i]; return v; }
  Matrix<T, Dynamic, 1> head(Index n) const { Matrix<T, Dynamic, 1> v(n); for (Index i = 0; i < n; ++i) v[i] = (*this)[i]; return v; }
  template <int N> Matrix<T, N, 1> segment(Index o) const { Matrix<T, N, 1> v; for (int i = 0; i < N; ++i) v[i] = (*this)[o + i]; return v; }
  template <int BR, int BC> Matrix<T, BR, BC> block(Index i0, Index j0) const {
    Matrix<T, BR, BC> m;
    for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) m(i, j) = (*this)(i0 + i, j0 + j);
    return m;
  }
  Matrix<T, Dynamic, Dynamic> block(Index i0, Index j0, Index br, Index bc) const {
    Matrix<T, Dynamic, Dynamic> m(br, bc);
    for (Index j = 0; j < bc; ++j) for (Index i = 0; i < br; ++i) m(i, j) = (*this)(i0 + i, j0 + j);
    return m;
  }
  template <int BR, int BC> Matrix<T, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
  template <int BR, int BC> Matrix<T, BR, BC> topRightCorner() const { return block<BR, BC>(0, cols() - BC); }
  Matrix<T, R == Dynamic ? Dynamic : R + 1, 1> homogeneous() const {
    Matrix<T, R == Dynamic ? Dynamic : R + 1, 1> v;
    v.s.resize(size() + 1, 1);
    for (Index i = 0; i < size(); ++i) v[i] = (*this)[i];
    v[size()] = T(1);
    return v;
  }
  Matrix<T, R == Dynamic ? Dynamic : R - 1, 1> hnormalized() const {
    Matrix<T, R == Dynamic ? Dynamic : R - 1, 1> v;
    v.s.resize(size() - 1, 1);
    for (Index i = 0; i + 1 < size(); ++i) v[i] = (*this)[i] / (*this)[size() - 1];
    return v;
  }

  // writable column / row / block access: the proxy IS the extracted value (so it takes part in every expression)
  // and writes through to its parent on assignment
  struct ColRef : public Matrix<T, R, 1> {
    typedef Matrix<T, R, 1> V;
    Matrix &m; Index j;
    ColRef(Matrix &mm, Index jj) : V(static_cast<const Matrix &>(mm).col(jj)), m(mm), j(jj) {}
    void flush() { for (Index i = 0; i < m.rows(); ++i) m(i, j) = (*this)[i]; }
    template <int R2, int C2> ColRef &operator=(const Matrix<T, R2, C2> &v) { assert(v.size() == m.rows()); for (Index i = 0; i < m.rows(); ++i) (*this)[i] = v[i]; flush(); return *this; }
    ColRef &operator=(const ColRef &v) { return *this = static_cast<const V &>(v); }
    template <int R2, int C2> ColRef &operator+=(const Matrix<T, R2, C2> &v) { for (Index i = 0; i < m.rows(); ++i) (*this)[i] += v[i]; flush(); return *this; }
    template <int R2, int C2> ColRef &operator-=(const Matrix<T, R2, C2> &v) { for (Index i = 0; i < m.rows(); ++i) (*this)[i] -= v[i]; flush(); return *this; }
    ColRef &operator*=(T v) { V::operator*=(v); flush(); return *this; }
    ColRef &operator/=(T v) { V::operator/=(v); flush(); return *this; }
  };
  ColRef col(Index j) { return ColRef(*this, j); }
  struct RowRef : public Matrix<T, 1, C> {
    typedef Matrix<T, 1, C> V;
    Matrix &m; Index i;
    RowRef(Matrix &mm, Index ii) : V(static_cast<const Matrix &>(mm).row(ii)), m(mm), i(ii) {}
    void flush() { for (Index j = 0; j < m.cols(); ++j) m(i, j) = (*this)[j]; }
    template <int R2, int C2> RowRef &operator=(const Matrix<T, R2, C2> &v) { assert(v.size() == m.cols()); for (Index j = 0; j < m.cols(); ++j) (*this)[j] = v[j]; flush(); return *this; }
    RowRef &operator=(const RowRef &v) { return *this = static_cast<const V &>(v); }
    template <int R2, int C2> RowRef &operator+=(const Matrix<T, R2, C2> &v) { for (Index j = 0; j < m.cols(); ++j) (*this)[j] += v[j]; flush(); return *this; }
    template <int R2, int C2> RowRef &operator-=(const Matrix<T, R2, C2> &v) { for (Index j = 0; j < m.cols(); ++j) (*this)[j] -= v[j]; flush(); return *this; }
    RowRef &operator*=(T v) { V::operator*=(v); flush(); return *this; }
    RowRef &operator/=(T v) { V::operator/=(v); flush(); return *this; }
  };
  RowRef row(Index i) { return RowRef(*this, i); }
  template <int BR, int BC> struct BlockRef : public Matrix<T, BR, BC> {
    typedef Matrix<T, BR, BC> V;
    Matrix &m; Index i0, j0;
    BlockRef(Matrix &mm, Index i, Index j) : V(static_cast<const Matrix &>(mm).template block<BR, BC>(i, j)), m(mm), i0(i), j0(j) {}
    void flush() { for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) m(i0 + i, j0 + j) = (*this)(i, j); }
    template <int R2, int C2> BlockRef &operator=(const Matrix<T, R2, C2> &v) {
      assert(v.size() == BR * BC);
      if (v.rows() == BR) for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) (*this)(i, j) = v(i, j);
      else for (int k = 0; k < BR * BC; ++k) (*this)[k] = v[k];
      flush(); return *this;
    }
    BlockRef &operator=(const BlockRef &v) { return *this = static_cast<const V &>(v); }
    template <int R2, int C2> BlockRef &operator+=(const Matrix<T, R2, C2> &v) { for (int k = 0; k < BR * BC; ++k) (*this)[k] += v[k]; flush(); return *this; }
    template <int R2, int C2> BlockRef &operator-=(const Matrix<T, R2, C2> &v) { for (int k = 0; k < BR * BC; ++k) (*this)[k] -= v[k]; flush(); return *this; }
    BlockRef &operator*=(T v) { V::operator*=(v); flush(); return *this; }
    BlockRef &operator/=(T v) { V::operator/=(v); flush(); return *this; }
  };
  template <int BR, int BC> BlockRef<BR, BC> block(Index i0, Index j0) { return BlockRef<BR, BC>(*this, i0, j0); }
  template <int BR, int BC> BlockRef<BR, BC> topLeftCorner() { return BlockRef<BR, BC>(*this, 0, 0); }
  template <int BR, int BC> BlockRef<BR, BC> topRightCorner() { return BlockRef<BR, BC>(*this, 0, cols() - BC); }
  template <int N> BlockRef<N, 1> head() { return BlockRef<N, 1>(*this, 0, 0); }

  T determinant() const;
  Matrix inverse() const;
};

template <typename T, int R, int C> std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C> &m) {
  for (Index i = 0; i < m.rows(); ++i) { for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j); if (i + 1 < m.rows()) os << "\n"; }
  return os;
}

// comma initialiser: row-major fill, like Eigen's
template <typename T, int R, int C> class CommaInit {
public:
  Matrix<T, R, C> &m;
  Index k = 0;
  explicit CommaInit(Matrix<T, R, C> &mm) : m(mm) {}
  void put(T v) { const Index i = k / m.cols(), j = k % m.cols(); m(i, j) = v; ++k; }
  CommaInit &operator,(T v) { put(v); return *this; }
  template <int R2, int C2> CommaInit &operator,(const Matrix<T, R2, C2> &b) { place(b); return *this; }
  template <int R2, int C2> void place(const Matrix<T, R2, C2> &b) {
    // blocks of full height laid side by side (columns), or of full width stacked (rows)
    if (b.rows() == m.rows() && m.rows() != 1) { const Index j0 = k; for (Index j = 0; j < b.cols(); ++j) for (Index i = 0; i < b.rows(); ++i) m(i, j0 + j) = b(i, j); k += b.cols(); colmode = true; }
    else if (m.cols() == 1 || m.rows() == 1) { for (Index i = 0; i < b.size(); ++i) m[k + i] = b[i]; k += b.size(); }
    else { assert(b.cols() == m.cols()); const Index i0 = k / m.cols(); for (Index j = 0; j < b.cols(); ++j) for (Index i = 0; i < b.rows(); ++i) m(i0 + i, j) = b(i, j); k += b.size(); }
  }
  bool colmode = false;
};
template <typename T, int R, int C, int Opt> CommaInit<T, R, C> Matrix<T, R, C, Opt>::operator<<(T v) { CommaInit<T, R, C> c(*this); c.put(v); return c; }
template <typename T, int R, int C, int Opt> template <int R2, int C2> CommaInit<T, R, C> Matrix<T, R, C, Opt>::operator<<(const Matrix<T, R2, C2> &b) { CommaInit<T, R, C> c(*this); c.place(b); return c; }

// ---- determinant / inverse (cofactors, as Eigen does for fixed sizes up to 4) --------------------------------------
namespace detail {
template <typename T, int R, int C> T det_rec(const Matrix<T, R, C> &a) {
  const Index n = a.rows();
  if (n == 1) return a(0, 0);
  if (n == 2) return a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1);
  if (n == 3)
    return a(0, 0) * (a(1, 1) * a(2, 2) - a(2, 1) * a(1, 2)) - a(0, 1) * (a(1, 0) * a(2, 2) - a(1, 2) * a(2, 0)) +
           a(0, 2) * (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0));
  T d = T(0);
  for (Index c = 0; c < n; ++c) {
    Matrix<T, Dynamic, Dynamic> sub(n - 1, n - 1);
    for (Index i = 1; i < n; ++i) { Index cc = 0; for (Index j = 0; j < n; ++j) { if (j == c) continue; sub(i - 1, cc++) = a(i, j); } }
    d += ((c % 2) ? T(-1) : T(1)) * a(0, c) * det_rec(sub);
  }
  return d;
}
} // namespace detail
template <typename T, int R, int C, int Opt> T Matrix<T, R, C, Opt>::determinant() const { assert(rows() == cols()); return detail::det_rec(*this); }
template <typename T, int R, int C, int Opt> Matrix<T, R, C, Opt> Matrix<T, R, C, Opt>::inverse() const {
  assert(rows() == cols());
  const Index n = rows();
  Matrix inv;
  inv.s.resize(n, n);
  if (n == 1) { inv(0, 0) = T(1) / (*this)(0, 0); return inv; }
  if (n == 2) {
    const T invdet = T(1) / determinant();
    inv(0, 0) = (*this)(1, 1) * invdet; inv(1, 0) = -(*this)(1, 0) * invdet;
    inv(0, 1) = -(*this)(0, 1) * invdet; inv(1, 1) = (*this)(0, 0) * invdet;
    return inv;
  }
  if (n == 3) {
    const Matrix &m = *this;
    auto cof = [&](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    T c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    const T det = m(0, 0) * c00 + m(1, 0) * c10 + m(2, 0) * c20; // cofactor expansion along the first column
    const T invdet = T(1) / det;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv(j, i) = cof(i, j) * invdet; // adjugate = cofactor^T
    return inv;
  }
  // general: Gauss-Jordan with partial pivoting
  Matrix<T, Dynamic, Dynamic> a(n, 2 * n);
  for (Index i = 0; i < n; ++i) for (Index j = 0; j < n; ++j) { a(i, j) = (*this)(i, j); a(i, n + j) = (i == j) ? T(1) : T(0); }
  for (Index c = 0; c < n; ++c) {
    Index p = c;
    for (Index i = c + 1; i < n; ++i) if (std::abs(a(i, c)) > std::abs(a(p, c))) p = i;
    if (p != c) for (Index j = 0; j < 2 * n; ++j) std::swap(a(p, j), a(c, j));
    const T d = a(c, c);
    for (Index j = 0; j < 2 * n; ++j) a(c, j) /= d;
    for (Index i = 0; i < n; ++i) { if (i == c) continue; const T f = a(i, c); if (f == T(0)) continue; for (Index j = 0; j < 2 * n; ++j) a(i, j) -= f * a(c, j); }
  }
  for (Index i = 0; i < n; ++i) for (Index j = 0; j < n; ++j) inv(i, j) = a(i, n + j);
  return inv;
}

// ---- LDLT (symmetric, diagonal pivoting) ------------------------------------------------------------------------------
template <typename M> class LDLT {
public:
  typedef typename M::Scalar T;
  M a;
  explicit LDLT(const M &m) : a(m) {}
  template <typename V> V solve(const V &b) const {
    const Index n = a.rows();
    Matrix<T, Dynamic, Dynamic> A(n, n);
    for (Index i = 0; i < n; ++i) for (Index j = 0; j < n; ++j) A(i, j) = a(i, j);
    std::vector<Index> perm(n);
    for (Index i = 0; i < n; ++i) perm[i] = i;
    Matrix<T, Dynamic, Dynamic> L = Matrix<T, Dynamic, Dynamic>::Identity(n, n);
    std::vector<T> D(n, T(0));
    for (Index k = 0; k < n; ++k) {
      Index p = k; // largest remaining diagonal entry
      for (Index i = k + 1; i < n; ++i) if (std::abs(A(i, i)) > std::abs(A(p, p))) p = i;
      if (p != k) {
        for (Index j = 0; j < n; ++j) std::swap(A(k, j), A(p, j));
        for (Index i = 0; i < n; ++i) std::swap(A(i, k), A(i, p));
        for (Index j = 0; j < k; ++j) std::swap(L(k, j), L(p, j));
        std::swap(perm[k], perm[p]);
      }
      D[k] = A(k, k);
      for (Index i = k + 1; i < n; ++i) {
        L(i, k) = (D[k] != T(0)) ? A(i, k) / D[k] : T(0);
      }
      for (Index i = k + 1; i < n; ++i) for (Index j = k + 1; j < n; ++j) A(i, j) -= L(i, k) * D[k] * L(j, k);
    }
    V x = b;
    std::vector<T> y(n);
    for (Index i = 0; i < n; ++i) y[i] = b[perm[i]];
    for (Index i = 0; i < n; ++i) for (Index k = 0; k < i; ++k) y[i] -= L(i, k) * y[k];
    for (Index i = 0; i < n; ++i) y[i] = (D[i] != T(0)) ? y[i] / D[i] : T(0);
    for (Index i = n - 1; i >= 0; --i) for (Index k = i + 1; k < n; ++k) y[i] -= L(k, i) * y[k];
    for (Index i = 0; i < n; ++i) x[perm[i]] = y[i];
    return x;
  }
};
template <typename T, int R, int C> LDLT<Matrix<T, R, C>> ldlt_of(const Matrix<T, R, C> &m) { return LDLT<Matrix<T, R, C>>(m); }

// ---- one-sided Jacobi SVD --------------------------------------------------------------------------------------------
template <typename M> class JacobiSVD {
public:
  typedef typename M::Scalar T;
  Matrix<T, Dynamic, Dynamic> U_, V_;
  Matrix<T, Dynamic, 1> S_;
  JacobiSVD() {}
  JacobiSVD(const M &m, unsigned = 0) { compute(m); }
  JacobiSVD &compute(const M &m, unsigned = 0) {
    const Index r = m.rows(), c = m.cols();
    Matrix<T, Dynamic, Dynamic> A(r, c), V = Matrix<T, Dynamic, Dynamic>::Identity(c, c);
    for (Index i = 0; i < r; ++i) for (Index j = 0; j < c; ++j) A(i, j) = m(i, j);
    for (int sweep = 0; sweep < 100; ++sweep) {
      bool rotated = false;
      for (Index p = 0; p + 1 < c; ++p)
        for (Index q = p + 1; q < c; ++q) {
          T alpha = 0, beta = 0, gamma = 0;
          for (Index i = 0; i < r; ++i) { alpha += A(i, p) * A(i, p); beta += A(i, q) * A(i, q); gamma += A(i, p) * A(i, q); }
          if (gamma == T(0) || std::abs(gamma) <= std::numeric_limits<T>::epsilon() * std::sqrt(alpha * beta)) continue;
          rotated = true;
          const T zeta = (beta - alpha) / (T(2) * gamma);
          const T t = (zeta >= 0 ? T(1) : T(-1)) / (std::abs(zeta) + std::sqrt(T(1) + zeta * zeta));
          const T cs = T(1) / std::sqrt(T(1) + t * t), sn = cs * t;
          for (Index i = 0; i < r; ++i) { const T a = A(i, p), b = A(i, q); A(i, p) = cs * a - sn * b; A(i, q) = sn * a + cs * b; }
          for (Index i = 0; i < c; ++i) { const T a = V(i, p), b = V(i, q); V(i, p) = cs * a - sn * b; V(i, q) = sn * a + cs * b; }
        }
      if (!rotated) break;
    }
    std::vector<T> sv(c);
    for (Index j = 0; j < c; ++j) { T s = 0; for (Index i = 0; i < r; ++i) s += A(i, j) * A(i, j); sv[j] = std::sqrt(s); }
    std::vector<Index> ord(c);
    for (Index j = 0; j < c; ++j) ord[j] = j;
    std::stable_sort(ord.begin(), ord.end(), [&](Index a, Index b) { return sv[a] > sv[b]; });
    const Index k = std::min(r, c);
    U_.resize(r, k); V_.resize(c, c); S_.resize(k);
    for (Index j = 0; j < c; ++j) for (Index i = 0; i < c; ++i) V_(i, j) = V(i, ord[j]);
    for (Index j = 0; j < k; ++j) {
      S_[j] = sv[ord[j]];
      for (Index i = 0; i < r; ++i) U_(i, j) = (sv[ord[j]] > T(0)) ? A(i, ord[j]) / sv[ord[j]] : T(0);
    }
    return *this;
  }
  const Matrix<T, Dynamic, Dynamic> &matrixU() const { return U_; }
  const Matrix<T, Dynamic, Dynamic> &matrixV() const { return V_; }
  const Matrix<T, Dynamic, 1> &singularValues() const { return S_; }
};

// ---- quaternion (Eigen/src/Geometry/Quaternion.h semantics; storage x,y,z,w, constructor (w,x,y,z)) -------------------
template <typename T> class Quaternion {
public:
  T x_, y_, z_, w_;
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}
  Quaternion(T w, T x, T y, T z) : x_(x), y_(y), z_(z), w_(w) {}
  explicit Quaternion(const Matrix<T, 3, 3> &m) {
    T t = m.trace();
    if (t > T(0)) {
      t = std::sqrt(t + T(1.0));
      w_ = T(0.5) * t;
      t = T(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t;
      y_ = (m(0, 2) - m(2, 0)) * t;
      z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
      T v[3];
      v[i] = T(0.5) * t;
      t = T(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t;
      v[j] = (m(j, i) + m(i, j)) * t;
      v[k] = (m(k, i) + m(i, k)) * t;
      x_ = v[0]; y_ = v[1]; z_ = v[2];
    }
  }
  T w() const { return w_; }
  T x() const { return x_; }
  T y() const { return y_; }
  T z() const { return z_; }
  T &w() { return w_; }
  T &x() { return x_; }
  T &y() { return y_; }
  T &z() { return z_; }
  T norm() const { return std::sqrt(x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_); }
  void normalize() { const T n = norm(); x_ /= n; y_ /= n; z_ /= n; w_ /= n; }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Matrix<T, 4, 1> coeffs() const { return Matrix<T, 4, 1>(x_, y_, z_, w_); }
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> res;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const { const T n2 = x_ * x_ + y_ * y_ + z_ * z_ + w_ * w_; return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2); }
  Quaternion operator*(const Quaternion &b) const {
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  Matrix<T, 3, 1> operator*(const Matrix<T, 3, 1> &v) const { return toRotationMatrix() * v; }
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

// Eigen::Map as the path uses it: a read view of a plain array. Here it is a VALUE (copied at construction), which is
// equivalent for the const / temporary uses in ceresbase/line_projection.h; RowMajor of the mapped type is honoured.
template <typename M> class Map;
template <typename T, int R, int C, int Opt> class Map<Matrix<T, R, C, Opt>> : public Matrix<T, R, C> {
public:
  explicit Map(const T *p) {
    for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) (*this)(i, j) = (Opt == RowMajor) ? p[i * C + j] : p[j * R + i];
  }
};
template <typename T, int R, int C, int Opt> class Map<const Matrix<T, R, C, Opt>> : public Matrix<T, R, C> {
public:
  explicit Map(const T *p) {
    for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) (*this)(i, j) = (Opt == RowMajor) ? p[i * C + j] : p[j * R + i];
  }
};

#define MINI_EIGEN_TYPEDEFS(T, S)                                                                                      \
  typedef Matrix<T, 2, 1> Vector2##S; typedef Matrix<T, 3, 1> Vector3##S; typedef Matrix<T, 4, 1> Vector4##S;            \
  typedef Matrix<T, Dynamic, 1> VectorX##S; typedef Matrix<T, 1, Dynamic> RowVectorX##S;                                \
  typedef Matrix<T, 1, 2> RowVector2##S; typedef Matrix<T, 1, 3> RowVector3##S; typedef Matrix<T, 1, 4> RowVector4##S;    \
  typedef Matrix<T, 2, 2> Matrix2##S; typedef Matrix<T, 3, 3> Matrix3##S; typedef Matrix<T, 4, 4> Matrix4##S;            \
  typedef Matrix<T, Dynamic, Dynamic> MatrixX##S; typedef Matrix<T, 3, 4> Matrix3x4##S;
MINI_EIGEN_TYPEDEFS(double, d)
MINI_EIGEN_TYPEDEFS(float, f)
MINI_EIGEN_TYPEDEFS(int, i)
#undef MINI_EIGEN_TYPEDEFS
typedef Matrix<double, 6, 1> Vector6d;
typedef Matrix<double, 6, 6> Matrix6d;
template <typename T, int N> using Vector = Matrix<T, N, 1>;

template <typename T, int R, int C, int Opt> LDLT<Matrix<T, R, C, Opt>> Matrix<T, R, C, Opt>::ldlt() const { return LDLT<Matrix<T, R, C, Opt>>(*this); }
template <typename T, int R, int C, int Opt> JacobiSVD<Matrix<T, R, C, Opt>> Matrix<T, R, C, Opt>::jacobiSvd(unsigned o) const { return JacobiSVD<Matrix<T, R, C, Opt>>(*this, o); }
} // namespace Eigen
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STL_VECTOR_SPECIALIZATION_H
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
