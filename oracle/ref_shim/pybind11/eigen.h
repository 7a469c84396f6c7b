// oracle/ref_shim: stands in for <pybind11/eigen.h> (TEST INFRASTRUCTURE). The real header needs Eigen's internals;
// the reference's .cc files only need casters to exist for their dict constructors / as_dict() to compile (they are
// never called from the oracle's C entry points). Matrices travel as numpy arrays of their scalar type.
#pragma once
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <Eigen/Core>
namespace pybind11 {
namespace detail {
template <typename T, int R, int C> struct type_caster<Eigen::Matrix<T, R, C>> {
  typedef Eigen::Matrix<T, R, C> MatT_;
  PYBIND11_TYPE_CASTER(MatT_, const_name("numpy.ndarray"));
  bool load(handle src, bool) {
    auto a = array_t<T, array::c_style | array::forcecast>::ensure(src);
    if (!a) return false;
    Eigen::Index r = a.ndim() >= 1 ? a.shape(0) : 1, c = a.ndim() >= 2 ? a.shape(1) : 1;
    if (a.ndim() == 1 && R == 1) { c = r; r = 1; }
    if ((R != Eigen::Dynamic && r != R) || (C != Eigen::Dynamic && c != C)) return false;
    value.resize(r, c);
    const T *d = a.data();
    for (Eigen::Index i = 0; i < r; ++i) for (Eigen::Index j = 0; j < c; ++j) value(i, j) = d[i * c + j];
    return true;
  }
  static handle cast(const Eigen::Matrix<T, R, C> &m, return_value_policy, handle) {
    const bool vec = (R == 1 || C == 1);
    array_t<T> a = vec ? array_t<T>((size_t)m.size()) : array_t<T>({(size_t)m.rows(), (size_t)m.cols()});
    T *d = a.mutable_data();
    if (vec) for (Eigen::Index i = 0; i < m.size(); ++i) d[i] = m[i];
    else for (Eigen::Index i = 0; i < m.rows(); ++i) for (Eigen::Index j = 0; j < m.cols(); ++j) d[i * m.cols() + j] = m(i, j);
    return a.release();
  }
};
} // namespace detail
} // namespace pybind11
