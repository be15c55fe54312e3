// STAND-IN -- this is NOT Eigen.  Test infrastructure only (see oracle/build_ref_pba.py).
//
// Eigen is not installed in this image and there is no network.  The reference's photometric bundle adjustment
// (evaluate_jacobians.hpp, hessian_block_evaluation.hpp, eigen_photometric_bundle_adjustment_problem.hpp, the pinhole
// ArrayReprojector, PixelMap, NormalLinearSystem::solve ...) is written against Eigen's API.  This header supplies just
// enough of that API -- same names, same semantics, every expression evaluated EAGERLY into a plain matrix, no
// vectorisation -- for those reference sources to compile UNCHANGED where they lie under /root/reference, so that the
// restatements in oracle/ can be pinned against the reference's own code.  Differences from real Eigen are limited to
// floating-point summation order inside products / norms (relative 1e-16 per operation in double); the golden vectors
// made through it are compared at 1e-9, far above that and far below anything an algorithmic difference would cause.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdlib>
#include <initializer_list>
#include <limits>
#include <memory>
#include <new>
#include <stdexcept>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

using Index = long;
constexpr int Dynamic = -1;
enum StorageOptions { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum UpLoType { Lower = 1, Upper = 2 };
enum DecompositionOptions { ComputeThinU = 4, ComputeThinV = 8, ComputeFullU = 16, ComputeFullV = 32 };

// Eigen::aligned_allocator: 32-byte aligned storage (EIGEN_DEFAULT_ALIGN_BYTES with AVX); the reference's AVX2
// calculate_pixelinfo uses aligned loads on vectors allocated through it (calculate_pixelinfo.cpp:386-392)
template <class T>
struct aligned_allocator {
  using value_type = T;
  aligned_allocator() = default;
  template <class U>
  aligned_allocator(const aligned_allocator<U>&) {}
  T* allocate(std::size_t n) { return static_cast<T*>(::operator new(n * sizeof(T), std::align_val_t(32))); }
  void deallocate(T* p, std::size_t) { ::operator delete(p, std::align_val_t(32)); }
  template <class U>
  bool operator==(const aligned_allocator<U>&) const {
    return true;
  }
  template <class U>
  bool operator!=(const aligned_allocator<U>&) const {
    return false;
  }
};

constexpr int default_options(int R, int C) { return (R == 1 && C != 1) ? RowMajor : ColMajor; }
constexpr int mul_dim(int a, int b) { return (a == Dynamic || b == Dynamic) ? Dynamic : a * b; }

template <class T, int R, int C, int Opt = default_options(R, C), int MR = R, int MC = C>
class Matrix;
template <class T, int R, int C, int Opt = default_options(R, C), int MR = R, int MC = C>
class Array;
template <class T, int R, int C, bool RowMaj>
class Block;
template <class X, int MapOptions = 0, class Stride = void>
class Map;
template <class X>
class Ref;
template <class D>
class MatrixBase;
template <class T>
class Quaternion;

template <class T, int N>
using Vector = Matrix<T, N, 1>;
template <class T, int N>
using RowVector = Matrix<T, 1, N>;
template <class T>
using Vector2 = Matrix<T, 2, 1>;
template <class T>
using Vector3 = Matrix<T, 3, 1>;
template <class T>
using Vector4 = Matrix<T, 4, 1>;
template <class T>
using VectorX = Matrix<T, Dynamic, 1>;
template <class T>
using MatrixX = Matrix<T, Dynamic, Dynamic>;
template <class T>
using Matrix2 = Matrix<T, 2, 2>;
template <class T>
using Matrix3 = Matrix<T, 3, 3>;
template <class T>
using Matrix4 = Matrix<T, 4, 4>;
using Vector2d = Vector2<double>;
using Vector3d = Vector3<double>;
using Vector4d = Vector4<double>;
using VectorXd = VectorX<double>;
using MatrixXd = MatrixX<double>;
using Matrix3d = Matrix3<double>;
using Matrix4d = Matrix4<double>;
using Vector2f = Vector2<float>;
using Vector3f = Vector3<float>;
using VectorXf = VectorX<float>;
using MatrixXf = MatrixX<float>;
using Vector2i = Vector2<int>;

namespace internal {
template <class D>
struct traits;
template <class T, int R, int C, int Opt, int MR, int MC>
struct traits<Matrix<T, R, C, Opt, MR, MC>> {
  using Scalar = T;
  static constexpr int Rows = R, Cols = C;
  static constexpr bool RowMaj = (Opt & RowMajor) != 0;
};
template <class T, int R, int C, bool RM>
struct traits<Block<T, R, C, RM>> {
  using Scalar = std::remove_const_t<T>;
  static constexpr int Rows = R, Cols = C;
  static constexpr bool RowMaj = RM;
};
template <class X, int O, class S>
struct traits<Map<X, O, S>> : traits<std::remove_const_t<X>> {};
template <class X>
struct traits<Ref<X>> : traits<std::remove_const_t<X>> {};

// fixed or dynamic storage
template <class T, int R, int C>
struct Storage {
  T d[R * C > 0 ? R * C : 1];
  Storage() {
    for (int i = 0; i < R * C; ++i) d[i] = T();
  }
  Storage(Index, Index) : Storage() {}
  T* data() { return d; }
  const T* data() const { return d; }
  static constexpr Index rows() { return R; }
  static constexpr Index cols() { return C; }
  void resize(Index r, Index c) {
    (void)r, (void)c;
    assert(r == R && c == C);
  }
};
template <class T, int R, int C>
  requires(R == Dynamic || C == Dynamic)
struct Storage<T, R, C> {
  std::vector<T> d;
  Index r = (R == Dynamic ? 0 : R), c = (C == Dynamic ? 0 : C);
  Storage() = default;
  Storage(Index rr, Index cc) : d(static_cast<size_t>(rr * cc), T()), r(rr), c(cc) {}
  T* data() { return d.data(); }
  const T* data() const { return d.data(); }
  Index rows() const { return r; }
  Index cols() const { return c; }
  void resize(Index rr, Index cc) {
    if (rr == r && cc == c) return;
    d.assign(static_cast<size_t>(rr * cc), T());
    r = rr;
    c = cc;
  }
};
}  // namespace internal

template <class D>
class ArrayBase;

// ------------------------------------------------------------------------------------------------------------------
// proxies
template <class D>
class ColwiseProxy;
template <class D>
class NoAliasProxy {
 public:
  explicit NoAliasProxy(D& d) : d_(d) {}
  template <class O>
  void operator=(const MatrixBase<O>& o) {
    d_ = o;
  }
  template <class O>
  void operator+=(const MatrixBase<O>& o) {
    d_ += o;
  }
  template <class O>
  void operator-=(const MatrixBase<O>& o) {
    d_ -= o;
  }

 private:
  D& d_;
};
template <class T, int N>
class DiagonalWrapper {
 public:
  Matrix<T, N, 1> v;
};
template <class T, int R>
class LDLT;
template <class T>
class CompleteOrthogonalDecomposition;

// arithmetic index sequences: Eigen::seq(first, last, increment) with Eigen::fix<N> constants (downscaleImage)
namespace internal {
template <int N>
struct FixedInt {
  constexpr operator Index() const { return N; }
};
}  // namespace internal
template <int N>
inline constexpr internal::FixedInt<N> fix{};
struct ArithmeticSequence {
  Index first, size, incr;
};
inline ArithmeticSequence seq(Index first, Index last, Index incr = 1) {
  return ArithmeticSequence{first, (last - first + incr) / incr, incr};
}

// ------------------------------------------------------------------------------------------------------------------
template <class Derived>
class MatrixBase {
 public:
  using Scalar = typename internal::traits<Derived>::Scalar;
  using RealScalar = Scalar;
  enum : int {
    RowsAtCompileTime = internal::traits<Derived>::Rows,
    ColsAtCompileTime = internal::traits<Derived>::Cols,
    SizeAtCompileTime = mul_dim(internal::traits<Derived>::Rows, internal::traits<Derived>::Cols),
    IsRowMajor = internal::traits<Derived>::RowMaj ? 1 : 0,
    IsVectorAtCompileTime = (internal::traits<Derived>::Rows == 1 || internal::traits<Derived>::Cols == 1) ? 1 : 0
  };
  static constexpr int R_ = RowsAtCompileTime, C_ = ColsAtCompileTime;
  static constexpr int PlainOpt_ = (IsRowMajor && !(R_ != 1 && C_ == 1)) ? RowMajor : default_options(R_, C_);
  using PlainObject = Matrix<Scalar, R_, C_, PlainOpt_>;

  MatrixBase() = default;
  MatrixBase(const MatrixBase&) = default;
  // assignment THROUGH the base (the reference writes `const_cast<MatrixBase<D>&>(x) = value`)
  MatrixBase& operator=(const MatrixBase& o) {
    derived() = o.derived();
    return *this;
  }
  template <class O>
  Derived& operator=(const MatrixBase<O>& o) {
    derived().operator=(o);
    return derived();
  }

  Derived& derived() { return *static_cast<Derived*>(this); }
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  // views handed out by const accessors are written through const_cast by the reference (Eigen's own idiom)
  Derived& const_cast_derived() const { return *const_cast<Derived*>(static_cast<const Derived*>(this)); }

  Index rows() const { return derived().rows(); }
  Index cols() const { return derived().cols(); }
  Index size() const { return rows() * cols(); }

  Scalar coeff(Index i, Index j) const { return derived().cref(i, j); }
  decltype(auto) operator()(Index i, Index j) { return derived().ref(i, j); }
  Scalar operator()(Index i, Index j) const { return derived().cref(i, j); }
  decltype(auto) operator()(Index i) { return lin(i); }
  Scalar operator()(Index i) const { return clin(i); }
  decltype(auto) operator[](Index i) { return lin(i); }
  Scalar operator[](Index i) const { return clin(i); }
  decltype(auto) x() { return lin(0); }
  decltype(auto) y() { return lin(1); }
  decltype(auto) z() { return lin(2); }
  decltype(auto) w() { return lin(3); }
  Scalar x() const { return clin(0); }
  Scalar y() const { return clin(1); }
  Scalar z() const { return clin(2); }
  Scalar w() const { return clin(3); }

  // index-list views (NormalLinearSystem::reduce_system): evaluated into a plain matrix
  Matrix<Scalar, Dynamic, Dynamic> operator()(const std::vector<int>& ri, const std::vector<int>& ci) const;
  Matrix<Scalar, Dynamic, 1> operator()(const std::vector<int>& ri) const;
  Matrix<Scalar, Dynamic, Dynamic, PlainOpt_ & RowMajor> operator()(const ArithmeticSequence& ri,
                                                                    const ArithmeticSequence& ci) const;

  // ---- blocks ------------------------------------------------------------------------------------------------
  template <int BR, int BC>
  auto block(Index i, Index j) {
    return derived().template mkblock<BR, BC>(i, j, BR, BC);
  }
  template <int BR, int BC>
  auto block(Index i, Index j) const {
    return derived().template mkcblock<BR, BC>(i, j, BR, BC);
  }
  auto block(Index i, Index j, Index r, Index c) { return derived().template mkblock<Dynamic, Dynamic>(i, j, r, c); }
  auto block(Index i, Index j, Index r, Index c) const {
    return derived().template mkcblock<Dynamic, Dynamic>(i, j, r, c);
  }
  auto row(Index i) { return derived().template mkblock<1, C_>(i, 0, 1, cols()); }
  auto row(Index i) const { return derived().template mkcblock<1, C_>(i, 0, 1, cols()); }
  auto col(Index j) { return derived().template mkblock<R_, 1>(0, j, rows(), 1); }
  auto col(Index j) const { return derived().template mkcblock<R_, 1>(0, j, rows(), 1); }
  template <int N>
  auto segment(Index i) {
    if constexpr (C_ == 1)
      return derived().template mkblock<N, 1>(i, 0, N, 1);
    else
      return derived().template mkblock<1, N>(0, i, 1, N);
  }
  template <int N>
  auto segment(Index i) const {
    if constexpr (C_ == 1)
      return derived().template mkcblock<N, 1>(i, 0, N, 1);
    else
      return derived().template mkcblock<1, N>(0, i, 1, N);
  }
  auto segment(Index i, Index n) {
    if constexpr (C_ == 1)
      return derived().template mkblock<Dynamic, 1>(i, 0, n, 1);
    else
      return derived().template mkblock<1, Dynamic>(0, i, 1, n);
  }
  auto segment(Index i, Index n) const {
    if constexpr (C_ == 1)
      return derived().template mkcblock<Dynamic, 1>(i, 0, n, 1);
    else
      return derived().template mkcblock<1, Dynamic>(0, i, 1, n);
  }
  template <int N>
  auto head() {
    return this->template segment<N>(0);
  }
  template <int N>
  auto head() const {
    return this->template segment<N>(0);
  }
  template <int N>
  auto tail() {
    return this->template segment<N>(size() - N);
  }
  template <int N>
  auto tail() const {
    return this->template segment<N>(size() - N);
  }
  auto head(Index n) { return segment(0, n); }
  auto head(Index n) const { return segment(0, n); }
  auto tail(Index n) { return segment(size() - n, n); }
  auto tail(Index n) const { return segment(size() - n, n); }
  template <int N>
  auto leftCols() {
    return derived().template mkblock<R_, N>(0, 0, rows(), N);
  }
  template <int N>
  auto leftCols() const {
    return derived().template mkcblock<R_, N>(0, 0, rows(), N);
  }
  template <int N>
  auto rightCols() {
    return derived().template mkblock<R_, N>(0, cols() - N, rows(), N);
  }
  template <int N>
  auto rightCols() const {
    return derived().template mkcblock<R_, N>(0, cols() - N, rows(), N);
  }
  auto leftCols(Index n) { return derived().template mkblock<R_, Dynamic>(0, 0, rows(), n); }
  auto rightCols(Index n) { return derived().template mkblock<R_, Dynamic>(0, cols() - n, rows(), n); }
  auto topRows(Index n) { return derived().template mkblock<Dynamic, C_>(0, 0, n, cols()); }
  auto bottomRows(Index n) { return derived().template mkblock<Dynamic, C_>(rows() - n, 0, n, cols()); }
  template <int N>
  auto topRows() {
    return derived().template mkblock<N, C_>(0, 0, N, cols());
  }
  template <int N>
  auto topRows() const {
    return derived().template mkcblock<N, C_>(0, 0, N, cols());
  }
  template <int N>
  auto bottomRows() {
    return derived().template mkblock<N, C_>(rows() - N, 0, N, cols());
  }
  template <int N>
  auto bottomRows() const {
    return derived().template mkcblock<N, C_>(rows() - N, 0, N, cols());
  }
  template <int BR, int BC>
  auto topLeftCorner() {
    return derived().template mkblock<BR, BC>(0, 0, BR, BC);
  }
  template <int BR, int BC>
  auto topLeftCorner() const {
    return derived().template mkcblock<BR, BC>(0, 0, BR, BC);
  }
  template <int BR, int BC>
  auto topRightCorner() {
    return derived().template mkblock<BR, BC>(0, cols() - BC, BR, BC);
  }
  template <int BR, int BC>
  auto topRightCorner() const {
    return derived().template mkcblock<BR, BC>(0, cols() - BC, BR, BC);
  }
  auto diagonal() {
    constexpr int N = (R_ == Dynamic || C_ == Dynamic) ? Dynamic : (R_ < C_ ? R_ : C_);
    return derived().template mkdiag<N>();
  }
  auto diagonal() const {
    constexpr int N = (R_ == Dynamic || C_ == Dynamic) ? Dynamic : (R_ < C_ ? R_ : C_);
    return derived().template mkcdiag<N>();
  }

  // ---- in-place ----------------------------------------------------------------------------------------------
  Derived& setZero() { return setConstant(Scalar(0)); }
  Derived& setOnes() { return setConstant(Scalar(1)); }
  // NAME ONLY: the general eigenvalue problem (the simple-radial camera's polynomial solver) is not on any tested path
  struct EigenvaluesNotImplemented {
    Index rows() const { std::abort(); }
    std::complex<Scalar> operator[](Index) const { std::abort(); }
    EigenvaluesNotImplemented eval() const { return *this; }
  };
  EigenvaluesNotImplemented eigenvalues() const { std::abort(); }
  Derived& setConstant(const Scalar& v) {
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) derived().ref(i, j) = v;
    return derived();
  }
  Derived& setIdentity() {
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) derived().ref(i, j) = Scalar(i == j ? 1 : 0);
    return derived();
  }
  NoAliasProxy<Derived> noalias() { return NoAliasProxy<Derived>(derived()); }

  template <class O>
  Derived& assign_from(const MatrixBase<O>& o) {
    // evaluated eagerly through a temporary so that aliasing right-hand sides behave as Eigen's do after .eval()
    if (o.rows() == rows() && o.cols() == cols()) {
      for (Index i = 0; i < rows(); ++i)
        for (Index j = 0; j < cols(); ++j) derived().ref(i, j) = Scalar(o.coeff(i, j));
    } else {
      // Eigen allows assigning a row vector to a column vector (and back)
      if (!((o.rows() == 1 || o.cols() == 1) && (rows() == 1 || cols() == 1) && o.size() == size()))
        throw std::logic_error("mini_eigen: size mismatch in assignment");
      for (Index k = 0; k < size(); ++k) lin(k) = Scalar(o.clin(k));
    }
    return derived();
  }
  template <class O>
  Derived& operator+=(const MatrixBase<O>& o) {
    auto t = o.eval();
    if (t.rows() == rows() && t.cols() == cols()) {
      for (Index i = 0; i < rows(); ++i)
        for (Index j = 0; j < cols(); ++j) derived().ref(i, j) += t.cref(i, j);
    } else {
      if (t.size() != size()) throw std::logic_error("mini_eigen: size mismatch in +=");
      for (Index k = 0; k < size(); ++k) lin(k) += t.clin(k);
    }
    return derived();
  }
  template <class O>
  Derived& operator-=(const MatrixBase<O>& o) {
    auto t = o.eval();
    if (t.rows() == rows() && t.cols() == cols()) {
      for (Index i = 0; i < rows(); ++i)
        for (Index j = 0; j < cols(); ++j) derived().ref(i, j) -= t.cref(i, j);
    } else {
      if (t.size() != size()) throw std::logic_error("mini_eigen: size mismatch in -=");
      for (Index k = 0; k < size(); ++k) lin(k) -= t.clin(k);
    }
    return derived();
  }
  template <class T2, int N2>
  Derived& operator+=(const DiagonalWrapper<T2, N2>& d) {
    for (Index i = 0; i < d.v.size(); ++i) derived().ref(i, i) += d.v[i];
    return derived();
  }
  Derived& operator*=(const Scalar& s) {
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) derived().ref(i, j) *= s;
    return derived();
  }
  Derived& operator/=(const Scalar& s) {
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) derived().ref(i, j) /= s;
    return derived();
  }

  // ---- value-returning ---------------------------------------------------------------------------------------
  PlainObject eval() const {
    PlainObject r(rows(), cols());
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) r.ref(i, j) = coeff(i, j);
    return r;
  }
  template <class F>
  auto unary(F f) const {
    using U = decltype(f(Scalar()));
    Matrix<U, R_, C_, PlainOpt_> r(rows(), cols());
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) r.ref(i, j) = f(coeff(i, j));
    return r;
  }
  template <class U>
  auto cast() const {
    return unary([](const Scalar& v) { return static_cast<U>(v); });
  }
  Matrix<Scalar, C_, R_> transpose() const {
    Matrix<Scalar, C_, R_> r(cols(), rows());
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) r.ref(j, i) = coeff(i, j);
    return r;
  }
  auto operator-() const {
    return unary([](const Scalar& v) { return -v; });
  }
  auto cwiseInverse() const {
    return unary([](const Scalar& v) { return Scalar(1) / v; });
  }
  auto cwiseSqrt() const {
    return unary([](const Scalar& v) {
      using std::sqrt;
      return sqrt(v);
    });
  }
  auto cwiseAbs() const {
    return unary([](const Scalar& v) {
      using std::abs;
      return abs(v);
    });
  }
  template <class O>
  PlainObject cwiseProduct(const MatrixBase<O>& o) const {
    PlainObject r(rows(), cols());
    if (o.rows() == rows() && o.cols() == cols()) {
      for (Index i = 0; i < rows(); ++i)
        for (Index j = 0; j < cols(); ++j) r.ref(i, j) = coeff(i, j) * o.coeff(i, j);
    } else {
      for (Index k = 0; k < size(); ++k) r.lin(k) = clin(k) * o.clin(k);
    }
    return r;
  }
  Scalar squaredNorm() const {
    Scalar s(0);
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  Scalar norm() const {
    using std::sqrt;
    return sqrt(squaredNorm());
  }
  PlainObject normalized() const { return (*this) / norm(); }
  void normalize() { (*this) /= norm(); }
  Scalar sum() const {
    Scalar s(0);
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) s += coeff(i, j);
    return s;
  }
  Scalar trace() const {
    Scalar s(0);
    for (Index i = 0; i < std::min(rows(), cols()); ++i) s += coeff(i, i);
    return s;
  }
  Scalar maxCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) m = std::max(m, coeff(i, j));
    return m;
  }
  Scalar minCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) m = std::min(m, coeff(i, j));
    return m;
  }
  template <class O>
  Scalar dot(const MatrixBase<O>& o) const {
    Scalar s(0);
    for (Index k = 0; k < size(); ++k) s += clin(k) * o.clin(k);
    return s;
  }
  template <class O>
  Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r[0] = clin(1) * o.clin(2) - clin(2) * o.clin(1);
    r[1] = clin(2) * o.clin(0) - clin(0) * o.clin(2);
    r[2] = clin(0) * o.clin(1) - clin(1) * o.clin(0);
    return r;
  }
  template <class O>
  auto lazyProduct(const MatrixBase<O>& o) const {
    return (*this) * o;
  }
  auto homogeneous() const {
    static_assert(C_ == 1);
    constexpr int N = R_ == Dynamic ? Dynamic : R_ + 1;
    Matrix<Scalar, N, 1> r(rows() + 1, 1);
    for (Index k = 0; k < rows(); ++k) r[k] = clin(k);
    r[rows()] = Scalar(1);
    return r;
  }
  auto hnormalized() const {
    static_assert(C_ == 1);
    constexpr int N = R_ == Dynamic ? Dynamic : R_ - 1;
    Matrix<Scalar, N, 1> r(rows() - 1, 1);
    for (Index k = 0; k + 1 < rows(); ++k) r[k] = clin(k) / clin(rows() - 1);
    return r;
  }
  DiagonalWrapper<Scalar, mul_dim(R_, C_)> asDiagonal() const {
    DiagonalWrapper<Scalar, mul_dim(R_, C_)> w;
    w.v = Matrix<Scalar, mul_dim(R_, C_), 1>(size(), 1);
    for (Index k = 0; k < size(); ++k) w.v[k] = clin(k);
    return w;
  }
  template <int UpLo>
  PlainObject selfadjointView() const {
    PlainObject r(rows(), cols());
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) {
        const bool stored = UpLo == Lower ? (i >= j) : (i <= j);
        r.ref(i, j) = stored ? coeff(i, j) : coeff(j, i);
      }
    return r;
  }
  bool allFinite() const {
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j)
        if (!std::isfinite(static_cast<double>(coeff(i, j)))) return false;
    return true;
  }
  bool hasNaN() const {
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j)
        if (std::isnan(static_cast<double>(coeff(i, j)))) return true;
    return false;
  }
  Array<Scalar, R_, C_, PlainOpt_> array() const;
  const Derived& matrix() const { return derived(); }
  ColwiseProxy<Derived> colwise() { return ColwiseProxy<Derived>(derived()); }
  ColwiseProxy<const Derived> colwise() const { return ColwiseProxy<const Derived>(derived()); }
  LDLT<Scalar, R_> ldlt() const;
  CompleteOrthogonalDecomposition<Scalar> completeOrthogonalDecomposition() const;
  PlainObject inverse() const;

  // ---- statics -----------------------------------------------------------------------------------------------
  static PlainObject Zero() { return PlainObject().setZero(); }
  static PlainObject Zero(Index r, Index c) { return PlainObject(r, c).setZero(); }
  static PlainObject Zero(Index n) { return PlainObject(n).setZero(); }
  static PlainObject Ones() { return PlainObject().setConstant(Scalar(1)); }
  static PlainObject Constant(const Scalar& v) { return PlainObject().setConstant(v); }
  static PlainObject Constant(Index r, Index c, const Scalar& v) { return PlainObject(r, c).setConstant(v); }
  static PlainObject Constant(Index n, const Scalar& v) { return PlainObject(n).setConstant(v); }
  static PlainObject Identity() { return PlainObject().setIdentity(); }
  static PlainObject Identity(Index r, Index c) { return PlainObject(r, c).setIdentity(); }

  // linear (vector) access; for matrices column-major order as in Eigen for column-major plain objects
  decltype(auto) lin(Index k) {
    if constexpr (C_ == 1)
      return derived().ref(k, 0);
    else if constexpr (R_ == 1)
      return derived().ref(0, k);
    else
      return rows() == 1 ? derived().ref(0, k) : (cols() == 1 ? derived().ref(k, 0) : derived().ref(k % rows(), k / rows()));
  }
  Scalar clin(Index k) const {
    if constexpr (C_ == 1)
      return derived().cref(k, 0);
    else if constexpr (R_ == 1)
      return derived().cref(0, k);
    else
      return rows() == 1 ? derived().cref(0, k) : (cols() == 1 ? derived().cref(k, 0) : derived().cref(k % rows(), k / rows()));
  }
};

// ------------------------------------------------------------------------------------------------------------------
// strided view on somebody else's coefficients
template <class T, int R, int C, bool RM>
class Block : public MatrixBase<Block<T, R, C, RM>> {
 public:
  using Base = MatrixBase<Block<T, R, C, RM>>;
  using Scalar = std::remove_const_t<T>;
  Block(T* p, Index r, Index c, Index rs, Index cs) : p_(p), r_(r), c_(c), rs_(rs), cs_(cs) {}
  Block(const Block&) = default;
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  T& ref(Index i, Index j) const { return p_[i * rs_ + j * cs_]; }
  Scalar cref(Index i, Index j) const { return p_[i * rs_ + j * cs_]; }
  T* data() const { return p_; }
  Index rowStride() const { return rs_; }
  Index colStride() const { return cs_; }

  template <int BR, int BC>
  Block<T, BR, BC, RM> mkblock(Index i, Index j, Index r, Index c) const {
    return Block<T, BR, BC, RM>(p_ + i * rs_ + j * cs_, r, c, rs_, cs_);
  }
  template <int BR, int BC>
  Block<const Scalar, BR, BC, RM> mkcblock(Index i, Index j, Index r, Index c) const {
    return Block<const Scalar, BR, BC, RM>(p_ + i * rs_ + j * cs_, r, c, rs_, cs_);
  }
  template <int N>
  Block<T, N, 1, false> mkdiag() const {
    return Block<T, N, 1, false>(p_, std::min(r_, c_), 1, rs_ + cs_, 0);
  }
  template <int N>
  Block<const Scalar, N, 1, false> mkcdiag() const {
    return Block<const Scalar, N, 1, false>(p_, std::min(r_, c_), 1, rs_ + cs_, 0);
  }

  // assignment writes THROUGH the view (also for a const view object: Eigen's blocks are handles)
  template <class O>
  const Block& operator=(const MatrixBase<O>& o) const {
    auto t = o.eval();
    const_cast<Block*>(this)->assign_from(t);
    return *this;
  }
  const Block& operator=(const Block& o) const {
    auto t = o.eval();
    const_cast<Block*>(this)->assign_from(t);
    return *this;
  }

 private:
  T* p_;
  Index r_, c_, rs_, cs_;
};

// ------------------------------------------------------------------------------------------------------------------
template <class T, int R, int C, int Opt, int MR, int MC>
class Matrix : public MatrixBase<Matrix<T, R, C, Opt, MR, MC>> {
 public:
  using Base = MatrixBase<Matrix>;
  using Scalar = T;
  static constexpr int Options_ = Opt;
  static constexpr bool RM = (Opt & RowMajor) != 0;

  Matrix() = default;
  Matrix(const Matrix&) = default;
  Matrix(Matrix&&) = default;
  // not defaulted: the base's copy assignment forwards to the derived one
  Matrix& operator=(const Matrix& o) {
    s_ = o.s_;
    return *this;
  }
  Matrix& operator=(Matrix&& o) {
    s_ = std::move(o.s_);
    return *this;
  }

  // (rows, cols) for dynamic types, two coefficients for fixed 2-vectors
  template <class A, class B>
    requires(std::is_arithmetic_v<A> && std::is_arithmetic_v<B>)
  Matrix(const A& a, const B& b) {
    if constexpr (R != Dynamic && C != Dynamic && R * C == 2) {
      s_.d[0] = T(a);
      s_.d[1] = T(b);
    } else {
      s_ = internal::Storage<T, R, C>(static_cast<Index>(a), static_cast<Index>(b));
    }
  }
  // size for dynamic vectors
  template <class A>
    requires(std::is_integral_v<A> && (R == Dynamic || C == Dynamic))
  explicit Matrix(const A& n) {
    if constexpr (C == 1)
      s_ = internal::Storage<T, R, C>(static_cast<Index>(n), 1);
    else if constexpr (R == 1)
      s_ = internal::Storage<T, R, C>(1, static_cast<Index>(n));
    else
      s_ = internal::Storage<T, R, C>(static_cast<Index>(n), static_cast<Index>(n));
  }
  Matrix(const T& a, const T& b, const T& c)
    requires(R != Dynamic && C != Dynamic && R * C == 3)
  {
    s_.d[0] = a, s_.d[1] = b, s_.d[2] = c;
  }
  Matrix(const T& a, const T& b, const T& c, const T& d)
    requires(R != Dynamic && C != Dynamic && R * C == 4)
  {
    s_.d[0] = a, s_.d[1] = b, s_.d[2] = c, s_.d[3] = d;
  }
  template <class O>
  Matrix(const MatrixBase<O>& o) {
    resize_like(o);
    this->assign_from(o);
  }
  template <int N>
  Matrix(const DiagonalWrapper<T, N>& d) {
    *this = d;
  }
  // an evaluated coefficient-wise expression assigned back to the matrix world (Eigen allows Matrix = Array)
  template <class U, int R2, int C2, int O2>
  Matrix(const Array<U, R2, C2, O2>& a) {
    s_.resize(a.rows(), a.cols());
    for (Index i = 0; i < a.rows(); ++i)
      for (Index j = 0; j < a.cols(); ++j) ref(i, j) = static_cast<T>(a(i, j));
  }

  template <class O>
  Matrix& operator=(const MatrixBase<O>& o) {
    if constexpr (std::is_same_v<O, Matrix>) {
      if (static_cast<const void*>(&o) == static_cast<const void*>(this)) return *this;
    }
    auto t = o.eval();
    resize_like(t);
    this->assign_from(t);
    return *this;
  }
  template <int N>
  Matrix& operator=(const DiagonalWrapper<T, N>& d) {
    const Index n = d.v.size();
    s_.resize(n, n);
    this->setZero();
    for (Index i = 0; i < n; ++i) ref(i, i) = d.v[i];
    return *this;
  }

  Index rows() const { return s_.rows(); }
  Index cols() const { return s_.cols(); }
  Index stride_r() const { return RM ? cols() : 1; }
  Index stride_c() const { return RM ? 1 : rows(); }
  T& ref(Index i, Index j) { return s_.data()[i * stride_r() + j * stride_c()]; }
  const T& ref(Index i, Index j) const { return s_.data()[i * stride_r() + j * stride_c()]; }
  T cref(Index i, Index j) const { return s_.data()[i * stride_r() + j * stride_c()]; }
  T* data() { return s_.data(); }
  const T* data() const { return s_.data(); }
  void resize(Index r, Index c) { s_.resize(r, c); }
  void resize(Index n) {
    if constexpr (C == 1)
      s_.resize(n, 1);
    else
      s_.resize(1, n);
  }
  void conservativeResize(Index r, Index c) {
    Matrix t(r, c);
    for (Index i = 0; i < std::min(r, rows()); ++i)
      for (Index j = 0; j < std::min(c, cols()); ++j) t.ref(i, j) = cref(i, j);
    *this = std::move(t);
  }
  void conservativeResize(Index n) {
    if constexpr (C == 1)
      conservativeResize(n, 1);
    else
      conservativeResize(1, n);
  }

  template <int BR, int BC>
  Block<T, BR, BC, RM> mkblock(Index i, Index j, Index r, Index c) {
    return Block<T, BR, BC, RM>(&ref(i, j), r, c, stride_r(), stride_c());
  }
  template <int BR, int BC>
  Block<const T, BR, BC, RM> mkcblock(Index i, Index j, Index r, Index c) const {
    return Block<const T, BR, BC, RM>(data() + i * stride_r() + j * stride_c(), r, c, stride_r(), stride_c());
  }
  template <int N>
  Block<T, N, 1, false> mkdiag() {
    return Block<T, N, 1, false>(data(), std::min(rows(), cols()), 1, stride_r() + stride_c(), 0);
  }
  template <int N>
  Block<const T, N, 1, false> mkcdiag() const {
    return Block<const T, N, 1, false>(data(), std::min(rows(), cols()), 1, stride_r() + stride_c(), 0);
  }

  // an inner product is usable as a scalar (Eigen: Product<...,InnerProduct> converts to Scalar)
  operator T() const
    requires(R == 1 && C == 1)
  {
    return s_.d[0];
  }

  template <class O>
  void resize_like(const MatrixBase<O>& o) {
    if constexpr (R == Dynamic || C == Dynamic) {
      Index r = o.rows(), c = o.cols();
      if constexpr (C == 1) {
        if (c != 1 && r == 1) std::swap(r, c);
      }
      if constexpr (R == 1) {
        if (r != 1 && c == 1) std::swap(r, c);
      }
      s_.resize(r, c);
    }
  }

 private:
  internal::Storage<T, R, C> s_;
};

// ------------------------------------------------------------------------------------------------------------------
// Map / Ref of a matrix type: a view on external memory with the plain type's layout
template <class X, int MO, class S>
class Map : public MatrixBase<Map<X, MO, S>> {
 public:
  using Plain = std::remove_const_t<X>;
  using Scalar = typename Plain::Scalar;
  using Ptr = std::conditional_t<std::is_const_v<X>, const Scalar*, Scalar*>;
  static constexpr int R = internal::traits<Plain>::Rows, C = internal::traits<Plain>::Cols;
  static constexpr bool RM = internal::traits<Plain>::RowMaj;
  Map(Ptr p) : p_(const_cast<Scalar*>(p)), r_(R), c_(C) { static_assert(R != Dynamic && C != Dynamic); }
  Map(Ptr p, Index n) : p_(const_cast<Scalar*>(p)), r_(C == 1 ? n : 1), c_(C == 1 ? 1 : n) {}
  Map(Ptr p, Index r, Index c) : p_(const_cast<Scalar*>(p)), r_(r), c_(c) {}
  Map(const Map&) = default;
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index stride_r() const { return RM ? c_ : 1; }
  Index stride_c() const { return RM ? 1 : r_; }
  Scalar& ref(Index i, Index j) const { return p_[i * stride_r() + j * stride_c()]; }
  Scalar cref(Index i, Index j) const { return p_[i * stride_r() + j * stride_c()]; }
  Ptr data() const { return p_; }
  template <int BR, int BC>
  auto mkblock(Index i, Index j, Index r, Index c) const {
    return Block<std::remove_pointer_t<Ptr>, BR, BC, RM>(p_ + i * stride_r() + j * stride_c(), r, c, stride_r(), stride_c());
  }
  template <int BR, int BC>
  auto mkcblock(Index i, Index j, Index r, Index c) const {
    return Block<const Scalar, BR, BC, RM>(p_ + i * stride_r() + j * stride_c(), r, c, stride_r(), stride_c());
  }
  template <int N>
  auto mkdiag() const {
    return Block<std::remove_pointer_t<Ptr>, N, 1, false>(p_, std::min(r_, c_), 1, stride_r() + stride_c(), 0);
  }
  template <int N>
  auto mkcdiag() const {
    return Block<const Scalar, N, 1, false>(p_, std::min(r_, c_), 1, stride_r() + stride_c(), 0);
  }
  template <class O>
  const Map& operator=(const MatrixBase<O>& o) const {
    auto t = o.eval();
    const_cast<Map*>(this)->assign_from(t);
    return *this;
  }
  const Map& operator=(const Map& o) const {
    auto t = o.eval();
    const_cast<Map*>(this)->assign_from(t);
    return *this;
  }

 private:
  Scalar* p_;
  Index r_, c_;
};

template <class X>
class Ref : public MatrixBase<Ref<X>> {
 public:
  using Plain = std::remove_const_t<X>;
  using Scalar = typename Plain::Scalar;
  static constexpr int R = internal::traits<Plain>::Rows, C = internal::traits<Plain>::Cols;
  static constexpr bool RM = internal::traits<Plain>::RowMaj;
  template <class T, int BR, int BC, bool BRM>
  Ref(const Block<T, BR, BC, BRM>& b)
      : p_(const_cast<Scalar*>(b.data())), r_(b.rows()), c_(b.cols()), rs_(b.rowStride()), cs_(b.colStride()) {}
  Ref(Plain& m) : p_(m.data()), r_(m.rows()), c_(m.cols()), rs_(m.stride_r()), cs_(m.stride_c()) {}
  Ref(const Plain& m)
    requires(std::is_const_v<X>)
      : p_(const_cast<Scalar*>(m.data())), r_(m.rows()), c_(m.cols()), rs_(m.stride_r()), cs_(m.stride_c()) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Scalar& ref(Index i, Index j) const { return p_[i * rs_ + j * cs_]; }
  Scalar cref(Index i, Index j) const { return p_[i * rs_ + j * cs_]; }
  template <int BR, int BC>
  auto mkblock(Index i, Index j, Index r, Index c) const {
    return Block<Scalar, BR, BC, RM>(p_ + i * rs_ + j * cs_, r, c, rs_, cs_);
  }
  template <int BR, int BC>
  auto mkcblock(Index i, Index j, Index r, Index c) const {
    return Block<const Scalar, BR, BC, RM>(p_ + i * rs_ + j * cs_, r, c, rs_, cs_);
  }
  template <int N>
  auto mkdiag() const {
    return Block<Scalar, N, 1, false>(p_, std::min(r_, c_), 1, rs_ + cs_, 0);
  }
  template <int N>
  auto mkcdiag() const {
    return Block<const Scalar, N, 1, false>(p_, std::min(r_, c_), 1, rs_ + cs_, 0);
  }
  template <class O>
  const Ref& operator=(const MatrixBase<O>& o) const {
    auto t = o.eval();
    const_cast<Ref*>(this)->assign_from(t);
    return *this;
  }

 private:
  Scalar* p_;
  Index r_, c_, rs_, cs_;
};

// ------------------------------------------------------------------------------------------------------------------
// free operators (all eager)
template <class A, class B>
auto operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  if (a.rows() == b.rows() && a.cols() == b.cols()) {
    for (Index i = 0; i < a.rows(); ++i)
      for (Index j = 0; j < a.cols(); ++j) r.ref(i, j) = a.coeff(i, j) + b.coeff(i, j);
  } else {
    if (a.size() != b.size()) throw std::logic_error("mini_eigen: size mismatch in +");
    for (Index k = 0; k < a.size(); ++k) r.lin(k) = a.clin(k) + b.clin(k);
  }
  return r;
}
template <class A, class B>
auto operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  if (a.rows() == b.rows() && a.cols() == b.cols()) {
    for (Index i = 0; i < a.rows(); ++i)
      for (Index j = 0; j < a.cols(); ++j) r.ref(i, j) = a.coeff(i, j) - b.coeff(i, j);
  } else {
    if (a.size() != b.size()) throw std::logic_error("mini_eigen: size mismatch in -");
    for (Index k = 0; k < a.size(); ++k) r.lin(k) = a.clin(k) - b.clin(k);
  }
  return r;
}
template <class A>
auto operator*(const MatrixBase<A>& a, const typename MatrixBase<A>::Scalar& s) {
  return a.unary([&](const typename MatrixBase<A>::Scalar& v) { return v * s; });
}
template <class A>
auto operator*(const typename MatrixBase<A>::Scalar& s, const MatrixBase<A>& a) {
  return a.unary([&](const typename MatrixBase<A>::Scalar& v) { return s * v; });
}
template <class A>
auto operator/(const MatrixBase<A>& a, const typename MatrixBase<A>::Scalar& s) {
  return a.unary([&](const typename MatrixBase<A>::Scalar& v) { return v / s; });
}
template <class A, class B>
auto operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  using T = typename MatrixBase<A>::Scalar;
  constexpr int RR = MatrixBase<A>::R_, CC = MatrixBase<B>::C_;
  if (a.cols() != b.rows()) throw std::logic_error("mini_eigen: size mismatch in product");
  Matrix<T, RR, CC> r(a.rows(), b.cols());
  for (Index i = 0; i < a.rows(); ++i)
    for (Index j = 0; j < b.cols(); ++j) {
      T s(0);
      for (Index k = 0; k < a.cols(); ++k) s += a.coeff(i, k) * b.coeff(k, j);
      r.ref(i, j) = s;
    }
  return r;
}
template <class T, int N, class B>
auto operator*(const DiagonalWrapper<T, N>& d, const MatrixBase<B>& b) {
  typename MatrixBase<B>::PlainObject r(b.rows(), b.cols());
  for (Index i = 0; i < b.rows(); ++i)
    for (Index j = 0; j < b.cols(); ++j) r.ref(i, j) = d.v[i] * b.coeff(i, j);
  return r;
}
template <class A, class T, int N>
auto operator*(const MatrixBase<A>& a, const DiagonalWrapper<T, N>& d) {
  typename MatrixBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index i = 0; i < a.rows(); ++i)
    for (Index j = 0; j < a.cols(); ++j) r.ref(i, j) = a.coeff(i, j) * d.v[j];
  return r;
}
template <class A, class B>
bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
  for (Index i = 0; i < a.rows(); ++i)
    for (Index j = 0; j < a.cols(); ++j)
      if (!(a.coeff(i, j) == b.coeff(i, j))) return false;
  return true;
}

template <class D>
Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::operator()(const std::vector<int>& ri,
                                                                                  const std::vector<int>& ci) const {
  Matrix<Scalar, Dynamic, Dynamic> r(static_cast<Index>(ri.size()), static_cast<Index>(ci.size()));
  for (size_t i = 0; i < ri.size(); ++i)
    for (size_t j = 0; j < ci.size(); ++j) r.ref(static_cast<Index>(i), static_cast<Index>(j)) = coeff(ri[i], ci[j]);
  return r;
}
template <class D>
Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic, MatrixBase<D>::PlainOpt_ & RowMajor> MatrixBase<D>::operator()(
    const ArithmeticSequence& ri, const ArithmeticSequence& ci) const {
  Matrix<Scalar, Dynamic, Dynamic, PlainOpt_ & RowMajor> r(ri.size, ci.size);
  for (Index i = 0; i < ri.size; ++i)
    for (Index j = 0; j < ci.size; ++j) r.ref(i, j) = coeff(ri.first + i * ri.incr, ci.first + j * ci.incr);
  return r;
}
template <class D>
Matrix<typename MatrixBase<D>::Scalar, Dynamic, 1> MatrixBase<D>::operator()(const std::vector<int>& ri) const {
  Matrix<Scalar, Dynamic, 1> r(static_cast<Index>(ri.size()), 1);
  for (size_t i = 0; i < ri.size(); ++i) r[static_cast<Index>(i)] = clin(ri[i]);
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// comma initialiser: scalars in row-major order; for vector targets sub-vectors are appended
template <class D>
class CommaInitializer {
 public:
  CommaInitializer(D& d) : d_(d) {}
  CommaInitializer& operator,(const typename D::Scalar& s) {
    put(s);
    return *this;
  }
  template <class O>
  CommaInitializer& operator,(const MatrixBase<O>& o) {
    for (Index k = 0; k < o.size(); ++k) put(o.clin(k));
    return *this;
  }
  void put(const typename D::Scalar& s) {
    if (d_.cols() == 1 || d_.rows() == 1)
      d_.lin(k_) = s;
    else
      d_.ref(k_ / d_.cols(), k_ % d_.cols()) = s;
    ++k_;
  }

 private:
  D& d_;
  Index k_ = 0;
};
template <class D>
CommaInitializer<D> operator<<(MatrixBase<D>& m, const typename MatrixBase<D>::Scalar& s) {
  CommaInitializer<D> c(m.derived());
  c.put(s);
  return c;
}
template <class D, class O>
CommaInitializer<D> operator<<(MatrixBase<D>& m, const MatrixBase<O>& o) {
  CommaInitializer<D> c(m.derived());
  c, o;
  return c;
}

// ------------------------------------------------------------------------------------------------------------------
// colwise()
template <class D>
class ColwiseProxy {
 public:
  using M = std::remove_const_t<D>;
  using Scalar = typename M::Scalar;
  explicit ColwiseProxy(D& d) : d_(d) {}
  template <class V>
  auto operator+(const MatrixBase<V>& v) const {
    auto r = d_.eval();
    for (Index i = 0; i < r.rows(); ++i)
      for (Index j = 0; j < r.cols(); ++j) r.ref(i, j) += v.clin(i);
    return r;
  }
  template <class V>
  void operator+=(const MatrixBase<V>& v) {
    auto t = v.eval();
    for (Index i = 0; i < d_.rows(); ++i)
      for (Index j = 0; j < d_.cols(); ++j) d_.ref(i, j) += t.clin(i);
  }
  auto hnormalized() const {
    constexpr int RR = M::R_ == Dynamic ? Dynamic : M::R_ - 1;
    Matrix<Scalar, RR, M::C_> r(d_.rows() - 1, d_.cols());
    for (Index j = 0; j < d_.cols(); ++j)
      for (Index i = 0; i + 1 < d_.rows(); ++i) r.ref(i, j) = d_.coeff(i, j) / d_.coeff(d_.rows() - 1, j);
    return r;
  }
  auto squaredNorm() const {
    Matrix<Scalar, 1, M::C_> r(1, d_.cols());
    for (Index j = 0; j < d_.cols(); ++j) {
      Scalar s(0);
      for (Index i = 0; i < d_.rows(); ++i) s += d_.coeff(i, j) * d_.coeff(i, j);
      r.ref(0, j) = s;
    }
    return r;
  }

 private:
  D& d_;
};

// ------------------------------------------------------------------------------------------------------------------
// Array: coefficient-wise world.  Plain storage; also usable with non-arithmetic element types (PixelInfo).
template <class T, int R, int C, int Opt, int MR, int MC>
class Array {
 public:
  using Scalar = T;
  static constexpr bool RM = (Opt & RowMajor) != 0;
  Array() = default;
  Array(Index r, Index c) : s_(r, c) {}
  Index rows() const { return s_.rows(); }
  Index cols() const { return s_.cols(); }
  Index size() const { return rows() * cols(); }
  T& operator()(Index i, Index j) { return s_.data()[RM ? i * cols() + j : i + j * rows()]; }
  const T& operator()(Index i, Index j) const { return s_.data()[RM ? i * cols() + j : i + j * rows()]; }
  T& lin(Index k) { return s_.data()[k]; }
  const T& lin(Index k) const { return s_.data()[k]; }
  T& operator()(Index k) { return s_.data()[k]; }
  const T& operator()(Index k) const { return s_.data()[k]; }
  T* data() { return s_.data(); }

  template <class F>
  auto map(F f) const {
    using U = decltype(f(T()));
    Array<U, R, C, Opt> r(rows(), cols());
    for (Index k = 0; k < size(); ++k) r.lin(k) = f(lin(k));
    return r;
  }
  template <class F, class O>
  auto zip(const O& o, F f) const {
    using U = decltype(f(T(), T()));
    Array<U, R, C, Opt> r(rows(), cols());
    for (Index k = 0; k < size(); ++k) r.lin(k) = f(lin(k), o.lin(k));
    return r;
  }
  bool all() const {
    for (Index k = 0; k < size(); ++k)
      if (!lin(k)) return false;
    return true;
  }
  bool any() const {
    for (Index k = 0; k < size(); ++k)
      if (lin(k)) return true;
    return false;
  }
  template <class F>
  auto unaryExpr(F f) const {
    return map(f);
  }
  template <class U>
  auto cast() const {
    return map([](const T& v) { return static_cast<U>(v); });
  }
  auto square() const {
    return map([](const T& v) { return v * v; });
  }
  auto sqrt() const {
    return map([](const T& v) {
      using std::sqrt;
      return sqrt(v);
    });
  }
  auto abs() const {
    return map([](const T& v) {
      using std::abs;
      return abs(v);
    });
  }
  auto inverse() const {
    return map([](const T& v) { return T(1) / v; });
  }
  T sum() const {
    T s(0);
    for (Index k = 0; k < size(); ++k) s += lin(k);
    return s;
  }
  Matrix<T, R, C, Opt> matrix() const {
    Matrix<T, R, C, Opt> m(rows(), cols());
    for (Index i = 0; i < rows(); ++i)
      for (Index j = 0; j < cols(); ++j) m.ref(i, j) = (*this)(i, j);
    return m;
  }

 private:
  internal::Storage<T, R, C> s_;
};
#define MINI_EIGEN_ARRAY_CMP(op)                                          \
  template <class T, int R, int C, int O>                                 \
  auto operator op(const Array<T, R, C, O>& a, const T& s) {              \
    return a.map([&](const T& v) { return v op s; });                     \
  }                                                                       \
  template <class T, int R, int C, int O>                                 \
  auto operator op(const Array<T, R, C, O>& a, const Array<T, R, C, O>& b) { \
    return a.zip(b, [](const T& x, const T& y) { return x op y; });       \
  }
MINI_EIGEN_ARRAY_CMP(>=)
MINI_EIGEN_ARRAY_CMP(<=)
MINI_EIGEN_ARRAY_CMP(>)
MINI_EIGEN_ARRAY_CMP(<)
MINI_EIGEN_ARRAY_CMP(+)
MINI_EIGEN_ARRAY_CMP(-)
MINI_EIGEN_ARRAY_CMP(*)
MINI_EIGEN_ARRAY_CMP(/)
#undef MINI_EIGEN_ARRAY_CMP
template <class T, int R, int C, int O>
auto operator+(const Array<T, R, C, O>& a, int s) {
  return a.map([&](const T& v) { return v + T(s); });
}
template <class T, int R, int C, int O>
auto operator*(const T& s, const Array<T, R, C, O>& a) {
  return a.map([&](const T& v) { return s * v; });
}
template <class T, int R, int C, int O>
auto round(const Array<T, R, C, O>& a) {  // std::round coefficient-wise: halves away from zero
  return a.map([](const T& v) {
    using std::round;
    return round(v);
  });
}
template <class T, int R, int C, int O>
auto operator/(const T& s, const Array<T, R, C, O>& a) {
  return a.map([&](const T& v) { return s / v; });
}
template <class T, int R, int C, int O>
auto operator-(const Array<T, R, C, O>& a) {
  return a.map([](const T& v) { return -v; });
}

template <class D>
Array<typename MatrixBase<D>::Scalar, MatrixBase<D>::R_, MatrixBase<D>::C_, MatrixBase<D>::PlainOpt_>
MatrixBase<D>::array() const {
  Array<Scalar, R_, C_, PlainOpt_> a(rows(), cols());
  for (Index i = 0; i < rows(); ++i)
    for (Index j = 0; j < cols(); ++j) a(i, j) = coeff(i, j);
  return a;
}

// Map of an Array (PixelMap's storage of PixelInfo records; photometricallyCorrectedImage's views of the rasters)
template <class T, class TP, int R, int C, int Opt>
class ArrayMapImpl {
 public:
  static constexpr bool RM = (Opt & RowMajor) != 0;
  ArrayMapImpl(TP* p, Index r, Index c) : p_(p), r_(r), c_(c) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Index size() const { return r_ * c_; }
  TP& operator()(Index i, Index j) const { return p_[RM ? i * c_ + j : i + j * r_]; }
  TP& operator()(Index k) const { return p_[k]; }
  TP* data() const { return p_; }
  template <class F>
  auto unaryExpr(F f) const {
    using U = std::decay_t<decltype(f(p_[0]))>;
    Array<U, R, C, Opt> a(r_, c_);
    for (Index k = 0; k < size(); ++k) a.lin(k) = f(p_[k]);
    return a;
  }
  template <class U>
  auto cast() const {
    return unaryExpr([](const T& v) { return static_cast<U>(v); });
  }
  // coefficient-wise assignment from an evaluated array of the same shape and storage order
  template <class U, int R2, int C2, int O2>
  const ArrayMapImpl& operator=(const Array<U, R2, C2, O2>& a) const {
    static_assert(((O2 & RowMajor) != 0) == RM, "storage orders differ");
    for (Index k = 0; k < size(); ++k) p_[k] = a.lin(k);
    return *this;
  }
  template <class U, int R2, int C2, int O2>
  const ArrayMapImpl& operator*=(const Array<U, R2, C2, O2>& a) const {
    static_assert(((O2 & RowMajor) != 0) == RM, "storage orders differ");
    for (Index k = 0; k < size(); ++k) p_[k] *= a.lin(k);
    return *this;
  }

 private:
  TP* p_;
  Index r_, c_;
};
template <class T, int R, int C, int Opt, int MR, int MC, int MO, class S>
class Map<Array<T, R, C, Opt, MR, MC>, MO, S> : public ArrayMapImpl<T, T, R, C, Opt> {
 public:
  using ArrayMapImpl<T, T, R, C, Opt>::ArrayMapImpl;
  using ArrayMapImpl<T, T, R, C, Opt>::operator=;
};
template <class T, int R, int C, int Opt, int MR, int MC, int MO, class S>
class Map<const Array<T, R, C, Opt, MR, MC>, MO, S> : public ArrayMapImpl<T, const T, R, C, Opt> {
 public:
  using ArrayMapImpl<T, const T, R, C, Opt>::ArrayMapImpl;
};

// ------------------------------------------------------------------------------------------------------------------
// LDL^T with symmetric diagonal pivoting (Eigen::LDLT's strategy: largest remaining diagonal entry first)
template <class T, int R>
class LDLT {
 public:
  template <class D>
  explicit LDLT(const MatrixBase<D>& a) : n_(a.rows()), l_(a.rows(), a.rows()), p_(static_cast<size_t>(a.rows())) {
    for (Index i = 0; i < n_; ++i)
      for (Index j = 0; j < n_; ++j) l_.ref(i, j) = a.coeff(i, j);
    for (Index i = 0; i < n_; ++i) p_[static_cast<size_t>(i)] = i;
    for (Index k = 0; k < n_; ++k) {
      Index piv = k;
      T best = std::abs(l_.ref(k, k));
      for (Index i = k + 1; i < n_; ++i)
        if (std::abs(l_.ref(i, i)) > best) best = std::abs(l_.ref(i, i)), piv = i;
      if (piv != k) {
        for (Index j = 0; j < n_; ++j) std::swap(l_.ref(k, j), l_.ref(piv, j));
        for (Index i = 0; i < n_; ++i) std::swap(l_.ref(i, k), l_.ref(i, piv));
        std::swap(p_[static_cast<size_t>(k)], p_[static_cast<size_t>(piv)]);
      }
      const T d = l_.ref(k, k);
      if (d == T(0)) continue;
      for (Index i = k + 1; i < n_; ++i) l_.ref(i, k) /= d;
      for (Index j = k + 1; j < n_; ++j)
        for (Index i = j; i < n_; ++i) l_.ref(i, j) -= l_.ref(i, k) * d * l_.ref(j, k);
      // keep the trailing block symmetric for the next pivot search / swaps
      for (Index j = k + 1; j < n_; ++j)
        for (Index i = j + 1; i < n_; ++i) l_.ref(j, i) = l_.ref(i, j);
    }
  }
  template <class D>
  Matrix<T, R, 1> solve(const MatrixBase<D>& b) const {
    std::vector<T> y(static_cast<size_t>(n_));
    for (Index i = 0; i < n_; ++i) y[static_cast<size_t>(i)] = b.clin(p_[static_cast<size_t>(i)]);
    for (Index i = 0; i < n_; ++i)
      for (Index k = 0; k < i; ++k) y[static_cast<size_t>(i)] -= l_.cref(i, k) * y[static_cast<size_t>(k)];
    for (Index i = 0; i < n_; ++i) {
      const T d = l_.cref(i, i);
      y[static_cast<size_t>(i)] = d == T(0) ? T(0) : y[static_cast<size_t>(i)] / d;
    }
    for (Index i = n_ - 1; i >= 0; --i)
      for (Index k = i + 1; k < n_; ++k) y[static_cast<size_t>(i)] -= l_.cref(k, i) * y[static_cast<size_t>(k)];
    Matrix<T, R, 1> x(n_, 1);
    for (Index i = 0; i < n_; ++i) x[p_[static_cast<size_t>(i)]] = y[static_cast<size_t>(i)];
    return x;
  }

 private:
  Index n_;
  Matrix<T, Dynamic, Dynamic> l_;
  std::vector<Index> p_;
};
template <class D>
LDLT<typename MatrixBase<D>::Scalar, MatrixBase<D>::R_> MatrixBase<D>::ldlt() const {
  return LDLT<Scalar, R_>(*this);
}
// Stand-in for Eigen's rank-revealing complete orthogonal decomposition, used by the reference only for
// `.pseudoInverse()` (normal_linear_system.cpp:33-36).  Computed here from a one-sided Jacobi SVD; singular values below
// eps * max(rows, cols) * sigma_max are treated as zero (Eigen's default rank threshold is of the same form, on the pivots
// of its column-pivoted QR), so the two agree whenever the rank decision is not borderline.
template <class T>
class CompleteOrthogonalDecomposition {
 public:
  template <class D>
  explicit CompleteOrthogonalDecomposition(const MatrixBase<D>& a) : a_(a) {}
  Matrix<T, Dynamic, Dynamic> pseudoInverse() const {
    const Index m = a_.rows(), n = a_.cols();
    if (m < n) {
      CompleteOrthogonalDecomposition<T> t(a_.transpose());
      return t.pseudoInverse().transpose();
    }
    Matrix<T, Dynamic, Dynamic> u = a_, v(n, n);
    v.setIdentity();
    for (int sweep = 0; sweep < 60; ++sweep) {
      T off = T(0);
      for (Index p = 0; p < n; ++p)
        for (Index q = p + 1; q < n; ++q) {
          T alpha = T(0), beta = T(0), gamma = T(0);
          for (Index i = 0; i < m; ++i) {
            alpha += u.ref(i, p) * u.ref(i, p);
            beta += u.ref(i, q) * u.ref(i, q);
            gamma += u.ref(i, p) * u.ref(i, q);
          }
          if (gamma == T(0)) continue;
          off = std::max(off, std::abs(gamma) / std::sqrt(alpha * beta + std::numeric_limits<T>::min()));
          const T zeta = (beta - alpha) / (T(2) * gamma);
          const T t = (zeta >= T(0) ? T(1) : T(-1)) / (std::abs(zeta) + std::sqrt(T(1) + zeta * zeta));
          const T c = T(1) / std::sqrt(T(1) + t * t), sn = c * t;
          for (Index i = 0; i < m; ++i) {
            const T up = u.ref(i, p), uq = u.ref(i, q);
            u.ref(i, p) = c * up - sn * uq;
            u.ref(i, q) = sn * up + c * uq;
          }
          for (Index i = 0; i < n; ++i) {
            const T vp = v.ref(i, p), vq = v.ref(i, q);
            v.ref(i, p) = c * vp - sn * vq;
            v.ref(i, q) = sn * vp + c * vq;
          }
        }
      if (off < std::numeric_limits<T>::epsilon()) break;
    }
    std::vector<T> sigma(static_cast<size_t>(n));
    T smax = T(0);
    for (Index j = 0; j < n; ++j) {
      T s2 = T(0);
      for (Index i = 0; i < m; ++i) s2 += u.ref(i, j) * u.ref(i, j);
      sigma[static_cast<size_t>(j)] = std::sqrt(s2);
      smax = std::max(smax, sigma[static_cast<size_t>(j)]);
    }
    const T thr = std::numeric_limits<T>::epsilon() * T(std::max(m, n)) * smax;
    Matrix<T, Dynamic, Dynamic> r(n, m);
    r.setZero();
    for (Index j = 0; j < n; ++j) {
      const T sj = sigma[static_cast<size_t>(j)];
      if (!(sj > thr)) continue;
      // A = sum_j (u_j / s_j) s_j v_j^T  ->  A^+ = sum_j v_j (u_j / s_j)^T / s_j
      for (Index a = 0; a < n; ++a)
        for (Index b = 0; b < m; ++b) r.ref(a, b) += v.ref(a, j) * u.ref(b, j) / (sj * sj);
    }
    return r;
  }

 private:
  Matrix<T, Dynamic, Dynamic> a_;
};
template <class D>
CompleteOrthogonalDecomposition<typename MatrixBase<D>::Scalar> MatrixBase<D>::completeOrthogonalDecomposition() const {
  return CompleteOrthogonalDecomposition<Scalar>(*this);
}
// general inverse by Gauss-Jordan with partial pivoting (small matrices only)
template <class D>
typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const {
  const Index n = rows();
  Matrix<Scalar, Dynamic, Dynamic> a(n, 2 * n);
  for (Index i = 0; i < n; ++i)
    for (Index j = 0; j < n; ++j) a.ref(i, j) = coeff(i, j), a.ref(i, n + j) = Scalar(i == j ? 1 : 0);
  for (Index k = 0; k < n; ++k) {
    Index piv = k;
    for (Index i = k + 1; i < n; ++i)
      if (std::abs(a.ref(i, k)) > std::abs(a.ref(piv, k))) piv = i;
    for (Index j = 0; j < 2 * n; ++j) std::swap(a.ref(k, j), a.ref(piv, j));
    const Scalar d = a.ref(k, k);
    for (Index j = 0; j < 2 * n; ++j) a.ref(k, j) /= d;
    for (Index i = 0; i < n; ++i)
      if (i != k) {
        const Scalar f = a.ref(i, k);
        for (Index j = 0; j < 2 * n; ++j) a.ref(i, j) -= f * a.ref(k, j);
      }
  }
  PlainObject r(n, n);
  for (Index i = 0; i < n; ++i)
    for (Index j = 0; j < n; ++j) r.ref(i, j) = a.ref(i, n + j);
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Quaternion (coefficients stored x, y, z, w as in Eigen)
template <class T>
class Quaternion {
 public:
  using Scalar = T;
  Quaternion() : c_{T(0), T(0), T(0), T(1)} {}
  Quaternion(const T& w, const T& x, const T& y, const T& z) : c_{x, y, z, w} {}
  template <class D>
  explicit Quaternion(const MatrixBase<D>& m) {
    if (m.rows() == 3 && m.cols() == 3) {
      // Eigen's quaternion-from-rotation-matrix (Shoemake)
      const T t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
      if (t > T(0)) {
        T s = std::sqrt(t + T(1));
        c_[3] = T(0.5) * s;
        s = T(0.5) / s;
        c_[0] = (m.coeff(2, 1) - m.coeff(1, 2)) * s;
        c_[1] = (m.coeff(0, 2) - m.coeff(2, 0)) * s;
        c_[2] = (m.coeff(1, 0) - m.coeff(0, 1)) * s;
      } else {
        int i = 0;
        if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
        if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        T s = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + T(1));
        c_[i] = T(0.5) * s;
        s = T(0.5) / s;
        c_[3] = (m.coeff(k, j) - m.coeff(j, k)) * s;
        c_[j] = (m.coeff(j, i) + m.coeff(i, j)) * s;
        c_[k] = (m.coeff(k, i) + m.coeff(i, k)) * s;
      }
    } else {
      for (int i = 0; i < 4; ++i) c_[i] = m.clin(i);
    }
  }
  T& x() { return c_[0]; }
  T& y() { return c_[1]; }
  T& z() { return c_[2]; }
  T& w() { return c_[3]; }
  T x() const { return c_[0]; }
  T y() const { return c_[1]; }
  T z() const { return c_[2]; }
  T w() const { return c_[3]; }
  Matrix<T, 4, 1> coeffs() const { return Matrix<T, 4, 1>(c_[0], c_[1], c_[2], c_[3]); }
  Matrix<T, 3, 1> vec() const { return Matrix<T, 3, 1>(c_[0], c_[1], c_[2]); }
  T squaredNorm() const { return c_[0] * c_[0] + c_[1] * c_[1] + c_[2] * c_[2] + c_[3] * c_[3]; }
  T norm() const { return std::sqrt(squaredNorm()); }
  void normalize() {
    const T n = norm();
    for (auto& v : c_) v /= n;
  }
  Quaternion conjugate() const { return Quaternion(c_[3], -c_[0], -c_[1], -c_[2]); }
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  // Eigen's QuaternionBase::_transformVector
  template <class D>
  Matrix<T, 3, 1> _transformVector(const MatrixBase<D>& v) const {
    Matrix<T, 3, 1> q = vec();
    Matrix<T, 3, 1> vv = v;
    Matrix<T, 3, 1> uv = q.cross(vv);
    uv = uv + uv;
    return vv + w() * uv + q.cross(uv);
  }
  template <class D>
  Matrix<T, 3, 1> operator*(const MatrixBase<D>& v) const {
    return _transformVector(v);
  }
  // Eigen's QuaternionBase::toRotationMatrix
  Matrix<T, 3, 3> toRotationMatrix() const {
    Matrix<T, 3, 3> res;
    const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
    const T twx = tx * w(), twy = ty * w(), twz = tz * w();
    const T txx = tx * x(), txy = ty * x(), txz = tz * x();
    const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res(0, 0) = T(1) - (tyy + tzz);
    res(0, 1) = txy - twz;
    res(0, 2) = txz + twy;
    res(1, 0) = txy + twz;
    res(1, 1) = T(1) - (txx + tzz);
    res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy;
    res(2, 1) = tyz + twx;
    res(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  template <class U>
  Quaternion<U> cast() const {
    return Quaternion<U>(U(w()), U(x()), U(y()), U(z()));
  }

 private:
  T c_[4];
};
using Quaterniond = Quaternion<double>;
using Quaternionf = Quaternion<float>;

template <class T>
struct NumTraits {
  static constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
  static constexpr T dummy_precision() { return T(1e-12); }
  static constexpr T highest() { return std::numeric_limits<T>::max(); }
  static constexpr T lowest() { return std::numeric_limits<T>::lowest(); }
};
template <class M>
class JacobiSVD;  // named by the reference's pseudoInverse (uncertainty path, not pinned here)

}  // namespace Eigen
