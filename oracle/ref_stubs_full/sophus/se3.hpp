// STAND-IN -- this is NOT Sophus.  Test infrastructure only (see oracle/build_ref_pba.py).
//
// The reference's motion type energy::motion::SE3<Scalar> derives from Sophus::SE3<Scalar> (se3_motion.hpp:17).  Sophus
// is not installed here; this header restates the part of its published interface the reference's photometric bundle
// adjustment touches, with Sophus' formulas: SO3 as a unit quaternion, exp through the half-angle with the Taylor branch
// below |theta| = 1e-10 (Sophus::SO3::expAndTheta), SE3::exp with the left Jacobian V of SO3 (Sophus::SE3::exp),
// Adj = [[R, hat(t) R], [0, R]], tangent order (translation, rotation).
#pragma once
#include <Eigen/Dense>
#include <cmath>

namespace Sophus {

template <class Scalar>
struct Constants {
  static Scalar epsilon() { return Scalar(1e-10); }
  static Scalar pi() { return Scalar(3.141592653589793238462643383279502884); }
};
template <>
struct Constants<float> {
  static float epsilon() { return 1e-5f; }
  static float pi() { return 3.141592653589793238462643383279502884f; }
};

template <class Derived>
class SO3Base {};
template <class Derived>
class SE3Base {};

template <class Scalar_>
class SO3 : public SO3Base<SO3<Scalar_>> {
 public:
  using Scalar = Scalar_;
  static constexpr int DoF = 3;
  static constexpr int num_parameters = 4;
  using Tangent = Eigen::Matrix<Scalar, 3, 1>;
  using Point = Eigen::Matrix<Scalar, 3, 1>;
  using Transformation = Eigen::Matrix<Scalar, 3, 3>;

  SO3() = default;
  explicit SO3(const Eigen::Quaternion<Scalar>& q) : q_(q) { q_.normalize(); }
  SO3(const Transformation& R) : q_(R) { q_.normalize(); }

  static Transformation hat(const Tangent& omega) {
    Transformation Omega;
    Omega(0, 0) = Scalar(0), Omega(0, 1) = -omega(2), Omega(0, 2) = omega(1);
    Omega(1, 0) = omega(2), Omega(1, 1) = Scalar(0), Omega(1, 2) = -omega(0);
    Omega(2, 0) = -omega(1), Omega(2, 1) = omega(0), Omega(2, 2) = Scalar(0);
    return Omega;
  }
  template <class D>
  static Transformation hat(const Eigen::MatrixBase<D>& omega) {
    return hat(Tangent(omega));
  }

  static SO3 expAndTheta(const Tangent& omega, Scalar* theta) {
    using std::abs;
    using std::cos;
    using std::sin;
    using std::sqrt;
    const Scalar theta_sq = omega.squaredNorm();
    Scalar imag_factor, real_factor;
    if (theta_sq < Constants<Scalar>::epsilon() * Constants<Scalar>::epsilon()) {
      *theta = Scalar(0);
      const Scalar theta_po4 = theta_sq * theta_sq;
      imag_factor = Scalar(0.5) - Scalar(1.0 / 48.0) * theta_sq + Scalar(1.0 / 3840.0) * theta_po4;
      real_factor = Scalar(1) - Scalar(1.0 / 8.0) * theta_sq + Scalar(1.0 / 384.0) * theta_po4;
    } else {
      *theta = sqrt(theta_sq);
      const Scalar half_theta = Scalar(0.5) * (*theta);
      const Scalar sin_half_theta = sin(half_theta);
      imag_factor = sin_half_theta / (*theta);
      real_factor = cos(half_theta);
    }
    SO3 q;
    q.q_ = Eigen::Quaternion<Scalar>(real_factor, imag_factor * omega(0), imag_factor * omega(1), imag_factor * omega(2));
    return q;
  }
  static SO3 exp(const Tangent& omega) {
    Scalar theta;
    return expAndTheta(omega, &theta);
  }
  static SO3 fitToSO3(const Transformation& R) { return SO3(R); }

  const Eigen::Quaternion<Scalar>& unit_quaternion() const { return q_; }
  Transformation matrix() const { return q_.toRotationMatrix(); }
  SO3 inverse() const {
    SO3 r;
    r.q_ = q_.conjugate();
    return r;
  }
  SO3 operator*(const SO3& o) const {
    // Sophus::SO3Base::operator*: quaternion product, renormalised only when the squared norm has drifted
    SO3 r;
    r.q_ = q_ * o.q_;
    const Scalar sn = r.q_.squaredNorm();
    if (sn != Scalar(1)) {
      const Scalar scale = Scalar(2.0) / (Scalar(1.0) + sn);
      r.q_ = Eigen::Quaternion<Scalar>(r.q_.w() * scale, r.q_.x() * scale, r.q_.y() * scale, r.q_.z() * scale);
    }
    return r;
  }
  template <class D>
  Point operator*(const Eigen::MatrixBase<D>& p) const {
    return q_._transformVector(p);
  }
  template <class U>
  SO3<U> cast() const {
    SO3<U> r;
    r.setQuaternionUnchecked(q_.template cast<U>());
    return r;
  }
  void setQuaternionUnchecked(const Eigen::Quaternion<Scalar>& q) { q_ = q; }
  Scalar* data() { return &q_.x(); }

 private:
  Eigen::Quaternion<Scalar> q_;
};

template <class Scalar_>
class SE3 : public SE3Base<SE3<Scalar_>> {
 public:
  using Scalar = Scalar_;
  static constexpr int DoF = 6;
  static constexpr int num_parameters = 7;
  using Tangent = Eigen::Matrix<Scalar, 6, 1>;
  using Point = Eigen::Matrix<Scalar, 3, 1>;
  using HomogeneousPoint = Eigen::Matrix<Scalar, 4, 1>;
  using Transformation = Eigen::Matrix<Scalar, 4, 4>;
  using Adjoint = Eigen::Matrix<Scalar, 6, 6>;

  SE3() { t_.setZero(); }
  SE3(const SO3<Scalar>& so3, const Point& t) : so3_(so3), t_(t) {}
  SE3(const Eigen::Matrix<Scalar, 3, 3>& R, const Point& t) : so3_(R), t_(t) {}
  SE3(const Eigen::Quaternion<Scalar>& q, const Point& t) : so3_(q), t_(t) {}
  template <class D>
  SE3(const SE3Base<D>& o) : SE3(static_cast<const D&>(o).template cast<Scalar>()) {}
  SE3(const SE3&) = default;
  SE3& operator=(const SE3&) = default;

  static SE3 exp(const Tangent& a) {
    using std::cos;
    using std::sin;
    const Eigen::Matrix<Scalar, 3, 1> upsilon = a.template head<3>();
    const Eigen::Matrix<Scalar, 3, 1> omega = a.template tail<3>();
    Scalar theta;
    const SO3<Scalar> so3 = SO3<Scalar>::expAndTheta(omega, &theta);
    const Eigen::Matrix<Scalar, 3, 3> Omega = SO3<Scalar>::hat(omega);
    const Eigen::Matrix<Scalar, 3, 3> Omega_sq = Omega * Omega;
    Eigen::Matrix<Scalar, 3, 3> V;
    if (theta < Constants<Scalar>::epsilon()) {
      V = so3.matrix();
    } else {
      const Scalar theta_sq = theta * theta;
      V = Eigen::Matrix<Scalar, 3, 3>::Identity() + (Scalar(1) - cos(theta)) / theta_sq * Omega +
          (theta - sin(theta)) / (theta_sq * theta) * Omega_sq;
    }
    return SE3(so3, Point(V * upsilon));
  }

  SO3<Scalar>& so3() { return so3_; }
  const SO3<Scalar>& so3() const { return so3_; }
  Point& translation() { return t_; }
  const Point& translation() const { return t_; }
  const Eigen::Quaternion<Scalar>& unit_quaternion() const { return so3_.unit_quaternion(); }
  Eigen::Matrix<Scalar, 3, 3> rotationMatrix() const { return so3_.matrix(); }
  Transformation matrix() const {
    Transformation m = Transformation::Identity();
    m.template block<3, 3>(0, 0) = so3_.matrix();
    m.template block<3, 1>(0, 3) = t_;
    return m;
  }
  Eigen::Matrix<Scalar, 3, 4> matrix3x4() const {
    Eigen::Matrix<Scalar, 3, 4> m;
    m.template block<3, 3>(0, 0) = so3_.matrix();
    m.template block<3, 1>(0, 3) = t_;
    return m;
  }
  Adjoint Adj() const {
    const Eigen::Matrix<Scalar, 3, 3> R = so3_.matrix();
    Adjoint res;
    res.template block<3, 3>(0, 0) = R;
    res.template block<3, 3>(3, 3) = R;
    res.template block<3, 3>(0, 3) = SO3<Scalar>::hat(t_) * R;
    res.template block<3, 3>(3, 0) = Eigen::Matrix<Scalar, 3, 3>::Zero();
    return res;
  }
  SE3 inverse() const {
    const SO3<Scalar> inv = so3_.inverse();
    return SE3(inv, Point(inv * (t_ * Scalar(-1))));
  }
  SE3 operator*(const SE3& o) const { return SE3(so3_ * o.so3_, Point(t_ + so3_ * o.t_)); }
  template <class D>
    requires(Eigen::MatrixBase<D>::RowsAtCompileTime == 3)
  Point operator*(const Eigen::MatrixBase<D>& p) const {
    return Point(so3_ * p + t_);
  }
  template <class D>
    requires(Eigen::MatrixBase<D>::RowsAtCompileTime == 4)
  HomogeneousPoint operator*(const Eigen::MatrixBase<D>& p) const {
    const Point p3 = p.template head<3>();
    const Point tp = so3_ * p3 + p(3) * t_;
    return HomogeneousPoint(tp(0), tp(1), tp(2), p(3));
  }
  template <class U>
  SE3<U> cast() const {
    return SE3<U>(so3_.template cast<U>(), t_.template cast<U>());
  }
  Scalar* data() { return so3_.data(); }

 private:
  SO3<Scalar> so3_;
  Point t_;
};

using SE3d = SE3<double>;
using SE3f = SE3<float>;
using SO3d = SO3<double>;

}  // namespace Sophus
