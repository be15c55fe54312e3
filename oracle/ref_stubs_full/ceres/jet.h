// STAND-IN -- this is NOT Ceres.  Test infrastructure only: the names the reference's headers mention on paths that the
// Eigen-backend bundle adjustment never instantiates (CeresGrid, the Jet overload of PixelMap::Evaluate, Ceres priors).
#pragma once
#include <Eigen/Dense>
#include <cmath>
namespace ceres {
using std::abs;
using std::cos;
using std::exp;
using std::pow;
using std::sin;
using std::sqrt;
template <class T, int N>
struct Jet {
  T a;
  Eigen::Matrix<T, N, 1> v;
};
template <class T, int C>
class Grid2D {
 public:
  enum { DATA_DIMENSION = C };
  Grid2D(const T*, int, int, int, int) {}
};
template <class Grid>
class BiCubicInterpolator {
 public:
  explicit BiCubicInterpolator(const Grid&) {}
  template <class... A>
  void Evaluate(A&&...) const {}
};
class CostFunction {
 public:
  virtual ~CostFunction() = default;
};
template <class F, int... N>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(F*) {}
};
class NormalPrior : public CostFunction {};
class Problem {};
}  // namespace ceres
