// Entry points that exercise the THIRD-PARTY STAND-INS themselves (oracle/ref_stubs_full/), so that tests can hold them against
// independent implementations (NumPy / SciPy): the pin of the oracle against the reference's code rests on these stand-ins
// computing what Eigen and Sophus are documented to compute.  Test infrastructure only.
#include <Eigen/Dense>
#include <sophus/se3.hpp>

extern "C" {

// A.ldlt().solve(b) for a symmetric n x n matrix (row-major)
void refstub_ldlt_solve(int n, const double* A, const double* b, double* x) {
  Eigen::MatrixX<double> M(n, n);
  Eigen::VectorX<double> v(n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) M(i, j) = A[i * n + j];
    v(i) = b[i];
  }
  const Eigen::VectorX<double> r = M.ldlt().solve(v);
  for (int i = 0; i < n; ++i) x[i] = r(i);
}

// A.completeOrthogonalDecomposition().pseudoInverse() for an m x n matrix (row-major in, n x m row-major out)
void refstub_pseudo_inverse(int m, int n, const double* A, double* P) {
  Eigen::MatrixX<double> M(m, n);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) M(i, j) = A[i * n + j];
  const Eigen::MatrixX<double> R = M.completeOrthogonalDecomposition().pseudoInverse();
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) P[i * m + j] = R(i, j);
}

// Sophus::SE3d::exp(xi).matrix3x4(), .Adj(), .inverse() and the group product exp(a) * exp(b)
void refstub_se3(const double* xi, const double* xi2, double* T34, double* Adj66, double* Tinv34, double* prod34) {
  Eigen::Matrix<double, 6, 1> a, b;
  for (int i = 0; i < 6; ++i) a(i) = xi[i], b(i) = xi2[i];
  const Sophus::SE3d A = Sophus::SE3d::exp(a), B = Sophus::SE3d::exp(b);
  const auto m = A.matrix3x4(), mi = A.inverse().matrix3x4(), mp = (A * B).matrix3x4();
  const auto adj = A.Adj();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) T34[4 * i + j] = m(i, j), Tinv34[4 * i + j] = mi(i, j), prod34[4 * i + j] = mp(i, j);
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) Adj66[6 * i + j] = adj(i, j);
}

// a few of the eager mini-Eigen's block / array / colwise semantics in one expression each; out[0..]:
//   0..5   (M.block<2,3>(1,0) * 2 - M.topRows<2>()) row-major          (M = 3x3 row-major input)
//   6..8   M.colwise().hnormalized() second column then ... see the test
void refstub_eigen_semantics(const double* m9, double* out) {
  Eigen::Matrix<double, 3, 3> M;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) M(i, j) = m9[3 * i + j];
  const Eigen::Matrix<double, 2, 3> a = M.block<2, 3>(1, 0) * 2.0 - M.topRows<2>();
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) out[3 * i + j] = a(i, j);
  const Eigen::Matrix<double, 2, 3> h = M.colwise().hnormalized();
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) out[6 + 3 * i + j] = h(i, j);
  Eigen::Matrix<double, 3, 3> C = M;
  C.colwise() += M.col(2) + M.topRightCorner<3, 1>() * 0.5;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[12 + 3 * i + j] = C(i, j);
  out[21] = (M.row(2).array() > 0.0).all() ? 1.0 : 0.0;
  out[22] = M.transpose().lazyProduct(M).trace();
  out[23] = M.selfadjointView<Eigen::Lower>()(0, 2);
  Eigen::Matrix<double, 3, 1> col;
  col = M.row(1);  // row vector assigned to a column vector (Eigen transposes vectors on assignment)
  out[24] = col(2);
  out[25] = (M.col(0).cwiseProduct(M.col(1))).sum();
  out[26] = M.diagonal().cwiseSqrt().cwiseInverse()(1);
}

}  // extern "C"
