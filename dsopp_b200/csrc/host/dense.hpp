// Minimal dense linear algebra for the 8N x 8N host-side systems (N <= 16), in place of the Eigen calls of the
// reference (Eigen is not available here): ldlt().solve, JacobiSVD / completeOrthogonalDecomposition().pseudoInverse
// of SYMMETRIC matrices (src/energy/problems/src/normal_linear_system.cpp:18-59,
// src/energy/problems/src/eigen_photometric_bundle_adjustment.cpp:31-45).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace dsopp_b200::dense {

using Vec = std::vector<double>;

struct Mat {
  int rows = 0, cols = 0;
  std::vector<double> a;
  Mat() = default;
  Mat(int r, int c) : rows(r), cols(c), a((size_t)r * c, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * cols + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * cols + j]; }
  void setZero() { std::fill(a.begin(), a.end(), 0.0); }
};

inline Vec matvec(const Mat& A, const Vec& x) {
  Vec y(A.rows, 0.0);
  for (int i = 0; i < A.rows; ++i) {
    double s = 0;
    for (int j = 0; j < A.cols; ++j) s += A(i, j) * x[j];
    y[i] = s;
  }
  return y;
}
inline double dot(const Vec& a, const Vec& b) {
  double s = 0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}
inline Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.rows, B.cols);
  for (int i = 0; i < A.rows; ++i)
    for (int k = 0; k < A.cols; ++k) {
      const double a = A(i, k);
      if (a == 0) continue;
      for (int j = 0; j < B.cols; ++j) C(i, j) += a * B(k, j);
    }
  return C;
}
inline Mat transpose(const Mat& A) {
  Mat T(A.cols, A.rows);
  for (int i = 0; i < A.rows; ++i)
    for (int j = 0; j < A.cols; ++j) T(j, i) = A(i, j);
  return T;
}

// LDL^T with symmetric (diagonal) pivoting -- what Eigen::LDLT does -- then solve.
inline Vec ldlt_solve(Mat A, const Vec& b) {
  const int n = A.rows;
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A(i, i)) > std::fabs(A(piv, piv))) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(A(k, j), A(piv, j));
      for (int i = 0; i < n; ++i) std::swap(A(i, k), A(i, piv));
      std::swap(perm[k], perm[piv]);
    }
    const double d = A(k, k);
    if (d == 0) continue;
    for (int i = k + 1; i < n; ++i) A(i, k) /= d;
    for (int i = k + 1; i < n; ++i) {
      const double lik = A(i, k);
      for (int j = k + 1; j <= i; ++j) A(i, j) -= lik * d * A(j, k);
    }
    for (int i = k + 1; i < n; ++i)
      for (int j = i + 1; j < n; ++j) A(i, j) = A(j, i);
  }
  Vec z(n), x(n);
  for (int i = 0; i < n; ++i) z[i] = b[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) z[i] -= A(i, j) * z[j];
  for (int i = 0; i < n; ++i) z[i] = A(i, i) != 0 ? z[i] / A(i, i) : 0.0;
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) z[i] -= A(j, i) * z[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = z[i];
  return x;
}

// cyclic Jacobi eigen-decomposition of a symmetric matrix: A = V diag(w) V^T
inline void sym_eig(Mat A, Vec& w, Mat& V) {
  const int n = A.rows;
  V = Mat(n, n);
  for (int i = 0; i < n; ++i) V(i, i) = 1.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; ++i) {
      diag += A(i, i) * A(i, i);
      for (int j = i + 1; j < n; ++j) off += A(i, j) * A(i, j);
    }
    if (off <= 1e-30 * (diag + off) || off == 0) break;
    for (int p = 0; p < n; ++p)
      for (int q = p + 1; q < n; ++q) {
        const double apq = A(p, q);
        if (apq == 0) continue;
        const double theta = (A(q, q) - A(p, p)) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < n; ++k) {
          const double akp = A(k, p), akq = A(k, q);
          A(k, p) = c * akp - s * akq;
          A(k, q) = s * akp + c * akq;
        }
        for (int k = 0; k < n; ++k) {
          const double apk = A(p, k), aqk = A(q, k);
          A(p, k) = c * apk - s * aqk;
          A(q, k) = s * apk + c * aqk;
        }
        for (int k = 0; k < n; ++k) {
          const double vkp = V(k, p), vkq = V(k, q);
          V(k, p) = c * vkp - s * vkq;
          V(k, q) = s * vkp + c * vkq;
        }
      }
  }
  w.resize(n);
  for (int i = 0; i < n; ++i) w[i] = A(i, i);
}

// pseudo-inverse of a symmetric matrix.  n_null >= 0: drop exactly the n_null smallest singular values
// (pseudoInverse(origin, number_of_nullspaces), eigen_photometric_bundle_adjustment.cpp:31-45); n_null < 0: rank by
// threshold, as completeOrthogonalDecomposition().pseudoInverse() does.
inline Mat sym_pinv(const Mat& A, int n_null) {
  const int n = A.rows;
  Vec w;
  Mat V;
  sym_eig(A, w, V);
  std::vector<int> order(n);
  for (int i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int x, int y) { return std::fabs(w[x]) > std::fabs(w[y]); });
  const double wmax = n ? std::fabs(w[order[0]]) : 0.0;
  Mat P(n, n);
  for (int r = 0; r < n; ++r) {
    const int k = order[r];
    const bool keep = n_null >= 0 ? r < n - n_null : std::fabs(w[k]) > wmax * n * 2.220446049250313e-16;
    if (!keep || w[k] == 0) continue;
    const double iw = 1.0 / w[k];
    for (int i = 0; i < n; ++i) {
      const double vi = V(i, k) * iw;
      for (int j = 0; j < n; ++j) P(i, j) += vi * V(j, k);
    }
  }
  return P;
}

}  // namespace dsopp_b200::dense
