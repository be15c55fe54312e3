// NormalLinearSystem<double, Dynamic> of the reference (src/energy/problems/include/energy/normal_linear_system.hpp:15-145,
// src/energy/problems/src/normal_linear_system.cpp:10-59), same member names and semantics, on dense.hpp.
#pragma once
#include <numeric>
#include <vector>

#include "dense.hpp"

namespace dsopp_b200 {

struct NormalLinearSystem {
  dense::Mat H;
  dense::Vec b;

  explicit NormalLinearSystem(int size = 0) : H(size, size), b(size, 0.0) {}
  int size() const { return (int)b.size(); }
  void setZero() {
    H.setZero();
    std::fill(b.begin(), b.end(), 0.0);
  }
  // conservativeResize + zero fill of the new rows/cols (eigen_photometric_bundle_adjustment.cpp:134-140)
  void resize(int n) {
    dense::Mat Hn(n, n);
    dense::Vec bn(n, 0.0);
    const int k = std::min(n, size());
    for (int i = 0; i < k; ++i) {
      bn[i] = b[i];
      for (int j = 0; j < k; ++j) Hn(i, j) = H(i, j);
    }
    H = Hn;
    b = bn;
  }
  NormalLinearSystem& operator+=(const NormalLinearSystem& o) {
    for (size_t i = 0; i < H.a.size(); ++i) H.a[i] += o.H.a[i];
    for (size_t i = 0; i < b.size(); ++i) b[i] += o.b[i];
    return *this;
  }
  NormalLinearSystem operator+(const NormalLinearSystem& o) const {
    NormalLinearSystem r = *this;
    r += o;
    return r;
  }
  NormalLinearSystem operator-(const NormalLinearSystem& o) const {
    NormalLinearSystem r = *this;
    for (size_t i = 0; i < H.a.size(); ++i) r.H.a[i] -= o.H.a[i];
    for (size_t i = 0; i < b.size(); ++i) r.b[i] -= o.b[i];
    return r;
  }
  NormalLinearSystem operator*(double s) const {
    NormalLinearSystem r = *this;
    for (auto& v : r.H.a) v *= s;
    for (auto& v : r.b) v *= s;
    return r;
  }

  static dense::Vec jacobiPreconditioner(const dense::Mat& H) {  // normal_linear_system.cpp:10-16
    const double kPreconditionerMinValue = 10;
    dense::Vec p(H.rows);
    for (int i = 0; i < H.rows; ++i) p[i] = 1.0 / std::sqrt(H(i, i) + kPreconditionerMinValue);
    return p;
  }

  dense::Vec solve() const {  // normal_linear_system.cpp:51-59
    const int n = size();
    const dense::Vec p = jacobiPreconditioner(H);
    dense::Mat Hp(n, n);
    dense::Vec bp(n);
    for (int i = 0; i < n; ++i) {
      bp[i] = p[i] * b[i];
      for (int j = 0; j < n; ++j) Hp(i, j) = p[i] * H(i, j) * p[j];
    }
    dense::Vec x = dense::ldlt_solve(Hp, bp);
    for (int i = 0; i < n; ++i) x[i] *= p[i];
    return x;
  }

  void reduce_system(const std::vector<int>& elim) {  // normal_linear_system.cpp:18-50
    const int n = size();
    std::vector<char> is_elim(n, 0);
    for (int i : elim) is_elim[i] = 1;
    std::vector<int> keep;
    for (int i = 0; i < n; ++i)
      if (!is_elim[i]) keep.push_back(i);
    const int nk = (int)keep.size(), ne = (int)elim.size();
    const dense::Vec p = jacobiPreconditioner(H);
    dense::Mat Hp(n, n);
    dense::Vec bp(n);
    for (int i = 0; i < n; ++i) {
      bp[i] = p[i] * b[i];
      for (int j = 0; j < n; ++j) Hp(i, j) = p[i] * H(i, j) * p[j];
    }
    dense::Mat Hee(ne, ne), Hke(nk, ne);
    for (int i = 0; i < ne; ++i)
      for (int j = 0; j < ne; ++j) Hee(i, j) = Hp(elim[i], elim[j]);
    for (int i = 0; i < nk; ++i)
      for (int j = 0; j < ne; ++j) Hke(i, j) = Hp(keep[i], elim[j]);
    const dense::Mat St = dense::matmul(Hke, dense::sym_pinv(Hee, -1));
    const dense::Mat SH = dense::matmul(St, dense::transpose(Hke));
    dense::Mat Hn(nk, nk);
    dense::Vec bn(nk);
    for (int i = 0; i < nk; ++i) {
      double s = bp[keep[i]];
      for (int j = 0; j < ne; ++j) s -= St(i, j) * bp[elim[j]];
      bn[i] = s;
      for (int j = 0; j < nk; ++j) Hn(i, j) = Hp(keep[i], keep[j]) - SH(i, j);
    }
    H = dense::Mat(nk, nk);
    b.assign(nk, 0.0);
    for (int i = 0; i < nk; ++i) {
      b[i] = bn[i] / p[keep[i]];
      for (int j = 0; j < nk; ++j) H(i, j) = 0.5 * (Hn(i, j) + Hn(j, i)) / (p[keep[i]] * p[keep[j]]);
    }
  }
};

}  // namespace dsopp_b200
