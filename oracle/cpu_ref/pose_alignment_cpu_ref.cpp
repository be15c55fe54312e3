// CPU restatement (plain C++17, double, SERIAL like the reference) of DSOPP's coarse-tracker direct image alignment.
// TEST INFRASTRUCTURE / TIMED CPU BASELINE ONLY (see oracle/pose_alignment_oracle.py for the NumPy twin and the
// statement of how it is pinned against the reference's own PoseAlignerProblem).  Follows, paths relative to /root/reference/src/:
//   PoseAlignerProblem                energy/problems/src/eigen_pose_alignment.cpp:28-241
//   EigenPoseAlignment::solve         energy/problems/src/eigen_pose_alignment.cpp:275-329
//   levenberg_marquardt_algorithm     energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:77-128
//   NormalLinearSystem::solve         energy/problems/src/normal_linear_system.cpp:10-59
// Dataflow as in the reference: calculateEnergy() caches the target samples and gradients per landmark, linearize()
// re-uses them (two passes over the landmarks per accepted iteration).
#include <cmath>
#include <cstring>
#include <vector>

namespace {
struct Frame {
  double T[12], exposure, ab0[2], intr[4];
  int W, H;
};
void se3_exp(const double* xi, double* R, double* t) {
  const double* v = xi;
  const double* w = xi + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
  double a, b, c;
  if (th < 1e-10) a = 1, b = 0.5, c = 1.0 / 6.0;
  else a = std::sin(th) / th, b = (1 - std::cos(th)) / th2, c = (th - std::sin(th)) / (th2 * th);
  const double Wm[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += Wm[i * 3 + k] * Wm[k * 3 + j];
      W2[i * 3 + j] = s;
    }
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + a * Wm[i] + b * W2[i];
    V[i] = I + b * Wm[i] + c * W2[i];
  }
  for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
}
void left_increment(const double* xi, double* T) {
  double R[9], t[3], out[12];
  se3_exp(xi, R, t);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = (j == 3) ? t[i] : 0.0;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * T[k * 4 + j];
      out[i * 4 + j] = s;
    }
  std::memcpy(T, out, sizeof(out));
}
// x = p * solve(pHp, p b), p = 1 / sqrt(diag + 10)  (dense Gaussian elimination on the SPD 8x8 system)
void normal_solve8(const double* H, const double* b, double* x) {
  double A[64], y[8], p[8];
  for (int i = 0; i < 8; ++i) p[i] = 1.0 / std::sqrt(H[i * 8 + i] + 10.0);
  for (int i = 0; i < 8; ++i) {
    for (int j = 0; j < 8; ++j) A[i * 8 + j] = H[i * 8 + j] * p[i] * p[j];
    y[i] = b[i] * p[i];
  }
  for (int k = 0; k < 8; ++k) {
    for (int i = k + 1; i < 8; ++i) {
      const double f = A[i * 8 + k] / A[k * 8 + k];
      for (int j = k; j < 8; ++j) A[i * 8 + j] -= f * A[k * 8 + j];
      y[i] -= f * y[k];
    }
  }
  for (int i = 7; i >= 0; --i) {
    double s = y[i];
    for (int j = i + 1; j < 8; ++j) s -= A[i * 8 + j] * y[j];
    y[i] = s / A[i * 8 + i];
  }
  for (int i = 0; i < 8; ++i) x[i] = y[i] * p[i];
}

struct Problem {
  Frame ref, tgt;
  const float* img;  // target {I,dx,dy} interleaved
  const unsigned char* mask;
  int n;
  const double *xy, *idepth, *patch;
  double sigma, ab_reg[2];
  double T[12], ab_eps[2], old_T[12], old_ab[2];
  double Hs[64], bs[8], step[8];
  std::vector<unsigned char> ok;
  std::vector<double> tI, dIu, dIv;

  void consts(double* A, double* M, double& s, double* ab_t) const {
    const double fx = ref.intr[0], fy = ref.intr[1], cx = ref.intr[2], cy = ref.intr[3];
    for (int i = 0; i < 3; ++i) {
      M[i * 4 + 0] = T[i * 4 + 0] / fx;
      M[i * 4 + 1] = T[i * 4 + 1] / fy;
      M[i * 4 + 2] = T[i * 4 + 0] * (-cx / fx) + T[i * 4 + 1] * (-cy / fy) + T[i * 4 + 2];
      M[i * 4 + 3] = T[i * 4 + 3];
    }
    for (int j = 0; j < 4; ++j) {
      A[0 + j] = tgt.intr[0] * M[0 + j] + tgt.intr[2] * M[8 + j];
      A[4 + j] = tgt.intr[1] * M[4 + j] + tgt.intr[3] * M[8 + j];
      A[8 + j] = M[8 + j];
    }
    ab_t[0] = tgt.ab0[0] + ab_eps[0];
    ab_t[1] = tgt.ab0[1] + ab_eps[1];
    s = (tgt.exposure / ref.exposure) * std::exp(ab_t[0] - ref.ab0[0]);
  }
  double calculate_energy(int& n_valid) {  // :55-108
    double A[12], M[12], s, ab_t[2];
    consts(A, M, s, ab_t);
    double energy = 0;
    n_valid = 0;
    const int W = tgt.W;
    for (int i = 0; i < n; ++i) {
      const double x = xy[2 * i], y = xy[2 * i + 1], rho = idepth[i];
      bool good = rho > -1e-4 && rho < 1010.0 && x >= 4 && y >= 4 && x <= ref.W - 5 && y <= ref.H - 5;
      const double X = A[0] * x + A[1] * y + (A[2] + A[3] * rho), Y = A[4] * x + A[5] * y + (A[6] + A[7] * rho),
                   Z = A[8] * x + A[9] * y + (A[10] + A[11] * rho);
      good = good && Z > 0;
      const double tu = X / Z, tv = Y / Z;
      good = good && tu >= 4 && tv >= 4 && tu <= tgt.W - 5 && tv <= tgt.H - 5;
      if (good && mask) good = mask[(int)std::lround(tv) * W + (int)std::lround(tu)] != 0;
      ok[i] = good;
      if (!good) continue;
      const int ix = (int)tu, iy = (int)tv;
      const double dx = tu - ix, dy = tv - iy, dxdy = dx * dy;
      const double w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = 1 - dx - dy + dxdy;
      const float* p00 = img + 3 * ((size_t)iy * W + ix);
      const float* p01 = p00 + 3;
      const float* p10 = p00 + 3 * (size_t)W;
      const float* p11 = p10 + 3;
      tI[i] = w11 * p11[0] + w10 * p10[0] + w01 * p01[0] + w00 * p00[0];
      dIu[i] = w11 * p11[1] + w10 * p10[1] + w01 * p01[1] + w00 * p00[1];
      dIv[i] = w11 * p11[2] + w10 * p10[2] + w01 * p01[2] + w00 * p00[2];
      const double r = (tI[i] - ab_t[1]) - s * (patch[i] - ref.ab0[1]);
      const double nrm = std::fabs(r);
      energy += nrm * nrm > sigma * sigma ? sigma * nrm - sigma * sigma / 2 : nrm * nrm / 2;
      ++n_valid;
    }
    energy += 0.5 * (ab_t[0] * ab_reg[0] * ab_t[0] + ab_t[1] * ab_reg[1] * ab_t[1]);
    return energy;
  }
  void linearize() {  // :110-192
    double A[12], M[12], s, ab_t[2];
    consts(A, M, s, ab_t);
    std::memset(Hs, 0, sizeof(Hs));
    std::memset(bs, 0, sizeof(bs));
    const double fx = tgt.intr[0], fy = tgt.intr[1];
    for (int i = 0; i < n; ++i) {
      if (!ok[i]) continue;
      const double x = xy[2 * i], y = xy[2 * i + 1], rho = idepth[i];
      const double qx = M[0] * x + M[1] * y + (M[2] + M[3] * rho), qy = M[4] * x + M[5] * y + (M[6] + M[7] * rho),
                   qz = M[8] * x + M[9] * y + (M[10] + M[11] * rho);
      const double sI = 1.0 / qz, b0 = qx * sI, b1 = qy * sI, nid = rho * sI;
      const double right = s * (patch[i] - ref.ab0[1]);
      const double r = (tI[i] - ab_t[1]) - right;
      const double w = r * r > sigma * sigma ? sigma / std::fabs(r) : 1.0;
      const double gu = dIu[i] * fx, gv = dIv[i] * fy;
      const double d[8] = {-(gu * nid),
                           -(gv * nid),
                           gu * nid * b0 + gv * nid * b1,
                           gu * b0 * b1 + gv * (b1 * b1 + 1),
                           -(gu * (b0 * b0 + 1) + gv * b0 * b1),
                           gu * b1 - gv * b0,
                           -right,
                           -1.0};
      for (int a = 0; a < 8; ++a) {
        for (int b = 0; b < 8; ++b) Hs[a * 8 + b] += w * d[a] * d[b];
        bs[a] += w * d[a] * r;
      }
    }
    for (int k = 0; k < 2; ++k) {
      Hs[(6 + k) * 8 + 6 + k] += ab_reg[k];
      bs[6 + k] += ab_reg[k] * ab_t[k];
    }
  }
  void calculate_step(double lam) {  // :194-206
    double H[64];
    std::memcpy(H, Hs, sizeof(H));
    for (int i = 0; i < 8; ++i) H[i * 8 + i] += Hs[i * 8 + i] * lam;
    normal_solve8(H, bs, step);
    std::memcpy(old_T, T, sizeof(T));
    old_ab[0] = ab_eps[0];
    old_ab[1] = ab_eps[1];
    left_increment(step, T);
    ab_eps[0] -= step[6];
    ab_eps[1] -= step[7];
  }
};
}  // namespace

extern "C" {
// Returns the energy; out: T_t_r[12], ab_eps[2], H[64], n_valid, iterations, converged.
double paref_solve(const double* ref_T, double ref_exposure, const double* ref_ab, const double* ref_intr, int ref_W, int ref_H,
                   const double* tgt_T, double tgt_exposure, const double* tgt_ab, const double* tgt_intr, int W, int H,
                   const float* tgt_image, const unsigned char* tgt_mask, int n, const double* xy, const double* idepth,
                   const double* patch, double sigma, const double* ab_reg, int max_it, double lambda0, double ftol,
                   double ptol, double dec, double inc, double* T_out, double* ab_out, double* H_out, int* n_valid_out,
                   int* iterations_out, int* converged_out) {
  Problem p;
  std::memcpy(p.ref.T, ref_T, sizeof(p.ref.T));
  std::memcpy(p.tgt.T, tgt_T, sizeof(p.tgt.T));
  p.ref.exposure = ref_exposure, p.tgt.exposure = tgt_exposure;
  for (int k = 0; k < 2; ++k) p.ref.ab0[k] = ref_ab[k], p.tgt.ab0[k] = tgt_ab[k], p.ab_reg[k] = ab_reg[k];
  for (int k = 0; k < 4; ++k) p.ref.intr[k] = ref_intr[k], p.tgt.intr[k] = tgt_intr[k];
  p.ref.W = ref_W, p.ref.H = ref_H, p.tgt.W = W, p.tgt.H = H;
  p.img = tgt_image, p.mask = tgt_mask, p.n = n, p.xy = xy, p.idepth = idepth, p.patch = patch, p.sigma = sigma;
  p.ok.assign(n, 0);
  p.tI.assign(n, 0), p.dIu.assign(n, 0), p.dIv.assign(n, 0);
  // t_t_r = T_w_t^-1 T_w_r  (:307-308)
  double Ti[12];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Ti[i * 4 + j] = tgt_T[j * 4 + i];
  for (int i = 0; i < 3; ++i) Ti[i * 4 + 3] = -(Ti[i * 4] * tgt_T[3] + Ti[i * 4 + 1] * tgt_T[7] + Ti[i * 4 + 2] * tgt_T[11]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = (j == 3) ? Ti[i * 4 + 3] : 0.0;
      for (int k = 0; k < 3; ++k) s += Ti[i * 4 + k] * ref_T[k * 4 + j];
      p.T[i * 4 + j] = s;
    }
  p.ab_eps[0] = p.ab_eps[1] = 0;
  // levenberg_marquardt_algorithm::solve
  double lam = lambda0;
  int nvalid = 0;
  double energy = p.calculate_energy(nvalid);
  bool converged = false, system_valid = false;
  int it = 0;
  for (; it < max_it && !converged && nvalid > 0; ++it) {
    if (!system_valid) p.linearize();
    p.calculate_step(lam);
    int n1 = 0;
    const double e1 = p.calculate_energy(n1);
    if (n1 == 0) {
      std::memcpy(p.T, p.old_T, sizeof(p.T));
      p.ab_eps[0] = p.old_ab[0], p.ab_eps[1] = p.old_ab[1];
      break;
    }
    if (std::fabs(energy - e1) / energy < ftol) converged = true;
    if (e1 < energy) {
      const double a0 = p.tgt.ab0[0] + p.old_ab[0], a1 = p.tgt.ab0[1] + p.old_ab[1];
      double sq = 0;
      for (int k = 0; k < 8; ++k) sq += p.step[k] * p.step[k];
      if (sq < ptol * (a0 * a0 + a1 * a1 + ptol)) converged = true;
      energy = e1;
      nvalid = n1;
      lam /= dec;
      system_valid = false;
    } else {
      std::memcpy(p.T, p.old_T, sizeof(p.T));
      p.ab_eps[0] = p.old_ab[0], p.ab_eps[1] = p.old_ab[1];
      lam *= inc;
      system_valid = true;
    }
  }
  int dummy;
  p.calculate_energy(dummy);
  std::memcpy(T_out, p.T, sizeof(p.T));
  ab_out[0] = p.ab_eps[0], ab_out[1] = p.ab_eps[1];
  std::memcpy(H_out, p.Hs, sizeof(p.Hs));
  *n_valid_out = nvalid;
  *iterations_out = it;
  *converged_out = converged;
  return energy;
}
}
