// CPU restatement (plain C++17 + OpenMP, no Eigen) of DSOPP's photometric bundle-adjustment hot path with the
// REFERENCE'S DATAFLOW: an array-of-structs ResidualPoint is materialised by the sweep, then the PosePose and
// Schur passes re-read it, Hessian sums are Kahan-compensated, and those two passes parallelise only over the
// reference frames -- exactly the structure of
//   src/energy/problems/internal/energy/problems/photometric_bundle_adjustment/{evaluate_jacobians,
//   first_estimate_jacobians,hessian_block_evaluation,eigen_photometric_bundle_adjustment_problem}.hpp
// in RoadlyInc/DSOPP @ a4af2aa ("PBA/" below; other paths relative to the reference's src/).
//
// TEST INFRASTRUCTURE: this is (a) the second, independent restatement that the NumPy oracle is cross-checked
// against, (b) the float32 oracle whose ROI / depth / mask predicates use the same operation order as the CUDA
// kernels (build with -ffp-contract=off), and (c) the CPU baseline that bench.py times ("port": the reference
// binary cannot be built here -- Eigen/Sophus/TBB are absent).  PARITY PINNED through the NumPy oracle, which equals the
// reference's own code compiled against stand-in third-party headers (oracle/build_ref_pba.py, tests/test_reference_pba.py)
// and which this file equals element-wise at 1e-9 (tests/test_cpu_ref_vs_numpy.py).
// Nothing under dsopp_b200/ links or includes this file.
//
// Scalar = double is the reference default (cmake/options.cmake:7), float is the CI build.  Per-pair constants
// are computed in double and rounded to Scalar (the float reference build computes them in float).
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <vector>

namespace {

constexpr int P = 8, B = 8, MAXF = 16;
constexpr int K_OK = 0, K_OUTLIER = 1, K_OOB = 3;
const double PAT[8][2] = {{0, 2}, {-1, 1}, {1, 1}, {-2, 0}, {0, 0}, {2, 0}, {-1, -1}, {0, -2}};

// ---- SE3 in double (Sophus closed forms) -------------------------------------------------------
struct SE3 {
  double R[9], t[3];
};
SE3 mul(const SE3& a, const SE3& b) {
  SE3 o;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += a.R[i * 3 + k] * b.R[k * 3 + j];
      o.R[i * 3 + j] = s;
    }
    double s = a.t[i];
    for (int k = 0; k < 3; ++k) s += a.R[i * 3 + k] * b.t[k];
    o.t[i] = s;
  }
  return o;
}
SE3 inv(const SE3& a) {
  SE3 o;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += o.R[i * 3 + k] * a.t[k];
    o.t[i] = -s;
  }
  return o;
}
SE3 expm(const double* xi, double sign) {
  SE3 o;
  const double v[3] = {sign * xi[0], sign * xi[1], sign * xi[2]};
  const double w[3] = {sign * xi[3], sign * xi[4], sign * xi[5]};
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double a, b, c;
  if (th < 1e-10) {
    a = 1, b = 0.5, c = 1.0 / 6.0;
  } else {
    a = sin(th) / th, b = (1 - cos(th)) / th2, c = (th - sin(th)) / (th2 * th);
  }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9], V[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += W[i * 3 + k] * W[k * 3 + j];
      W2[i * 3 + j] = s;
    }
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    o.R[i] = I + a * W[i] + b * W2[i];
    V[i] = I + b * W[i] + c * W2[i];
  }
  for (int i = 0; i < 3; ++i) o.t[i] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
  return o;
}
void adjoint(const SE3& T, double* A) {
  const double th[9] = {0, -T.t[2], T.t[1], T.t[2], 0, -T.t[0], -T.t[1], T.t[0], 0};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += th[i * 3 + k] * T.R[k * 3 + j];
      A[i * 6 + j] = T.R[i * 3 + j];
      A[i * 6 + 3 + j] = s;
      A[(3 + i) * 6 + j] = 0;
      A[(3 + i) * 6 + 3 + j] = T.R[i * 3 + j];
    }
}

// ---- data model (PBA/local_frame.hpp:173-300) ---------------------------------------------------
template <typename S>
struct ResidualPoint {
  uint8_t status = K_OK, cand = K_OK;
  S residuals[P] = {};
  S d_u_idepth[P] = {}, d_v_idepth[P] = {};
  S d_u_t[P * 6] = {}, d_v_t[P * 6] = {};  // [pixel][6]
  bool jac_valid = false, was_estimated = false;
  S Jref[P * B] = {}, Jtgt[P * B] = {};  // row-major 8x8
  S d_idepth[P] = {};
  S huber_weight = 1, energy = 0, bcs = 0;
};

template <typename S>
struct Landmark {
  S proj[2], idepth = 0, idepth_step = 0;
  S patch[P];
  bool is_marg = false, to_marg = false, is_outlier = false, ill = false;
  S ref_pattern[P][2];
  S corrected[P] = {};
  S rel_baseline = 0, inv_hdd = 0, b_d = 0;
  std::vector<S> hpd;
  uint32_t n_inliers = 0;
};

template <typename S>
struct Frame {
  int id = 0;
  SE3 T_lin;
  double exposure = 1, ab0[2] = {0, 0}, intr[4];
  double eps[8] = {}, step[8] = {};
  int W = 0, H = 0;
  std::vector<S> image;  // H*W*3 {I,dx,dy}
  std::vector<uint8_t> mask;
  bool fixed = false, to_marg = false, is_marg = false;
  std::vector<Landmark<S>> lms;
  std::vector<std::vector<ResidualPoint<S>>> res;  // [target frame index]
};

// per ordered pair prologue (PBA/evaluate_jacobians.hpp:36-66; projector/.../camera_reproject.hpp:235-260)
template <typename S>
struct Pair {
  S A[12], M[12], tr[3], M0[12], t0[3], adj[36], adj0[36];
  S s, s0, b_t, b_r, b_r0, fx, fy, cx, cy;
};

void make_proj(const SE3& T, const double* ir, const double* it, double* M, double* A) {
  const double fx = ir[0], fy = ir[1], cx = ir[2], cy = ir[3];
  for (int i = 0; i < 3; ++i) {
    M[i * 4 + 0] = T.R[i * 3 + 0] * (1.0 / fx);
    M[i * 4 + 1] = T.R[i * 3 + 1] * (1.0 / fy);
    M[i * 4 + 2] = T.R[i * 3 + 0] * (-cx / fx) + T.R[i * 3 + 1] * (-cy / fy) + T.R[i * 3 + 2];
    M[i * 4 + 3] = T.t[i];
  }
  for (int j = 0; j < 4; ++j) {
    A[0 + j] = it[0] * M[0 + j] + it[2] * M[8 + j];
    A[4 + j] = it[1] * M[4 + j] + it[3] * M[8 + j];
    A[8 + j] = M[8 + j];
  }
}

template <typename S>
Pair<S> make_pair(const Frame<S>& R, const Frame<S>& T) {
  Pair<S> p;
  const SE3 T0 = mul(inv(T.T_lin), R.T_lin);
  double er[6], et[6];
  for (int k = 0; k < 6; ++k) {
    er[k] = R.eps[k] + R.step[k];
    et[k] = T.eps[k] + T.step[k];
  }
  const SE3 Tc = mul(expm(et, -1.0), mul(T0, expm(er, 1.0)));
  double M[12], A[12], M0[12], A0[12], a[36], a0[36];
  make_proj(Tc, R.intr, T.intr, M, A);
  make_proj(T0, R.intr, T.intr, M0, A0);
  adjoint(Tc, a);
  adjoint(T0, a0);
  for (int i = 0; i < 12; ++i) {
    p.A[i] = (S)A[i];
    p.M[i] = (S)M[i];
    p.M0[i] = (S)M0[i];
  }
  for (int i = 0; i < 3; ++i) {
    p.tr[i] = (S)Tc.t[i];
    p.t0[i] = (S)T0.t[i];
  }
  for (int i = 0; i < 36; ++i) {
    p.adj[i] = (S)a[i];
    p.adj0[i] = (S)a0[i];
  }
  const double a_r = R.ab0[0] + R.eps[6] + R.step[6], b_r = R.ab0[1] + R.eps[7] + R.step[7];
  const double a_t = T.ab0[0] + T.eps[6] + T.step[6], b_t = T.ab0[1] + T.eps[7] + T.step[7];
  p.s = (S)((T.exposure / R.exposure) * exp(a_t - a_r));
  p.s0 = (S)((T.exposure / R.exposure) * exp(T.ab0[0] - R.ab0[0]));
  p.b_t = (S)b_t;
  p.b_r = (S)b_r;
  p.b_r0 = (S)R.ab0[1];
  p.fx = (S)T.intr[0];
  p.fy = (S)T.intr[1];
  p.cx = (S)T.intr[2];
  p.cy = (S)T.intr[3];
  return p;
}

template <typename S>
inline bool in_roi(S x, S y, S xmax, S ymax) {
  return x >= S(4) && y >= S(4) && x <= xmax && y <= ymax;
}
template <typename S>
inline bool valid_idepth(S r) {
  return r > S(-1e-4) && r < S(1010.0);
}
// (a0 u + a1 v) + (a2 + a3 rho): camera_reproject.hpp:283-284 -- same association as the CUDA kernels
template <typename S>
inline S row_apply(const S* a, S u, S v, S rho) {
  return (a[0] * u + a[1] * v) + (a[2] + a[3] * rho);
}

// values-only reprojection, camera_reproject.hpp:270-293
template <typename S>
bool reproject_values(const Pair<S>& pc, const Landmark<S>& lm, S rho, int W, int H, S tp[P][2]) {
  const S xmax = S(W - 5), ymax = S(H - 5);
  bool ok = valid_idepth(rho);
  for (int i = 0; i < P; ++i) ok = ok && in_roi(lm.ref_pattern[i][0], lm.ref_pattern[i][1], xmax, ymax);
  bool zpos = true, roi = true;
  for (int i = 0; i < P; ++i) {
    const S u = lm.ref_pattern[i][0], v = lm.ref_pattern[i][1];
    const S X = row_apply(pc.A + 0, u, v, rho), Y = row_apply(pc.A + 4, u, v, rho), Z = row_apply(pc.A + 8, u, v, rho);
    zpos = zpos && (Z > S(0));
    const S rz = S(1) / Z;  // hnormalized() as reciprocal + products: the operation order of the CUDA kernels
    tp[i][0] = X * rz;
    tp[i][1] = Y * rz;
    roi = roi && in_roi(tp[i][0], tp[i][1], xmax, ymax);
  }
  return ok && zpos && roi;
}

// reprojection with Jacobians, camera_reproject.hpp:305-367
template <typename S>
bool reproject_jac(const S* M, const S* t, const Pair<S>& pc, const Landmark<S>& lm, S rho, int W, int H, S tp[P][2],
                   ResidualPoint<S>& r) {
  const S xmax = S(W - 5), ymax = S(H - 5);
  bool ok = valid_idepth(rho);
  for (int i = 0; i < P; ++i) ok = ok && in_roi(lm.ref_pattern[i][0], lm.ref_pattern[i][1], xmax, ymax);
  bool zpos = true, roi = true;
  for (int i = 0; i < P; ++i) {
    const S u = lm.ref_pattern[i][0], v = lm.ref_pattern[i][1];
    const S qx = row_apply(M + 0, u, v, rho), qy = row_apply(M + 4, u, v, rho), qz = row_apply(M + 8, u, v, rho);
    zpos = zpos && (qz > S(0));
    const S rz = S(1) / qz;
    tp[i][0] = (pc.fx * qx + pc.cx * qz) * rz;
    tp[i][1] = (pc.fy * qy + pc.cy * qz) * rz;
    roi = roi && in_roi(tp[i][0], tp[i][1], xmax, ymax);
    const S sI = S(1) / qz, b0 = qx * sI, b1 = qy * sI, nid = rho * sI;
    r.d_u_idepth[i] = pc.fx * (t[0] * sI - t[2] * sI * b0);
    r.d_v_idepth[i] = pc.fy * (t[1] * sI - t[2] * sI * b1);
    S* du = r.d_u_t + i * 6;
    S* dv = r.d_v_t + i * 6;
    dv[0] = 0, dv[1] = pc.fy * nid, dv[2] = pc.fy * (-nid * b1), dv[3] = pc.fy * (-(b1 * b1 + 1)), dv[4] = pc.fy * (b0 * b1),
    dv[5] = pc.fy * b0;
    du[0] = pc.fx * nid, du[1] = 0, du[2] = pc.fx * (-nid * b0), du[3] = pc.fx * (-(b0 * b1)), du[4] = pc.fx * (b0 * b0 + 1),
    du[5] = pc.fx * (-b1);
  }
  return ok && zpos && roi;
}

template <typename S>
inline void sample(const Frame<S>& f, S x, S y, S out[3]) {  // features/.../pixel_map.hpp:20-40
  const int ix = (int)x, iy = (int)y;
  const S dx = x - S(ix), dy = y - S(iy), dxdy = dx * dy;
  const S w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = S(1) - dx - dy + dxdy;
  const S* p = f.image.data() + ((size_t)iy * f.W + ix) * 3;
  const S* q = p + (size_t)f.W * 3;
  for (int c = 0; c < 3; ++c) out[c] = w11 * q[3 + c] + w10 * q[c] + w01 * p[3 + c] + w00 * p[c];
}

// Intensity sample with the CUDA kernels' exact operation sequence (eval_pixel in dsopp_b200/csrc/pba_kernels.cu: products
// of the weights rounded separately, the four taps folded by a chain of fused multiply-adds).  Only the `device_ops`
// flavour of the float build uses it, so that residual energies -- and with them the 75 % quantile threshold of
// updatePointStatuses -- can be compared with the device bit for bit.
inline float sample_intensity_device_ops(const Frame<float>& f, float x, float y) {
  const int ix = (int)x, iy = (int)y;
  const float dx = x - (float)ix, dy = y - (float)iy, dxdy = dx * dy;
  const float w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = ((1.f - dx) - dy) + dxdy;
  const float* p = f.image.data() + ((size_t)iy * f.W + ix) * 3;
  const float* q = p + (size_t)f.W * 3;
  return fmaf(w00, p[0], fmaf(w01, p[3], fmaf(w10, q[0], w11 * q[3])));
}
inline double sample_intensity_device_ops(const Frame<double>& f, double x, double y) {
  (void)f, (void)x, (void)y;
  return 0;  // the flavour exists for the float build only
}

template <typename S>
struct Window {
  std::vector<std::unique_ptr<Frame<S>>> frames;
  int threads = 1;
  bool device_ops = false;  // float build: residual / energy arithmetic in the CUDA kernels' operation order
  std::vector<double> Hpose, bpose, Hschur, bschur;  // last linearisation

  int N() const { return (int)frames.size(); }

  // K6 -- PBA/first_estimate_jacobians.hpp:14-71
  void first_estimate() {
    const int n = N();
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (int r = 0; r < n; ++r) {
      Frame<S>& R = *frames[r];
      for (int t = 0; t < n; ++t) {
        if (t == r) continue;
        Frame<S>& T = *frames[t];
        const Pair<S> pc = make_pair(R, T);
        auto& res = R.res[t];
        for (size_t l = 0; l < R.lms.size(); ++l) {
          Landmark<S>& lm = R.lms[l];
          if (lm.is_marg && !lm.to_marg) continue;
          S tp[P][2];
          res[l].jac_valid = reproject_jac(pc.M0, pc.t0, pc, lm, lm.idepth, T.W, T.H, tp, res[l]);
          for (int i = 0; i < P; ++i) lm.corrected[i] = pc.s0 * (lm.patch[i] - pc.b_r0);
          res[l].bcs = pc.s0;
        }
      }
    }
  }

  // K1 / K2 -- PBA/evaluate_jacobians.hpp:20-202
  void evaluate(double sigma_d, bool fej, bool eval_jac, bool huber) {
    const int n = N();
    const S sigma = (S)sigma_d, sig2 = (S)(sigma_d * sigma_d);
    std::vector<Pair<S>> pairs((size_t)n * n);
    for (int r = 0; r < n; ++r)
      for (int t = 0; t < n; ++t)
        if (r != t) pairs[(size_t)r * n + t] = make_pair(*frames[r], *frames[t]);
    struct Job {
      int r, t;
      size_t l0, l1;
    };
    std::vector<Job> jobs;
    const size_t CH = 64;
    for (int r = 0; r < n; ++r)
      for (int t = 0; t < n; ++t)
        if (r != t)
          for (size_t l0 = 0; l0 < frames[r]->lms.size(); l0 += CH)
            jobs.push_back({r, t, l0, std::min(l0 + CH, frames[r]->lms.size())});
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (size_t j = 0; j < jobs.size(); ++j) {
      const Job jb = jobs[j];
      Frame<S>& R = *frames[jb.r];
      Frame<S>& T = *frames[jb.t];
      const Pair<S>& pc = pairs[(size_t)jb.r * n + jb.t];
      const S* adj = fej ? pc.adj0 : pc.adj;
      auto& resv = R.res[jb.t];
      for (size_t l = jb.l0; l < jb.l1; ++l) {
        const Landmark<S>& lm = R.lms[l];
        if (lm.is_marg && !lm.to_marg) continue;
        ResidualPoint<S>& res = resv[l];
        res.was_estimated = true;
        S tp[P][2], corrected[P], dshift;
        const S rho = lm.idepth + lm.idepth_step;
        bool ok;
        if (fej || !eval_jac) {
          ok = reproject_values(pc, lm, rho, T.W, T.H, tp);
          ok = ok && (!fej || res.jac_valid);
          dshift = res.bcs;
          for (int i = 0; i < P; ++i) corrected[i] = lm.corrected[i];
        } else {
          ok = res.jac_valid = reproject_jac(pc.M, pc.tr, pc, lm, rho, T.W, T.H, tp, res);
          for (int i = 0; i < P; ++i) corrected[i] = pc.s * (lm.patch[i] - pc.b_r);
          dshift = pc.s;
        }
        if (ok) {  // CameraMask::valid<false>, camera_mask.hpp:48-89
          for (int i = 0; i < P && ok; ++i)
            ok = T.mask[(size_t)((int)std::round(tp[i][1])) * T.W + (int)std::round(tp[i][0])] != 0;
        }
        if (!ok) res.cand = K_OOB;
        if (ok && res.status == K_OK) {
          res.cand = K_OK;
          S dIu[P], dIv[P], n2 = 0;
          const bool dev = device_ops && sizeof(S) == 4;
          for (int i = 0; i < P; ++i) {
            S smp[3];
            sample(T, tp[i][0], tp[i][1], smp);
            dIu[i] = smp[1];
            dIv[i] = smp[2];
            if (dev) {  // r = fma(-s, patch - b_r, I - b_t) with I from the kernels' tap chain
              const S I = (S)sample_intensity_device_ops(T, tp[i][0], tp[i][1]);
              res.residuals[i] = (S)fmaf(-(float)pc.s, (float)(lm.patch[i] - pc.b_r), (float)(I - pc.b_t));
            } else {
              res.residuals[i] = (smp[0] - pc.b_t) - pc.s * (lm.patch[i] - pc.b_r);
            }
            n2 += res.residuals[i] * res.residuals[i];
          }
          S sig2_cmp = sig2;
          if (dev) {  // the kernels' 8-lane butterfly: (i, i^4), then (i, i^2), then (i, i^1); sigma^2 in float
            S q[P];
            for (int i = 0; i < P; ++i) q[i] = res.residuals[i] * res.residuals[i];
            const S a0 = q[0] + q[4], a1 = q[1] + q[5], a2 = q[2] + q[6], a3 = q[3] + q[7];
            n2 = (a0 + a2) + (a1 + a3);
            sig2_cmp = sigma * sigma;
          }
          res.energy = n2 * S(0.5);
          res.huber_weight = 1;
          if (huber && n2 > sig2_cmp) {
            const S nrm = std::sqrt(n2);
            res.huber_weight = sigma / nrm;
            res.energy = dev ? (S)fmaf((float)sigma, (float)nrm, -(float)(sig2_cmp * S(0.5))) : sigma * nrm - sig2 * S(0.5);
          }
          if (eval_jac) {
            for (int i = 0; i < P; ++i) {
              S g[6];
              for (int k = 0; k < 6; ++k) g[k] = dIv[i] * res.d_v_t[i * 6 + k] + dIu[i] * res.d_u_t[i * 6 + k];
              for (int k = 0; k < 6; ++k) {
                S a = 0;
                for (int m = 0; m < 6; ++m) a += g[m] * adj[m * 6 + k];
                res.Jref[i * 8 + k] = a;
                res.Jtgt[i * 8 + k] = -g[k];
              }
              res.d_idepth[i] = dIu[i] * res.d_u_idepth[i] + dIv[i] * res.d_v_idepth[i];
              res.Jref[i * 8 + 6] = corrected[i];
              res.Jref[i * 8 + 7] = dshift;
              res.Jtgt[i * 8 + 6] = -corrected[i];
              res.Jtgt[i * 8 + 7] = -1;
            }
          }
        } else {
          for (int i = 0; i < P; ++i) res.residuals[i] = 0;
          res.energy = 0;
          if (eval_jac) {
            memset(res.Jref, 0, sizeof(res.Jref));
            memset(res.Jtgt, 0, sizeof(res.Jtgt));
            memset(res.d_idepth, 0, sizeof(res.d_idepth));
          }
        }
      }
    }
  }

  static inline void kahan(double& sum, double& comp, double x) {  // internal/matrix_accumulator.hpp:39-45
    const double y = x - comp;
    const double t = sum + y;
    comp = (t - sum) - y;
    sum = t;
  }

  // K3 -- PBA/hessian_block_evaluation.hpp:38-164 (parallel over reference frames only, Kahan sums)
  void pose_pose(bool for_marg, double* H, double* b) {
    const int n = N(), D = 8 * n;
    std::fill(H, H + (size_t)D * D, 0.0);
    std::fill(b, b + D, 0.0);
#pragma omp parallel for schedule(dynamic) num_threads(std::min(threads, n))
    for (int r = 0; r < n; ++r) {
      const Frame<S>& R = *frames[r];
      for (int t = 0; t < n; ++t) {
        if (t == r) continue;
        S acc[208] = {}, cmp[208] = {};  // accumulators are Precision (= Scalar) in the reference
        const auto& resv = R.res[t];
        for (size_t l = 0; l < R.lms.size(); ++l) {
          const Landmark<S>& lm = R.lms[l];
          if (for_marg ? !lm.to_marg : lm.is_marg) continue;
          const ResidualPoint<S>& rp = resv[l];
          const S w = rp.huber_weight;
          S v[208];
          for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) {
              S rr = 0, rt = 0, tt = 0;
              for (int p = 0; p < P; ++p) {
                rr += rp.Jref[p * 8 + i] * rp.Jref[p * 8 + j];
                rt += rp.Jref[p * 8 + i] * rp.Jtgt[p * 8 + j];
                tt += rp.Jtgt[p * 8 + i] * rp.Jtgt[p * 8 + j];
              }
              v[i * 8 + j] = w * rr;
              v[64 + i * 8 + j] = w * rt;
              v[128 + i * 8 + j] = w * tt;
            }
          for (int i = 0; i < 8; ++i) {
            S br = 0, bt = 0;
            for (int p = 0; p < P; ++p) {
              br += rp.Jref[p * 8 + i] * rp.residuals[p];
              bt += rp.Jtgt[p * 8 + i] * rp.residuals[p];
            }
            v[192 + i] = w * br;
            v[200 + i] = w * bt;
          }
          for (int k = 0; k < 208; ++k) {
            const S y = v[k] - cmp[k];
            const S tt = acc[k] + y;
            cmp[k] = (tt - acc[k]) - y;
            acc[k] = tt;
          }
        }
#pragma omp critical(posepose)
        {
          for (int i = 0; i < 8; ++i) {
            for (int j = 0; j < 8; ++j) {
              H[(size_t)(8 * r + i) * D + 8 * r + j] += (double)acc[i * 8 + j];
              H[(size_t)(8 * r + i) * D + 8 * t + j] = (double)acc[64 + i * 8 + j];  // assignment (quirk Q3)
              H[(size_t)(8 * t + i) * D + 8 * t + j] += (double)acc[128 + i * 8 + j];
            }
            b[8 * r + i] += (double)acc[192 + i];
            b[8 * t + i] += (double)acc[200 + i];
          }
        }
      }
    }
    for (int bi = 0; bi < n; ++bi) {
      for (int i = 0; i < 8; ++i)
        for (int j = i + 1; j < 8; ++j) H[(size_t)(8 * bi + i) * D + 8 * bi + j] = H[(size_t)(8 * bi + j) * D + 8 * bi + i];
      for (int bj = bi + 1; bj < n; ++bj)
        for (int i = 0; i < 8; ++i)
          for (int j = 0; j < 8; ++j) {
            double& a = H[(size_t)(8 * bi + i) * D + 8 * bj + j];
            double& c = H[(size_t)(8 * bj + j) * D + 8 * bi + i];
            a += c;
            c = a;
          }
    }
  }

  // K4 -- PBA/hessian_block_evaluation.hpp:169-236
  void schur(bool for_marg, double* Hs, double* bs) {
    const int n = N(), D = 8 * n;
    std::fill(Hs, Hs + (size_t)D * D, 0.0);
    std::fill(bs, bs + D, 0.0);
#pragma omp parallel for schedule(dynamic) num_threads(std::min(threads, n))
    for (int r = 0; r < n; ++r) {
      Frame<S>& R = *frames[r];
      std::vector<S> hacc((size_t)D * D, 0), hcmp((size_t)D * D, 0), bacc(D, 0), bcmp(D, 0), hpd(D);
      for (size_t l = 0; l < R.lms.size(); ++l) {
        Landmark<S>& lm = R.lms[l];
        if (for_marg ? !lm.to_marg : lm.is_marg) continue;
        std::fill(hpd.begin(), hpd.end(), S(0));
        S bd = 0, hdd = 0;
        for (int t = 0; t < n; ++t) {
          if (t == r) continue;
          const ResidualPoint<S>& rp = R.res[t][l];
          const S w = rp.huber_weight;
          S dd = 0, dr = 0;
          for (int p = 0; p < P; ++p) {
            dd += rp.d_idepth[p] * rp.d_idepth[p];
            dr += rp.d_idepth[p] * rp.residuals[p];
          }
          for (int i = 0; i < 8; ++i) {
            S a = 0, c = 0;
            for (int p = 0; p < P; ++p) {
              a += rp.Jref[p * 8 + i] * rp.d_idepth[p];
              c += rp.Jtgt[p * 8 + i] * rp.d_idepth[p];
            }
            hpd[8 * r + i] += w * a;
            hpd[8 * t + i] += w * c;
          }
          hdd += w * dd;
          bd += w * dr;
        }
        lm.b_d = bd;
        lm.hpd = hpd;
        if (hdd > S(1e-15)) {
          if (for_marg && R.fixed) hdd += S(1e8);
          lm.inv_hdd = S(1) / hdd;
          lm.ill = false;
          const S ib = lm.inv_hdd * bd;
          for (int i = 0; i < D; ++i) {
            const S y = ib * hpd[i] - bcmp[i];
            const S tt = bacc[i] + y;
            bcmp[i] = (tt - bacc[i]) - y;
            bacc[i] = tt;
            const S hi = lm.inv_hdd * hpd[i];
            S* ha = hacc.data() + (size_t)i * D;
            S* hc = hcmp.data() + (size_t)i * D;
            for (int j = 0; j < D; ++j) {
              const S y2 = hi * hpd[j] - hc[j];
              const S t2 = ha[j] + y2;
              hc[j] = (t2 - ha[j]) - y2;
              ha[j] = t2;
            }
          }
        } else {
          lm.ill = true;
        }
      }
#pragma omp critical(schur)
      {
        for (size_t i = 0; i < (size_t)D * D; ++i) Hs[i] += (double)hacc[i];
        for (int i = 0; i < D; ++i) bs[i] += (double)bacc[i];
      }
    }
  }

  // K5 -- PBA/hessian_block_evaluation.hpp:238-263
  void calculate_idepths(const double* step, double lambda) {
    const int n = N(), D = 8 * n;
    const S k = (S)(1.0 / (1.0 + lambda));
    for (int r = 0; r < n; ++r) {
      Frame<S>& R = *frames[r];
#pragma omp parallel for schedule(static) num_threads(threads)
      for (size_t l = 0; l < R.lms.size(); ++l) {
        Landmark<S>& lm = R.lms[l];
        if (lm.is_marg || lm.ill || (int)lm.hpd.size() != D) continue;
        S dot = 0;
        for (int i = 0; i < D; ++i) dot += lm.hpd[i] * (S)step[i];
        lm.idepth_step = -((lm.b_d - dot) * k * lm.inv_hdd);
      }
    }
  }

  double landmarks_energy(bool for_marg, int* nvalid) {  // problem.hpp:93-144
    double e = 0;
    int nv = 0;
    const int n = N();
    for (int r = 0; r < n; ++r)
      for (int t = 0; t < n; ++t) {
        if (r == t) continue;
        const Frame<S>& R = *frames[r];
        for (size_t l = 0; l < R.lms.size(); ++l) {
          const Landmark<S>& lm = R.lms[l];
          if (for_marg ? !lm.to_marg : lm.is_marg) continue;
          e += (double)R.res[t][l].energy;
          nv += R.res[t][l].energy > 0;
        }
      }
    *nvalid = nv;
    return e;
  }

  void change_statuses(bool accept) {
    for (auto& f : frames)
      for (auto& v : f->res)
        for (auto& r : v) {
          if (accept) r.status = r.cand;
          else r.cand = r.status;
        }
  }

  void accept(double* state_sq, double* step_sq) {  // problem.hpp:366-388
    double a = 0, b = 0;
    for (auto& f : frames) {
      for (int k = 0; k < 8; ++k) a += f->eps[k] * f->eps[k];
      a += f->ab0[0] * f->ab0[0] + f->ab0[1] * f->ab0[1];
      for (int k = 0; k < 8; ++k) {
        f->eps[k] += f->step[k];
        b += f->step[k] * f->step[k];
        f->step[k] = 0;
      }
      for (auto& lm : f->lms) {
        a += (double)lm.idepth * lm.idepth;
        lm.idepth += lm.idepth_step;
        b += (double)lm.idepth_step * lm.idepth_step;
        lm.idepth_step = 0;
      }
    }
    change_statuses(true);
    *state_sq = a;
    *step_sq = b;
  }

  // updatePointStatuses, energy/problems/src/photometric_bundle_adjustment.cpp:322-406: 75 % quantile of the energies of
  // kOk residuals of active landmarks (+ sigma^2 / 2) -> residuals above it are reset to {kOutlier} (quirk Q6), inlier
  // counts, relative baseline, outlier flag.  Arithmetic in Scalar like the reference (Precision).
  double update_point_statuses(int min_valid, double sigma_d) {
    const int n = N();
    std::vector<S> energies;
    for (int r = 0; r < n; ++r) {
      const Frame<S>& R = *frames[r];
      for (int t = 0; t < n; ++t) {
        if (t == r || frames[t]->is_marg) continue;
        for (size_t l = 0; l < R.lms.size(); ++l)
          if (!R.lms[l].is_marg && R.res[t][l].status == K_OK) energies.push_back(R.res[t][l].energy);
      }
    }
    S thr = 0;
    if (!energies.empty()) {
      const size_t k = (size_t)((double)energies.size() * 0.75);
      std::nth_element(energies.begin(), energies.begin() + (long)k, energies.end());
      thr = energies[k] + (S)(sigma_d * sigma_d / 2);
    }
    std::vector<double> tw((size_t)n * 3);
    for (int f = 0; f < n; ++f) {  // translation of tWorldAgent() = T_lin exp(eps[0:6])  (local_frame.hpp:525-527)
      const SE3 T = mul(frames[f]->T_lin, expm(frames[f]->eps, 1.0));
      for (int k = 0; k < 3; ++k) tw[3 * f + k] = T.t[k];
    }
    for (int r = 0; r < n; ++r) {
      Frame<S>& R = *frames[r];
      for (size_t l = 0; l < R.lms.size(); ++l) {
        Landmark<S>& lm = R.lms[l];
        if (lm.is_marg) continue;
        uint32_t valid = 0;
        for (int t = 0; t < n; ++t) {
          if (t == r || frames[t]->is_marg) continue;
          ResidualPoint<S>& rp = R.res[t][l];
          if (rp.energy > thr) {
            rp = ResidualPoint<S>();
            rp.status = rp.cand = K_OUTLIER;
          }
          if (rp.status == K_OK) {
            const double dx = tw[3 * r] - tw[3 * t], dy = tw[3 * r + 1] - tw[3 * t + 1], dz = tw[3 * r + 2] - tw[3 * t + 2];
            const S dist = (S)sqrt(dx * dx + dy * dy + dz * dz);
            lm.rel_baseline = std::max(lm.rel_baseline, lm.idepth * dist);
            ++valid;
          }
        }
        lm.n_inliers = valid;
        if ((int)valid < min_valid) lm.is_outlier = true;
      }
    }
    return (double)thr;
  }

  void reject() {
    for (auto& f : frames) {
      for (int k = 0; k < 8; ++k) f->step[k] = 0;
      for (auto& lm : f->lms) lm.idepth_step = 0;
    }
    change_statuses(false);
  }
};

// ---- NormalLinearSystem::solve (energy/problems/src/normal_linear_system.cpp:10-59) --------------
// Jacobi preconditioner (+10 floor) and a diagonally pivoted LDL^T (Eigen's LDLT restated)
void normal_solve(int n, const double* H, const double* b, double* x) {
  std::vector<double> p(n), A((size_t)n * n), y(n);
  for (int i = 0; i < n; ++i) p[i] = 1.0 / sqrt(H[(size_t)i * n + i] + 10.0);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = p[i] * H[(size_t)i * n + j] * p[j];
  for (int i = 0; i < n; ++i) y[i] = p[i] * b[i];
  std::vector<int> perm(n);
  for (int i = 0; i < n; ++i) perm[i] = i;
  for (int k = 0; k < n; ++k) {  // in-place LDL^T with symmetric pivoting on the largest |diagonal|
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (fabs(A[(size_t)i * n + i]) > fabs(A[(size_t)piv * n + piv])) piv = i;
    if (piv != k) {
      for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)piv * n + j]);
      for (int i = 0; i < n; ++i) std::swap(A[(size_t)i * n + k], A[(size_t)i * n + piv]);
      std::swap(perm[k], perm[piv]);
    }
    const double d = A[(size_t)k * n + k];
    if (d == 0) continue;
    for (int i = k + 1; i < n; ++i) A[(size_t)i * n + k] /= d;
    for (int i = k + 1; i < n; ++i) {
      const double lik = A[(size_t)i * n + k];
      for (int j = k + 1; j <= i; ++j) A[(size_t)i * n + j] -= lik * d * A[(size_t)j * n + k];
    }
    for (int i = k + 1; i < n; ++i)
      for (int j = i + 1; j < n; ++j) A[(size_t)i * n + j] = A[(size_t)j * n + i];
  }
  std::vector<double> z(n);
  for (int i = 0; i < n; ++i) z[i] = y[perm[i]];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) z[i] -= A[(size_t)i * n + j] * z[j];
  for (int i = 0; i < n; ++i) {
    const double d = A[(size_t)i * n + i];
    z[i] = d != 0 ? z[i] / d : 0;
  }
  for (int i = n - 1; i >= 0; --i)
    for (int j = i + 1; j < n; ++j) z[i] -= A[(size_t)j * n + i] * z[j];
  for (int i = 0; i < n; ++i) x[perm[i]] = p[perm[i]] * z[i];
}

struct AnyWindow {
  bool use_float;
  Window<float> wf;
  Window<double> wd;
};

template <typename S>
int push_frame(Window<S>& w, int id, const float* img, const uint8_t* mask, int W, int H, const double* T,
               double exposure, const double* ab0, const double* intr, int fixed) {
  auto f = std::make_unique<Frame<S>>();
  f->id = id;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) f->T_lin.R[i * 3 + j] = T[i * 4 + j];
    f->T_lin.t[i] = T[i * 4 + 3];
  }
  f->exposure = exposure;
  f->ab0[0] = ab0[0];
  f->ab0[1] = ab0[1];
  memcpy(f->intr, intr, sizeof(f->intr));
  f->W = W;
  f->H = H;
  f->image.resize((size_t)W * H * 3);
  for (size_t i = 0; i < f->image.size(); ++i) f->image[i] = (S)img[i];
  f->mask.assign((size_t)W * H, 255);
  if (mask) memcpy(f->mask.data(), mask, (size_t)W * H);
  f->fixed = fixed != 0;
  w.frames.push_back(std::move(f));
  const int n = w.N();
  for (auto& fr : w.frames) fr->res.resize(n);
  for (int r = 0; r < n; ++r)
    for (int t = 0; t < n; ++t)
      if (r != t && w.frames[r]->res[t].size() != w.frames[r]->lms.size())
        w.frames[r]->res[t].resize(w.frames[r]->lms.size());
  return n - 1;
}

template <typename S>
void set_landmarks(Window<S>& w, int slot, int n, const double* uv, const double* idepth, const double* patch,
                   const uint8_t* flags) {
  Frame<S>& f = *w.frames[slot];
  f.lms.assign(n, Landmark<S>());
  for (int l = 0; l < n; ++l) {
    Landmark<S>& lm = f.lms[l];
    lm.proj[0] = (S)uv[2 * l];
    lm.proj[1] = (S)uv[2 * l + 1];
    lm.idepth = (S)idepth[l];
    for (int i = 0; i < P; ++i) {
      lm.patch[i] = (S)patch[8 * l + i];
      lm.ref_pattern[i][0] = lm.proj[0] + (S)PAT[i][0];
      lm.ref_pattern[i][1] = lm.proj[1] + (S)PAT[i][1];
    }
    const int fl = flags ? flags[l] : 0;
    lm.is_marg = fl & 1;
    lm.to_marg = fl & 2;
    lm.is_outlier = fl & 4;
  }
  for (int t = 0; t < w.N(); ++t)
    if (t != slot) f.res[t].assign(n, ResidualPoint<S>());
}

#define DISPATCH(expr_f, expr_d) \
  if (a->use_float) {            \
    auto& w = a->wf;             \
    (void)w;                     \
    expr_f;                      \
  } else {                       \
    auto& w = a->wd;             \
    (void)w;                     \
    expr_d;                      \
  }

}  // namespace

extern "C" {

void* cpuref_create(int use_float, int threads) {
  AnyWindow* a = new AnyWindow();
  a->use_float = use_float != 0;
  a->wf.threads = a->wd.threads = threads > 0 ? threads : 1;
  return a;
}
void cpuref_destroy(void* h) { delete (AnyWindow*)h; }
int cpuref_max_threads() { return omp_get_max_threads(); }

int cpuref_push_frame(void* h, int id, const float* img, const uint8_t* mask, int W, int H, const double* T,
                      double exposure, const double* ab0, const double* intr, int fixed) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(return push_frame(w, id, img, mask, W, H, T, exposure, ab0, intr, fixed),
           return push_frame(w, id, img, mask, W, H, T, exposure, ab0, intr, fixed));
}
void cpuref_set_landmarks(void* h, int slot, int n, const double* uv, const double* idepth, const double* patch,
                          const uint8_t* flags) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(set_landmarks(w, slot, n, uv, idepth, patch, flags), set_landmarks(w, slot, n, uv, idepth, patch, flags));
}
void cpuref_set_statuses(void* h, int r, int t, int n, const uint8_t* st) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(for (int l = 0; l < n; ++l) w.frames[r]->res[t][l].status = w.frames[r]->res[t][l].cand = st[l],
           for (int l = 0; l < n; ++l) w.frames[r]->res[t][l].status = w.frames[r]->res[t][l].cand = st[l]);
}
void cpuref_set_state(void* h, const double* eps, const double* step) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(
      for (int f = 0; f < w.N(); ++f) for (int k = 0; k < 8; ++k) {
        if (eps) w.frames[f]->eps[k] = eps[8 * f + k];
        if (step) w.frames[f]->step[k] = step[8 * f + k];
      },
      for (int f = 0; f < w.N(); ++f) for (int k = 0; k < 8; ++k) {
        if (eps) w.frames[f]->eps[k] = eps[8 * f + k];
        if (step) w.frames[f]->step[k] = step[8 * f + k];
      });
}
void cpuref_get_state(void* h, double* eps, double* step) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(
      for (int f = 0; f < w.N(); ++f) for (int k = 0; k < 8; ++k) {
        eps[8 * f + k] = w.frames[f]->eps[k];
        step[8 * f + k] = w.frames[f]->step[k];
      },
      for (int f = 0; f < w.N(); ++f) for (int k = 0; k < 8; ++k) {
        eps[8 * f + k] = w.frames[f]->eps[k];
        step[8 * f + k] = w.frames[f]->step[k];
      });
}
void cpuref_first_estimate(void* h) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.first_estimate(), w.first_estimate());
}
void cpuref_evaluate(void* h, double sigma, int fej, int eval_jac, int huber) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.evaluate(sigma, fej, eval_jac, huber), w.evaluate(sigma, fej, eval_jac, huber));
}
void cpuref_pose_pose(void* h, int for_marg, double* H, double* b) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.pose_pose(for_marg, H, b), w.pose_pose(for_marg, H, b));
}
void cpuref_schur(void* h, int for_marg, double* H, double* b) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.schur(for_marg, H, b), w.schur(for_marg, H, b));
}
void cpuref_calculate_idepths(void* h, const double* step, double lambda) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.calculate_idepths(step, lambda), w.calculate_idepths(step, lambda));
}
double cpuref_landmarks_energy(void* h, int for_marg, int* n) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(return w.landmarks_energy(for_marg, n), return w.landmarks_energy(for_marg, n));
}
void cpuref_accept(void* h, double* a_, double* b_) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.accept(a_, b_), w.accept(a_, b_));
}
void cpuref_reject(void* h) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.reject(), w.reject());
}
void cpuref_change_statuses(void* h, int accept) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(w.change_statuses(accept), w.change_statuses(accept));
}
void cpuref_normal_solve(int n, const double* H, const double* b, double* x) { normal_solve(n, H, b, x); }
void cpuref_set_device_ops(void* h, int on) {
  AnyWindow* a = (AnyWindow*)h;
  a->wf.device_ops = a->wd.device_ops = on != 0;
}
double cpuref_update_point_statuses(void* h, int min_valid, double sigma) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(return w.update_point_statuses(min_valid, sigma), return w.update_point_statuses(min_valid, sigma));
}
void cpuref_get_jac_valid(void* h, int r, int t, uint8_t* out) {
  AnyWindow* a = (AnyWindow*)h;
  DISPATCH(for (size_t l = 0; l < w.frames[r]->res[t].size(); ++l) out[l] = w.frames[r]->res[t][l].jac_valid,
           for (size_t l = 0; l < w.frames[r]->res[t].size(); ++l) out[l] = w.frames[r]->res[t][l].jac_valid);
}
// plant a landmark state (e.g. the device's, float bits carried exactly by the doubles)
void cpuref_set_idepths(void* h, int slot, int n, const double* idepth, const double* idepth_step) {
  AnyWindow* a = (AnyWindow*)h;
#define BODY                                                   \
  auto& v = w.frames[slot]->lms;                               \
  using SS = decltype(v[0].idepth);                            \
  for (int l = 0; l < n && l < (int)v.size(); ++l) {           \
    if (idepth) v[l].idepth = (SS)idepth[l];                   \
    if (idepth_step) v[l].idepth_step = (SS)idepth_step[l];    \
  }
  DISPATCH(BODY, BODY);
#undef BODY
}
void cpuref_get_landmark_flags(void* h, int slot, uint8_t* outlier, uint32_t* n_inliers, double* rel_baseline) {
  AnyWindow* a = (AnyWindow*)h;
#define BODY                                                   \
  const auto& v = w.frames[slot]->lms;                         \
  for (size_t l = 0; l < v.size(); ++l) {                      \
    if (outlier) outlier[l] = v[l].is_outlier;                 \
    if (n_inliers) n_inliers[l] = v[l].n_inliers;              \
    if (rel_baseline) rel_baseline[l] = v[l].rel_baseline;     \
  }
  DISPATCH(BODY, BODY);
#undef BODY
}

// residual block download in the same layout as dpba_download_residual_block (doubles)
void cpuref_get_residuals(void* h, int r, int t, double* res, double* jref, double* jtgt, double* did, double* wgt,
                          double* energy, uint8_t* status, uint8_t* cand) {
  AnyWindow* a = (AnyWindow*)h;
#define BODY                                                               \
  const auto& v = w.frames[r]->res[t];                                     \
  for (size_t l = 0; l < v.size(); ++l) {                                  \
    for (int i = 0; i < 8; ++i) {                                          \
      if (res) res[l * 8 + i] = v[l].residuals[i];                         \
      if (did) did[l * 8 + i] = v[l].d_idepth[i];                          \
    }                                                                      \
    for (int i = 0; i < 64; ++i) {                                         \
      if (jref) jref[l * 64 + i] = v[l].Jref[i];                           \
      if (jtgt) jtgt[l * 64 + i] = v[l].Jtgt[i];                           \
    }                                                                      \
    if (wgt) wgt[l] = v[l].huber_weight;                                   \
    if (energy) energy[l] = v[l].energy;                                   \
    if (status) status[l] = v[l].status;                                   \
    if (cand) cand[l] = v[l].cand;                                         \
  }
  DISPATCH(BODY, BODY);
#undef BODY
}
void cpuref_get_landmarks(void* h, int slot, double* idepth, double* idepth_step, double* inv_hdd, double* b_d,
                          uint8_t* ill, double* hpd, int hpd_stride) {
  AnyWindow* a = (AnyWindow*)h;
#define BODY                                                                        \
  const auto& v = w.frames[slot]->lms;                                              \
  for (size_t l = 0; l < v.size(); ++l) {                                           \
    if (idepth) idepth[l] = v[l].idepth;                                            \
    if (idepth_step) idepth_step[l] = v[l].idepth_step;                             \
    if (inv_hdd) inv_hdd[l] = v[l].inv_hdd;                                         \
    if (b_d) b_d[l] = v[l].b_d;                                                     \
    if (ill) ill[l] = v[l].ill;                                                     \
    if (hpd)                                                                        \
      for (int i = 0; i < hpd_stride; ++i)                                          \
        hpd[l * hpd_stride + i] = i < (int)v[l].hpd.size() ? (double)v[l].hpd[i] : 0.0; \
  }
  DISPATCH(BODY, BODY);
#undef BODY
}

// One Gauss-Newton / LM iteration exactly as the loop body of levenberg_marquardt_algorithm::solve with
// force_accept (levenberg_marquardt_algorithm.hpp:85-114): linearize, calculateStep, calculateEnergy, acceptStep.
// Priors: affine-brightness regulariser + fixed-frame regulariser (problem.hpp:37-77); no marginalised prior.
// Returns the new energy; times[0..5] = sweep(K1), posepose(K3), schur(K4), solve+K5, energy sweep(K2), total [s].
double cpuref_gn_iteration(void* h, double sigma, int fej, double lambda, const double* ab_reg, double fixed_reg,
                           double* times, double* step_out) {
  AnyWindow* a = (AnyWindow*)h;
  using clk = std::chrono::steady_clock;
  auto sec = [](clk::time_point x, clk::time_point y) { return std::chrono::duration<double>(y - x).count(); };
  double energy = 0;
#define BODY                                                                                              \
  const int n = w.N(), D = 8 * n;                                                                         \
  std::vector<double> Hp((size_t)D * D), bp(D), Hs((size_t)D * D), bs(D), H((size_t)D * D), b(D), step(D); \
  auto t0 = clk::now();                                                                                   \
  w.evaluate(sigma, fej, true, true);                                                                     \
  auto t1 = clk::now();                                                                                   \
  w.pose_pose(false, Hp.data(), bp.data());                                                               \
  auto t2 = clk::now();                                                                                   \
  w.schur(false, Hs.data(), bs.data());                                                                   \
  auto t3 = clk::now();                                                                                   \
  for (int f = 0; f < n; ++f) {                                                                           \
    const auto& F = *w.frames[f];                                                                         \
    if (F.fixed) {                                                                                        \
      for (int k = 0; k < 8; ++k) {                                                                       \
        Hp[(size_t)(8 * f + k) * D + 8 * f + k] += fixed_reg;                                             \
        bp[8 * f + k] += fixed_reg * F.eps[k];                                                            \
      }                                                                                                   \
    } else {                                                                                              \
      for (int k = 0; k < 2; ++k) {                                                                       \
        Hp[(size_t)(8 * f + 6 + k) * D + 8 * f + 6 + k] += ab_reg[k];                                     \
        bp[8 * f + 6 + k] += ab_reg[k] * (F.ab0[k] + F.eps[6 + k]);                                       \
      }                                                                                                   \
    }                                                                                                     \
  }                                                                                                       \
  const double ks = -1.0 / (1.0 + lambda);                                                                \
  for (size_t i = 0; i < (size_t)D * D; ++i) H[i] = Hp[i] + ks * Hs[i];                                   \
  for (int i = 0; i < D; ++i) {                                                                           \
    H[(size_t)i * D + i] += lambda * Hp[(size_t)i * D + i];                                               \
    b[i] = bp[i] + ks * bs[i];                                                                            \
  }                                                                                                       \
  normal_solve(D, H.data(), b.data(), step.data());                                                       \
  for (int f = 0; f < n; ++f)                                                                             \
    for (int k = 0; k < 8; ++k) w.frames[f]->step[k] = -step[8 * f + k];                                  \
  w.calculate_idepths(step.data(), lambda);                                                               \
  auto t4 = clk::now();                                                                                   \
  w.evaluate(sigma, fej, false, true);                                                                    \
  int nv = 0;                                                                                             \
  energy = w.landmarks_energy(false, &nv);                                                                \
  for (int f = 0; f < n; ++f) {                                                                           \
    const auto& F = *w.frames[f];                                                                         \
    for (int k = 0; k < 2; ++k) {                                                                         \
      const double ab = F.ab0[k] + F.eps[6 + k] + F.step[6 + k];                                          \
      energy += 0.5 * ab_reg[k] * ab * ab;                                                                \
    }                                                                                                     \
  }                                                                                                       \
  auto t5 = clk::now();                                                                                   \
  double sa, sb;                                                                                          \
  w.accept(&sa, &sb);                                                                                     \
  auto t6 = clk::now();                                                                                   \
  if (times) {                                                                                            \
    times[0] = sec(t0, t1);                                                                               \
    times[1] = sec(t1, t2);                                                                               \
    times[2] = sec(t2, t3);                                                                               \
    times[3] = sec(t3, t4);                                                                               \
    times[4] = sec(t4, t5);                                                                               \
    times[5] = sec(t0, t6);                                                                               \
  }                                                                                                       \
  if (step_out) memcpy(step_out, step.data(), D * sizeof(double));
  DISPATCH(BODY, BODY);
#undef BODY
  return energy;
}

}  // extern "C"
