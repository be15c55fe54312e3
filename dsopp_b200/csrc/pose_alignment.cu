// B200-native coarse-tracker direct image alignment (include/dsopp_cuda_pose_alignment.h).
//
// Reference being replaced (paths relative to /root/reference/src/):
//   PoseAlignerProblem::calculateEnergy / linearize / calculateStep / acceptStep / rejectStep
//                                      energy/problems/src/eigen_pose_alignment.cpp:55-218
//   EigenPoseAlignment::solve          energy/problems/src/eigen_pose_alignment.cpp:275-329
//   levenberg_marquardt_algorithm::solve
//                                      energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:77-128
//   depth-map LocalFrame constructor   energy/problems/internal/energy/problems/photometric_bundle_adjustment/local_frame.hpp:367-392
//
// Design.  The reference runs the whole problem serially on one core: per LM iteration one sweep over the depth-map
// landmarks (1-pixel residuals) for the energy, one for the 8x8 system, and an 8x8 solve.  Here ONE kernel launch runs
// the complete LM solve of a pyramid level on ONE thread-block cluster (8 CTAs x 512 threads, distributed shared
// memory): every sweep evaluates energy AND the 44 sums of the 8x8 system at the same state (the reference's
// linearize() re-uses exactly the samples its last calculateEnergy() cached, so the two are the same numbers), the
// CTAs' partial sums meet through DSMEM after a cluster barrier, and every CTA then takes the identical fp64 LM
// decision redundantly -- no grid-wide barrier, no host round trip, no second launch per iteration.
// Arithmetic: fp32 per point, fp64 for every sum across points and for the whole LM / SE3 / 8x8 algebra.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/dsopp_cuda_pba.h"
#include "../../include/dsopp_cuda_pose_alignment.h"
#include "pba_internal.h"

namespace cg = cooperative_groups;

namespace {

constexpr int PA_CLUSTER = 8;     // CTAs of the single cluster (portable maximum)
constexpr int PA_THREADS = 512;
constexpr int PA_SUMS = 46;       // 36 (upper triangle of H) + 8 (b) + energy + n_valid
constexpr unsigned FULL = 0xffffffffu;

struct PaFrame {
  double T[12];  // world <- agent, 3x4 row-major
  double exposure, ab0[2], intr[4];
  int W, H;
};

struct PaProblem {   // kernel argument (by value)
  const float4* lm;          // {x, y, idepth, patch}
  int n;
  const float4* img;         // target image, 32-byte records {texel(x), texel(x+1)}
  const uint8_t* mask;       // null when the mask has no zero
  PaFrame ref, tgt;
  double T0[12];             // initial t_t_r (prior rotation already applied)
  int max_it;
  double lambda0, ftol, ptol, sigma, ab_reg[2], dec, inc;
};

constexpr int PA_TRACE = 64;
struct PaOut {
  double energy, T[12], ab_eps[2], H[64];
  int n_valid, converged, iterations;
  double trace_energy[PA_TRACE];   // trial energy of every loop body
  double trace_lambda[PA_TRACE];
  int trace_accept[PA_TRACE];
};

// fp32 constants of one sweep, derived in fp64 from the current t_t_r (camera_reproject.hpp:235-260)
struct PaConst {
  float A[12];   // reproject_ = K_t [R|t] K_r^-1
  float M[12];   // transform_unproject_ = [R|t] K_r^-1
  float t[3];
  float fx, fy, s, b_t, b_r;
};

__device__ __forceinline__ void ldg256_nc(const float4* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}

__device__ void se3_exp_d(const double* xi, double* R, double* t) {  // Sophus SE3::exp, tangent [upsilon; omega]
  const double* v = xi;
  const double* w = xi + 3;
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double a, b, c;
  if (th < 1e-10) {
    a = 1.0, b = 0.5, c = 1.0 / 6.0;
  } else {
    double sn, cs;
    sincos(th, &sn, &cs);
    a = sn / th, b = (1.0 - cs) / th2, c = (th - sn) / (th2 * th);
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
    R[i] = I + a * W[i] + b * W2[i];
    V[i] = I + b * W[i] + c * W2[i];
  }
  for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
}

// T <- exp(xi) * T   (leftIncrement, energy/motion/include/energy/motion/se3_motion.hpp:231-236), T 3x4 row-major
__device__ void left_increment(const double* xi, double* T) {
  double R[9], t[3], out[12];
  se3_exp_d(xi, R, t);
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 4; ++j) {
      double s = (j == 3) ? t[i] : 0.0;
      for (int k = 0; k < 3; ++k) s += R[i * 3 + k] * T[k * 4 + j];
      out[i * 4 + j] = s;
    }
  }
  for (int i = 0; i < 12; ++i) T[i] = out[i];
}

__device__ void make_consts(const double* T, const PaProblem& p, double ab_eps0, double ab_eps1, PaConst& c) {
  const double fx = p.ref.intr[0], fy = p.ref.intr[1], cx = p.ref.intr[2], cy = p.ref.intr[3];
  double m[12];
  for (int i = 0; i < 3; ++i) {
    m[i * 4 + 0] = T[i * 4 + 0] * (1.0 / fx);
    m[i * 4 + 1] = T[i * 4 + 1] * (1.0 / fy);
    m[i * 4 + 2] = T[i * 4 + 0] * (-cx / fx) + T[i * 4 + 1] * (-cy / fy) + T[i * 4 + 2];
    m[i * 4 + 3] = T[i * 4 + 3];
  }
  const double* it = p.tgt.intr;
  for (int j = 0; j < 4; ++j) {
    c.M[0 + j] = (float)m[0 + j];
    c.M[4 + j] = (float)m[4 + j];
    c.M[8 + j] = (float)m[8 + j];
    c.A[0 + j] = (float)(it[0] * m[0 + j] + it[2] * m[8 + j]);
    c.A[4 + j] = (float)(it[1] * m[4 + j] + it[3] * m[8 + j]);
    c.A[8 + j] = (float)m[8 + j];
  }
  for (int i = 0; i < 3; ++i) c.t[i] = (float)T[i * 4 + 3];
  c.fx = (float)it[0];
  c.fy = (float)it[1];
  const double a_t = p.tgt.ab0[0] + ab_eps0;
  c.s = (float)((p.tgt.exposure / p.ref.exposure) * exp(a_t - p.ref.ab0[0]));  // eigen_pose_alignment.cpp:71-72
  c.b_t = (float)(p.tgt.ab0[1] + ab_eps1);
  c.b_r = (float)p.ref.ab0[1];
}

// NormalLinearSystem::solve (energy/problems/src/normal_linear_system.cpp:10-59) of the 8x8 system
//   (H + lambda diag(H)) x = b,  Jacobi preconditioner 1/sqrt(diag + 10), LDL^T.  One thread.
__device__ void solve8(const double* H, const double* b, double lambda, double* x) {
  double A[64], pre[8], y[8], d[8];
  for (int i = 0; i < 8; ++i) pre[i] = 1.0 / sqrt(H[i * 8 + i] * (1.0 + lambda) + 10.0);
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) A[i * 8 + j] = (H[i * 8 + j] + (i == j ? H[i * 8 + i] * lambda : 0.0)) * pre[i] * pre[j];
  for (int i = 0; i < 8; ++i) y[i] = b[i] * pre[i];
  // LDL^T in place (lower), forward / diagonal / backward substitution
  for (int j = 0; j < 8; ++j) {
    double dj = A[j * 8 + j];
    for (int k = 0; k < j; ++k) dj -= A[j * 8 + k] * A[j * 8 + k] * d[k];
    d[j] = dj;
    const double inv = dj != 0.0 ? 1.0 / dj : 0.0;
    for (int i = j + 1; i < 8; ++i) {
      double v = A[i * 8 + j];
      for (int k = 0; k < j; ++k) v -= A[i * 8 + k] * A[j * 8 + k] * d[k];
      A[i * 8 + j] = v * inv;
    }
  }
  for (int i = 0; i < 8; ++i)
    for (int k = 0; k < i; ++k) y[i] -= A[i * 8 + k] * y[k];
  for (int i = 0; i < 8; ++i) y[i] = d[i] != 0.0 ? y[i] / d[i] : 0.0;
  for (int i = 7; i >= 0; --i)
    for (int k = i + 1; k < 8; ++k) y[i] -= A[k * 8 + i] * y[k];
  for (int i = 0; i < 8; ++i) x[i] = y[i] * pre[i];
}

struct PaShared {
  PaConst c;
  double part[2][PA_SUMS];        // this CTA's sums of the current / previous sweep (read by the peers through DSMEM)
  double tot[PA_SUMS];            // cluster-wide sums of the last sweep
  float warp_part[PA_THREADS / 32][PA_SUMS];
  // LM state, kept identically by every CTA
  double T[12], T_old[12], ab_eps[2], ab_old[2];
  double Hc[64], bc[8];           // current linear system (data terms), the reference's system_
  double step[8];                 // the last step_
  double energy, lambda;
  int n_valid, converged, iteration, go;
};

// one sweep over the landmarks at the state in sh.T / sh.ab_eps; leaves the sums over ALL CTAs in sh.tot.
// GRID = false: the CTAs are the 8 of one thread-block cluster, their partial sums meet through distributed shared
// memory after a cluster barrier.  GRID = true (round 2, dense depth maps): the CTAs are a cooperative grid with one CTA
// per SM, the partial sums meet in global memory after a grid barrier -- the whole chip sweeps, where the cluster
// version leaves 140 of 148 SMs idle (0.46 ms for 298 k points).  Either way every CTA adds the
// same partials in the same order and takes the identical LM decision.
template <bool GRID>
__device__ void pa_sweep(const PaProblem& p, PaShared& sh, int parity, double* __restrict__ gpart) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int blk = GRID ? (int)blockIdx.x : (int)cg::this_cluster().block_rank();
  const int nblk = GRID ? (int)gridDim.x : PA_CLUSTER;
  if (tid == 0) make_consts(sh.T, p, sh.ab_eps[0], sh.ab_eps[1], sh.c);
  __syncthreads();
  const PaConst& c = sh.c;
  const float xmaxr = (float)(p.ref.W - 5), ymaxr = (float)(p.ref.H - 5);
  const float xmax = (float)(p.tgt.W - 5), ymax = (float)(p.tgt.H - 5);
  const int W = p.tgt.W;
  const float sigma = (float)p.sigma, sigma2 = sigma * sigma;
  float acc[PA_SUMS];
#pragma unroll
  for (int k = 0; k < PA_SUMS; ++k) acc[k] = 0.f;
  const int stride = nblk * PA_THREADS;
  for (int i = blk * PA_THREADS + tid; i < p.n; i += stride) {
    const float4 lm = p.lm[i];
    const float x = lm.x, y = lm.y, rho = lm.z, patch = lm.w;
    // reprojectPattern, values (camera_reproject.hpp:270-293): success = validIdepth, ROI(reference), z > 0, ROI(target)
    bool ok = (rho > -1e-4f && rho < 1010.f) && x >= 4.f && y >= 4.f && x <= xmaxr && y <= ymaxr;
    const float X = c.A[0] * x + c.A[1] * y + (c.A[2] + c.A[3] * rho);
    const float Y = c.A[4] * x + c.A[5] * y + (c.A[6] + c.A[7] * rho);
    const float Z = c.A[8] * x + c.A[9] * y + (c.A[10] + c.A[11] * rho);
    ok = ok && Z > 0.f;
    const float rz = 1.f / Z;
    const float tu = X * rz, tv = Y * rz;
    ok = ok && tu >= 4.f && tv >= 4.f && tu <= xmax && tv <= ymax;
    if (p.mask && ok) ok = p.mask[(int)roundf(tv) * W + (int)roundf(tu)] != 0;  // mask_.valid(), :77
    // interpolateLinear (pixel_map.hpp:20-40); texel (8, 8) when the point failed
    const float su = ok ? tu : 8.f, sv = ok ? tv : 8.f;
    const int ix = (int)su, iy = (int)sv;
    const float dx = su - (float)ix, dy = sv - (float)iy, dxdy = dx * dy;
    const float w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = 1.f - dx - dy + dxdy;
    const float4* q = p.img + ((size_t)iy * W + ix) * 2;
    float4 t00, t01, t10, t11;
    ldg256_nc(q, t00, t01);
    ldg256_nc(q + 2 * (size_t)W, t10, t11);
    const float I = w11 * t11.x + w10 * t10.x + w01 * t01.x + w00 * t00.x;
    const float dIu = w11 * t11.y + w10 * t10.y + w01 * t01.y + w00 * t00.y;
    const float dIv = w11 * t11.z + w10 * t10.z + w01 * t01.z + w00 * t00.z;
    const float right = c.s * (patch - c.b_r);
    const float r = ok ? (I - c.b_t) - right : 0.f;  // :79-85
    const float r2 = r * r;
    const bool lin = r2 > sigma2;                    // :88-93
    const float nrm = fabsf(r);
    const float e = ok ? (lin ? sigma * nrm - 0.5f * sigma2 : 0.5f * r2) : 0.f;
    const float wgt = ok ? (lin ? sigma / nrm : 1.f) : 0.f;  // :148-149
    // reprojection Jacobians at the current t_t_r (camera_reproject.hpp:305-367, kCheckSuccess = false)
    const float qx = c.M[0] * x + c.M[1] * y + (c.M[2] + c.M[3] * rho);
    const float qy = c.M[4] * x + c.M[5] * y + (c.M[6] + c.M[7] * rho);
    const float qz = c.M[8] * x + c.M[9] * y + (c.M[10] + c.M[11] * rho);
    const float sI = 1.f / (ok ? qz : 1.f);
    const float b0 = qx * sI, b1 = qy * sI, nid = rho * sI, b0b1 = b0 * b1;
    const float gu = dIu * c.fx, gv = dIv * c.fy;
    float u[8];  // d_state row (:151-169): -(dI/du du/dxi + dI/dv dv/dxi), -residuals_right, -1
    u[0] = -(gu * nid);
    u[1] = -(gv * nid);
    u[2] = gu * (nid * b0) + gv * (nid * b1);
    u[3] = gu * b0b1 + gv * (b1 * b1 + 1.f);
    u[4] = -(gu * (b0 * b0 + 1.f) + gv * b0b1);
    u[5] = gu * b1 - gv * b0;
    u[6] = -right;
    u[7] = -1.f;
    int idx = 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const float wa = wgt * u[a];
#pragma unroll
      for (int b = a; b < 8; ++b) acc[idx++] += wa * u[b];
      acc[36 + a] += wa * r;
    }
    acc[44] += e;
    acc[45] += ok ? 1.f : 0.f;
  }
  // warp -> CTA -> cluster reduction (fixed order: deterministic)
#pragma unroll
  for (int k = 0; k < PA_SUMS; ++k) {
    float v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0) sh.warp_part[warp][k] = v;
  }
  __syncthreads();
  if (tid < PA_SUMS) {
    double s = 0;
    for (int wv = 0; wv < PA_THREADS / 32; ++wv) s += (double)sh.warp_part[wv][tid];
    sh.part[parity][tid] = s;
  }
  if (GRID) {
    double* mine = gpart + ((size_t)parity * nblk + blk) * PA_SUMS;
    if (tid < PA_SUMS) mine[tid] = sh.part[parity][tid];
    __threadfence();
    cg::this_grid().sync();  // every CTA's partial is in global memory
    // 8 thread groups stride over the CTAs, then the groups are added in a fixed order: the same sum on every CTA
    const int k = tid & 63, g = tid >> 6;
    if (k < PA_SUMS) {
      const double* src = gpart + (size_t)parity * nblk * PA_SUMS + k;
      double a = 0;
      for (int b = g; b < nblk; b += PA_THREADS / 64) a += __ldcg(src + (size_t)b * PA_SUMS);
      reinterpret_cast<double*>(sh.warp_part)[g * PA_SUMS + k] = a;  // warp_part is free again (16 x 46 floats >= 8 x 46 doubles)
    }
    __syncthreads();
    if (tid < PA_SUMS) {
      double t = 0;
      for (int g2 = 0; g2 < PA_THREADS / 64; ++g2) t += reinterpret_cast<double*>(sh.warp_part)[g2 * PA_SUMS + tid];
      sh.tot[tid] = t;
    }
    __syncthreads();
  } else {
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();  // every CTA's part[parity] is complete and visible
    if (tid < PA_SUMS) {
      double s = 0;
      for (unsigned rk = 0; rk < PA_CLUSTER; ++rk) {
        const PaShared* peer = cluster.map_shared_rank(&sh, rk);
        s += peer->part[parity][tid];
      }
      sh.tot[tid] = s;
    }
    __syncthreads();
  }
  // part[parity] is rewritten two sweeps from now; the barrier of the next sweep lies in between
}

template <bool GRID>
__device__ __forceinline__ void pose_align_body(const PaProblem& p, PaOut* __restrict__ out, double* __restrict__ gpart) {
  __shared__ PaShared sh;
  const int tid = threadIdx.x;
  const bool first_cta = GRID ? blockIdx.x == 0 : cg::this_cluster().block_rank() == 0;
  if (tid == 0) {
    for (int i = 0; i < 12; ++i) sh.T[i] = p.T0[i];
    sh.ab_eps[0] = sh.ab_eps[1] = 0.0;
    sh.lambda = p.lambda0;
    sh.converged = 0;
    sh.iteration = 0;
  }
  __syncthreads();
  int parity = 0;
  // energy (+ prior) of the sums in sh.tot at the current affine state
  auto total_energy = [&]() {
    const double a = p.tgt.ab0[0] + sh.ab_eps[0], b = p.tgt.ab0[1] + sh.ab_eps[1];
    return sh.tot[44] + 0.5 * (a * p.ab_reg[0] * a + b * p.ab_reg[1] * b);  // AffineBrightnessPrior::energyTerm
  };
  auto adopt_system = [&]() {  // the sums of the last sweep become the reference's system_ (linearize())
    int idx = 0;
    for (int a = 0; a < 8; ++a)
      for (int b = a; b < 8; ++b) {
        sh.Hc[a * 8 + b] = sh.Hc[b * 8 + a] = sh.tot[idx];
        ++idx;
      }
    for (int a = 0; a < 8; ++a) sh.bc[a] = sh.tot[36 + a];
  };
  // result.energy = problem.calculateEnergy(); the first linearize() sees the same samples
  pa_sweep<GRID>(p, sh, parity, gpart);
  parity ^= 1;
  if (tid == 0) {
    sh.energy = total_energy();
    sh.n_valid = (int)llrint(sh.tot[45]);
    adopt_system();
    sh.go = sh.iteration < p.max_it && !sh.converged && sh.n_valid > 0;
  }
  __syncthreads();
  while (sh.go) {
    if (tid == 0) {
      // calculateStep(lambda), eigen_pose_alignment.cpp:194-206: priors were added to system_ by linearize() (:178-190)
      double H[64], b[8], step[8];
      for (int i = 0; i < 64; ++i) H[i] = sh.Hc[i];
      for (int i = 0; i < 8; ++i) b[i] = sh.bc[i];
      const double ab[2] = {p.tgt.ab0[0] + sh.ab_eps[0], p.tgt.ab0[1] + sh.ab_eps[1]};
      for (int k = 0; k < 2; ++k) {
        H[(6 + k) * 8 + 6 + k] += p.ab_reg[k];
        b[6 + k] += p.ab_reg[k] * ab[k];
      }
      solve8(H, b, sh.lambda, step);
      for (int i = 0; i < 12; ++i) sh.T_old[i] = sh.T[i];
      sh.ab_old[0] = sh.ab_eps[0];
      sh.ab_old[1] = sh.ab_eps[1];
      left_increment(step, sh.T);
      sh.ab_eps[0] -= step[6];
      sh.ab_eps[1] -= step[7];
      for (int i = 0; i < 8; ++i) sh.step[i] = step[i];
    }
    __syncthreads();
    pa_sweep<GRID>(p, sh, parity, gpart);  // calculateEnergy() at the trial state (+ the system, should it be accepted)
    parity ^= 1;
    if (tid == 0) {
      const double next_energy = total_energy();
      const int next_n = (int)llrint(sh.tot[45]);
      bool stop_now = false;
      if (first_cta && sh.iteration < PA_TRACE) {
        out->trace_energy[sh.iteration] = next_energy;
        out->trace_lambda[sh.iteration] = sh.lambda;
        out->trace_accept[sh.iteration] = next_n != 0 && next_energy < sh.energy;
      }
      if (next_n == 0) {  // rejectStep(); break  (lm.hpp:95-98)
        for (int i = 0; i < 12; ++i) sh.T[i] = sh.T_old[i];
        sh.ab_eps[0] = sh.ab_old[0];
        sh.ab_eps[1] = sh.ab_old[1];
        stop_now = true;
      } else {
        if (fabs(sh.energy - next_energy) / sh.energy < p.ftol) sh.converged = 1;  // before the accept test (Q7)
        if (next_energy < sh.energy) {
          // acceptStep(), :208-213: (|ab0 + old eps|^2, |step|^2)
          const double a0 = p.tgt.ab0[0] + sh.ab_old[0], a1 = p.tgt.ab0[1] + sh.ab_old[1];
          double sq = 0;
          for (int i = 0; i < 8; ++i) sq += sh.step[i] * sh.step[i];
          if (sq < p.ptol * (a0 * a0 + a1 * a1 + p.ptol)) sh.converged = 1;
          sh.energy = next_energy;
          sh.n_valid = next_n;
          sh.lambda /= p.dec;
          adopt_system();  // linear_system_valid = false -> linearize() at the accepted state
        } else {
          for (int i = 0; i < 12; ++i) sh.T[i] = sh.T_old[i];  // rejectStep(): the previous system stays valid
          sh.ab_eps[0] = sh.ab_old[0];
          sh.ab_eps[1] = sh.ab_old[1];
          sh.lambda *= p.inc;
        }
      }
      sh.iteration += 1;
      sh.go = !stop_now && sh.iteration < p.max_it && !sh.converged && sh.n_valid > 0;
    }
    __syncthreads();
  }
  if (first_cta && tid == 0) {
    out->energy = sh.energy;
    out->n_valid = sh.n_valid;
    out->converged = sh.converged;
    out->iterations = sh.iteration;
    for (int i = 0; i < 12; ++i) out->T[i] = sh.T[i];
    out->ab_eps[0] = sh.ab_eps[0];
    out->ab_eps[1] = sh.ab_eps[1];
    for (int i = 0; i < 64; ++i) out->H[i] = sh.Hc[i];
    out->H[6 * 8 + 6] += p.ab_reg[0];  // problem.hessian() = system_.H incl. the affine prior (:178-185)
    out->H[7 * 8 + 7] += p.ab_reg[1];
  }
  if (!GRID) cg::this_cluster().sync();  // no CTA may exit while a peer could still read its shared memory
}

__global__ void __cluster_dims__(PA_CLUSTER, 1, 1) __launch_bounds__(PA_THREADS)
    k_pose_align(const PaProblem p, PaOut* __restrict__ out) {
  pose_align_body<false>(p, out, nullptr);
}

// whole-chip variant: cooperative launch, one CTA per SM, partial sums through global memory (gpart: [2][grid][PA_SUMS])
__global__ void __launch_bounds__(PA_THREADS) k_pose_align_grid(const PaProblem p, PaOut* __restrict__ out,
                                                                double* __restrict__ gpart) {
  pose_align_body<true>(p, out, gpart);
}

// ---- depth map -> landmarks on the device (local_frame.hpp:367-392), order preserved -------------------------------
__device__ __forceinline__ bool dm_keep(const float* __restrict__ ids, const float* __restrict__ wgt, int x, int y, int W,
                                        float& idepth) {
  const float w = wgt[y * W + x];
  idepth = w > 0.f ? ids[y * W + x] / w : 0.f;
  return w > 0.f && idepth >= 1e-6f;
}
__global__ void k_dm_count(const float* __restrict__ ids, const float* __restrict__ wgt, int W, int H, int* __restrict__ row_count) {
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y >= H) return;
  int cnt = 0;
  if (y >= 4 && y < H - 4)
    for (int x0 = 0; x0 < W; x0 += 32) {
      const int x = x0 + lane;
      float id;
      const bool k = x >= 4 && x < W - 4 && dm_keep(ids, wgt, x, y, W, id);
      cnt += __popc(__ballot_sync(FULL, k));
    }
  if (lane == 0) row_count[y] = cnt;
}
__global__ void k_dm_scan(int* __restrict__ row_count, int H, int* __restrict__ total) {  // exclusive scan, one thread
  int s = 0;
  for (int y = 0; y < H; ++y) {
    const int c = row_count[y];
    row_count[y] = s;
    s += c;
  }
  *total = s;
}
__global__ void k_dm_write(const float* __restrict__ ids, const float* __restrict__ wgt, const float* __restrict__ img3,
                           int W, int H, const int* __restrict__ row_off, int cap, float4* __restrict__ lm) {
  const int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (y < 4 || y >= H - 4) return;
  int off = row_off[y];
  for (int x0 = 0; x0 < W; x0 += 32) {
    const int x = x0 + lane;
    float id = 0.f;
    const bool k = x >= 4 && x < W - 4 && dm_keep(ids, wgt, x, y, W, id);
    const unsigned b = __ballot_sync(FULL, k);
    const int pos = off + __popc(b & ((1u << lane) - 1u));
    if (k && pos < cap) lm[pos] = make_float4((float)x, (float)y, id, img3[3 * (y * W + x)]);
    off += __popc(b);
  }
}
__global__ void k_pack_lm(const float* __restrict__ xy, const float* __restrict__ idepth, const float* __restrict__ patch,
                          int n, float4* __restrict__ lm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lm[i] = make_float4(xy[2 * i], xy[2 * i + 1], idepth[i], patch[i]);
}

}  // namespace

struct dpa_handle {
  dpa_config cfg;
  cudaStream_t stream = nullptr;
  std::string err;
  float4* lm = nullptr;
  int n = 0;
  float4* img = nullptr;
  uint8_t* mask = nullptr;
  bool mask_all = true;
  float* stage = nullptr;       // device staging: image {I,dx,dy} or landmark arrays / depth map
  float* stage2 = nullptr;      // depth-map accumulators
  int* rows = nullptr;          // [max_height + 1]
  PaFrame ref{}, tgt{};
  bool have_ref = false, have_tgt = false;
  PaOut* out_dev = nullptr;
  PaOut* out_h = nullptr;       // pinned
  int* total_h = nullptr;       // pinned
  double* gpart = nullptr;      // whole-chip variant: [2][SMs][PA_SUMS] partial sums (lazily)
  int sm_count = 0;
  int grid_min_points = 32768;  // from this many landmarks on the cooperative whole-chip kernel is used (dpa_set_grid_threshold)
};

namespace {
int pfail(dpa_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}
#define PCK(expr)                                                                          \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) return pfail(h, DPBA_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
  } while (0)
#define PREQ(cond, msg)                                     \
  do {                                                      \
    if (!(cond)) return pfail(h, DPBA_E_INVALID, (msg));    \
  } while (0)

void fill_frame(PaFrame& f, const double* T, double exposure, const double* ab, const double* intr, int W, int H) {
  memcpy(f.T, T, sizeof(f.T));
  f.exposure = exposure;
  f.ab0[0] = ab[0];
  f.ab0[1] = ab[1];
  memcpy(f.intr, intr, sizeof(f.intr));
  f.W = W;
  f.H = H;
}
void inv34(const double* T, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o[i * 4 + j] = T[j * 4 + i];
  for (int i = 0; i < 3; ++i) o[i * 4 + 3] = -(o[i * 4] * T[3] + o[i * 4 + 1] * T[7] + o[i * 4 + 2] * T[11]);
}
void mul34(const double* a, const double* b, double* o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = (j == 3) ? a[i * 4 + 3] : 0.0;
      for (int k = 0; k < 3; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      o[i * 4 + j] = s;
    }
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" {

const char* dpa_last_error(const dpa_handle* h) { return h ? h->err.c_str() : "null handle"; }
void* dpa_stream(dpa_handle* h) { return h ? (void*)h->stream : nullptr; }

int dpa_create(const dpa_config* cfg, dpa_handle** out) {
  if (!cfg || !out) return DPBA_E_INVALID;
  *out = nullptr;
  if (cfg->max_points < 1 || cfg->max_width < 16 || cfg->max_height < 16) return DPBA_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || cfg->device < 0 || cfg->device >= ndev) {
    fprintf(stderr, "dpa_create: CUDA device %d not available (%d devices); there is no CPU fallback\n", cfg->device, ndev);
    return DPBA_E_CUDA;
  }
  dpa_handle* h = new dpa_handle();
  h->cfg = *cfg;
  const size_t npx = (size_t)cfg->max_width * cfg->max_height;
  bool ok = cudaSetDevice(cfg->device) == cudaSuccess;
  ok = ok && cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
  ok = ok && cudaMalloc(&h->lm, sizeof(float4) * (size_t)cfg->max_points) == cudaSuccess;
  ok = ok && cudaMalloc(&h->img, npx * 2 * sizeof(float4)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->mask, npx) == cudaSuccess;
  ok = ok && cudaMalloc(&h->stage, std::max(npx * 3, (size_t)cfg->max_points * 4) * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->stage2, npx * 2 * sizeof(float)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->rows, (cfg->max_height + 1) * sizeof(int)) == cudaSuccess;
  ok = ok && cudaMalloc(&h->out_dev, sizeof(PaOut)) == cudaSuccess;
  ok = ok && cudaMallocHost(&h->out_h, sizeof(PaOut)) == cudaSuccess;
  ok = ok && cudaMallocHost(&h->total_h, sizeof(int)) == cudaSuccess;
  if (!ok) {
    fprintf(stderr, "dpa_create: %s\n", cudaGetErrorString(cudaGetLastError()));
    dpa_destroy(h);
    return DPBA_E_CUDA;
  }
  *out = h;
  return DPBA_SUCCESS;
}

int dpa_destroy(dpa_handle* h) {
  if (!h) return DPBA_E_INVALID;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->lm);
  cudaFree(h->img);
  cudaFree(h->mask);
  cudaFree(h->stage);
  cudaFree(h->stage2);
  cudaFree(h->rows);
  cudaFree(h->out_dev);
  cudaFreeHost(h->out_h);
  cudaFreeHost(h->total_h);
  if (h->stream) cudaStreamDestroy(h->stream);
  cudaFree(h->gpart);
  delete h;
  return DPBA_SUCCESS;
}

int dpa_num_landmarks(const dpa_handle* h) { return h ? h->n : DPBA_E_INVALID; }

int dpa_set_reference_landmarks(dpa_handle* h, int32_t n, const float* xy, const float* idepth, const float* patch,
                                const double T[12], double exposure, const double ab[2], const double intr[4],
                                int32_t width, int32_t height) {
  PREQ(h, "null handle");
  PREQ(n >= 0 && (n == 0 || (xy && idepth && patch)) && T && ab && intr, "null argument");
  PREQ(exposure > 0 && width >= 16 && height >= 16, "bad frame");
  if (n > h->cfg.max_points) return pfail(h, DPBA_E_CAPACITY, "too many landmarks for this handle");
  if (n) {
    PCK(cudaMemcpyAsync(h->stage, xy, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    PCK(cudaMemcpyAsync(h->stage + 2 * (size_t)n, idepth, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    PCK(cudaMemcpyAsync(h->stage + 3 * (size_t)n, patch, sizeof(float) * n, cudaMemcpyHostToDevice, h->stream));
    pba::add_launches(1);
    k_pack_lm<<<(n + 255) / 256, 256, 0, h->stream>>>(h->stage, h->stage + 2 * (size_t)n, h->stage + 3 * (size_t)n, n, h->lm);
    PCK(cudaGetLastError());
    PCK(cudaStreamSynchronize(h->stream));  // the caller's buffers may be pageable and are free again on return
  }
  h->n = n;
  fill_frame(h->ref, T, exposure, ab, intr, width, height);
  h->have_ref = true;
  return DPBA_SUCCESS;
}

int dpa_set_reference_depth_map(dpa_handle* h, const float* image, const float* idepth_sum, const float* weight,
                                const double T[12], double exposure, const double ab[2], const double intr[4],
                                int32_t width, int32_t height) {
  PREQ(h, "null handle");
  PREQ(image && idepth_sum && weight && T && ab && intr, "null argument");
  PREQ(exposure > 0 && width >= 16 && height >= 16 && width <= h->cfg.max_width && height <= h->cfg.max_height, "bad frame");
  const size_t npx = (size_t)width * height;
  PCK(cudaMemcpyAsync(h->stage, image, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  PCK(cudaMemcpyAsync(h->stage2, idepth_sum, npx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  PCK(cudaMemcpyAsync(h->stage2 + npx, weight, npx * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  const int rows_per_cta = 8;
  pba::add_launches(3);
  k_dm_count<<<(height + rows_per_cta - 1) / rows_per_cta, 32 * rows_per_cta, 0, h->stream>>>(h->stage2, h->stage2 + npx, width, height, h->rows);
  k_dm_scan<<<1, 1, 0, h->stream>>>(h->rows, height, h->rows + height);
  k_dm_write<<<(height + rows_per_cta - 1) / rows_per_cta, 32 * rows_per_cta, 0, h->stream>>>(
      h->stage2, h->stage2 + npx, h->stage, width, height, h->rows, h->cfg.max_points, h->lm);
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(h->total_h, h->rows + height, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  PCK(cudaStreamSynchronize(h->stream));
  if (*h->total_h > h->cfg.max_points) return pfail(h, DPBA_E_CAPACITY, "depth map holds more landmarks than max_points");
  h->n = *h->total_h;
  fill_frame(h->ref, T, exposure, ab, intr, width, height);
  h->have_ref = true;
  return h->n;
}

int dpa_get_reference_landmarks(dpa_handle* h, int32_t n, float* xy, float* idepth, float* patch) {
  PREQ(h, "null handle");
  PREQ(n >= 0 && n <= h->n, "n exceeds the landmark count");
  if (!n) return DPBA_SUCCESS;
  std::vector<float4> tmp(n);
  PCK(cudaMemcpyAsync(tmp.data(), h->lm, sizeof(float4) * n, cudaMemcpyDeviceToHost, h->stream));
  PCK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < n; ++i) {
    if (xy) xy[2 * i] = tmp[i].x, xy[2 * i + 1] = tmp[i].y;
    if (idepth) idepth[i] = tmp[i].z;
    if (patch) patch[i] = tmp[i].w;
  }
  return DPBA_SUCCESS;
}

int dpa_set_target(dpa_handle* h, const float* image, const uint8_t* mask, const double T[12], double exposure,
                   const double ab[2], const double intr[4], int32_t width, int32_t height) {
  PREQ(h, "null handle");
  PREQ(image && T && ab && intr, "null argument");
  PREQ(exposure > 0 && width >= 16 && height >= 16 && width <= h->cfg.max_width && height <= h->cfg.max_height, "bad frame");
  const size_t npx = (size_t)width * height;
  PCK(cudaMemcpyAsync(h->stage, image, npx * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
  pba::launch_pack_image(h->stage, h->img, (int)npx, width, h->stream);
  PCK(cudaGetLastError());
  h->mask_all = true;
  if (mask) {
    h->mask_all = memchr(mask, 0, npx) == nullptr;
    if (!h->mask_all) PCK(cudaMemcpyAsync(h->mask, mask, npx, cudaMemcpyHostToDevice, h->stream));
  }
  PCK(cudaStreamSynchronize(h->stream));  // pageable caller buffers are free again on return
  fill_frame(h->tgt, T, exposure, ab, intr, width, height);
  h->have_tgt = true;
  return DPBA_SUCCESS;
}

int dpa_set_grid_threshold(dpa_handle* h, int32_t min_points) {
  PREQ(h, "null handle");
  h->grid_min_points = min_points < 0 ? 0x7fffffff : min_points;
  return DPBA_SUCCESS;
}

int dpa_get_trace(dpa_handle* h, int32_t capacity, double* energies, double* lambdas, int32_t* accepted) {
  PREQ(h, "null handle");
  const int n = std::min(std::min(h->out_h->iterations, (int)PA_TRACE), (int)capacity);
  for (int i = 0; i < n; ++i) {
    if (energies) energies[i] = h->out_h->trace_energy[i];
    if (lambdas) lambdas[i] = h->out_h->trace_lambda[i];
    if (accepted) accepted[i] = h->out_h->trace_accept[i];
  }
  return n;
}

int dpa_solve(dpa_handle* h, const dpa_options* o, const double* prior_rotation, dpa_result* res) {
  PREQ(h, "null handle");
  PREQ(o && res, "null argument");
  PREQ(h->have_ref && h->have_tgt, "push the reference and the target frame first");
  PREQ(o->max_num_iterations >= 0 && o->initial_trust_region_radius > 0, "bad options");
  PaProblem p;
  memset(&p, 0, sizeof(p));
  p.lm = h->lm;
  p.n = h->n;
  p.img = h->img;
  p.mask = h->mask_all ? nullptr : h->mask;
  p.ref = h->ref;
  p.tgt = h->tgt;
  double Ti[12];
  inv34(h->tgt.T, Ti);
  mul34(Ti, h->ref.T, p.T0);  // t_t_r = T_w_target^-1 T_w_reference (:307-308)
  if (prior_rotation)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) p.T0[i * 4 + j] = prior_rotation[i * 3 + j];  // :309-311
  p.max_it = o->max_num_iterations;
  p.lambda0 = 1.0 / o->initial_trust_region_radius;
  p.ftol = o->function_tolerance;
  p.ptol = o->parameter_tolerance;
  p.sigma = o->sigma_huber_loss;
  p.ab_reg[0] = o->affine_brightness_regularizer[0];
  p.ab_reg[1] = o->affine_brightness_regularizer[1];
  p.dec = o->regularizer_decrease_on_accept;
  p.inc = o->regularizer_increase_on_reject;
  pba::add_launches(1);
  if (h->n >= h->grid_min_points) {
    // dense depth map: one CTA per SM (cooperative launch), grid barrier between the sweeps
    if (!h->sm_count) {
      int dev = 0, coop = 0;
      PCK(cudaGetDevice(&dev));
      PCK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, dev));
      PCK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
      PREQ(coop, "cooperative launch not supported");
      PCK(cudaMalloc(&h->gpart, 2 * (size_t)h->sm_count * PA_SUMS * sizeof(double)));
    }
    int grid = std::min(h->sm_count, (h->n + PA_THREADS - 1) / PA_THREADS);
    void* args[] = {(void*)&p, (void*)&h->out_dev, (void*)&h->gpart};
    PCK(cudaLaunchCooperativeKernel((const void*)k_pose_align_grid, dim3(grid), dim3(PA_THREADS), args, 0, h->stream));
  } else {
    k_pose_align<<<PA_CLUSTER, PA_THREADS, 0, h->stream>>>(p, h->out_dev);
  }
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(h->out_h, h->out_dev, sizeof(PaOut), cudaMemcpyDeviceToHost, h->stream));
  PCK(cudaStreamSynchronize(h->stream));
  const PaOut& r = *h->out_h;
  res->energy = r.energy;
  res->number_of_valid_residuals = r.n_valid;
  res->converged = r.converged;
  res->iterations = r.iterations;
  res->rmse = r.n_valid > 0 ? sqrt(r.energy / r.n_valid / 1.0) : INFINITY;
  memcpy(res->T_target_reference, r.T, sizeof(r.T));
  double inv[12];
  inv34(r.T, inv);
  mul34(h->ref.T, inv, res->T_world_target);  // :325
  res->affine_brightness_eps[0] = r.ab_eps[0];
  res->affine_brightness_eps[1] = r.ab_eps[1];
  memcpy(res->hessian, r.H, sizeof(r.H));
  return DPBA_SUCCESS;
}

}  // extern "C"
#pragma GCC visibility pop

// ------------------------------------------------------------------------------------------------------------------------
// Mean-square optical flow of the reference landmarks under a relative pose (keyframe decision of the tracker):
// calculateMeanSquareOpticalFlow, src/tracker/tracker/src/monocular_tracker.cpp:104-133, evaluated by the tracker on
// level 0 of the reference depth map with the pose the aligner just returned, and once more with the rotation removed
// (:474-480).  The landmark list compacted by dpa_set_reference_depth_map is exactly the pixel set the reference loops
// over (4-px border, weight > 0, idepth >= 1e-6), so this is one grid-stride reduction over data that is already
// resident: fp32 per landmark (optical_flow_body.h), fp64 sums, one (sum, count) pair per CTA added up on the host in
// CTA order (deterministic).  STATUS: written after the round-1 GPU minutes were spent; the per-landmark body runs on the
// CPU against the oracle (tests/test_kernel_emulation.py), the kernel itself has not run on hardware yet.
// ------------------------------------------------------------------------------------------------------------------------
#include "optical_flow_body.h"

namespace {
constexpr int OF_THREADS = 256;

__global__ void __launch_bounds__(OF_THREADS) k_optical_flow(const float4* __restrict__ lm, int n, pba::FlowConst c,
                                                             double2* __restrict__ partial) {
  double sum = 0.0, cnt = 0.0;
  for (int i = blockIdx.x * OF_THREADS + threadIdx.x; i < n; i += gridDim.x * OF_THREADS) {
    float sq;
    if (pba::flow_term(c, lm[i], sq)) {
      sum += (double)sq;
      cnt += 1.0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ double2 red[OF_THREADS / 32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = make_double2(sum, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    double2 t = red[0];
    for (int k = 1; k < OF_THREADS / 32; ++k) {
      t.x += red[k].x;
      t.y += red[k].y;
    }
    partial[blockIdx.x] = t;
  }
}
}  // namespace

#pragma GCC visibility push(default)
extern "C" int dpa_mean_square_optical_flow(dpa_handle* h, const double T_target_reference[12], double* flow,
                                            int32_t* n_used) {
  PREQ(h, "null handle");
  PREQ(T_target_reference && flow, "null argument");
  PREQ(h->have_ref, "dpa_set_reference_* first");
  const pba::FlowConst c = pba::make_flow_const(T_target_reference, h->ref.intr, h->ref.W, h->ref.H);
  double sum = 0.0, cnt = 0.0;
  if (h->n > 0) {
    const int ctas = std::min(2 * pba::sm_count(), (h->n + OF_THREADS - 1) / OF_THREADS);
    double2* partial = reinterpret_cast<double2*>(h->stage2);  // depth-map accumulators are consumed by now
    PREQ((size_t)ctas * sizeof(double2) <= (size_t)2 * h->cfg.max_width * h->cfg.max_height * sizeof(float), "staging too small");
    pba::add_launches(1);
    k_optical_flow<<<ctas, OF_THREADS, 0, h->stream>>>(h->lm, h->n, c, partial);
    PCK(cudaGetLastError());
    std::vector<double2> host(ctas);
    PCK(cudaMemcpyAsync(host.data(), partial, sizeof(double2) * ctas, cudaMemcpyDeviceToHost, h->stream));
    PCK(cudaStreamSynchronize(h->stream));
    for (const double2& p : host) {
      sum += p.x;
      cnt += p.y;
    }
  }
  *flow = sqrt(sum / cnt);  // 0 / 0 = NaN when nothing reprojects, as in the reference (:132)
  if (n_used) *n_used = (int32_t)cnt;
  return DPBA_SUCCESS;
}
#pragma GCC visibility pop
