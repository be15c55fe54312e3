// Hand-written sm_100a kernels of the photometric bundle-adjustment hot path.
//
// Reference loops replaced (paths relative to /root/reference/src/, "PBA/" =
// energy/problems/internal/energy/problems/photometric_bundle_adjustment/):
//   K1/K2  evaluateJacobians                              PBA/evaluate_jacobians.hpp:20-202
//   K3     evaluateLinearSystemPosePose[Block]            PBA/hessian_block_evaluation.hpp:38-164
//   K4     evaluateLinearSystemPoseDepthSchurComplement   PBA/hessian_block_evaluation.hpp:169-236
//   K5     calculateIdepths                               PBA/hessian_block_evaluation.hpp:238-263
//   K6     firstEstimateJacobians_                        PBA/first_estimate_jacobians.hpp:14-71
//   accept/reject, changeResidualStatuses, calculateLandmarksEnergy
//                                                         PBA/eigen_photometric_bundle_adjustment_problem.hpp:20-35,93-144,366-402
//
// Thread mapping of the sweeps: 8 lanes per patch-residual (lane = pattern pixel), 4 patch-residuals per warp;
// the 8-pixel sums are warp-shuffle reductions.  Arithmetic is fp32 per residual, fp64 for every sum that
// crosses landmarks.  The ROI / depth / mask predicates that decide connection statuses are evaluated with
// explicitly rounded, non-contracted fp32 operations (__fmul_rn/__fadd_rn/__fdiv_rn) so the fp32 CPU oracle
// reproduces them bit for bit.
#include <math.h>
#include <stdio.h>

#include <atomic>

#include "pba_internal.h"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int K_OK = 0, K_OUTLIER = 1, K_OOB = 3;
constexpr int LM_MARG = 1, LM_TO_MARG = 2, LM_OUTLIER = 4, LM_ILL = 8;

// 8-point DSO pattern, common/pattern/include/common/pattern/pattern.hpp:22-33, packed as nibbles (+2)
__device__ __forceinline__ float pat_x(int i) { return (float)((0x21420312u >> (4 * i)) & 15u) - 2.f; }
__device__ __forceinline__ float pat_y(int i) { return (float)((0x01222334u >> (4 * i)) & 15u) - 2.f; }

// camera_model_base.hpp:52-60 (border 4 px) -- NaN compares false
__device__ __forceinline__ bool in_roi(float x, float y, float xmax, float ymax) {
  return x >= 4.f && y >= 4.f && x <= xmax && y <= ymax;
}
// camera_model_base.hpp:68-74
__device__ __forceinline__ bool valid_idepth(float r) { return r > -1e-4f && r < 1010.f; }

// rows of a 3x4 matrix applied to [u, v, 1, rho] with the reference's association
// (A[:, :2] uv) + (A[:,2] + A[:,3] rho)   (camera_reproject.hpp:283-284,323-325), no FMA contraction
__device__ __forceinline__ float row_apply(const float* a, float u, float v, float rho) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], u), __fmul_rn(a[1], v)), __fadd_rn(a[2], __fmul_rn(a[3], rho)));
}

__device__ __forceinline__ bool group_all(bool p, int lane) {
  unsigned b = __ballot_sync(FULL, p);
  return ((b >> (lane & 24)) & 0xffu) == 0xffu;
}

__device__ __forceinline__ float group_sum(float v) {
  v += __shfl_xor_sync(FULL, v, 4);
  v += __shfl_xor_sync(FULL, v, 2);
  v += __shfl_xor_sync(FULL, v, 1);
  return v;
}

// 8 values per lane, 8 lanes per group -> lane px returns sum over the group of v[px]  (7 shuffles)
__device__ __forceinline__ float group_transpose_reduce(const float (&v)[8], int px) {
  float a[4], b[2];
  bool hi = px & 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    float send = hi ? v[k] : v[k + 4];
    float keep = hi ? v[k + 4] : v[k];
    a[k] = keep + __shfl_xor_sync(FULL, send, 4);
  }
  hi = px & 2;
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    float send = hi ? a[k] : a[k + 2];
    float keep = hi ? a[k + 2] : a[k];
    b[k] = keep + __shfl_xor_sync(FULL, send, 2);
  }
  hi = px & 1;
  float send = hi ? b[0] : b[1];
  float keep = hi ? b[1] : b[0];
  return keep + __shfl_xor_sync(FULL, send, 1);
}

// one halving step of a warp-wide transpose-reduce: N values per lane in, ceil(N/2) out
template <int N, int XOR>
__device__ __forceinline__ void tr_step(const float (&in)[N], float (&out)[(N + 1) / 2], int lane) {
  constexpr int LO = (N + 1) / 2;
  constexpr int HI = N - LO;
  const bool hi = lane & XOR;
#pragma unroll
  for (int k = 0; k < LO; ++k) {
    float upper = (k < HI) ? in[(k < HI) ? LO + k : 0] : 0.f;
    float send = hi ? in[k] : upper;
    float keep = hi ? upper : in[k];
    out[k] = keep + __shfl_xor_sync(FULL, send, XOR);
  }
}

// logical frame slot -> physical storage slot (dpba_remove_frame frees a physical slot without moving data)
__device__ __forceinline__ int lm_index(const WindowDev& w, int f, int l) { return w.phys[f] * w.max_pts + l; }
__device__ __forceinline__ size_t res_index(const WindowDev& w, int r, int t, int l) {
  return ((size_t)(w.phys[r] * PBA_MAXF + w.phys[t])) * w.max_pts + l;
}

// device-resident LM: kernels of a loop body return immediately once the loop has terminated (mode 1), and the
// linearisation kernels also when the previous linear system is still valid (mode 2)
__device__ __forceinline__ bool lm_skip(const LmCtl* ctl, int mode) {
  if (!ctl || mode == 0) return false;
  if (ctl->done) return true;
  return mode == 2 && ctl->system_valid;
}

struct LandmarkIn {
  float u, v, rho, rho0, patch;
  int flags;
};

struct PixelOut {
  float r;     // residual (0 when not evaluated)
  float g[6];  // dI/d(xi) of T_t_r  (row of d_target_reference_state, evaluate_jacobians.hpp:149-157)
  float d;     // d r / d idepth
  float c;     // corrected reference intensity (affine `a` column)
  float e;     // energy of the patch (identical on the 8 lanes)
  float w;     // huber weight
  bool ok;     // reprojection + mask succeeded for the whole pattern
  bool ev;     // ok && committed status == kOk  -> residual was evaluated
};

// Evaluates one pattern pixel of one patch-residual.  All 8 lanes of a group must call it together.
//   FEJ : first-estimate Jacobians (production)   JAC : evaluate Jacobians
template <bool FEJ, bool JAC>
__device__ __forceinline__ void eval_pixel(const PairConst& pc, const LandmarkIn& lm, const float4* __restrict__ img,
                                           const uint8_t* __restrict__ mask, int W, int H, int status, float sigma,
                                           int huber, int lane, PixelOut& o) {
  const int px = lane & 7;
  const float xmax = (float)(W - 5), ymax = (float)(H - 5);
  const float ur = lm.u + pat_x(px), vr = lm.v + pat_y(px);

  bool ok = valid_idepth(lm.rho) && in_roi(ur, vr, xmax, ymax);
  float tu, tv;
  float qx = 0.f, qy = 0.f, qz = 1.f, rho_j = 0.f;  // point used for the reprojection Jacobians
  if (FEJ || !JAC) {
    // values-only reprojection at the current state (camera_reproject.hpp:270-293)
    const float X = row_apply(pc.A + 0, ur, vr, lm.rho);
    const float Y = row_apply(pc.A + 4, ur, vr, lm.rho);
    const float Z = row_apply(pc.A + 8, ur, vr, lm.rho);
    ok = ok && (Z > 0.f);
    tu = __fdiv_rn(X, Z);
    tv = __fdiv_rn(Y, Z);
    ok = ok && in_roi(tu, tv, xmax, ymax);
    if (FEJ) {
      // reprojection_jacobians_valid of firstEstimateJacobians_ (first_estimate_jacobians.hpp:52-54):
      // the Jacobian variant of reproject() at the linearisation point and the snapshot idepth
      qx = row_apply(pc.M0 + 0, ur, vr, lm.rho0);
      qy = row_apply(pc.M0 + 4, ur, vr, lm.rho0);
      qz = row_apply(pc.M0 + 8, ur, vr, lm.rho0);
      rho_j = lm.rho0;
      bool okj = valid_idepth(lm.rho0) && (qz > 0.f);
      const float u0 = __fdiv_rn(__fadd_rn(__fmul_rn(pc.fx_t, qx), __fmul_rn(pc.cx_t, qz)), qz);
      const float v0 = __fdiv_rn(__fadd_rn(__fmul_rn(pc.fy_t, qy), __fmul_rn(pc.cy_t, qz)), qz);
      okj = okj && in_roi(u0, v0, xmax, ymax);
      ok = ok && okj;
    }
  } else {
    // Jacobian variant at the current state (camera_reproject.hpp:305-367)
    qx = row_apply(pc.M + 0, ur, vr, lm.rho);
    qy = row_apply(pc.M + 4, ur, vr, lm.rho);
    qz = row_apply(pc.M + 8, ur, vr, lm.rho);
    rho_j = lm.rho;
    ok = ok && (qz > 0.f);
    tu = __fdiv_rn(__fadd_rn(__fmul_rn(pc.fx_t, qx), __fmul_rn(pc.cx_t, qz)), qz);
    tv = __fdiv_rn(__fadd_rn(__fmul_rn(pc.fy_t, qy), __fmul_rn(pc.cy_t, qz)), qz);
    ok = ok && in_roi(tu, tv, xmax, ymax);
  }
  ok = group_all(ok, lane);
  // CameraMask::valid<false>: round() + lookup, only meaningful after the ROI test (quirk Q5)
  bool mok = false;
  if (ok) mok = mask[(int)roundf(tv) * W + (int)roundf(tu)] != 0;
  ok = group_all(ok && mok, lane);

  o.ok = ok;
  o.ev = ok && (status == K_OK);
  o.r = 0.f;
  o.d = 0.f;
  o.c = 0.f;
  o.e = 0.f;
  o.w = 1.f;
#pragma unroll
  for (int k = 0; k < 6; ++k) o.g[k] = 0.f;

  float r = 0.f, dIu = 0.f, dIv = 0.f;
  if (o.ev) {
    // interpolateLinear, features/include/features/camera/pixel_map.hpp:20-40
    const int ix = (int)tu, iy = (int)tv;
    const float dx = tu - (float)ix, dy = tv - (float)iy;
    const float dxdy = dx * dy;
    const float w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = 1.f - dx - dy + dxdy;
    const float4* p = img + (size_t)iy * W + ix;
    const float4 t00 = __ldg(p), t01 = __ldg(p + 1), t10 = __ldg(p + W), t11 = __ldg(p + W + 1);
    const float I = w11 * t11.x + w10 * t10.x + w01 * t01.x + w00 * t00.x;
    if (JAC) {
      dIu = w11 * t11.y + w10 * t10.y + w01 * t01.y + w00 * t00.y;
      dIv = w11 * t11.z + w10 * t10.z + w01 * t01.z + w00 * t00.z;
    }
    // r = (I_t - b_t) - s (patch - b_r), evaluate_jacobians.hpp:124-135
    r = (I - pc.b_t) - pc.s * (lm.patch - pc.b_r);
  }
  const float n2 = group_sum(r * r);
  if (o.ev) {
    o.r = r;
    o.e = 0.5f * n2;
    if (huber && n2 > sigma * sigma) {  // evaluate_jacobians.hpp:139-146
      const float nrm = sqrtf(n2);
      o.w = sigma / nrm;
      o.e = sigma * nrm - sigma * sigma * 0.5f;
    }
    if (JAC) {
      // camera_reproject.hpp:339-365
      const float sI = 1.f / qz;
      const float b0 = qx * sI, b1 = qy * sI;
      const float nid = rho_j * sI;
      const float* tt = FEJ ? pc.t0 : pc.tr;
      const float fx = pc.fx_t, fy = pc.fy_t;
      const float du_id = fx * (tt[0] * sI - tt[2] * sI * b0);
      const float dv_id = fy * (tt[1] * sI - tt[2] * sI * b1);
      const float b0b1 = b0 * b1;
      const float gu = dIu * fx, gv = dIv * fy;
      // Jg = dIv * dv/dxi + dIu * du/dxi   (evaluate_jacobians.hpp:149-157)
      o.g[0] = gu * nid;
      o.g[1] = gv * nid;
      o.g[2] = -gu * (nid * b0) - gv * (nid * b1);
      o.g[3] = -gu * b0b1 - gv * (b1 * b1 + 1.f);
      o.g[4] = gu * (b0 * b0 + 1.f) + gv * b0b1;
      o.g[5] = -gu * b1 + gv * b0;
      o.d = dIu * du_id + dIv * dv_id;  // evaluate_jacobians.hpp:165-174
      // corrected_reference_intensities: FEJ -> landmark.corrected_intensities (last target wins, Q1),
      // else s (patch - b_r)   (evaluate_jacobians.hpp:96,103-106)
      o.c = FEJ ? pc.s0_last * (lm.patch - pc.b_r0) : pc.s * (lm.patch - pc.b_r);
    }
  }
}

__device__ __forceinline__ LandmarkIn load_landmark(const WindowDev& w, int gl, int px) {
  LandmarkIn lm;
  const float2 uv = w.uv[gl];
  lm.u = uv.x;
  lm.v = uv.y;
  lm.rho = w.idepth[gl] + w.idepth_step[gl];
  lm.rho0 = w.idepth_fej[gl];
  lm.patch = w.patch[(size_t)gl * 8 + px];
  lm.flags = w.flags[gl];
  return lm;
}

// ------------------------------------------------------------------------------------------------
// K2: residual-only sweep + energy reduction.  grid = (chunks of 32 landmarks, ordered pairs)
// ------------------------------------------------------------------------------------------------
template <bool FEJ>
__global__ void __launch_bounds__(256) k_residual_sweep(const __grid_constant__ WindowDev w, float sigma, int huber,
                                                        double* __restrict__ scal, const LmCtl* __restrict__ ctl,
                                                        int ctl_mode) {
  if (lm_skip(ctl, ctl_mode)) return;
  __shared__ PairConst pcs;
  __shared__ float s_e[8];
  __shared__ int s_n[8];
  const int N = w.n_frames;
  const int r = blockIdx.y / (N - 1);
  int t = blockIdx.y % (N - 1);
  t += (t >= r);
  const int M = w.n_lm[r];
  if ((int)blockIdx.x * 32 >= M) return;
  for (int i = threadIdx.x; i < (int)(sizeof(PairConst) / 4); i += blockDim.x)
    ((float*)&pcs)[i] = ((const float*)&w.pairs[r * PBA_MAXF + t])[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int l = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int px = lane & 7;
  float e_acc = 0.f;
  int n_acc = 0;
  const bool inb = l < M;
  const int gl = lm_index(w, r, inb ? l : 0);
  LandmarkIn lm = load_landmark(w, gl, px);
  const bool skip = !inb || ((lm.flags & LM_MARG) && !(lm.flags & LM_TO_MARG));  // evaluate_jacobians.hpp:83
  const size_t res = res_index(w, r, t, inb ? l : 0);
  const int status = skip ? K_OUTLIER : w.status[res];
  if (skip) lm.rho = -1.f;  // forces !ok without touching memory
  PixelOut o;
  eval_pixel<FEJ, false>(pcs, lm, w.img[t], w.mask[t], w.W, w.H, status, sigma, huber, lane, o);
  if (!skip && px == 0) {
    if (!o.ok) w.cand[res] = K_OOB;   // evaluate_jacobians.hpp:111-113
    else if (o.ev) w.cand[res] = K_OK;  // :115
    w.energy[res] = o.e;
    if (!(lm.flags & LM_MARG)) {  // calculateLandmarksEnergy, problem.hpp:124-133
      e_acc = o.e;
      n_acc = o.e > 0.f;
    }
  }
  // block reduction -> one fp64 atomic per block
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    e_acc += __shfl_xor_sync(FULL, e_acc, s);
    n_acc += __shfl_xor_sync(FULL, n_acc, s);
  }
  if (lane == 0) {
    s_e[warp] = e_acc;
    s_n[warp] = n_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double e = 0;
    int n = 0;
    for (int i = 0; i < 8; ++i) {
      e += (double)s_e[i];
      n += s_n[i];
    }
    if (n | (e != 0)) {
      atomicAdd(&scal[0], e);
      atomicAdd(&scal[1], (double)n);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K1 (reference-surface mode): materialise every ResidualPoint.  Same grid as K2.
// Per patch-residual: 146 floats written (r[8], J_ref[8x8], J_tgt[8x8], d_idepth[8], w, e) + statuses.
// ------------------------------------------------------------------------------------------------
template <bool FEJ>
__global__ void __launch_bounds__(256) k_materialise_sweep(const __grid_constant__ WindowDev w, float sigma, int huber) {
  __shared__ PairConst pcs;
  const int N = w.n_frames;
  const int r = blockIdx.y / (N - 1);
  int t = blockIdx.y % (N - 1);
  t += (t >= r);
  const int M = w.n_lm[r];
  if ((int)blockIdx.x * 32 >= M) return;
  for (int i = threadIdx.x; i < (int)(sizeof(PairConst) / 4); i += blockDim.x)
    ((float*)&pcs)[i] = ((const float*)&w.pairs[r * PBA_MAXF + t])[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int l = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int px = lane & 7;
  const bool inb = l < M;
  const int gl = lm_index(w, r, inb ? l : 0);
  LandmarkIn lm = load_landmark(w, gl, px);
  const bool skip = !inb || ((lm.flags & LM_MARG) && !(lm.flags & LM_TO_MARG));
  const size_t res = res_index(w, r, t, inb ? l : 0);
  const int status = skip ? K_OUTLIER : w.status[res];
  if (skip) lm.rho = -1.f;
  PixelOut o;
  eval_pixel<FEJ, true>(pcs, lm, w.img[t], w.mask[t], w.W, w.H, status, sigma, huber, lane, o);
  if (skip) return;
  if (px == 0) {
    if (!o.ok) w.cand[res] = K_OOB;
    else if (o.ev) w.cand[res] = K_OK;
    w.energy[res] = o.e;
    if (o.ev) w.m_w[res] = o.w;  // huber_weight is left untouched when not evaluated (evaluate_jacobians.hpp:184-194)
  }
  w.m_r[res * 8 + px] = o.r;
  w.m_did[res * 8 + px] = o.d;
  const float* adj = FEJ ? pcs.adj0 : pcs.adj;
  float jr[8], jt[8];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) a += o.g[k] * adj[k * 6 + j];  // J_ref[:,0:6] = Jg Adj  (:162-163)
    jr[j] = a;
    jt[j] = -o.g[j];  // J_tgt[:,0:6] = -Jg leftLog (= I)  (:159-160)
  }
  const float sp = FEJ ? pcs.s0 : pcs.s;  // d_reference_affineBrightnessShift (:95,107)
  jr[6] = o.c;
  jr[7] = o.ev ? sp : 0.f;
  jt[6] = -o.c;
  jt[7] = o.ev ? -1.f : 0.f;
  float4* pr = reinterpret_cast<float4*>(w.m_jref + res * 64 + px * 8);
  float4* pt = reinterpret_cast<float4*>(w.m_jtgt + res * 64 + px * 8);
  pr[0] = make_float4(jr[0], jr[1], jr[2], jr[3]);
  pr[1] = make_float4(jr[4], jr[5], jr[6], jr[7]);
  pt[0] = make_float4(jt[0], jt[1], jt[2], jt[3]);
  pt[1] = make_float4(jt[4], jt[5], jt[6], jt[7]);
}

// ------------------------------------------------------------------------------------------------
// Fused linearise: K1 + K3 + the per-landmark half of K4, nothing materialised.
//
// With u_i = [g_i(6), c_i, 1] the two Jacobian rows of a pixel are  J_tgt_i = -u_i  and  J_ref_i = u_i B,
// B = blockdiag(Adj, 1, s').  So per ordered pair only the 8x8 "core" C = sum w u^T u (36 unique) and
// q = sum w u r (8) are accumulated (44 sums instead of 3*64+16 = 208); H_rr = B^T C B, H_rt = -B^T C,
// H_tt = C, b_r = B^T q, b_t = -q are formed once per pair by k_assemble in fp64.
//
// grid = (landmark chunks, host frames); one warp per target frame; each warp walks the chunk 4 landmarks at a
// time keeping its pair's 44 running sums in registers, so the per-landmark quantities that couple the targets
// (H_pd, H_dd, b_d) meet in shared memory.
// ------------------------------------------------------------------------------------------------
template <bool FEJ>
__global__ void __launch_bounds__(32 * (PBA_MAXF - 1))
    k_linearize_fused(const __grid_constant__ WindowDev w, float sigma, int huber, int for_marg, int lpb,
                      double* __restrict__ core, const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  extern __shared__ float smem[];
  const int N = w.n_frames;
  const int D = 8 * N;
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l0 = blockIdx.x * lpb;
  if (l0 >= M) return;
  float* hpd_s = smem;                 // [lpb][D]
  float* hdd_s = hpd_s + lpb * D;      // [lpb]
  float* bd_s = hdd_s + lpb;           // [lpb]
  PairConst* pcs = reinterpret_cast<PairConst*>(bd_s + lpb);  // [nwarps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = lane & 7, grp = lane >> 3;
  const int t = warp + (warp >= f);

  for (int i = threadIdx.x; i < lpb * (D + 2); i += blockDim.x) smem[i] = 0.f;
  for (int i = lane; i < (int)(sizeof(PairConst) / 4); i += 32)
    ((float*)&pcs[warp])[i] = ((const float*)&w.pairs[f * PBA_MAXF + t])[i];
  __syncthreads();
  const PairConst& pc = pcs[warp];
  const float* adj = FEJ ? pc.adj0 : pc.adj;
  const float sp = FEJ ? pc.s0 : pc.s;
  const float4* img = w.img[t];
  const uint8_t* mask = w.mask[t];

  float acc[PBA_CORE];
#pragma unroll
  for (int k = 0; k < PBA_CORE; ++k) acc[k] = 0.f;

  for (int it = 0; it < lpb; it += 4) {
    const int ls = it + grp;  // slot in the chunk
    const int l = l0 + ls;
    const bool inb = l < M;
    const int gl = lm_index(w, f, inb ? l : 0);
    LandmarkIn lm = load_landmark(w, gl, px);
    const bool skip = !inb || ((lm.flags & LM_MARG) && !(lm.flags & LM_TO_MARG));
    const size_t res = res_index(w, f, t, inb ? l : 0);
    const int status = skip ? K_OUTLIER : w.status[res];
    if (skip) lm.rho = -1.f;
    PixelOut o;
    eval_pixel<FEJ, true>(pc, lm, img, mask, w.W, w.H, status, sigma, huber, lane, o);
    if (!skip && px == 0) {
      if (!o.ok) w.cand[res] = K_OOB;
      else if (o.ev) w.cand[res] = K_OK;
      w.energy[res] = o.e;
    }
    // landmark selection of K3/K4 (hessian_block_evaluation.hpp:68-72,190-194)
    const bool sel = !skip && (for_marg ? (lm.flags & LM_TO_MARG) != 0 : (lm.flags & LM_MARG) == 0);
    const float wgt = (sel && o.ev) ? o.w : 0.f;
    float u[8];
#pragma unroll
    for (int k = 0; k < 6; ++k) u[k] = o.g[k];
    u[6] = o.c;
    u[7] = 1.f;
    float wu[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) wu[k] = wgt * u[k];
    {
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = a; b < 8; ++b) acc[idx++] += wu[a] * u[b];
#pragma unroll
      for (int a = 0; a < 8; ++a) acc[36 + a] += wu[a] * o.r;
    }
    // per landmark: H_pd blocks, H_dd, b_d  (hessian_block_evaluation.hpp:198-212)
    float pv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pv[k] = wu[k] * o.d;
    const float P = group_transpose_reduce(pv, px);  // lane px holds sum_i w d_i u_i[px]
    const float hdd = group_sum(wgt * o.d * o.d);
    const float bd = group_sum(wgt * o.d * o.r);
    // reference block: B^T-row px of P  ->  sum_k adj[k][px] P_k for px < 6
    float Pk[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) Pk[k] = __shfl_sync(FULL, P, (lane & 24) + k);
    float refv;
    if (px < 6) {
      refv = 0.f;
#pragma unroll
      for (int k = 0; k < 6; ++k) refv += adj[k * 6 + px] * Pk[k];
    } else {
      refv = (px == 6) ? P : sp * P;
    }
    if (sel) {
      hpd_s[ls * D + 8 * t + px] = -P;                 // target block: this warp is its only writer
      atomicAdd(&hpd_s[ls * D + 8 * f + px], refv);    // reference block: summed over the target warps
      if (px == 0) {
        atomicAdd(&hdd_s[ls], hdd);
        atomicAdd(&bd_s[ls], bd);
      }
    }
  }

  // warp-wide transpose-reduce of the 44(48) running sums: 24+12+6+3+2 = 47 shuffles, then <= 2 atomics per lane
  {
    float v24[24], v12[12], v6[6], v3[3], v2[2];
    tr_step<48, 16>(acc, v24, lane);
    tr_step<24, 8>(v24, v12, lane);
    tr_step<12, 4>(v12, v6, lane);
    tr_step<6, 2>(v6, v3, lane);
    tr_step<3, 1>(v3, v2, lane);
    const int off = ((lane & 16) ? 24 : 0) + ((lane & 8) ? 12 : 0) + ((lane & 4) ? 6 : 0) + ((lane & 2) ? 3 : 0) +
                    ((lane & 1) ? 2 : 0);
    double* dst = core + (size_t)(f * PBA_MAXF + t) * PBA_CORE;
    if (off < 44 && v2[0] != 0.f) atomicAdd(dst + off, (double)v2[0]);
    if (!(lane & 1) && off + 1 < 44 && v2[1] != 0.f) atomicAdd(dst + off + 1, (double)v2[1]);
  }
  __syncthreads();

  // finalise the chunk's landmarks (hessian_block_evaluation.hpp:213-227)
  for (int ls = threadIdx.x; ls < lpb; ls += blockDim.x) {
    const int l = l0 + ls;
    if (l >= M) continue;
    const int gl = lm_index(w, f, l);
    const int fl = w.flags[gl];
    const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
    const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
    if (!sel) continue;
    float hdd = hdd_s[ls];
    w.b_d[gl] = bd_s[ls];
    if (hdd > 1e-15f) {
      if (for_marg && w.fixed[f]) hdd += 1e8f;  // kScaleNullspaceRegularizer
      w.inv_hdd[gl] = 1.f / hdd;
      w.flags[gl] = (uint8_t)(fl & ~LM_ILL);
    } else {
      w.flags[gl] = (uint8_t)(fl | LM_ILL);
    }
  }
  for (int i = threadIdx.x; i < lpb * D; i += blockDim.x) {
    const int ls = i / D;
    const int l = l0 + ls;
    if (l >= M) break;
    const int gl = lm_index(w, f, l);
    const int fl = w.flags[gl];
    const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
    const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
    if (sel) w.hpd[(size_t)gl * w.hpd_stride + (i - ls * D)] = hpd_s[i];
  }
}

// ------------------------------------------------------------------------------------------------
// Reference three-pass dataflow on the device (cross-check + materialising-sweep measurement):
// K3 from the materialised arrays: per ordered pair H_rr, H_rt, H_tt, b_r, b_t written straight into H / b.
// grid = (chunks, pairs), 208 threads = one per output element.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(224) k_posepose_from_materialized(const __grid_constant__ WindowDev w, int for_marg,
                                                                    int chunk, double* __restrict__ Hp,
                                                                    double* __restrict__ bp) {
  const int N = w.n_frames, D = 8 * N;
  const int r = blockIdx.y / (N - 1);
  int t = blockIdx.y % (N - 1);
  t += (t >= r);
  const int M = w.n_lm[r];
  const int l0 = blockIdx.x * chunk;
  if (l0 >= M) return;
  const int e = threadIdx.x;
  if (e >= 208) return;
  const int which = e < 192 ? e / 64 : 3 + (e - 192) / 8;  // 0 rr, 1 rt, 2 tt, 3 br, 4 bt
  const int i = e < 192 ? (e % 64) / 8 : (e - 192) % 8;
  const int j = e % 8;
  double acc = 0;
  const int l1 = min(l0 + chunk, M);
  for (int l = l0; l < l1; ++l) {
    const int fl = w.flags[lm_index(w, r, l)];
    const bool sel = for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0;
    if (!sel) continue;
    const size_t res = res_index(w, r, t, l);
    const float wg = w.m_w[res];
    const float* jr = w.m_jref + res * 64;
    const float* jt = w.m_jtgt + res * 64;
    const float* rr = w.m_r + res * 8;
    float s = 0.f;
    for (int p = 0; p < 8; ++p) {
      const float a = (which == 0 || which == 1 || which == 3) ? jr[p * 8 + i] : jt[p * 8 + i];
      const float b = which == 0 ? jr[p * 8 + j] : (which <= 2 ? jt[p * 8 + j] : rr[p]);
      s += a * b;
    }
    acc += (double)(wg * s);
  }
  if (acc == 0) return;
  if (which == 0) atomicAdd(&Hp[(size_t)(8 * r + i) * D + 8 * r + j], acc);
  else if (which == 1) atomicAdd(&Hp[(size_t)(8 * r + i) * D + 8 * t + j], acc);
  else if (which == 2) atomicAdd(&Hp[(size_t)(8 * t + i) * D + 8 * t + j], acc);
  else if (which == 3) atomicAdd(&bp[8 * r + i], acc);
  else atomicAdd(&bp[8 * t + i], acc);
}

// per-landmark half of K4 from the materialised arrays.  One thread per (landmark, output column)
__global__ void __launch_bounds__(128) k_schur_prep_from_materialized(const __grid_constant__ WindowDev w, int for_marg) {
  const int N = w.n_frames, D = 8 * N;
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l = blockIdx.x;
  if (l >= M) return;
  const int gl = lm_index(w, f, l);
  const int fl = w.flags[gl];
  const bool sel = for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0;
  if (!sel) return;
  const int c = threadIdx.x;
  if (c < D) {
    const int blk = c / 8, j = c % 8;
    float acc = 0.f;
    for (int t = 0; t < N; ++t) {
      if (t == f) continue;
      if (blk != f && blk != t) continue;
      const size_t res = res_index(w, f, t, l);
      const float* J = (blk == f ? w.m_jref : w.m_jtgt) + res * 64;
      const float* d = w.m_did + res * 8;
      float s = 0.f;
      for (int p = 0; p < 8; ++p) s += J[p * 8 + j] * d[p];
      acc += w.m_w[res] * s;
    }
    w.hpd[(size_t)gl * w.hpd_stride + c] = acc;
  }
  if (c == 0) {
    float hdd = 0.f, bd = 0.f;
    for (int t = 0; t < N; ++t) {
      if (t == f) continue;
      const size_t res = res_index(w, f, t, l);
      const float* d = w.m_did + res * 8;
      const float* rr = w.m_r + res * 8;
      float a = 0.f, b = 0.f;
      for (int p = 0; p < 8; ++p) {
        a += d[p] * d[p];
        b += d[p] * rr[p];
      }
      hdd += w.m_w[res] * a;
      bd += w.m_w[res] * b;
    }
    w.b_d[gl] = bd;
    if (hdd > 1e-15f) {
      if (for_marg && w.fixed[f]) hdd += 1e8f;
      w.inv_hdd[gl] = 1.f / hdd;
      w.flags[gl] = (uint8_t)(fl & ~LM_ILL);
    } else {
      w.flags[gl] = (uint8_t)(fl | LM_ILL);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K4 second half: H_s += inv H_pd H_pd^T, b_s += inv b_d H_pd over the selected, well-conditioned landmarks.
// A (8N x L) x (L x 8N) SYRK.  Persistent grid; each block streams tiles of 32 landmarks through shared memory,
// every thread owns a 4x4 output tile: fp32 products per tile, fp64 running sums, one fp64 atomic per output
// element per block at the end.
// ------------------------------------------------------------------------------------------------
constexpr int SCHUR_TL = 32;
__global__ void __launch_bounds__(1024) k_schur(const __grid_constant__ WindowDev w, int for_marg,
                                                double* __restrict__ Hs, double* __restrict__ bs,
                                                const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  extern __shared__ float sm[];
  const int N = w.n_frames, D = 8 * N;
  const int T4 = D / 4;  // tiles per side
  float* Ps = sm;                   // [TL][D]  H_pd
  float* Qs = Ps + SCHUR_TL * D;    // [TL][D]  inv * H_pd
  float* Bs = Qs + SCHUR_TL * D;    // [TL]     inv * b_d
  const int tid = threadIdx.x;
  const int ty = tid / T4, tx = tid % T4;
  const bool active = tid < T4 * T4;
  double dacc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) dacc[k] = 0;
  double bacc = 0;  // thread tid < D accumulates b_s[tid]

  // flattened tile list over all frames
  int tiles_of[PBA_MAXF + 1];
  tiles_of[0] = 0;
  for (int f = 0; f < N; ++f) tiles_of[f + 1] = tiles_of[f] + (w.n_lm[f] + SCHUR_TL - 1) / SCHUR_TL;
  const int total = tiles_of[N];
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int f = 0;
    while (tile >= tiles_of[f + 1]) ++f;
    const int l0 = (tile - tiles_of[f]) * SCHUR_TL;
    const int M = w.n_lm[f];
    __syncthreads();
    for (int i = tid; i < SCHUR_TL * D; i += blockDim.x) {
      const int ls = i / D, c = i - ls * D;
      const int l = l0 + ls;
      float p = 0.f, q = 0.f;
      if (l < M) {
        const int gl = lm_index(w, f, l);
        const int fl = w.flags[gl];
        const bool sel = (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0) && !(fl & LM_ILL);
        if (sel) {
          p = w.hpd[(size_t)gl * w.hpd_stride + c];
          q = p * w.inv_hdd[gl];
        }
      }
      Ps[i] = p;
      Qs[i] = q;
    }
    for (int ls = tid; ls < SCHUR_TL; ls += blockDim.x) {
      const int l = l0 + ls;
      float v = 0.f;
      if (l < M) {
        const int gl = lm_index(w, f, l);
        const int fl = w.flags[gl];
        const bool sel = (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0) && !(fl & LM_ILL);
        if (sel) v = w.inv_hdd[gl] * w.b_d[gl];
      }
      Bs[ls] = v;
    }
    __syncthreads();
    if (active) {
      float a[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = 0.f;
#pragma unroll 4
      for (int ls = 0; ls < SCHUR_TL; ++ls) {
        const float4 qv = *reinterpret_cast<const float4*>(Qs + ls * D + 4 * ty);
        const float4 pv = *reinterpret_cast<const float4*>(Ps + ls * D + 4 * tx);
        const float qa[4] = {qv.x, qv.y, qv.z, qv.w};
        const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i * 4 + j] += qa[i] * pa[j];
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) dacc[k] += (double)a[k];
    }
    if (tid < D) {
      float b = 0.f;
      for (int ls = 0; ls < SCHUR_TL; ++ls) b += Bs[ls] * Ps[ls * D + tid];
      bacc += (double)b;
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (dacc[i * 4 + j] != 0) atomicAdd(&Hs[(size_t)(4 * ty + i) * D + 4 * tx + j], dacc[i * 4 + j]);
  }
  if (tid < D && bacc != 0) atomicAdd(&bs[tid], bacc);
}

// ------------------------------------------------------------------------------------------------
// assembly of H_pp / b_p from the per-pair cores (fp64) with the reference's scatter + symmetrisation
// (hessian_block_evaluation.hpp:118-163, quirk Q3).  grid = ordered pairs, 64 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_assemble(const __grid_constant__ WindowDev w, int fej,
                                                 const double* __restrict__ core, double* __restrict__ Hp,
                                                 double* __restrict__ bp, const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  const int N = w.n_frames, D = 8 * N;
  const int r = blockIdx.x / (N - 1);
  int t = blockIdx.x % (N - 1);
  t += (t >= r);
  __shared__ double C[8][8], Bm[8][8], BC[8][8], q[8];
  const int i = threadIdx.x / 8, j = threadIdx.x % 8;
  const double* c = core + (size_t)(r * PBA_MAXF + t) * PBA_CORE;
  const PairAssemble& pa = w.pairs_asm[r * PBA_MAXF + t];
  {
    const int a = min(i, j), b = max(i, j);
    const int idx = a * 8 - (a * (a - 1)) / 2 + (b - a);  // upper-triangle row-major index
    C[i][j] = c[idx];
    double bm = 0;
    if (i < 6 && j < 6) bm = (fej ? pa.adj_fej : pa.adj_cur)[i * 6 + j];
    else if (i == 6 && j == 6) bm = 1.0;
    else if (i == 7 && j == 7) bm = fej ? pa.s0 : pa.s;
    Bm[i][j] = bm;
    if (threadIdx.x < 8) q[threadIdx.x] = c[36 + threadIdx.x];
  }
  __syncthreads();
  double s = 0;
  for (int k = 0; k < 8; ++k) s += Bm[k][i] * C[k][j];  // (B^T C)[i][j]
  BC[i][j] = s;
  __syncthreads();
  double hrr = 0;
  for (int k = 0; k < 8; ++k) hrr += BC[i][k] * Bm[k][j];
  atomicAdd(&Hp[(size_t)(8 * r + i) * D + 8 * r + j], hrr);
  Hp[(size_t)(8 * r + i) * D + 8 * t + j] = -BC[i][j];  // assignment (Q3); each (r,t) block has one writer
  atomicAdd(&Hp[(size_t)(8 * t + i) * D + 8 * t + j], C[i][j]);
  if (j == 0) {
    double br = 0;
    for (int k = 0; k < 8; ++k) br += Bm[k][i] * q[k];
    atomicAdd(&bp[8 * r + i], br);
    atomicAdd(&bp[8 * t + i], -q[i]);
  }
}

__global__ void k_symmetrise(int D, double* __restrict__ Hp, const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= D * D) return;
  const int i = a / D, j = a % D;
  const int bi = i / 8, bj = j / 8;
  if (bi == bj) {
    if (i < j) Hp[(size_t)i * D + j] = Hp[(size_t)j * D + i];  // selfadjointView<Lower>
  } else if (bi < bj) {
    const double v = Hp[(size_t)i * D + j] + Hp[(size_t)j * D + i];
    Hp[(size_t)i * D + j] = v;
    Hp[(size_t)j * D + i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// K5: calculateIdepths.  8 lanes per landmark, each lane 1/8 of the 8N-long dot product.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_back_substitute(const __grid_constant__ WindowDev w,
                                                         const double* __restrict__ step_pose, float inv_lambda,
                                                         const LmCtl* __restrict__ ctl, double* __restrict__ norms) {
  if (lm_skip(ctl, 1)) return;
  if (ctl) inv_lambda = (float)(1.0 / (1.0 + ctl->lambda));
  __shared__ float sp[PBA_MAXF * 8];
  const int N = w.n_frames, D = 8 * N;
  for (int i = threadIdx.x; i < D; i += blockDim.x) sp[i] = (float)step_pose[i];
  __syncthreads();
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l = blockIdx.x * 32 + (threadIdx.x >> 3);
  const int px = threadIdx.x & 7;
  const bool inb = l < M;
  const int gl = lm_index(w, f, inb ? l : 0);
  float dot = 0.f;
  if (inb) {
    const float* h = w.hpd + (size_t)gl * w.hpd_stride;
    for (int c = px; c < D; c += 8) dot += h[c] * sp[c];
  }
  dot = group_sum(dot);
  double n_state = 0, n_step = 0;
  if (inb && px == 0) {
    const int fl = w.flags[gl];
    float stp = w.idepth_step[gl];
    if (!(fl & LM_MARG) && !(fl & LM_ILL)) {
      stp = -((w.b_d[gl] - dot) * inv_lambda * w.inv_hdd[gl]);
      w.idepth_step[gl] = stp;
    }
    if (norms) {  // landmark part of acceptStep's norms (problem.hpp:377-382), used by the device LM
      const float id = w.idepth[gl];
      n_state = (double)id * id;
      n_step = (double)stp * stp;
    }
  }
  if (norms) {
    __shared__ double sa[8], sb[8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      n_state += __shfl_xor_sync(FULL, n_state, s);
      n_step += __shfl_xor_sync(FULL, n_step, s);
    }
    if ((threadIdx.x & 31) == 0) {
      sa[threadIdx.x >> 5] = n_state;
      sb[threadIdx.x >> 5] = n_step;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int i = 0; i < 8; ++i) {
        a += sa[i];
        b += sb[i];
      }
      if (a != 0 || b != 0) {
        atomicAdd(&norms[2], a);
        atomicAdd(&norms[3], b);
      }
    }
  }
}

// acceptStep / rejectStep over landmarks (problem.hpp:377-384,395-399); norms in fp64
__global__ void __launch_bounds__(256) k_accept_landmarks(const __grid_constant__ WindowDev w, int accept,
                                                          double* __restrict__ scal, const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 1)) return;
  if (ctl) accept = ctl->accept;
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  double st = 0, sp = 0;
  if (l < M) {
    const int gl = lm_index(w, f, l);
    const float s = w.idepth_step[gl];
    if (accept > 0) {
      const float id = w.idepth[gl];
      st = (double)id * id;
      sp = (double)s * s;
      w.idepth[gl] = id + s;
    }
    w.idepth_step[gl] = 0.f;
  }
  if (accept <= 0 || !scal) return;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    st += __shfl_xor_sync(FULL, st, s);
    sp += __shfl_xor_sync(FULL, sp, s);
  }
  __shared__ double a[8], b[8];
  if ((threadIdx.x & 31) == 0) {
    a[threadIdx.x >> 5] = st;
    b[threadIdx.x >> 5] = sp;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double x = 0, y = 0;
    for (int i = 0; i < 8; ++i) {
      x += a[i];
      y += b[i];
    }
    atomicAdd(&scal[2], x);
    atomicAdd(&scal[3], y);
  }
}

// changeResidualStatuses (problem.hpp:20-35)
__global__ void __launch_bounds__(256) k_change_statuses(const __grid_constant__ WindowDev w, int accept,
                                                         const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 1)) return;
  if (ctl) accept = ctl->accept;
  const int N = w.n_frames;
  const int r = blockIdx.y / (N - 1);
  int t = blockIdx.y % (N - 1);
  t += (t >= r);
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= w.n_lm[r]) return;
  const size_t res = res_index(w, r, t, l);
  if (accept > 0) w.status[res] = w.cand[res];
  else w.cand[res] = w.status[res];
}

// calculateLandmarksEnergy (problem.hpp:93-144)
__global__ void __launch_bounds__(256) k_landmarks_energy(const __grid_constant__ WindowDev w, int for_marg,
                                                          double* __restrict__ scal) {
  const int N = w.n_frames;
  const int r = blockIdx.y / (N - 1);
  int t = blockIdx.y % (N - 1);
  t += (t >= r);
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  double e = 0;
  int n = 0;
  if (l < w.n_lm[r]) {
    const int fl = w.flags[lm_index(w, r, l)];
    const bool sel = for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0;
    if (sel) {
      const float v = w.energy[res_index(w, r, t, l)];
      e = v;
      n = v > 0.f;
    }
  }
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) {
    e += __shfl_xor_sync(FULL, e, s);
    n += __shfl_xor_sync(FULL, n, s);
  }
  if ((threadIdx.x & 31) == 0 && (n | (e != 0))) {
    atomicAdd(&scal[0], e);
    atomicAdd(&scal[1], (double)n);
  }
}

__global__ void k_snapshot_fej(const __grid_constant__ WindowDev w) {
  const int f = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= w.n_lm[f]) return;
  const int gl = lm_index(w, f, l);
  const int fl = w.flags[gl];
  if ((fl & LM_MARG) && !(fl & LM_TO_MARG)) return;  // first_estimate_jacobians.hpp:49
  w.idepth_fej[gl] = w.idepth[gl];                   // quirk Q9: the CURRENT idepth, no step
}

// second half of updatePointStatuses (photometric_bundle_adjustment.cpp:363-405)
__global__ void k_apply_point_statuses(const __grid_constant__ WindowDev w, float thr, int min_valid,
                                       const float* __restrict__ pair_dist) {
  const int N = w.n_frames;
  const int f = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= w.n_lm[f]) return;
  const int gl = lm_index(w, f, l);
  int fl = w.flags[gl];
  if (fl & LM_MARG) return;
  unsigned valid = 0;
  float rb = w.rel_baseline[gl];
  const float id = w.idepth[gl];
  for (int t = 0; t < N; ++t) {
    if (t == f || w.frame_marg[t]) continue;
    const size_t res = res_index(w, f, t, l);
    if (w.energy[res] > thr) {  // residual = {kOutlier}  (quirk Q6)
      w.status[res] = K_OUTLIER;
      w.cand[res] = K_OUTLIER;
      w.energy[res] = 0.f;
    }
    if (w.status[res] == K_OK) {
      rb = fmaxf(rb, id * pair_dist[f * PBA_MAXF + t]);
      ++valid;
    }
  }
  w.rel_baseline[gl] = rb;
  w.n_inliers[gl] = valid;
  if ((int)valid < min_valid) w.flags[gl] = (uint8_t)(fl | LM_OUTLIER);
}

// ------------------------------------------------------------------------------------------------
// per-pair constants in fp64 (evaluate_jacobians.hpp:36-66, camera_reproject.hpp:235-260,
// first_estimate_jacobians.hpp:28-37).  Sophus exp / Adj restated from the closed forms.
// ------------------------------------------------------------------------------------------------
struct SE3d {
  double R[9];
  double t[3];
};
__device__ void se3_mul(const SE3d& a, const SE3d& b, SE3d& o) {
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
}
__device__ void se3_inv(const SE3d& a, SE3d& o) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) o.R[i * 3 + j] = a.R[j * 3 + i];
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += o.R[i * 3 + k] * a.t[k];
    o.t[i] = -s;
  }
}
__device__ void se3_exp(const double* xi, double sign, SE3d& o) {
  const double v[3] = {sign * xi[0], sign * xi[1], sign * xi[2]};
  const double wv[3] = {sign * xi[3], sign * xi[4], sign * xi[5]};
  const double th2 = wv[0] * wv[0] + wv[1] * wv[1] + wv[2] * wv[2];
  const double th = sqrt(th2);
  double a, b, c;
  if (th < 1e-10) {
    a = 1.0;
    b = 0.5;
    c = 1.0 / 6.0;
  } else {
    double sn, cs;
    sincos(th, &sn, &cs);
    a = sn / th;
    b = (1.0 - cs) / th2;
    c = (th - sn) / (th2 * th);
  }
  const double W[9] = {0, -wv[2], wv[1], wv[2], 0, -wv[0], -wv[1], wv[0], 0};
  double W2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double s = 0;
      for (int k = 0; k < 3; ++k) s += W[i * 3 + k] * W[k * 3 + j];
      W2[i * 3 + j] = s;
    }
  double V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    o.R[i] = I + a * W[i] + b * W2[i];
    V[i] = I + b * W[i] + c * W2[i];
  }
  for (int i = 0; i < 3; ++i) o.t[i] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
}
__device__ void se3_adj(const SE3d& T, double* A) {  // [[R, hat(t) R], [0, R]] row-major 6x6
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
// transform_unproject_ = [R|t] Kr^-1 (3x4) and reproject_ = K_t * that
__device__ void make_proj(const SE3d& T, const double* ir, const double* it, float* M, float* A) {
  const double fx = ir[0], fy = ir[1], cx = ir[2], cy = ir[3];
  double m[12];
  for (int i = 0; i < 3; ++i) {
    m[i * 4 + 0] = T.R[i * 3 + 0] * (1.0 / fx);
    m[i * 4 + 1] = T.R[i * 3 + 1] * (1.0 / fy);
    m[i * 4 + 2] = T.R[i * 3 + 0] * (-cx / fx) + T.R[i * 3 + 1] * (-cy / fy) + T.R[i * 3 + 2];
    m[i * 4 + 3] = T.t[i];
  }
  for (int i = 0; i < 12; ++i) M[i] = (float)m[i];
  if (A) {
    for (int j = 0; j < 4; ++j) {
      A[0 + j] = (float)(it[0] * m[0 + j] + it[2] * m[8 + j]);
      A[4 + j] = (float)(it[1] * m[4 + j] + it[3] * m[8 + j]);
      A[8 + j] = (float)m[8 + j];
    }
  }
}

__global__ void __launch_bounds__(256) k_pair_setup(const FrameParams* __restrict__ fr, int N,
                                                     PairConst* __restrict__ pairs, PairAssemble* __restrict__ pasm) {
  // phase 1 (one thread per frame): exp(+eps), exp(-eps) and the inverse linearisation pose
  __shared__ SE3d s_er[PBA_MAXF], s_et[PBA_MAXF], s_tl[PBA_MAXF], s_ti[PBA_MAXF];
  __shared__ double s_a[PBA_MAXF], s_b[PBA_MAXF];
  const int tid = threadIdx.x;
  if (tid < N) {
    const FrameParams& F = fr[tid];
    double e[6];
    for (int k = 0; k < 6; ++k) e[k] = F.eps[k] + F.step[k];
    se3_exp(e, 1.0, s_er[tid]);
    se3_exp(e, -1.0, s_et[tid]);
    SE3d T;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T.R[i * 3 + j] = F.T_lin[i * 4 + j];
      T.t[i] = F.T_lin[i * 4 + 3];
    }
    s_tl[tid] = T;
    se3_inv(T, s_ti[tid]);
    s_a[tid] = F.ab0[0] + F.eps[6] + F.step[6];
    s_b[tid] = F.ab0[1] + F.eps[7] + F.step[7];
  }
  __syncthreads();
  // phase 2 (one thread per ordered pair)
  if (tid >= N * N) return;
  const int r = tid / N, t = tid % N;
  if (r == t) return;
  const FrameParams& R = fr[r];
  const FrameParams& T = fr[t];
  SE3d T0, tmp, Tc;
  se3_mul(s_ti[t], s_tl[r], T0);  // t_t_r0 (evaluate_jacobians.hpp:47-48)
  se3_mul(T0, s_er[r], tmp);
  se3_mul(s_et[t], tmp, Tc);      // t_t_r = exp(-eps_t) T0 exp(eps_r)  (:49)

  PairConst pc;
  make_proj(Tc, R.intr, T.intr, pc.M, pc.A);
  make_proj(T0, R.intr, T.intr, pc.M0, nullptr);
  for (int i = 0; i < 3; ++i) {
    pc.tr[i] = (float)Tc.t[i];
    pc.t0[i] = (float)T0.t[i];
  }
  PairAssemble pa;
  se3_adj(Tc, pa.adj_cur);
  se3_adj(T0, pa.adj_fej);
  for (int i = 0; i < 36; ++i) {
    pc.adj[i] = (float)pa.adj_cur[i];
    pc.adj0[i] = (float)pa.adj_fej[i];
  }
  const double ratio = T.exposure / R.exposure;
  pa.s = ratio * exp(s_a[t] - s_a[r]);
  pa.s0 = ratio * exp(T.ab0[0] - R.ab0[0]);
  const int last = (r == N - 1) ? N - 2 : N - 1;  // last target in deque order (quirk Q1)
  const double s0_last = (fr[last].exposure / R.exposure) * exp(fr[last].ab0[0] - R.ab0[0]);
  pc.s = (float)pa.s;
  pc.s0 = (float)pa.s0;
  pc.s0_last = (float)s0_last;
  pc.b_t = (float)s_b[t];
  pc.b_r = (float)s_b[r];
  pc.b_r0 = (float)R.ab0[1];
  pc.fx_t = (float)T.intr[0];
  pc.fy_t = (float)T.intr[1];
  pc.cx_t = (float)T.intr[2];
  pc.cy_t = (float)T.intr[3];
  pc.pad[0] = pc.pad[1] = 0.f;
  pairs[r * PBA_MAXF + t] = pc;
  pasm[r * PBA_MAXF + t] = pa;
}

// {I,dx,dy} float3 -> float4 texels
__global__ void k_pack_image(const float* __restrict__ src, float4* __restrict__ dst, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
}

// gradient packing from the intensity plane, features/src/calculate_pixelinfo.cpp:340-374
__global__ void k_pixelinfo(const float* __restrict__ I, float4* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const float c = I[y * W + x];
  float dx, dy;
  if (x == 0) dx = 1.0f * (I[y * W + 1] - c);
  else if (x == W - 1) dx = 1.0f * (c - I[y * W + x - 1]);
  else dx = 0.5f * (I[y * W + x + 1] - I[y * W + x - 1]);
  const int yu = y == 0 ? y : y - 1, yb = y == H - 1 ? y : y + 1;
  dy = ((y == 0 || y == H - 1) ? 1.0f : 0.5f) * (I[yb * W + x] - I[yu * W + x]);
  dst[y * W + x] = make_float4(c, dx, dy, 0.f);
}


// ------------------------------------------------------------------------------------------------
// Device-resident Levenberg-Marquardt: the control flow of levenberg_marquardt_algorithm::solve
// (levenberg_marquardt_algorithm.hpp:77-128) and the host half of PhotometricBundleAdjustmentProblem
// (problem.hpp:290-402) as single-CTA fp64 kernels, so that a whole solve is one stream of launches with no
// host round trip.  Decisions live in LmCtl; loop-body kernels early-out on ctl->done.
// ------------------------------------------------------------------------------------------------
__global__ void k_lm_init(LmCtl* ctl, const LmOptionsDev* opt) {
  if (threadIdx.x) return;
  ctl->lambda = opt->lambda0;
  ctl->energy = 0;
  ctl->next_energy = 0;
  ctl->state_sq = ctl->step_sq = 0;
  ctl->n_valid = ctl->next_n = 0;
  ctl->converged = 0;
  ctl->done = 0;
  ctl->system_valid = 0;
  ctl->accept = 0;
  ctl->iteration = 0;
  ctl->iterations_executed = 0;
}

__global__ void k_lm_zero(const LmCtl* ctl, double* p, int n, int mode) {
  if (lm_skip(ctl, mode)) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}

// calculateEnergy() tail (problem.hpp:293-316) + the accept decision (levenberg_marquardt_algorithm.hpp:95-104)
__global__ void __launch_bounds__(128) k_lm_energy(LmCtl* ctl, const LmOptionsDev* opt, const FrameParams* fr,
                                                   int N, const double* scal, const double* Hmarg,
                                                   const double* bmarg, int kind) {
  if (kind == pba::LM_ENERGY_TRIAL && ctl->done) return;
  __shared__ double s[PBA_MAXF * 8];
  __shared__ double red[128];
  const int D = 8 * N, i = threadIdx.x;
  if (i < D) s[i] = fr[i / 8].eps[i % 8] + fr[i / 8].step[i % 8];
  __syncthreads();
  double acc = 0;
  if (i < D) {
    if (Hmarg) {
      double t = 0;
      for (int j = 0; j < D; ++j) t += Hmarg[(size_t)i * D + j] * s[j];
      acc += bmarg[i] * s[i] + 0.5 * s[i] * t;  // DSO eq 8.19
    }
    const int k = i % 8;
    if (k >= 6) {  // AffineBrightnessPrior::energyTerm for every frame, fixed included (quirk Q8)
      const double ab = fr[i / 8].ab0[k - 6] + s[i];
      acc += 0.5 * ab * opt->ab_reg[k - 6] * ab;
    }
  }
  red[i] = acc;
  __syncthreads();
  for (int st = 64; st > 0; st >>= 1) {
    if (i < st) red[i] += red[i + st];
    __syncthreads();
  }
  if (i) return;
  const double E = opt->energy_marg + red[0] + scal[0];
  const int n = (int)llrint(scal[1]);
  if (kind == pba::LM_ENERGY_INITIAL) {
    ctl->energy = E;
    ctl->n_valid = n;
    if (n <= 0 || opt->max_it <= 0) ctl->done = 1;
  } else if (kind == pba::LM_ENERGY_TRIAL) {
    ctl->next_energy = E;
    ctl->next_n = n;
    ctl->iterations_executed += 1;
    // landmark parts of the norms, accumulated by k_back_substitute (problem.hpp:377-382)
    ctl->state_sq = scal[2];
    ctl->step_sq = scal[3];
    if (n == 0) {
      ctl->accept = -1;  // rejectStep(); break
    } else {
      if (fabs(ctl->energy - E) / ctl->energy < opt->ftol) ctl->converged = 1;  // before the accept test (Q7)
      ctl->accept = (E < ctl->energy || (opt->force_accept && ctl->iteration < opt->min_it)) ? 1 : 0;
    }
  }
}

// acceptStep / rejectStep for the frame state + the loop bookkeeping (problem.hpp:366-402, lm.hpp:104-122)
__global__ void k_lm_finish(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, int N) {
  if (threadIdx.x || ctl->done) return;
  if (ctl->accept > 0) {
    double st = ctl->state_sq, sp = ctl->step_sq;
    for (int f = 0; f < N; ++f) {
      for (int k = 0; k < 8; ++k) st += fr[f].eps[k] * fr[f].eps[k];
      st += fr[f].ab0[0] * fr[f].ab0[0] + fr[f].ab0[1] * fr[f].ab0[1];
      for (int k = 0; k < 8; ++k) {
        fr[f].eps[k] += fr[f].step[k];
        sp += fr[f].step[k] * fr[f].step[k];
        fr[f].step[k] = 0;
      }
    }
    ctl->state_sq = st;
    ctl->step_sq = sp;
    if (sp < opt->ptol * (st + opt->ptol)) ctl->converged = 1;
    ctl->energy = ctl->next_energy;
    ctl->n_valid = ctl->next_n;
    ctl->lambda /= opt->dec;
    ctl->system_valid = 0;
  } else {
    for (int f = 0; f < N; ++f)
      for (int k = 0; k < 8; ++k) fr[f].step[k] = 0;
    if (ctl->accept < 0 || opt->force_accept) {
      ctl->done = 1;
      return;
    }
    ctl->lambda *= opt->inc;
    ctl->system_valid = 1;
  }
  ctl->iteration += 1;
  if (ctl->iteration >= opt->max_it || ctl->converged || ctl->n_valid <= 0) ctl->done = 1;
}

// calculateStep (problem.hpp:342-357): priors (problem.hpp:37-77), full system, Jacobi preconditioner + LDL^T
// (normal_linear_system.cpp:10-59), all fp64 in one CTA; A lives in dynamic shared memory [D][D+1].
__global__ void __launch_bounds__(256) k_lm_step(const LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr,
                                                 const int* fixed, int N, const double* __restrict__ Hp,
                                                 const double* __restrict__ bp, const double* __restrict__ Hs,
                                                 const double* __restrict__ bs, const double* __restrict__ Hmarg,
                                                 const double* __restrict__ bmarg, double* __restrict__ step_dev) {
  if (ctl->done) return;
  extern __shared__ double sh[];
  const int D = 8 * N, LD = D + 1, tid = threadIdx.x, nt = blockDim.x;
  double* A = sh;              // [D][LD]
  double* b = A + D * LD;      // [D]
  double* pre = b + D;         // [D]
  double* st = pre + D;        // [D] state eps
  const double lambda = ctl->lambda, ks = -1.0 / (1.0 + lambda);
  for (int i = tid; i < D; i += nt) st[i] = fr[i / 8].eps[i % 8];
  __syncthreads();
  for (int idx = tid; idx < D * D; idx += nt) {
    const int i = idx / D, j = idx - i * D;
    double hp = Hp[idx];
    if (i == j) {
      const int f = i / 8, k = i % 8;
      if (fixed[f]) hp += opt->fixed_reg;
      else if (k >= 6) hp += opt->ab_reg[k - 6];
      hp += hp * lambda;
    }
    A[i * LD + j] = hp + (Hmarg ? Hmarg[idx] : 0.0) + ks * Hs[idx];
  }
  for (int i = tid; i < D; i += nt) {
    const int f = i / 8, k = i % 8;
    double v = bp[i] + ks * bs[i];
    if (fixed[f]) v += opt->fixed_reg * st[i];
    else if (k >= 6) v += opt->ab_reg[k - 6] * (fr[f].ab0[k - 6] + st[i]);
    if (Hmarg) {
      double t = 0;
      for (int j = 0; j < D; ++j) t += Hmarg[(size_t)i * D + j] * st[j];
      v += bmarg[i] + t;
    }
    b[i] = v;
  }
  __syncthreads();
  for (int i = tid; i < D; i += nt) pre[i] = 1.0 / sqrt(A[i * LD + i] + 10.0);
  __syncthreads();
  for (int idx = tid; idx < D * D; idx += nt) {
    const int i = idx / D, j = idx - i * D;
    A[i * LD + j] *= pre[i] * pre[j];
  }
  for (int i = tid; i < D; i += nt) b[i] *= pre[i];
  __syncthreads();
  // LDL^T, right-looking, forward substitution folded in; row k keeps d * L^T so the update is A_ij -= L_ik A_kj
  for (int k = 0; k < D; ++k) {
    const double d = A[k * LD + k];
    const double inv = d != 0.0 ? 1.0 / d : 0.0;
    for (int i = k + 1 + tid; i < D; i += nt) A[i * LD + k] *= inv;
    __syncthreads();
    const int m = D - k - 1;
    for (int idx = tid; idx < m * m; idx += nt) {
      const int i = k + 1 + idx / m, j = k + 1 + idx % m;
      A[i * LD + j] -= A[i * LD + k] * A[k * LD + j];
    }
    for (int i = k + 1 + tid; i < D; i += nt) b[i] -= A[i * LD + k] * b[k];
    __syncthreads();
  }
  for (int i = tid; i < D; i += nt) {
    const double d = A[i * LD + i];
    b[i] = d != 0.0 ? b[i] / d : 0.0;
  }
  __syncthreads();
  for (int k = D - 1; k > 0; --k) {
    const double xk = b[k];
    for (int i = tid; i < k; i += nt) b[i] -= A[k * LD + i] * xk;
    __syncthreads();
  }
  for (int i = tid; i < D; i += nt) {
    const double x = b[i] * pre[i];
    step_dev[i] = x;
    fr[i / 8].step[i % 8] = -x;  // frame.state_eps_step = -frame_step (problem.hpp:353-357)
  }
}

int max_landmarks(const WindowDev& w) {
  int m = 0;
  for (int f = 0; f < w.n_frames; ++f) m = w.n_lm[f] > m ? w.n_lm[f] : m;
  return m;
}

}  // namespace

namespace pba {

std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void add_launches(long long n) { g_launches += n; }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}


void launch_lm_init(LmCtl* ctl, const LmOptionsDev* opt, cudaStream_t s) {
  ++g_launches;
  k_lm_init<<<1, 32, 0, s>>>(ctl, opt);
}

void launch_lm_zero(const LmCtl* ctl, double* p, int n, int mode, cudaStream_t s) {
  ++g_launches;
  k_lm_zero<<<(n + 255) / 256, 256, 0, s>>>(ctl, p, n, mode);
}

void launch_lm_energy(LmCtl* ctl, const LmOptionsDev* opt, const FrameParams* fr, int N, const double* scal,
                      const double* Hmarg, const double* bmarg, int kind, cudaStream_t s) {
  ++g_launches;
  k_lm_energy<<<1, 128, 0, s>>>(ctl, opt, fr, N, scal, Hmarg, bmarg, kind);
}

void launch_lm_step(const LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, const int* fixed, int N, ReduceBuf rb,
                    const double* Hmarg, const double* bmarg, double* step_dev, cudaStream_t s) {
  const int D = 8 * N;
  const size_t smem = (size_t)(D * (D + 1) + 3 * D) * sizeof(double);
  static bool attr_set = false;
  if (smem > 48 * 1024 && !attr_set) {
    cudaFuncSetAttribute(k_lm_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(128 * 129 + 3 * 128) * 8));
    attr_set = true;
  }
  ++g_launches;
  k_lm_step<<<1, 256, smem, s>>>(ctl, opt, fr, fixed, N, rb.Hp, rb.bp, rb.Hs, rb.bs, Hmarg, bmarg, step_dev);
}

void launch_lm_finish(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, int N, cudaStream_t s) {
  ++g_launches;
  k_lm_finish<<<1, 32, 0, s>>>(ctl, opt, fr, N);
}

void launch_pair_setup(const FrameParams* frames, int n_frames, PairConst* pairs, PairAssemble* pasm, cudaStream_t s) {
  ++g_launches;
  k_pair_setup<<<1, 256, 0, s>>>(frames, n_frames, pairs, pasm);
}

void launch_pack_image(const float* src3, float4* dst, int n_px, cudaStream_t s) {
  ++g_launches;
  k_pack_image<<<(n_px + 255) / 256, 256, 0, s>>>(src3, dst, n_px);
}

void launch_pixelinfo(const float* I, float4* dst, int W, int H, cudaStream_t s) {
  dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
  ++g_launches;
  k_pixelinfo<<<g, b, 0, s>>>(I, dst, W, H);
}

void launch_residual_sweep(const WindowDev& w, float sigma, int huber, int fej, double* scal, cudaStream_t s,
                           const LmCtl* ctl, int ctl_mode) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  dim3 g((m + 31) / 32, w.n_frames * (w.n_frames - 1));
  ++g_launches;
  if (fej) k_residual_sweep<true><<<g, 256, 0, s>>>(w, sigma, huber, scal, ctl, ctl_mode);
  else k_residual_sweep<false><<<g, 256, 0, s>>>(w, sigma, huber, scal, ctl, ctl_mode);
}

void launch_materialise_sweep(const WindowDev& w, float sigma, int huber, int fej, cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  dim3 g((m + 31) / 32, w.n_frames * (w.n_frames - 1));
  ++g_launches;
  if (fej) k_materialise_sweep<true><<<g, 256, 0, s>>>(w, sigma, huber);
  else k_materialise_sweep<false><<<g, 256, 0, s>>>(w, sigma, huber);
}

void launch_linearize_fused(const WindowDev& w, float sigma, int huber, int fej, int for_marg, ReduceBuf rb,
                            cudaStream_t s, const LmCtl* ctl) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  const int N = w.n_frames, D = 8 * N;
  // landmarks per block: small chunks while the window is small (fill the 148 SMs), larger once it is not
  int lpb = 16;
  while (lpb < 64 && (long)((m + lpb - 1) / lpb) * N > 8L * sm_count()) lpb *= 2;
  const size_t smem = (size_t)lpb * (D + 2) * sizeof(float) + (size_t)(N - 1) * sizeof(PairConst);
  dim3 g((m + lpb - 1) / lpb, N);
  const int threads = 32 * (N - 1);
  if (fej) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_linearize_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ++g_launches;
    k_linearize_fused<true><<<g, threads, smem, s>>>(w, sigma, huber, for_marg, lpb, rb.core, ctl);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_linearize_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    ++g_launches;
    k_linearize_fused<false><<<g, threads, smem, s>>>(w, sigma, huber, for_marg, lpb, rb.core, ctl);
  }
}

void launch_linearize_from_materialized(const WindowDev& w, int for_marg, ReduceBuf rb, cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  const int N = w.n_frames;
  const int chunk = 64;
  dim3 g((m + chunk - 1) / chunk, N * (N - 1));
  ++g_launches;
  k_posepose_from_materialized<<<g, 224, 0, s>>>(w, for_marg, chunk, rb.Hp, rb.bp);
  dim3 g2(m, N);
  ++g_launches;
  k_schur_prep_from_materialized<<<g2, 128, 0, s>>>(w, for_marg);
}

void launch_schur(const WindowDev& w, int for_marg, ReduceBuf rb, cudaStream_t s, const LmCtl* ctl) {
  const int N = w.n_frames, D = 8 * N;
  int tiles = 0;
  for (int f = 0; f < N; ++f) tiles += (w.n_lm[f] + SCHUR_TL - 1) / SCHUR_TL;
  if (tiles == 0) return;
  const int T4 = D / 4;
  int threads = ((T4 * T4 + 31) / 32) * 32;
  if (threads < D) threads = ((D + 31) / 32) * 32;
  const size_t smem = (size_t)(2 * SCHUR_TL * D + SCHUR_TL) * sizeof(float);
  const int grid = tiles < 2 * sm_count() ? tiles : 2 * sm_count();
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  ++g_launches;
  k_schur<<<grid, threads, smem, s>>>(w, for_marg, rb.Hs, rb.bs, ctl);
}

void launch_assemble(const WindowDev& w, int fej, ReduceBuf rb, cudaStream_t s, const LmCtl* ctl) {
  const int N = w.n_frames, D = 8 * N;
  if (N < 2) return;
  ++g_launches;
  k_assemble<<<N * (N - 1), 64, 0, s>>>(w, fej, rb.core, rb.Hp, rb.bp, ctl);
  ++g_launches;
  k_symmetrise<<<(D * D + 255) / 256, 256, 0, s>>>(D, rb.Hp, ctl);
}

void launch_symmetrise_only(int D, double* Hp, cudaStream_t s) {
  ++g_launches;
  k_symmetrise<<<(D * D + 255) / 256, 256, 0, s>>>(D, Hp, nullptr);
}

void launch_back_substitute(const WindowDev& w, const double* step_pose_dev, double lambda, cudaStream_t s,
                            const LmCtl* ctl, double* norms) {
  const int m = max_landmarks(w);
  if (m == 0) return;
  dim3 g((m + 31) / 32, w.n_frames);
  ++g_launches;
  k_back_substitute<<<g, 256, 0, s>>>(w, step_pose_dev, (float)(1.0 / (1.0 + lambda)), ctl, norms);
}

void launch_accept(const WindowDev& w, int accept, double* scal, cudaStream_t s, const LmCtl* ctl) {
  const int m = max_landmarks(w);
  if (m == 0) return;
  dim3 g((m + 255) / 256, w.n_frames);
  ++g_launches;
  k_accept_landmarks<<<g, 256, 0, s>>>(w, accept, scal, ctl);
}

void launch_change_statuses(const WindowDev& w, int accept, cudaStream_t s, const LmCtl* ctl) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  dim3 g((m + 255) / 256, w.n_frames * (w.n_frames - 1));
  ++g_launches;
  k_change_statuses<<<g, 256, 0, s>>>(w, accept, ctl);
}

void launch_landmarks_energy(const WindowDev& w, int for_marg, double* scal, cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  dim3 g((m + 255) / 256, w.n_frames * (w.n_frames - 1));
  ++g_launches;
  k_landmarks_energy<<<g, 256, 0, s>>>(w, for_marg, scal);
}

void launch_snapshot_fej(const WindowDev& w, cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0) return;
  dim3 g((m + 255) / 256, w.n_frames);
  ++g_launches;
  k_snapshot_fej<<<g, 256, 0, s>>>(w);
}

void launch_apply_point_statuses(const WindowDev& w, float threshold, int min_valid, const float* pair_dist,
                                 cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0) return;
  dim3 g((m + 255) / 256, w.n_frames);
  ++g_launches;
  k_apply_point_statuses<<<g, 256, 0, s>>>(w, threshold, min_valid, pair_dist);
}

}  // namespace pba
