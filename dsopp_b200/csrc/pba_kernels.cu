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

#include <algorithm>
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
  if (mode == 3) return !ctl->relin;  // re-linearisation at the restored state after a rejected speculative step
  if (ctl->done) return true;
  return mode == 2 && ctl->system_valid;
}

// clock64() stamps of the single-CTA LM kernels (diagnostics: dpba_debug_stamps); written by thread 0 only
__device__ long long g_stamps[64];
__device__ int g_stamps_on = 0;
// per-CTA timeline of the fused sweep (diagnostics: dpba_debug_cta_times): %globaltimer (ns) at entry, end of the sweep
// proper, end of the CTA, and the SM it ran on -- written by thread 0 of every CTA while the stamps are on
__device__ long long g_cta_times[1024 * 4];
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// entry (block 0) and exit (latest block) of the LM-loop kernels in %globaltimer ns (diagnostics: dpba_debug_cta_times reads
// them behind the CTA table): where an iteration's time goes BETWEEN the kernels
__device__ unsigned long long g_kst[32];
struct KStamp {
  int id;
  __device__ __forceinline__ explicit KStamp(int i) : id(i) {
    if (g_stamps_on && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_kst[2 * id] = (unsigned long long)global_ns();
  }
  __device__ __forceinline__ ~KStamp() {
    if (g_stamps_on && threadIdx.x == 0) atomicMax(&g_kst[2 * id + 1], (unsigned long long)global_ns());
  }
};
__device__ __forceinline__ void cta_stamp(int slot) {
  if (g_stamps_on && threadIdx.x == 0) {
    const int c = blockIdx.y * gridDim.x + blockIdx.x;
    if (c < 1024) {
      if (slot == 3) {
        unsigned sm;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        g_cta_times[c * 4 + 3] = sm;
      } else {
        g_cta_times[c * 4 + slot] = global_ns();
      }
    }
  }
}
__device__ __forceinline__ void stamp(int i) {
  if (g_stamps_on && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) g_stamps[i] = clock64();
}

// ---- fused exchange over NVLink peer memory (round 2) -------------------------------------------------------------------
// The sharded solve sums ONE block per iteration over the ranks ([H_pp | b_p | H_s | b_s | 8 scalars], 66.6 KB at 8
// keyframes).  As a separate collective (ncclAllReduce, or the stand-alone mailbox kernel of peer_exchange.cu) it costs
// 22-52 us per iteration on 2-8 B200s -- launch boundaries and a latency-bound kernel between the producers and the LM
// step.  Here the exchange has no kernel of its own: the PRODUCERS (k_assemble, k_finish_fused, k_reduce_scal) store every
// result into slot [parity][rank] of every rank's mailbox while they write it locally (posted NVLink stores), fence, and
// add one to their arrival counter in every mailbox (red.release.sys); the CONSUMERS (k_lm_energy, k_lm_step) wait for the
// counters of all ranks (ld.acquire.sys, with a time-out), sum the W slots IN RANK ORDER from local L2 -- every rank adds
// the same values in the same order, so the replicated LM step stays bitwise identical -- and bump the epoch.  Parity and
// epoch are those of peer_exchange.cu (one counter for both kinds of exchange), so the two can interleave.
__device__ pba::PeerDev g_peer;
__device__ __forceinline__ unsigned peer_epoch() { return *reinterpret_cast<volatile unsigned*>(g_peer.seq) + 1u; }
__device__ __forceinline__ void red_store(double* p, double v, int peer_push, unsigned epoch) {
  *p = v;
  if (peer_push) {
    const size_t idx = (size_t)(p - g_peer.red_base);
    const size_t off = ((size_t)(epoch & 1u) * pba::PEER_MAXW + (size_t)g_peer.rank) * g_peer.slot + idx;
#pragma unroll
    for (int r = 0; r < pba::PEER_MAXW; ++r)
      if (r < g_peer.world) g_peer.data[r][off] = v;
  }
}
// all threads of a producer CTA, after their red_store()s
__device__ __forceinline__ void peer_arrive(int kind, unsigned epoch, bool stored) {
  // only the threads that pushed something fence: a system-scope fence waits for the thread's NVLink stores to be
  // acknowledged, and a thousand threads doing that for nothing cost the producers ~14 us each (r02j)
  if (stored) __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < g_peer.world) {
    unsigned* c = nullptr;
#pragma unroll
    for (int r = 0; r < pba::PEER_MAXW; ++r)
      if (r == (int)threadIdx.x) c = g_peer.cnt[r];
    c += ((epoch & 1u) * 2 + kind) * pba::PEER_MAXW + g_peer.rank;
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
  }
}
// all threads of the (single) consumer CTA: returns once `expected` arrivals of kind `kind` from every rank are visible
__device__ __forceinline__ void peer_wait_arrivals(int kind, unsigned epoch, unsigned expected) {
  if ((int)threadIdx.x < g_peer.world) {
    const unsigned* c = nullptr;
#pragma unroll
    for (int r = 0; r < pba::PEER_MAXW; ++r)
      if (r == g_peer.rank) c = g_peer.cnt[r];
    c += ((epoch & 1u) * 2 + kind) * pba::PEER_MAXW + threadIdx.x;
    const long long t0 = clock64();
    for (;;) {
      unsigned v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
      if (v >= expected || *reinterpret_cast<volatile int*>(g_peer.error)) break;
      if (clock64() - t0 > 30000000000LL) {  // ~15 s: a peer that has not arrived by then has failed (DPBA_E_COMM)
        *reinterpret_cast<volatile int*>(g_peer.error) = 1;
        *reinterpret_cast<volatile int*>(g_peer.error_host) = 1;
        break;
      }
    }
  }
  __syncthreads();
}
// all threads: out[i] = sum over ranks (in rank order) of slot [parity][r][off + i], i < n
__device__ __forceinline__ void peer_collect(double* out, size_t off, int n, unsigned epoch) {
  const double* box = nullptr;
#pragma unroll
  for (int r = 0; r < pba::PEER_MAXW; ++r)
    if (r == g_peer.rank) box = g_peer.data[r];
  box += (size_t)(epoch & 1u) * pba::PEER_MAXW * g_peer.slot + off;
  const int W = g_peer.world;
  for (int i = 2 * threadIdx.x; i < n; i += 2 * blockDim.x) {  // off and n are even
    double2 acc = __ldcg(reinterpret_cast<const double2*>(box + i));
    for (int r = 1; r < W; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(box + (size_t)r * g_peer.slot + i));
      acc.x += v.x;
      acc.y += v.y;
    }
    *reinterpret_cast<double2*>(out + i) = acc;
  }
}
// thread 0 of the consumer, after a barrier: counters of this parity back to zero, epoch published
__device__ __forceinline__ void peer_close(unsigned epoch) {
  unsigned* c = nullptr;
#pragma unroll
  for (int r = 0; r < pba::PEER_MAXW; ++r)
    if (r == g_peer.rank) c = g_peer.cnt[r];
  c += (epoch & 1u) * 2 * pba::PEER_MAXW;
  for (int i = 0; i < 2 * pba::PEER_MAXW; ++i) *reinterpret_cast<volatile unsigned*>(c + i) = 0u;
  __threadfence();
  *reinterpret_cast<volatile unsigned*>(g_peer.seq) = epoch;
}

// MUFU.RSQ / MUFU.RCP (<= 2 ulp) without the denormal fix-up code of rsqrtf() / 1.f / x; used only where no connection
// status depends on the result (Huber weight, Jacobians)
__device__ __forceinline__ float rsqrt_approx(float x) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
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

__device__ __forceinline__ float4 ldf4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// sm_100 256-bit global accesses (LDG.E.256 / STG.E.256): one instruction per 32-byte record.  The address must be
// 32-byte aligned.
__device__ __forceinline__ void ldg256_nc(const float4* p, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}
__device__ __forceinline__ void stg256_cs(float* p, const float (&v)[8]) {  // streaming (evict-first) store
  asm volatile("st.global.cs.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// row of a 3x4 matrix applied to [u, v, 1, rho] with the reference's association
// (A[:, :2] uv) + (A[:,2] + A[:,3] rho)   (camera_reproject.hpp:283-284,323-325), no FMA contraction
__device__ __forceinline__ float row_apply4(const float4 a, float u, float v, float rho) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a.x, u), __fmul_rn(a.y, v)), __fadd_rn(a.z, __fmul_rn(a.w, rho)));
}

// Evaluates one pattern pixel of one patch-residual.  All 8 lanes of a group must call it together.
//   FEJ : first-estimate Jacobians (production)   JAC : evaluate Jacobians
//   jv  : reprojection_jacobians_valid of the FEJ pass (ignored unless FEJ)
// The body is branch-free: when the residual is not evaluated every load goes to a safe texel and every
// output is selected to zero, so the warp never diverges around the shuffles.
template <bool FEJ, bool JAC>
__device__ __forceinline__ void eval_pixel(const PairConst& pc, const LandmarkIn& lm, bool jv,
                                           const float4* __restrict__ img, const uint8_t* __restrict__ mask, int W,
                                           int H, int status, float sigma, int huber, int lane, float pox, float poy,
                                           PixelOut& o) {
  const float xmax = (float)(W - 5), ymax = (float)(H - 5);
  const float ur = lm.u + pox, vr = lm.v + poy;  // pattern offsets of this lane's pixel, hoisted by the caller

  bool ok = valid_idepth(lm.rho) && in_roi(ur, vr, xmax, ymax);
  float tu, tv;
  float qx = 0.f, qy = 0.f, qz = 1.f, rho_j = 0.f;  // point used for the reprojection Jacobians
  if (FEJ || !JAC) {
    // values-only reprojection at the current state (camera_reproject.hpp:270-293); hnormalized() as one
    // correctly rounded reciprocal and two products (the fp32 oracle does the same)
    const float X = row_apply4(ldf4(pc.A + 0), ur, vr, lm.rho);
    const float Y = row_apply4(ldf4(pc.A + 4), ur, vr, lm.rho);
    const float Z = row_apply4(ldf4(pc.A + 8), ur, vr, lm.rho);
    ok = ok && (Z > 0.f);
    const float rz = __frcp_rn(Z);
    tu = __fmul_rn(X, rz);
    tv = __fmul_rn(Y, rz);
    ok = ok && in_roi(tu, tv, xmax, ymax);
    if (FEJ) ok = ok && jv;  // evaluate_jacobians.hpp:94
    if (FEJ && JAC) {
      // FEJ: the reprojection Jacobians are those of firstEstimateJacobians_, i.e. taken at the linearisation
      // pose and the snapshot idepth; they are recomputed here instead of being stored (112 scalars per residual)
      qx = row_apply4(ldf4(pc.M0 + 0), ur, vr, lm.rho0);
      qy = row_apply4(ldf4(pc.M0 + 4), ur, vr, lm.rho0);
      qz = row_apply4(ldf4(pc.M0 + 8), ur, vr, lm.rho0);
      rho_j = lm.rho0;
    }
  } else {
    // Jacobian variant at the current state (camera_reproject.hpp:305-367)
    qx = row_apply4(ldf4(pc.M + 0), ur, vr, lm.rho);
    qy = row_apply4(ldf4(pc.M + 4), ur, vr, lm.rho);
    qz = row_apply4(ldf4(pc.M + 8), ur, vr, lm.rho);
    rho_j = lm.rho;
    ok = ok && (qz > 0.f);
    const float rz = __frcp_rn(qz);
    tu = __fmul_rn(__fadd_rn(__fmul_rn(pc.fx_t, qx), __fmul_rn(pc.cx_t, qz)), rz);
    tv = __fmul_rn(__fadd_rn(__fmul_rn(pc.fy_t, qy), __fmul_rn(pc.cy_t, qz)), rz);
    ok = ok && in_roi(tu, tv, xmax, ymax);
  }
  // CameraMask::valid<false>: round() + lookup, only meaningful after the ROI test (quirk Q5).  mask == nullptr
  // means the frame's mask has no zero (checked once at upload): the lookup and its dependent load are skipped.
  if (mask) {
    ok = group_all(ok, lane);
    const int midx = ok ? (int)roundf(tv) * W + (int)roundf(tu) : 0;
    ok = ok && (mask[midx] != 0);
  }
  ok = group_all(ok, lane);
  const bool ev = ok && (status == K_OK);
  o.ok = ok;
  o.ev = ev;

  // interpolateLinear, features/include/features/camera/pixel_map.hpp:20-40 (texel (8,8) when not evaluated)
  const float su = ev ? tu : 8.f, sv = ev ? tv : 8.f;
  const int ix = (int)su, iy = (int)sv;
  const float dx = su - (float)ix, dy = sv - (float)iy;
  // The intensity, the residual and the patch energy feed the 75 % energy quantile of updatePointStatuses, i.e. connection
  // statuses: their operation sequence is pinned (explicit roundings / fused multiply-adds) and restated by the float
  // build of oracle/cpu_ref (`device_ops`), so that the bookkeeping can be compared bit for bit.
  const float dxdy = __fmul_rn(dx, dy);
  const float w11 = dxdy, w10 = __fsub_rn(dy, dxdy), w01 = __fsub_rn(dx, dxdy),
              w00 = __fadd_rn(__fsub_rn(__fsub_rn(1.f, dx), dy), dxdy);
  // the image is stored as 32-byte records {texel(x), texel(x + 1)}: the two horizontal taps of a row are ONE
  // 256-bit load (half the gather instructions and L1 wavefronts of four 128-bit loads)
  const float4* p = img + ((size_t)iy * W + ix) * 2;
  float4 t00, t01, t10, t11;
  ldg256_nc(p, t00, t01);
  ldg256_nc(p + 2 * (size_t)W, t10, t11);
  const float I = __fmaf_rn(w00, t00.x, __fmaf_rn(w01, t01.x, __fmaf_rn(w10, t10.x, __fmul_rn(w11, t11.x))));
  // r = (I_t - b_t) - s (patch - b_r), evaluate_jacobians.hpp:124-135
  const float r = ev ? __fmaf_rn(-pc.s, __fsub_rn(lm.patch, pc.b_r), __fsub_rn(I, pc.b_t)) : 0.f;
  const float n2 = group_sum(__fmul_rn(r, r));
  const float sig2 = __fmul_rn(sigma, sigma);
  float e = 0.5f * n2, wgt = 1.f;
  if (huber && n2 > sig2) {  // evaluate_jacobians.hpp:139-146
    wgt = sigma * rsqrt_approx(n2);  // weight: MUFU.RSQ (2 ulp), no status depends on it
    e = __fmaf_rn(sigma, __fsqrt_rn(n2), -(0.5f * sig2));  // energy: correctly rounded (it is compared with the quantile)
  }
  o.r = r;
  o.e = ev ? e : 0.f;
  o.w = wgt;
  if (JAC) {
    const float dIu = w11 * t11.y + w10 * t10.y + w01 * t01.y + w00 * t00.y;
    const float dIv = w11 * t11.z + w10 * t10.z + w01 * t01.z + w00 * t00.z;
    // camera_reproject.hpp:339-365
    const float sI = rcp_approx(ev ? qz : 1.f);  // Jacobians only (no status depends on it): MUFU.RCP
    const float b0 = qx * sI, b1 = qy * sI;
    const float nid = rho_j * sI;
    const float* tt = FEJ ? pc.t0 : pc.tr;
    const float evf = ev ? 1.f : 0.f;
    const float gu = evf * dIu * pc.fx_t, gv = evf * dIv * pc.fy_t;
    const float du_id = tt[0] * sI - tt[2] * sI * b0;
    const float dv_id = tt[1] * sI - tt[2] * sI * b1;
    const float b0b1 = b0 * b1;
    // Jg = dIv * dv/dxi + dIu * du/dxi   (evaluate_jacobians.hpp:149-157)
    o.g[0] = gu * nid;
    o.g[1] = gv * nid;
    o.g[2] = -gu * (nid * b0) - gv * (nid * b1);
    o.g[3] = -gu * b0b1 - gv * (b1 * b1 + 1.f);
    o.g[4] = gu * (b0 * b0 + 1.f) + gv * b0b1;
    o.g[5] = -gu * b1 + gv * b0;
    o.d = gu * du_id + gv * dv_id;  // evaluate_jacobians.hpp:165-174
    // corrected_reference_intensities: FEJ -> landmark.corrected_intensities (last target wins, Q1),
    // else s (patch - b_r)   (evaluate_jacobians.hpp:96,103-106)
    o.c = evf * (FEJ ? pc.s0_last * (lm.patch - pc.b_r0) : pc.s * (lm.patch - pc.b_r));
  } else {
#pragma unroll
    for (int k = 0; k < 6; ++k) o.g[k] = 0.f;
    o.d = 0.f;
    o.c = 0.f;
  }
}

__device__ __forceinline__ LandmarkIn load_landmark(const WindowDev& w, int gl, int px) {
  LandmarkIn lm;
  const float4 k = w.lmk[gl];
  lm.u = k.x;
  lm.v = k.y;
  lm.rho = k.z + w.idepth_step[gl];
  lm.rho0 = k.w;
  lm.patch = w.patch[(size_t)gl * 8 + px];
  lm.flags = w.flags[gl];
  return lm;
}

// ------------------------------------------------------------------------------------------------
// K6: firstEstimateJacobians_ (first_estimate_jacobians.hpp:14-71).  Freezes the idepth used by the FEJ
// Jacobians (quirk Q9: the CURRENT idepth) and stores reprojection_jacobians_valid per residual; the Jacobians
// themselves, corrected_intensities and brightness_change_scale are functions of the frozen point and are
// recomputed by the sweeps.  grid = (chunks of 32 landmarks, ordered pairs)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_first_estimate(const __grid_constant__ WindowDev w) {
  const int N = w.n_frames;
  const int r = blockIdx.y / (N - 1);
  int t = blockIdx.y % (N - 1);
  t += (t >= r);
  const int M = w.n_lm[r];
  const int l = blockIdx.x * 32 + (threadIdx.x >> 3);
  if ((int)blockIdx.x * 32 >= M) return;
  const int lane = threadIdx.x & 31, px = lane & 7;
  const bool inb = l < M;
  const int gl = lm_index(w, r, inb ? l : 0);
  const float4 k = w.lmk[gl];
  const int fl = w.flags[gl];
  const bool skip = !inb || ((fl & LM_MARG) && !(fl & LM_TO_MARG));  // first_estimate_jacobians.hpp:49
  const PairConst& pc = w.pairs[r * PBA_MAXF + t];
  const float xmax = (float)(w.W - 5), ymax = (float)(w.H - 5);
  const float ur = k.x + pat_x(px), vr = k.y + pat_y(px);
  const float rho0 = k.z;  // the snapshot taken by this pass
  const float qx = row_apply4(ldf4(pc.M0 + 0), ur, vr, rho0);
  const float qy = row_apply4(ldf4(pc.M0 + 4), ur, vr, rho0);
  const float qz = row_apply4(ldf4(pc.M0 + 8), ur, vr, rho0);
  bool ok = valid_idepth(rho0) && in_roi(ur, vr, xmax, ymax) && (qz > 0.f);
  const float rz = __frcp_rn(qz);
  const float u0 = __fmul_rn(__fadd_rn(__fmul_rn(pc.fx_t, qx), __fmul_rn(pc.cx_t, qz)), rz);
  const float v0 = __fmul_rn(__fadd_rn(__fmul_rn(pc.fy_t, qy), __fmul_rn(pc.cy_t, qz)), rz);
  ok = group_all(ok && in_roi(u0, v0, xmax, ymax), lane);
  if (!skip && px == 0) {
    w.jac_valid[res_index(w, r, t, l)] = ok ? 1 : 0;
    if (t == (r == 0 ? 1 : 0)) w.lmk[gl].w = rho0;  // one writer per landmark
  }
}

// ------------------------------------------------------------------------------------------------
// K2: residual-only sweep + energy reduction.  grid = (landmark chunks, host frames), one warp per target frame (the
// shape of the fused linearise): the chunk's landmark records are staged in shared memory once and shared by all
// target warps, the pair constants once per warp, and every CTA leaves ONE (energy, n_valid) partial -- ~6x fewer
// CTAs, prologues and partials than one CTA per (32 landmarks, ordered pair).
// ------------------------------------------------------------------------------------------------
struct __align__(16) LandmarkRec {  // 64 bytes
  float4 k;        // u, v, idepth, idepth at the FEJ point
  float patch[8];
  float step;
  int flags;
  float pad0, pad1;
};

template <bool FEJ>
__global__ void __launch_bounds__(480) k_residual_sweep(const __grid_constant__ WindowDev w, float sigma, int huber,
                                                        int lpb, double2* __restrict__ part,
                                                        const LmCtl* __restrict__ ctl, int ctl_mode) {
  if (lm_skip(ctl, ctl_mode)) return;
  extern __shared__ __align__(16) float smem[];
  // every CTA owns one slot of `part` (energy, n_valid): no same-address atomics, deterministic sum afterwards
  double2* my_part = part + (size_t)blockIdx.y * gridDim.x + blockIdx.x;
  const int N = w.n_frames;
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l0 = blockIdx.x * lpb;
  if (l0 >= M) {
    if (threadIdx.x == 0) *my_part = make_double2(0.0, 0.0);
    return;
  }
  const int nwarps = N - 1;
  PairConst* pcs = reinterpret_cast<PairConst*>(smem);                                   // [nwarps]
  LandmarkRec* recs = reinterpret_cast<LandmarkRec*>(smem + (size_t)nwarps * (sizeof(PairConst) / 4));  // [lpb]
  float* s_e = reinterpret_cast<float*>(recs + lpb);                                     // [nwarps]
  int* s_n = reinterpret_cast<int*>(s_e + nwarps);                                       // [nwarps]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = lane & 7, grp = lane >> 3;
  const int t = warp + (warp >= f);
  const int lm_base = lm_index(w, f, 0);
  reinterpret_cast<float4*>(&pcs[warp])[lane] = reinterpret_cast<const float4*>(&w.pairs[f * PBA_MAXF + t])[lane];
  for (int i = threadIdx.x; i < lpb; i += blockDim.x) {
    const int l = min(l0 + i, M - 1);
    const int gl = lm_base + l;
    LandmarkRec r;
    r.k = w.lmk[gl];
    const float4 p0 = ldf4(w.patch + (size_t)gl * 8), p1 = ldf4(w.patch + (size_t)gl * 8 + 4);
    r.patch[0] = p0.x, r.patch[1] = p0.y, r.patch[2] = p0.z, r.patch[3] = p0.w;
    r.patch[4] = p1.x, r.patch[5] = p1.y, r.patch[6] = p1.z, r.patch[7] = p1.w;
    r.step = w.idepth_step[gl];
    r.flags = w.flags[gl];
    r.pad0 = r.pad1 = 0.f;
    recs[i] = r;
  }
  __syncthreads();
  const PairConst& pc = pcs[warp];
  const float4* img = w.img[t];
  const uint8_t* mask = w.mask_all[t] ? nullptr : w.mask[t];
  const float pox = pat_x(px), poy = pat_y(px);
  const size_t res_base = res_index(w, f, t, 0);
  float e_acc = 0.f;
  int n_acc = 0;
  for (int it = 0; it < lpb; it += 4) {
    const int ls = it + grp;
    const int l = l0 + ls;
    const bool inb = ls < lpb && l < M;
    const LandmarkRec& rec = recs[inb ? ls : 0];
    LandmarkIn lm;
    lm.u = rec.k.x;
    lm.v = rec.k.y;
    lm.rho = rec.k.z + rec.step;
    lm.rho0 = rec.k.w;
    lm.patch = rec.patch[px];
    lm.flags = rec.flags;
    const bool skip = !inb || ((lm.flags & LM_MARG) && !(lm.flags & LM_TO_MARG));  // evaluate_jacobians.hpp:83
    const size_t res = res_base + (inb ? l : 0);
    const int status = skip ? K_OUTLIER : w.status[res];
    const bool jv = FEJ ? (w.jac_valid[res] != 0) : true;
    if (skip) lm.rho = -1.f;  // forces !ok
    PixelOut o;
    eval_pixel<FEJ, false>(pc, lm, jv, img, mask, w.W, w.H, status, sigma, huber, lane, pox, poy, o);
    if (!skip && px == 0) {
      if (!o.ok) w.cand[res] = K_OOB;   // evaluate_jacobians.hpp:111-113
      else if (o.ev) w.cand[res] = K_OK;  // :115
      w.energy[res] = o.e;
      if (!(lm.flags & LM_MARG)) {  // calculateLandmarksEnergy, problem.hpp:124-133
        e_acc += o.e;
        n_acc += o.e > 0.f;
      }
    }
  }
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
    for (int i = 0; i < nwarps; ++i) {
      e += (double)s_e[i];
      n += s_n[i];
    }
    *my_part = make_double2(e, (double)n);
  }
}

// second stage of the (energy, n_valid) and landmark-norm reductions: one CTA sums the per-CTA partials
__global__ void __launch_bounds__(1024) k_reduce_scal(const LmCtl* __restrict__ ctl, int ctl_mode,
                                                      const double2* __restrict__ e_part, int n_e,
                                                      const double2* __restrict__ n_part, int n_n,
                                                      double* __restrict__ scal, int core_frames, int peer_push) {
  KStamp kstamp_(9);
  if (lm_skip(ctl, ctl_mode)) return;
  const unsigned epoch = peer_push ? peer_epoch() : 0u;
  __shared__ double s[4][32];
  double a = 0, b = 0, c = 0, d = 0;
  for (int i = threadIdx.x; i < n_e; i += blockDim.x) {
    size_t idx = i;
    if (core_frames) {  // e_part is the per-pair core array: (energy, n_valid) in slots 44 / 45 of pair (r, t)
      const int r = i / (core_frames - 1);
      int t = i % (core_frames - 1);
      t += (t >= r);
      idx = ((size_t)(r * PBA_MAXF + t) * PBA_CORE + 44) / 2;
    }
    const double2 v = e_part[idx];
    a += v.x;
    b += v.y;
  }
  for (int i = threadIdx.x; i < n_n; i += blockDim.x) {
    const double2 v = n_part[i];
    c += v.x;
    d += v.y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(FULL, a, o);
    b += __shfl_xor_sync(FULL, b, o);
    c += __shfl_xor_sync(FULL, c, o);
    d += __shfl_xor_sync(FULL, d, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s[0][warp] = a;
    s[1][warp] = b;
    s[2][warp] = c;
    s[3][warp] = d;
  }
  __syncthreads();
  if (warp == 0) {
    const int nw = blockDim.x >> 5;
    a = lane < nw ? s[0][lane] : 0;
    b = lane < nw ? s[1][lane] : 0;
    c = lane < nw ? s[2][lane] : 0;
    d = lane < nw ? s[3][lane] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(FULL, a, o);
      b += __shfl_xor_sync(FULL, b, o);
      c += __shfl_xor_sync(FULL, c, o);
      d += __shfl_xor_sync(FULL, d, o);
    }
    if (lane == 0) {
      if (e_part) {
        red_store(&scal[0], a, peer_push, epoch);
        red_store(&scal[1], b, peer_push, epoch);
      }
      // the fused exchange always carries all four scalars (a consumer sums whatever is in the slot)
      if (n_part || peer_push) {
        red_store(&scal[2], c, peer_push, epoch);
        red_store(&scal[3], d, peer_push, epoch);
      }
    }
  }
  if (peer_push) peer_arrive(0, epoch, threadIdx.x == 0);
}

// ------------------------------------------------------------------------------------------------
// K1 (reference-surface mode): materialise every ResidualPoint.  Same grid as K2.
// Per patch-residual: 146 floats written (r[8], J_ref[8x8], J_tgt[8x8], d_idepth[8], w, e) + statuses.
// ------------------------------------------------------------------------------------------------
struct PairOrder {  // dispatch order of the ordered pairs (blockIdx.y -> reference, target)
  uint8_t r[PBA_MAXF * (PBA_MAXF - 1)], t[PBA_MAXF * (PBA_MAXF - 1)];
};

template <bool FEJ>
__global__ void __launch_bounds__(256, 6) k_materialise_sweep(const __grid_constant__ WindowDev w,
                                                           const __grid_constant__ PairOrder order, float sigma, int huber,
                                                           int subs) {
  __shared__ PairConst pcs;
  // Pair order (see launch_materialise_sweep): CTAs are dispatched in blockIdx order, so pairs that share a target
  // image -- and, inside a tile of targets, a reference frame's landmark arrays -- run back to back and stay in L2
  // instead of being evicted by the ~12 MB of Jacobians every pair streams out (r01f capture: 502 MB read in
  // reference-major order against 80 MB of images + landmarks)
  const int r = order.r[blockIdx.y], t = order.t[blockIdx.y];
  const int M = w.n_lm[r];
  if ((int)blockIdx.x * 32 * subs >= M) return;
  if (threadIdx.x < 32)
    reinterpret_cast<float4*>(&pcs)[threadIdx.x] = reinterpret_cast<const float4*>(&w.pairs[r * PBA_MAXF + t])[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int px = lane & 7;
  const float pox = pat_x(px), poy = pat_y(px);
  const float4* img = w.img[t];
  const uint8_t* mask = w.mask_all[t] ? nullptr : w.mask[t];
  const float* adj = FEJ ? pcs.adj0 : pcs.adj;
  const float sp = FEJ ? pcs.s0 : pcs.s;  // d_reference_affineBrightnessShift (:95,107)
  const int lm_base = lm_index(w, r, 0);
  const size_t res_base = res_index(w, r, t, 0);
  // a CTA walks `subs` consecutive groups of 32 landmarks of its pair: the pair constants are staged once
  // (fetching group k+1 while group k is sampled was tried: the extra registers cost more occupancy than the
  // saved round trip gains -- 0.51 -> 0.48 of the HBM roofline at 1.12 M units)
  for (int sub = 0; sub < subs; ++sub) {
    const int l = (blockIdx.x * subs + sub) * 32 + (threadIdx.x >> 3);
    if ((l & ~31) >= M) break;  // uniform over the CTA
    const bool inb = l < M;
    const int li = inb ? l : 0;
    LandmarkIn lm = load_landmark(w, lm_base + li, px);
    const bool skip = !inb || ((lm.flags & LM_MARG) && !(lm.flags & LM_TO_MARG));
    const size_t res = res_base + li;
    const int status = skip ? K_OUTLIER : w.status[res];
    const bool jv = FEJ ? (w.jac_valid[res] != 0) : true;
    if (skip) lm.rho = -1.f;
    PixelOut o;
    eval_pixel<FEJ, true>(pcs, lm, jv, img, mask, w.W, w.H, status, sigma, huber, lane, pox, poy, o);
    if (skip) continue;
    if (px == 0) {
      if (!o.ok) w.cand[res] = K_OOB;
      else if (o.ev) w.cand[res] = K_OK;
      w.energy[res] = o.e;
      if (o.ev) w.m_w[res] = o.w;  // huber_weight is left untouched when not evaluated (evaluate_jacobians.hpp:184-194)
    }
    __stcs(w.m_r + res * 8 + px, o.r);
    __stcs(w.m_did + res * 8 + px, o.d);
    float jr[8], jt[8];
#pragma unroll
    for (int j = 0; j < 6; ++j) jr[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {  // J_ref[:,0:6] = Jg Adj  (:162-163); rows of Adj as float2 pairs from shared memory
      const float2 a0 = *reinterpret_cast<const float2*>(adj + k * 6);
      const float2 a1 = *reinterpret_cast<const float2*>(adj + k * 6 + 2);
      const float2 a2 = *reinterpret_cast<const float2*>(adj + k * 6 + 4);
      jr[0] += o.g[k] * a0.x;
      jr[1] += o.g[k] * a0.y;
      jr[2] += o.g[k] * a1.x;
      jr[3] += o.g[k] * a1.y;
      jr[4] += o.g[k] * a2.x;
      jr[5] += o.g[k] * a2.y;
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) jt[j] = -o.g[j];  // J_tgt[:,0:6] = -Jg leftLog (= I)  (:159-160)
    jr[6] = o.c;
    jr[7] = o.ev ? sp : 0.f;
    jt[6] = -o.c;
    jt[7] = o.ev ? -1.f : 0.f;
    // one 256-bit streaming store per Jacobian row: a warp writes 1 KB contiguous per instruction
    stg256_cs(w.m_jref + res * 64 + px * 8, jr);
    stg256_cs(w.m_jtgt + res * 64 + px * 8, jt);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused linearise: K1 + K3 + the per-landmark half of K4, nothing materialised.
//
// With u_i = [g_i(6), c_i, 1] the two Jacobian rows of a pixel are  J_tgt_i = -u_i  and  J_ref_i = u_i B,
// B = blockdiag(Adj, 1, s').  So per ordered pair only the 8x8 "core" C = sum w u^T u (36 unique) and
// q = sum w u r (8) are accumulated (44 sums instead of 3*64+16 = 208); H_rr = B^T C B, H_rt = -B^T C,
// H_tt = C, b_r = B^T q, b_t = -q are formed once per pair by k_assemble in fp64.  Likewise the target block of a
// landmark's H_pd is -p_t (p_t = sum w d u) and its reference block is sum_t B_t^T p_t, formed once per landmark
// when the chunk is finalised.
//
// grid = (landmark chunks, host frames); one warp per target frame; each warp walks the chunk 4 landmarks at a
// time keeping its pair's 44 running sums in registers, so the per-landmark quantities that couple the targets
// (H_pd, H_dd, b_d) meet in shared memory.
// ------------------------------------------------------------------------------------------------
template <bool FEJ, int NWMAX, int MINB, bool PREFETCH>
__global__ void __launch_bounds__(32 * NWMAX, MINB)
    k_linearize_fused(const __grid_constant__ WindowDev w, float sigma, int huber, int for_marg, int lpb,
                      float* __restrict__ core_part, float* __restrict__ schur_part, const LmCtl* __restrict__ ctl,
                      int ctl_mode) {
  if (lm_skip(ctl, ctl_mode)) return;
  extern __shared__ __align__(16) float smem[];
  const int N = w.n_frames;
  const int D = 8 * N;
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l0 = blockIdx.x * lpb;
  if (l0 >= M) return;
  const int nwarps = N - 1;
  PairConst* pcs = reinterpret_cast<PairConst*>(smem);               // [nwarps]
  float* hpd_s = smem + (size_t)nwarps * (sizeof(PairConst) / 4);    // [lpb][D]
  float* hdd_s = hpd_s + lpb * D;                                    // [lpb] Schur weight 1/H_dd after the finalise
  float* bd_s = hdd_s + lpb;                                         // [lpb] weight * b_d
  float* hdd_w = bd_s + lpb;                                         // [lpb][nwarps] per-target H_dd terms
  float* bd_w = hdd_w + lpb * nwarps;                                // [lpb][nwarps] per-target b_d terms
  // the chunk's landmark records, staged once and shared by the N - 1 target warps (16-byte aligned: every term above
  // is a multiple of 4 floats when lpb is)
  LandmarkRec* recs = reinterpret_cast<LandmarkRec*>(bd_w + lpb * nwarps + ((4 - ((lpb * (D + 2 + 2 * nwarps)) & 3)) & 3));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int px = lane & 7, grp = lane >> 3;
  const int t = warp + (warp >= f);
  const int lm_base = lm_index(w, f, 0);

  for (int i = threadIdx.x; i < lpb * (D + 2 + 2 * nwarps); i += blockDim.x) hpd_s[i] = 0.f;
  reinterpret_cast<float4*>(&pcs[warp])[lane] = reinterpret_cast<const float4*>(&w.pairs[f * PBA_MAXF + t])[lane];
  for (int i = threadIdx.x; i < lpb; i += blockDim.x) {
    const int gl = lm_base + min(l0 + i, M - 1);
    LandmarkRec r;
    r.k = w.lmk[gl];
    const float4 p0 = ldf4(w.patch + (size_t)gl * 8), p1 = ldf4(w.patch + (size_t)gl * 8 + 4);
    r.patch[0] = p0.x, r.patch[1] = p0.y, r.patch[2] = p0.z, r.patch[3] = p0.w;
    r.patch[4] = p1.x, r.patch[5] = p1.y, r.patch[6] = p1.z, r.patch[7] = p1.w;
    r.step = w.idepth_step[gl];
    r.flags = w.flags[gl];
    r.pad0 = r.pad1 = 0.f;
    recs[i] = r;
  }
  __syncthreads();
  const PairConst& pc = pcs[warp];
  const float4* img = w.img[t];
  const uint8_t* mask = w.mask_all[t] ? nullptr : w.mask[t];
  const float pox = pat_x(px), poy = pat_y(px);
  const size_t res_base = res_index(w, f, t, 0);

  float acc[44];  // 36 (upper triangle of the core) + 8 (core^T r); padded to PBA_CORE only for the final reduction
#pragma unroll
  for (int k = 0; k < 44; ++k) acc[k] = 0.f;
  // calculateLandmarksEnergy of this pair's residuals (problem.hpp:124-133) rides in the two spare slots of the core
  // record, so that a linearisation at a trial state also IS the energy evaluation there (device LM, speculative mode)
  float e_acc = 0.f, n_acc = 0.f;

  for (int it = 0; it < lpb; it += 4) {
    const int ls = it + grp;  // slot in the chunk
    const int l = l0 + ls;
    const bool inb = l < M;
    const int li = inb ? l : 0;
    if (PREFETCH && it + 4 < lpb) {
      // The warp stalls longest on the L2 gathers of the bilinear taps (r01f capture: the first consumer of the taps
      // holds ~20 % of the stall samples).  The NEXT group's records are already in shared memory, so its tap
      // addresses cost a few FMAs: prefetch those two 32-byte records into L1 while this group's arithmetic runs.
      const LandmarkRec& nx = recs[min(ls + 4, lpb - 1)];
      const float un = nx.k.x + pox, vn = nx.k.y + poy, rn = nx.k.z + nx.step;
      const float Xn = pc.A[0] * un + pc.A[1] * vn + (pc.A[2] + pc.A[3] * rn);
      const float Yn = pc.A[4] * un + pc.A[5] * vn + (pc.A[6] + pc.A[7] * rn);
      const float Zn = pc.A[8] * un + pc.A[9] * vn + (pc.A[10] + pc.A[11] * rn);
      const float rzn = rcp_approx(Zn);
      const int ixn = min(max((int)(Xn * rzn), 0), w.W - 2), iyn = min(max((int)(Yn * rzn), 0), w.H - 2);
      const float4* pn = img + ((size_t)iyn * w.W + ixn) * 2;
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pn));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + 2 * (size_t)w.W));
    }
    const LandmarkRec& rec = recs[inb ? ls : 0];
    LandmarkIn lm;
    lm.u = rec.k.x;
    lm.v = rec.k.y;
    lm.rho = rec.k.z + rec.step;
    lm.rho0 = rec.k.w;
    lm.patch = rec.patch[px];
    lm.flags = rec.flags;
    const bool skip = !inb || ((lm.flags & LM_MARG) && !(lm.flags & LM_TO_MARG));
    const size_t res = res_base + li;
    const int status = skip ? K_OUTLIER : w.status[res];
    const bool jv = FEJ ? (w.jac_valid[res] != 0) : true;
    if (skip) lm.rho = -1.f;
    PixelOut o;
    eval_pixel<FEJ, true>(pc, lm, jv, img, mask, w.W, w.H, status, sigma, huber, lane, pox, poy, o);
    if (!skip && px == 0) {
      if (!o.ok) w.cand[res] = K_OOB;
      else if (o.ev) w.cand[res] = K_OK;
      w.energy[res] = o.e;
      if (!(lm.flags & LM_MARG)) {
        e_acc += o.e;
        n_acc += o.e > 0.f ? 1.f : 0.f;
      }
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
    // per landmark: target block of H_pd, H_dd, b_d  (hessian_block_evaluation.hpp:198-212)
    float pv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pv[k] = wu[k] * o.d;
    const float P = group_transpose_reduce(pv, px);  // lane px holds sum_i w d_i u_i[px]
    const float wd = wgt * o.d;
    const float hdd = group_sum(wd * o.d);
    const float bd = group_sum(wd * o.r);
    if (sel) {
      hpd_s[ls * D + 8 * t + px] = -P;  // this warp is the only writer of target block t
      if (px == 0) {  // own slot per target warp, summed in a fixed order below: bit-reproducible runs
        hdd_w[ls * nwarps + warp] = hdd;
        bd_w[ls * nwarps + warp] = bd;
      }
    }
  }

  // warp-wide transpose-reduce of the 44(48) running sums: 24+12+6+3+2 = 47 shuffles, then <= 2 stores per lane
  {
    float v48[PBA_CORE], v24[24], v12[12], v6[6], v3[3], v2[2];
#pragma unroll
    for (int k = 0; k < PBA_CORE; ++k) v48[k] = k < 44 ? acc[k] : (k == 44 ? e_acc : (k == 45 ? n_acc : 0.f));
    tr_step<48, 16>(v48, v24, lane);
    tr_step<24, 8>(v24, v12, lane);
    tr_step<12, 4>(v12, v6, lane);
    tr_step<6, 2>(v6, v3, lane);
    tr_step<3, 1>(v3, v2, lane);
    const int off = ((lane & 16) ? 24 : 0) + ((lane & 8) ? 12 : 0) + ((lane & 4) ? 6 : 0) + ((lane & 2) ? 3 : 0) +
                    ((lane & 1) ? 2 : 0);
    // this warp's 48 partial sums go to its own slot [host frame][chunk][target warp][48]: plain coalesced stores,
    // summed over the chunks in fp64 by k_assemble (no atomics, deterministic)
    float* dst = core_part + (((size_t)f * gridDim.x + blockIdx.x) * nwarps + warp) * PBA_CORE;
    dst[off] = v2[0];
    if (!(lane & 1)) dst[off + 1] = v2[1];
  }
  __syncthreads();

  // reference block of H_pd:  sum_t B_t^T p_t  with p_t = -(target block t)   (J_ref = U B)
  for (int i = threadIdx.x; i < lpb * 8; i += blockDim.x) {
    const int ls = i >> 3, j = i & 7;
    float refv = 0.f;
    for (int wi = 0; wi < nwarps; ++wi) {
      const int tt = wi + (wi >= f);
      const float* pt = hpd_s + ls * D + 8 * tt;
      const PairConst& pw = pcs[wi];
      if (j < 6) {
        const float* adj = FEJ ? pw.adj0 : pw.adj;
#pragma unroll
        for (int k = 0; k < 6; ++k) refv -= adj[k * 6 + j] * pt[k];
      } else if (j == 6) {
        refv -= pt[6];
      } else {
        refv -= (FEJ ? pw.s0 : pw.s) * pt[7];
      }
    }
    hpd_s[ls * D + 8 * f + j] = refv;
  }
  __syncthreads();

  // finalise the chunk's landmarks (hessian_block_evaluation.hpp:213-227); hdd_s / bd_s are overwritten with the
  // Schur weights  s = 1/H_dd (0 when not selected or ill-conditioned)  and  s * b_d
  for (int ls = threadIdx.x; ls < lpb; ls += blockDim.x) {
    const int l = l0 + ls;
    float sc = 0.f, sb = 0.f;
    if (l < M) {
      const int gl = lm_base + l;
      const int fl = w.flags[gl];
      const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
      const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
      if (sel) {
        float hdd = 0.f, bd = 0.f;
        for (int wi = 0; wi < nwarps; ++wi) {
          hdd += hdd_w[ls * nwarps + wi];
          bd += bd_w[ls * nwarps + wi];
        }
        w.b_d[gl] = bd;
        if (hdd > 1e-15f) {
          if (for_marg && w.fixed[f]) hdd += 1e8f;  // kScaleNullspaceRegularizer
          sc = 1.f / hdd;
          sb = sc * bd;
          w.inv_hdd[gl] = sc;
          w.flags[gl] = (uint8_t)(fl & ~LM_ILL);
        } else {
          w.flags[gl] = (uint8_t)(fl | LM_ILL);
        }
      }
    }
    hdd_s[ls] = sc;
    bd_s[ls] = sb;
  }
  __syncthreads();
  const int D4 = D / 4;
  for (int i = threadIdx.x; i < lpb * D4; i += blockDim.x) {
    const int ls = i / D4;
    const int l = l0 + ls;
    if (l >= M) break;
    const int gl = lm_base + l;
    const int fl = w.flags[gl];
    const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
    const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
    if (sel)
      reinterpret_cast<float4*>(w.hpd + (size_t)gl * w.hpd_stride)[i - ls * D4] =
          reinterpret_cast<const float4*>(hpd_s + ls * D)[i - ls * D4];
  }
  // K4 second half for this chunk while its H_pd rows are still in shared memory:
  //   S = sum_l s_l H_pd_l H_pd_l^T (upper triangle as 4x4 tiles),  b = sum_l s_l b_d_l H_pd_l
  // stored as this CTA's partial; k_finish_fused adds the CTAs up in fp64.
  {
    const int T4 = D / 4, ntri = T4 * (T4 + 1) / 2, nout = ntri * 16 + D;
    float* part = schur_part + ((size_t)f * gridDim.x + blockIdx.x) * nout;
    for (int tile = threadIdx.x; tile < ntri; tile += blockDim.x) {
      int ty = 0, rem = tile;
      while (rem >= T4 - ty) {
        rem -= T4 - ty;
        ++ty;
      }
      const int tx = ty + rem;
      float a[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = 0.f;
      for (int ls = 0; ls < lpb; ++ls) {
        const float sc = hdd_s[ls];
        const float4 qv = ldf4(hpd_s + ls * D + 4 * ty);
        const float4 pv = ldf4(hpd_s + ls * D + 4 * tx);
        const float qa[4] = {sc * qv.x, sc * qv.y, sc * qv.z, sc * qv.w};
        const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i * 4 + j] += qa[i] * pa[j];
      }
      float4* dst = reinterpret_cast<float4*>(part + tile * 16);
      dst[0] = make_float4(a[0], a[1], a[2], a[3]);
      dst[1] = make_float4(a[4], a[5], a[6], a[7]);
      dst[2] = make_float4(a[8], a[9], a[10], a[11]);
      dst[3] = make_float4(a[12], a[13], a[14], a[15]);
    }
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float b = 0.f;
      for (int ls = 0; ls < lpb; ++ls) b += bd_s[ls] * hpd_s[ls * D + c];
      part[ntri * 16 + c] = b;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused linearise, second generation (round 2): ONE THREAD PER PATCH-RESIDUAL.
//
// The first generation spreads a patch over 8 lanes (lane = pattern pixel): every per-residual quantity is computed 8
// times, every 8-pixel sum costs shuffles, and the kernel executes ~670 warp instructions per 4 residuals.  Here a lane owns
// a landmark and walks the 8 pattern pixels itself: the patch sums, the Huber weight and the per-landmark H_pd / H_dd / b_d
// terms are plain register arithmetic, the per-residual work is done once, and the pixels of one residual give the
// thread 8 independent gathers to keep in flight (instruction-level parallelism instead of occupancy).
//
//   grid  = (chunks of `lpb` landmarks -- a multiple of 32 --, host frames), block = 32 (N - 1): one warp per target
//   pass 1 (per pixel)  reproject with the current state (pinned operation order: the status predicates),
//                       gather the four taps (two LDG.E.256), intensity / gradient, raw residual
//   between             all-pixel validity, mask, committed status -> evaluated?, patch norm, Huber weight, energy
//   pass 2 (per pixel)  reprojection Jacobians at the (FEJ) linearisation point, u = [g(6), c, 1], running sums of the
//                       pair's 8x8 core / core^T r in 44 registers, per-landmark p_t = sum w d u, H_dd, b_d in registers
// Everything after the sweep (reference block of H_pd, 1x1 Schur inverse, H_pd rows to HBM, the chunk's rank-k update)
// is the first generation's epilogue on the same shared-memory layout, so the second-stage kernels are unchanged.
// The arithmetic that decides statuses and energies is the same sequence of rounded operations as in eval_pixel.
// ------------------------------------------------------------------------------------------------
template <bool FEJ, int NWMAX, int MINB>
__global__ void __launch_bounds__(32 * NWMAX, MINB)
    k_linearize_fused2(const __grid_constant__ WindowDev w, float sigma, int huber, int for_marg, int lpb,
                       float* __restrict__ core_part, float* __restrict__ schur_part, const LmCtl* __restrict__ ctl,
                       int ctl_mode, int fold, const double* __restrict__ step_pose, double2* __restrict__ norms, int epi) {
  KStamp kstamp_(0);
  // epi = 1 (round 2, default): second-generation epilogue -- the reference block of H_pd is formed per target inside the
  // sweep (B_t^T p_t from the registers that hold p_t) and only summed over the targets afterwards, the landmark selection is
  // staged in shared memory once, and the H_pd row store, the chunk's rank-k update and its b vector run side by side in one
  // phase (two block barriers after the sweep instead of four).  epi = 0: the first generation's epilogue (A/B).
  // fold (device LM, speculative sequence): this sweep evaluates the trial state of step k + 1, so it first closes
  // step k for ITS landmarks and residuals -- acceptStep() / rejectStep() incl. changeResidualStatuses (problem.hpp:20-35,
  // 377-384,395-399; k_accept_landmarks) -- and then back-substitutes the new pose step into their inverse depths
  // (calculateIdepths, hessian_block_evaluation.hpp:238-263; k_back_substitute).  Two launches and their boundaries
  // leave the critical path of an iteration; every landmark / residual is touched by exactly one CTA / thread.
  const int N = w.n_frames;
  const int D = 8 * N;
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l0 = blockIdx.x * lpb;
  const int nwarps = N - 1;
  const int lm_base0 = lm_index(w, f, 0);
  cudaGridDependencySynchronize();  // programmatic dependent launch: everything above ran beside the predecessor's tail
  if (fold && ctl->apply) {
    const int accept = ctl->accept;
    const int tw = threadIdx.x >> 5;
    const int tt = tw + (tw >= f);
    const size_t rb0 = res_index(w, f, tt, 0);
    for (int ls = threadIdx.x & 31; ls < lpb; ls += 32) {
      const int l = l0 + ls;
      if (l >= M) break;
      if (accept > 0) w.status[rb0 + l] = w.cand[rb0 + l];
      else w.cand[rb0 + l] = w.status[rb0 + l];
    }
    for (int ls = threadIdx.x; ls < lpb; ls += blockDim.x) {
      const int l = l0 + ls;
      if (l >= M) break;
      const int gl = lm_base0 + l;
      if (accept > 0) w.lmk[gl].z += w.idepth_step[gl];
      w.idepth_step[gl] = 0.f;
    }
  }
  if (lm_skip(ctl, ctl_mode)) return;
  extern __shared__ __align__(16) float smem[];
  if (fold) {
    float* sp = smem;  // the pose step, staged in the (not yet used) dynamic shared memory
    __syncthreads();   // the commits above
    for (int i = threadIdx.x; i < D; i += blockDim.x) sp[i] = (float)step_pose[i];
    __syncthreads();
    const float inv_lambda = (float)(1.0 / (1.0 + ctl->lambda));
    double n_state = 0, n_step = 0;
    const int px = threadIdx.x & 7;
    for (int lsb = 0; lsb < lpb; lsb += blockDim.x >> 3) {  // 8 lanes per landmark, as k_back_substitute
      const int ls = lsb + (threadIdx.x >> 3);
      const int l = l0 + ls;
      const bool inb = ls < lpb && l < M;
      const int gl = lm_base0 + (inb ? l : 0);
      float dot = 0.f;
      if (inb) {
        const float* hrow = w.hpd + (size_t)gl * w.hpd_stride;
        for (int c = px; c < D; c += 8) dot += hrow[c] * sp[c];
      }
      dot = group_sum(dot);
      if (inb && px == 0) {
        const int fl = w.flags[gl];
        float stp = w.idepth_step[gl];
        if (!(fl & LM_MARG) && !(fl & LM_ILL)) {
          stp = -((w.b_d[gl] - dot) * inv_lambda * w.inv_hdd[gl]);
          w.idepth_step[gl] = stp;
        }
        const float id = w.lmk[gl].z;  // landmark part of acceptStep's norms (problem.hpp:377-382)
        n_state += (double)id * id;
        n_step += (double)stp * stp;
      }
    }
#pragma unroll
    for (int sft = 16; sft > 0; sft >>= 1) {
      n_state += __shfl_xor_sync(FULL, n_state, sft);
      n_step += __shfl_xor_sync(FULL, n_step, sft);
    }
    __syncthreads();  // sp is dead; the same words now collect the warps' norm partials
    double* nw = reinterpret_cast<double*>(smem);
    if ((threadIdx.x & 31) == 0) {
      nw[2 * (threadIdx.x >> 5)] = n_state;
      nw[2 * (threadIdx.x >> 5) + 1] = n_step;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
        a += nw[2 * i];
        b += nw[2 * i + 1];
      }
      norms[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = make_double2(a, b);
    }
    __syncthreads();  // idepth_step of this chunk is final; smem is free for the sweep
  }
  if (l0 >= M) return;
  PairConst* pcs = reinterpret_cast<PairConst*>(smem);               // [nwarps]
  float* hpd_s = smem + (size_t)nwarps * (sizeof(PairConst) / 4);    // [lpb][D]
  float* hdd_s = hpd_s + lpb * D;                                    // [lpb] Schur weight 1/H_dd after the finalise
  float* bd_s = hdd_s + lpb;                                         // [lpb] weight * b_d
  float* hdd_w = bd_s + lpb;                                         // [lpb][nwarps] per-target H_dd terms
  float* bd_w = hdd_w + lpb * nwarps;                                // [lpb][nwarps] per-target b_d terms
  float* ref_w = bd_w + lpb * nwarps;                                // epi: [nwarps][lpb][8] per-target B_t^T p_t
  int* fl_s = reinterpret_cast<int*>(ref_w + (size_t)nwarps * lpb * 8);  // epi: [lpb] landmark flags (-1: past the end)
  // epi & 2: the 44 + 2 per-lane sums of the landmark groups are accumulated in shared memory ([warp][46][32], conflict-free)
  // and reduced over the lanes ONCE per CTA instead of once per group (47 shuffles and ~200 selects each)
  float* racc = reinterpret_cast<float*>(fl_s + lpb);
  const bool ronce = (epi & 2) && lpb > 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = warp + (warp >= f);
  const int lm_base = lm_index(w, f, 0);

  stamp(30);
  cta_stamp(0);
  cta_stamp(3);
  for (int i = threadIdx.x; i < lpb * (D + 2 + 2 * nwarps); i += blockDim.x) hpd_s[i] = 0.f;
  reinterpret_cast<float4*>(&pcs[warp])[lane] = reinterpret_cast<const float4*>(&w.pairs[f * PBA_MAXF + t])[lane];
  __syncthreads();
  stamp(31);
  const PairConst& pc = pcs[warp];
  const float4* __restrict__ img = w.img[t];
  const uint8_t* __restrict__ mask = w.mask_all[t] ? nullptr : w.mask[t];
  const size_t res_base = res_index(w, f, t, 0);
  const int W = w.W;
  const float xmax = (float)(w.W - 5), ymax = (float)(w.H - 5);
  const float sig2 = __fmul_rn(sigma, sigma);
  // where this lane's two slots of the warp-reduced 48-vector live (see the transpose-reduce below)
  const int off = ((lane & 16) ? 24 : 0) + ((lane & 8) ? 12 : 0) + ((lane & 4) ? 6 : 0) + ((lane & 2) ? 3 : 0) +
                  ((lane & 1) ? 2 : 0);
  // running totals of the pair's core record: only TWO registers per lane live across the landmark groups (the 44 sums
  // of a group are reduced over the warp right after its pass 2, so they do not occupy registers during the gathers)
  float tot0 = 0.f, tot1 = 0.f;

  for (int g0 = 0; g0 < lpb; g0 += 32) {
    const int ls = g0 + lane;  // slot in the chunk
    const int l = l0 + ls;
    const bool inb = l < M;
    const int gl = lm_base + (inb ? l : 0);
    const float4 k4 = w.lmk[gl];
    const int flags = w.flags[gl];
    const bool skip = !inb || ((flags & LM_MARG) && !(flags & LM_TO_MARG));  // evaluate_jacobians.hpp:83
    const size_t res = res_base + (inb ? l : 0);
    const int status = skip ? K_OUTLIER : w.status[res];
    const bool jv = FEJ ? (w.jac_valid[res] != 0) : true;
    const float u = k4.x, v = k4.y, rho0 = k4.w;
    const float rho = skip ? -1.f : k4.z + w.idepth_step[gl];  // -1 forces !ok
    const float* __restrict__ patch = w.patch + (size_t)gl * 8;

    // ---- pass 1: reprojection at the current state, taps, raw residuals -------------------------------------------------
    // row . [u, v, 1, rho] = (a0 u + a1 v) + (a2 + a3 rho): the second bracket is the same for the 8 pixels
    const float* PA = (FEJ) ? pc.A : pc.M;  // non-FEJ: the Jacobian variant projects through M and K_t (see eval_pixel)
    const float c0 = __fadd_rn(PA[2], __fmul_rn(PA[3], rho));
    const float c1 = __fadd_rn(PA[6], __fmul_rn(PA[7], rho));
    const float c2 = __fadd_rn(PA[10], __fmul_rn(PA[11], rho));
    bool ok = valid_idepth(rho);
    float rr[8], gu[8], gv[8];
    // two batches of four pixels: addresses first, then the eight 256-bit gathers back to back, then the arithmetic --
    // the thread keeps 8 independent L2 round trips in flight instead of one
#pragma unroll
    for (int h4 = 0; h4 < 8; h4 += 4) {
      float fdx[4], fdy[4];
      const float4* tp[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int p = h4 + q4;
        const float ox = (float)((int)((0x21420312u >> (4 * p)) & 15u) - 2), oy = (float)((int)((0x01222334u >> (4 * p)) & 15u) - 2);
        const float ur = u + ox, vr = v + oy;
        bool okp = in_roi(ur, vr, xmax, ymax);
        const float X = __fadd_rn(__fadd_rn(__fmul_rn(PA[0], ur), __fmul_rn(PA[1], vr)), c0);
        const float Y = __fadd_rn(__fadd_rn(__fmul_rn(PA[4], ur), __fmul_rn(PA[5], vr)), c1);
        const float Z = __fadd_rn(__fadd_rn(__fmul_rn(PA[8], ur), __fmul_rn(PA[9], vr)), c2);
        okp = okp && (Z > 0.f);
        const float rz = __frcp_rn(Z);
        float tu, tv;
        if (FEJ) {
          tu = __fmul_rn(X, rz);
          tv = __fmul_rn(Y, rz);
        } else {
          tu = __fmul_rn(__fadd_rn(__fmul_rn(pc.fx_t, X), __fmul_rn(pc.cx_t, Z)), rz);
          tv = __fmul_rn(__fadd_rn(__fmul_rn(pc.fy_t, Y), __fmul_rn(pc.cy_t, Z)), rz);
        }
        okp = okp && in_roi(tu, tv, xmax, ymax);
        if (mask) {  // CameraMask::valid<false>: round() + lookup, only meaningful after the ROI test (quirk Q5)
          const int midx = okp ? (int)roundf(tv) * W + (int)roundf(tu) : 0;
          okp = okp && (mask[midx] != 0);
        }
        ok = ok && okp;
        // gather speculatively (a pixel that failed goes to texel (8, 8)); whether the residual is evaluated at all is
        // only known after the 8th pixel, and an unevaluated one contributes through weights that are exactly zero
        const float su = okp ? tu : 8.f, sv = okp ? tv : 8.f;
        const int ix = (int)su, iy = (int)sv;
        fdx[q4] = su - (float)ix;
        fdy[q4] = sv - (float)iy;
        tp[q4] = img + ((size_t)iy * W + ix) * 2;
      }
      float4 t00[4], t01[4], t10[4], t11[4];
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        ldg256_nc(tp[q4], t00[q4], t01[q4]);
        ldg256_nc(tp[q4] + 2 * (size_t)W, t10[q4], t11[q4]);
      }
      const float4 pq = ldf4(patch + h4);
      const float pv4[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        const int p = h4 + q4;
        const float dx = fdx[q4], dy = fdy[q4];
        const float dxdy = __fmul_rn(dx, dy);
        const float w11 = dxdy, w10 = __fsub_rn(dy, dxdy), w01 = __fsub_rn(dx, dxdy),
                    w00 = __fadd_rn(__fsub_rn(__fsub_rn(1.f, dx), dy), dxdy);
        const float I = __fmaf_rn(w00, t00[q4].x, __fmaf_rn(w01, t01[q4].x, __fmaf_rn(w10, t10[q4].x, __fmul_rn(w11, t11[q4].x))));
        rr[p] = __fmaf_rn(-pc.s, __fsub_rn(pv4[q4], pc.b_r), __fsub_rn(I, pc.b_t));  // evaluate_jacobians.hpp:124-135
        gu[p] = (w11 * t11[q4].y + w10 * t10[q4].y + w01 * t01[q4].y + w00 * t00[q4].y) * pc.fx_t;
        gv[p] = (w11 * t11[q4].z + w10 * t10[q4].z + w01 * t01[q4].z + w00 * t00[q4].z) * pc.fy_t;
      }
    }
    if (FEJ) ok = ok && jv;  // evaluate_jacobians.hpp:94
    const bool ev = ok && (status == K_OK);
    if (g0 == 0) stamp(32);
    // patch norm in the 8-lane butterfly order of the first generation (and of oracle/cpu_ref `device_ops`)
    float q[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      rr[p] = ev ? rr[p] : 0.f;
      gu[p] = ev ? gu[p] : 0.f;
      gv[p] = ev ? gv[p] : 0.f;
      q[p] = __fmul_rn(rr[p], rr[p]);
    }
    const float n2 = __fadd_rn(__fadd_rn(__fadd_rn(q[0], q[4]), __fadd_rn(q[2], q[6])),
                               __fadd_rn(__fadd_rn(q[1], q[5]), __fadd_rn(q[3], q[7])));
    float e = 0.5f * n2, wgt = 1.f;
    if (huber && n2 > sig2) {  // evaluate_jacobians.hpp:139-146
      wgt = sigma * rsqrt_approx(n2);
      e = __fmaf_rn(sigma, __fsqrt_rn(n2), -(0.5f * sig2));
    }
    e = ev ? e : 0.f;
    float e_add = 0.f, n_add = 0.f;
    if (!skip) {
      if (!ok) w.cand[res] = K_OOB;       // evaluate_jacobians.hpp:111-113
      else if (ev) w.cand[res] = K_OK;    // :115
      w.energy[res] = e;
      if (!(flags & LM_MARG)) {           // calculateLandmarksEnergy, problem.hpp:124-133
        e_add = e;
        n_add = e > 0.f ? 1.f : 0.f;
      }
    }
    // landmark selection of K3/K4 (hessian_block_evaluation.hpp:68-72,190-194)
    const bool sel = !skip && (for_marg ? (flags & LM_TO_MARG) != 0 : (flags & LM_MARG) == 0);
    const float wq = (sel && ev) ? wgt : 0.f;

    // ---- pass 2: Jacobians at the linearisation point, the group's sums ---------------------------------------------------
    const float* PM = FEJ ? pc.M0 : pc.M;
    const float rho_j = FEJ ? rho0 : rho;
    const float* tt = FEJ ? pc.t0 : pc.tr;
    const float m0 = PM[2] + PM[3] * rho_j, m1 = PM[6] + PM[7] * rho_j, m2 = PM[10] + PM[11] * rho_j;
    const float cs = ev ? (FEJ ? pc.s0_last : pc.s) : 0.f, cb = FEJ ? pc.b_r0 : pc.b_r;
    float acc[44];
#pragma unroll
    for (int k = 0; k < 44; ++k) acc[k] = 0.f;
    float pt[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pt[k] = 0.f;
    float hdd = 0.f, bd = 0.f;
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const float ox = (float)((int)((0x21420312u >> (4 * p)) & 15u) - 2), oy = (float)((int)((0x01222334u >> (4 * p)) & 15u) - 2);
      const float ur = u + ox, vr = v + oy;
      const float qx = PM[0] * ur + PM[1] * vr + m0;
      const float qy = PM[4] * ur + PM[5] * vr + m1;
      const float qz = PM[8] * ur + PM[9] * vr + m2;
      const float sI = rcp_approx(ev ? qz : 1.f);  // Jacobians only (no status depends on it): MUFU.RCP
      const float b0 = (ev ? qx : 0.f) * sI, b1 = (ev ? qy : 0.f) * sI;
      const float nid = rho_j * sI;
      const float gup = gu[p], gvp = gv[p];
      const float du_id = tt[0] * sI - tt[2] * sI * b0;
      const float dv_id = tt[1] * sI - tt[2] * sI * b1;
      const float b0b1 = b0 * b1;
      float uu[8];
      uu[0] = gup * nid;
      uu[1] = gvp * nid;
      uu[2] = -gup * (nid * b0) - gvp * (nid * b1);
      uu[3] = -gup * b0b1 - gvp * (b1 * b1 + 1.f);
      uu[4] = gup * (b0 * b0 + 1.f) + gvp * b0b1;
      uu[5] = -gup * b1 + gvp * b0;
      uu[6] = cs * (patch[p] - cb);
      uu[7] = 1.f;
      const float d = gup * du_id + gvp * dv_id;  // evaluate_jacobians.hpp:165-174
      float wu[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) wu[k] = wq * uu[k];
      int idx = 0;
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = a; b < 8; ++b) acc[idx++] += wu[a] * uu[b];
#pragma unroll
      for (int a = 0; a < 8; ++a) acc[36 + a] += wu[a] * rr[p];
#pragma unroll
      for (int k = 0; k < 8; ++k) pt[k] += wu[k] * d;
      const float wd = wq * d;
      hdd += wd * d;
      bd += wd * rr[p];
    }
    if (g0 == 0) stamp(33);
    if (sel) {  // this thread is the only writer of target block t of its landmark
      float4* hp = reinterpret_cast<float4*>(hpd_s + ls * D + 8 * t);
      hp[0] = make_float4(-pt[0], -pt[1], -pt[2], -pt[3]);
      hp[1] = make_float4(-pt[4], -pt[5], -pt[6], -pt[7]);
      hdd_w[ls * nwarps + warp] = hdd;
      bd_w[ls * nwarps + warp] = bd;
    }
    {
      if (epi) {  // unconditional: an unselected landmark has p_t = 0 (its weight is zero), and the slots are not pre-zeroed
        if (warp == 0) fl_s[ls] = inb ? flags : -1;  // the finalise phase reads the flags from here, not from HBM
        // this target's share of the reference block: B_t^T p_t, B = blockdiag(Adj, 1, s')  (J_ref = U B)
        const float* adj = FEJ ? pc.adj0 : pc.adj;
        float rc[8];
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < 6; ++k) a += adj[k * 6 + j] * pt[k];
          rc[j] = a;
        }
        rc[6] = pt[6];
        rc[7] = (FEJ ? pc.s0 : pc.s) * pt[7];
        float4* rp = reinterpret_cast<float4*>(ref_w + ((size_t)warp * lpb + ls) * 8);
        rp[0] = make_float4(rc[0], rc[1], rc[2], rc[3]);
        rp[1] = make_float4(rc[4], rc[5], rc[6], rc[7]);
      }
    }
    // warp-wide transpose-reduce of the group's 44 sums + (energy, n): 24+12+6+3+2 = 47 shuffles; every lane ends up with
    // (at most) two entries of the 48-vector and adds them to its running totals
    if (ronce) {
      float* my = racc + (size_t)warp * (46 * 32) + lane;
      if (g0 == 0) {
#pragma unroll
        for (int k = 0; k < 44; ++k) my[k * 32] = acc[k];
        my[44 * 32] = e_add;
        my[45 * 32] = n_add;
      } else {
#pragma unroll
        for (int k = 0; k < 44; ++k) my[k * 32] += acc[k];
        my[44 * 32] += e_add;
        my[45 * 32] += n_add;
      }
    } else {
      float v48[PBA_CORE], v24[24], v12[12], v6[6], v3[3], v2[2];
#pragma unroll
      for (int k = 0; k < PBA_CORE; ++k) v48[k] = k < 44 ? acc[k] : (k == 44 ? e_add : (k == 45 ? n_add : 0.f));
      tr_step<48, 16>(v48, v24, lane);
      tr_step<24, 8>(v24, v12, lane);
      tr_step<12, 4>(v12, v6, lane);
      tr_step<6, 2>(v6, v3, lane);
      tr_step<3, 1>(v3, v2, lane);
      tot0 += v2[0];
      tot1 += v2[1];
    }
    if (g0 == 0) stamp(34);
  }
  stamp(35);

  {
    // this warp's 48 partial sums go to its own slot [host frame][chunk][target warp][48]: plain coalesced stores, summed
    // over the chunks in fp64 by k_core_reduce (no atomics, deterministic)
    float* dst = core_part + (((size_t)f * gridDim.x + blockIdx.x) * nwarps + warp) * PBA_CORE;
    if (ronce) {
      // lane L sums rows L and L + 32 of its warp's [46][32] block over the 32 columns, starting at column L (every lane on its
      // own bank); rows 46, 47 of the record stay zero
      __syncwarp();
      const float* base = racc + (size_t)warp * (46 * 32);
      const bool two = lane + 32 < 46;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const int c = (j + lane) & 31;
        s0 += base[lane * 32 + c];
        if (two) s1 += base[(lane + 32) * 32 + c];
      }
      dst[lane] = s0;
      if (lane + 32 < PBA_CORE) dst[lane + 32] = two ? s1 : 0.f;
    } else {
      dst[off] = tot0;
      if (!(lane & 1)) dst[off + 1] = tot1;
    }
  }
  __syncthreads();
  stamp(36);
  cta_stamp(1);

  if (epi) {
    // ---- phase A: finalise the chunk's landmarks (hessian_block_evaluation.hpp:213-227) and sum the reference block ------
    // hdd_s[ls] = 1 / H_dd (0: ill conditioned), bd_s[ls] = b_d / H_dd; a NEGATIVE hdd_s marks a landmark that is not
    // selected at all (its H_pd row is not stored)
    for (int ls = threadIdx.x; ls < lpb; ls += blockDim.x) {
      const int l = l0 + ls;
      float sc = -1.f, sb = 0.f;
      const int fl = fl_s[ls];
      if (fl >= 0) {
        const int gl = lm_base + l;
        const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
        const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
        if (sel) {
          float hdd = 0.f, bd = 0.f;
          for (int wi = 0; wi < nwarps; ++wi) {
            hdd += hdd_w[ls * nwarps + wi];
            bd += bd_w[ls * nwarps + wi];
          }
          w.b_d[gl] = bd;
          sc = 0.f;
          if (hdd > 1e-15f) {
            if (for_marg && w.fixed[f]) hdd += 1e8f;  // kScaleNullspaceRegularizer
            sc = 1.f / hdd;
            sb = sc * bd;
            w.inv_hdd[gl] = sc;
            w.flags[gl] = (uint8_t)(fl & ~LM_ILL);
          } else {
            w.flags[gl] = (uint8_t)(fl | LM_ILL);
          }
        }
      }
      hdd_s[ls] = sc;
      bd_s[ls] = sb;
    }
    for (int i = threadIdx.x; i < lpb * 8; i += blockDim.x) {
      float refv = 0.f;
      for (int wi = 0; wi < nwarps; ++wi) refv += ref_w[(size_t)wi * lpb * 8 + i];  // consecutive lanes, consecutive words
      hpd_s[(i >> 3) * D + 8 * f + (i & 7)] = refv;
    }
    __syncthreads();
    stamp(38);
    // ---- phase B: H_pd rows to HBM, the chunk's rank-k update, its b vector -- no barrier between them ---------------------
    const int D4 = D / 4;
    for (int i = threadIdx.x; i < lpb * D4; i += blockDim.x) {
      const int ls = i / D4;
      if (l0 + ls >= M) break;
      if (hdd_s[ls] >= 0.f)
        reinterpret_cast<float4*>(w.hpd + (size_t)(lm_base + l0 + ls) * w.hpd_stride)[i - ls * D4] =
            reinterpret_cast<const float4*>(hpd_s + ls * D)[i - ls * D4];
    }
    stamp(39);
    const int T4 = D / 4, ntri = T4 * (T4 + 1) / 2, nout = ntri * 16 + D;
    float* part = schur_part + ((size_t)f * gridDim.x + blockIdx.x) * nout;
    // the tiles are dealt from the LAST thread downwards and the b entries from the first thread upwards, so that with
    // ntri + D <= blockDim.x (N <= 8: 136 + 64 <= 224) nobody does both
    for (int tile = (int)blockDim.x - 1 - (int)threadIdx.x; tile < ntri; tile += blockDim.x) {
      int ty = 0, rem = tile;
      while (rem >= T4 - ty) {
        rem -= T4 - ty;
        ++ty;
      }
      const int tx = ty + rem;
      float a[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = 0.f;
      const float* qp = hpd_s + 4 * ty;
      const float* pp = hpd_s + 4 * tx;
#pragma unroll 4
      for (int ls = 0; ls < lpb; ++ls) {
        const float sc = fmaxf(hdd_s[ls], 0.f);
        const float4 qv = ldf4(qp + ls * D);
        const float4 pv = ldf4(pp + ls * D);
        const float qa[4] = {sc * qv.x, sc * qv.y, sc * qv.z, sc * qv.w};
        const float pa2[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i * 4 + j] += qa[i] * pa2[j];
      }
      float4* dst = reinterpret_cast<float4*>(part + tile * 16);
      dst[0] = make_float4(a[0], a[1], a[2], a[3]);
      dst[1] = make_float4(a[4], a[5], a[6], a[7]);
      dst[2] = make_float4(a[8], a[9], a[10], a[11]);
      dst[3] = make_float4(a[12], a[13], a[14], a[15]);
    }
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;  // four independent chains (lpb is a multiple of 32)
      for (int ls = 0; ls < lpb; ls += 4) {
        b0 += bd_s[ls] * hpd_s[ls * D + c];
        b1 += bd_s[ls + 1] * hpd_s[(ls + 1) * D + c];
        b2 += bd_s[ls + 2] * hpd_s[(ls + 2) * D + c];
        b3 += bd_s[ls + 3] * hpd_s[(ls + 3) * D + c];
      }
      part[ntri * 16 + c] = (b0 + b1) + (b2 + b3);
    }
    stamp(41);
    cta_stamp(2);
    return;
  }

  // reference block of H_pd:  sum_t B_t^T p_t  with p_t = -(target block t)   (J_ref = U B)
  for (int i = threadIdx.x; i < lpb * 8; i += blockDim.x) {
    const int ls = i >> 3, j = i & 7;
    float refv = 0.f;
    for (int wi = 0; wi < nwarps; ++wi) {
      const int tt2 = wi + (wi >= f);
      const float* ptv = hpd_s + ls * D + 8 * tt2;
      const PairConst& pw = pcs[wi];
      if (j < 6) {
        const float* adj = FEJ ? pw.adj0 : pw.adj;
#pragma unroll
        for (int k = 0; k < 6; ++k) refv -= adj[k * 6 + j] * ptv[k];
      } else if (j == 6) {
        refv -= ptv[6];
      } else {
        refv -= (FEJ ? pw.s0 : pw.s) * ptv[7];
      }
    }
    hpd_s[ls * D + 8 * f + j] = refv;
  }
  __syncthreads();
  stamp(37);

  // finalise the chunk's landmarks (hessian_block_evaluation.hpp:213-227)
  for (int ls = threadIdx.x; ls < lpb; ls += blockDim.x) {
    const int l = l0 + ls;
    float sc = 0.f, sb = 0.f;
    if (l < M) {
      const int gl = lm_base + l;
      const int fl = w.flags[gl];
      const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
      const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
      if (sel) {
        float hdd = 0.f, bd = 0.f;
        for (int wi = 0; wi < nwarps; ++wi) {
          hdd += hdd_w[ls * nwarps + wi];
          bd += bd_w[ls * nwarps + wi];
        }
        w.b_d[gl] = bd;
        if (hdd > 1e-15f) {
          if (for_marg && w.fixed[f]) hdd += 1e8f;  // kScaleNullspaceRegularizer
          sc = 1.f / hdd;
          sb = sc * bd;
          w.inv_hdd[gl] = sc;
          w.flags[gl] = (uint8_t)(fl & ~LM_ILL);
        } else {
          w.flags[gl] = (uint8_t)(fl | LM_ILL);
        }
      }
    }
    hdd_s[ls] = sc;
    bd_s[ls] = sb;
  }
  __syncthreads();
  stamp(38);
  const int D4 = D / 4;
  for (int i = threadIdx.x; i < lpb * D4; i += blockDim.x) {
    const int ls = i / D4;
    const int l = l0 + ls;
    if (l >= M) break;
    const int gl = lm_base + l;
    const int fl = w.flags[gl];
    const bool skip = (fl & LM_MARG) && !(fl & LM_TO_MARG);
    const bool sel = !skip && (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0);
    if (sel)
      reinterpret_cast<float4*>(w.hpd + (size_t)gl * w.hpd_stride)[i - ls * D4] =
          reinterpret_cast<const float4*>(hpd_s + ls * D)[i - ls * D4];
  }
  stamp(39);
  // K4 second half for this chunk: S = sum_l s_l H_pd_l H_pd_l^T (upper triangle as 4x4 tiles), b = sum_l s_l b_d_l H_pd_l
  {
    const int T4 = D / 4, ntri = T4 * (T4 + 1) / 2, nout = ntri * 16 + D;
    float* part = schur_part + ((size_t)f * gridDim.x + blockIdx.x) * nout;
    for (int tile = threadIdx.x; tile < ntri; tile += blockDim.x) {
      int ty = 0, rem = tile;
      while (rem >= T4 - ty) {
        rem -= T4 - ty;
        ++ty;
      }
      const int tx = ty + rem;
      float a[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = 0.f;
      for (int ls = 0; ls < lpb; ++ls) {
        const float sc = hdd_s[ls];
        const float4 qv = ldf4(hpd_s + ls * D + 4 * ty);
        const float4 pv = ldf4(hpd_s + ls * D + 4 * tx);
        const float qa[4] = {sc * qv.x, sc * qv.y, sc * qv.z, sc * qv.w};
        const float pa2[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i * 4 + j] += qa[i] * pa2[j];
      }
      float4* dst = reinterpret_cast<float4*>(part + tile * 16);
      dst[0] = make_float4(a[0], a[1], a[2], a[3]);
      dst[1] = make_float4(a[4], a[5], a[6], a[7]);
      dst[2] = make_float4(a[8], a[9], a[10], a[11]);
      dst[3] = make_float4(a[12], a[13], a[14], a[15]);
    }
    stamp(40);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float b = 0.f;
      for (int ls = 0; ls < lpb; ++ls) b += bd_s[ls] * hpd_s[ls * D + c];
      part[ntri * 16 + c] = b;
    }
  }
  stamp(41);
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
// K4 second half: H_s += inv H_pd H_pd^T, b_s += inv b_d H_pd over the selected, well-conditioned landmarks:
// a (8N x L)(L x 8N) SYRK.  Persistent grid (<= one CTA per SM); every CTA streams tiles of 32 landmarks through a
// cp.async double buffer, each thread owns one 4x4 tile of the UPPER triangle of H_s (fp32 products per tile, fp64
// running sums) and stores its CTA's partial sums; k_finish_system adds the CTAs up and mirrors the lower triangle.
// ------------------------------------------------------------------------------------------------
constexpr int SCHUR_TL = 32;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(1024) k_schur(const __grid_constant__ WindowDev w, int for_marg, int G, int gthreads,
                                                double* __restrict__ Hs, double* __restrict__ bs,
                                                const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  extern __shared__ __align__(16) float sm[];
  const int N = w.n_frames, D = 8 * N;
  const int T4 = D / 4;               // 4x4 tiles per side
  const int ntri = T4 * (T4 + 1) / 2; // tiles of the upper triangle
  float* Ps[2] = {sm, sm + SCHUR_TL * D};           // [TL][D] H_pd, two stages
  float* Sc = sm + 2 * SCHUR_TL * D;                // [2][TL] inv_hdd (0 when not selected)
  float* Bs = Sc + 2 * SCHUR_TL;                    // [2][TL] inv_hdd * b_d
  double* red = reinterpret_cast<double*>(Bs + 2 * SCHUR_TL);  // [ntri*16 + D] cross-group reduction
  const int tid = threadIdx.x, nt = blockDim.x;
  // G groups of gthreads threads split the landmarks of a tile (K-split); inside a group thread t owns one 4x4
  // tile (ty <= tx) of the upper triangle and, for t < D, entry t of b_s
  const int g = tid / gthreads, t = tid - g * gthreads;
  int ty = 0, tx = 0;
  const bool active = t < ntri;
  if (active) {
    int rem = t;
    while (rem >= T4 - ty) {
      rem -= T4 - ty;
      ++ty;
    }
    tx = ty + rem;
  }
  double dacc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) dacc[k] = 0;
  double bacc = 0;

  int tiles_of[PBA_MAXF + 1];
  tiles_of[0] = 0;
  for (int f = 0; f < N; ++f) tiles_of[f + 1] = tiles_of[f] + (w.n_lm[f] + SCHUR_TL - 1) / SCHUR_TL;
  const int total = tiles_of[N];
  const int D4 = D / 4;

  auto issue = [&](int tile, int stage) {
    int f = 0;
    while (tile >= tiles_of[f + 1]) ++f;
    const int l0 = (tile - tiles_of[f]) * SCHUR_TL;
    const int M = w.n_lm[f];
    const int base = lm_index(w, f, 0);
    for (int i = tid; i < SCHUR_TL * D4; i += nt) {
      const int ls = i / D4, c = i - ls * D4;
      const int l = min(l0 + ls, M - 1);  // rows past the end re-read the last row, their scale is 0
      cp_async16(Ps[stage] + ls * D + 4 * c, w.hpd + (size_t)(base + l) * w.hpd_stride + 4 * c);
    }
    if (tid < SCHUR_TL) {
      const int l = l0 + tid;
      float sc = 0.f, bv = 0.f;
      if (l < M) {
        const int gl = base + l;
        const int fl = w.flags[gl];
        const bool sel = (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0) && !(fl & LM_ILL);
        if (sel) {
          sc = w.inv_hdd[gl];
          bv = sc * w.b_d[gl];
        }
      }
      Sc[stage * SCHUR_TL + tid] = sc;
      Bs[stage * SCHUR_TL + tid] = bv;
    }
    cp_async_commit();
  };

  int stage = 0;
  if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < total) {
      issue(next, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* P = Ps[stage];
    const float* S = Sc + stage * SCHUR_TL;
    if (active) {
      float a[16];
#pragma unroll
      for (int k = 0; k < 16; ++k) a[k] = 0.f;
#pragma unroll 4
      for (int ls = g; ls < SCHUR_TL; ls += G) {
        const float sc = S[ls];
        const float4 qv = ldf4(P + ls * D + 4 * ty);
        const float4 pv = ldf4(P + ls * D + 4 * tx);
        const float qa[4] = {sc * qv.x, sc * qv.y, sc * qv.z, sc * qv.w};
        const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) a[i * 4 + j] += qa[i] * pa[j];
      }
#pragma unroll
      for (int k = 0; k < 16; ++k) dacc[k] += (double)a[k];
    }
    if (t < D) {
      const float* B = Bs + stage * SCHUR_TL;
      float b = 0.f;
      for (int ls = g; ls < SCHUR_TL; ls += G) b += B[ls] * P[ls * D + t];
      bacc += (double)b;
    }
    __syncthreads();
    stage ^= 1;
  }
  // cross-group reduction in shared memory, then one fp64 atomic per output element per CTA
  for (int gg = 0; gg < G; ++gg) {
    if (g == gg) {
      if (active) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (gg == 0) red[t * 16 + k] = dacc[k];
          else red[t * 16 + k] += dacc[k];
        }
      }
      if (t < D) {
        if (gg == 0) red[ntri * 16 + t] = bacc;
        else red[ntri * 16 + t] += bacc;
      }
    }
    __syncthreads();
  }
  if (g == 0) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 4 * ty + i, c = 4 * tx + j;
          if (r <= c) Hs[(size_t)blockIdx.x * D * D + (size_t)r * D + c] = red[t * 16 + i * 4 + j];
        }
    }
    if (t < D) bs[(size_t)blockIdx.x * D + t] = red[ntri * 16 + t];
  }
}

// ------------------------------------------------------------------------------------------------
// K4 second half on the tensor cores.  H_s = P^T diag(s) P with P = [H_pd rows] is a true dense contraction
// (8N x L) x (L x 8N), L = landmarks, so it runs as warp-level mma.sync m16n8k8 TF32 with the 3xTF32 split
// (x = hi + lo, x y ~= hi hi' + hi lo' + lo hi': fp32-grade products, fp32 accumulators in the MMA, fp64 atomics
// once per CTA).  The problem is far too small for a tcgen05/TMEM pipeline to pay (64x64 outputs, K = ~100 per
// CTA); what matters is that the inner product leaves the FFMA/LDS-bound CUDA-core path.
//   grid <= one CTA per SM, 8 warps; tiles of 32 landmarks stream through a cp.async double buffer [32][DP+8]
//   (row padding of 8 floats makes every fragment load bank-conflict free); warp tile = 16 x 32 outputs.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned f2tf32(float x) {
  unsigned r;
  asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

constexpr int SCHUR_MAXWT = 4;  // warp tiles per warp (8N = 128 -> 32 warp tiles over 8 warps)

__global__ void __launch_bounds__(256) k_schur_mma(const __grid_constant__ WindowDev w, int for_marg,
                                                   double* __restrict__ Hs, double* __restrict__ bs,
                                                   const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  extern __shared__ __align__(16) float sm[];
  const int N = w.n_frames, D = 8 * N;
  const int DP = (D + 31) & ~31, LDP = DP + 8;
  float* Ps[2] = {sm, sm + SCHUR_TL * LDP};
  float* Sc = sm + 2 * SCHUR_TL * LDP;  // [2][TL] inv_hdd (0 when not selected)
  float* Bs = Sc + 2 * SCHUR_TL;        // [2][TL] inv_hdd * b_d
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, lane = tid & 31;
  const int gid = lane >> 2, tig = lane & 3;
  const int nct = DP / 32, nwt = (DP / 16) * nct;
  for (int i = tid; i < 2 * SCHUR_TL * LDP; i += nt) sm[i] = 0.f;  // pad columns stay zero

  float acc[SCHUR_MAXWT][4][4];
#pragma unroll
  for (int a = 0; a < SCHUR_MAXWT; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;
  float bacc = 0.f;

  int tiles_of[PBA_MAXF + 1];
  tiles_of[0] = 0;
  for (int f = 0; f < N; ++f) tiles_of[f + 1] = tiles_of[f] + (w.n_lm[f] + SCHUR_TL - 1) / SCHUR_TL;
  const int total = tiles_of[N];
  const int D4 = D / 4;
  __syncthreads();

  auto issue = [&](int tile, int stage) {
    int f = 0;
    while (tile >= tiles_of[f + 1]) ++f;
    const int l0 = (tile - tiles_of[f]) * SCHUR_TL;
    const int M = w.n_lm[f];
    const int base = lm_index(w, f, 0);
    for (int i = tid; i < SCHUR_TL * D4; i += nt) {
      const int ls = i / D4, c = i - ls * D4;
      const int l = min(l0 + ls, M - 1);  // rows past the end re-read the last row, their scale is 0
      cp_async16(Ps[stage] + ls * LDP + 4 * c, w.hpd + (size_t)(base + l) * w.hpd_stride + 4 * c);
    }
    if (tid < SCHUR_TL) {
      const int l = l0 + tid;
      float sc = 0.f, bv = 0.f;
      if (l < M) {
        const int gl = base + l;
        const int fl = w.flags[gl];
        const bool sel = (for_marg ? (fl & LM_TO_MARG) != 0 : (fl & LM_MARG) == 0) && !(fl & LM_ILL);
        if (sel) {
          sc = w.inv_hdd[gl];
          bv = sc * w.b_d[gl];
        }
      }
      Sc[stage * SCHUR_TL + tid] = sc;
      Bs[stage * SCHUR_TL + tid] = bv;
    }
    cp_async_commit();
  };

  int stage = 0;
  if ((int)blockIdx.x < total) issue(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int next = tile + gridDim.x;
    if (next < total) {
      issue(next, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* P = Ps[stage];
    const float* S = Sc + stage * SCHUR_TL;
#pragma unroll
    for (int a = 0; a < SCHUR_MAXWT; ++a) {
      const int wt = warp + 8 * a;
      if (wt >= nwt) break;
      const int row0 = 16 * (wt / nct), col0 = 32 * (wt % nct);
      if (col0 + 31 < row0) continue;  // entirely below the diagonal
#pragma unroll
      for (int kk = 0; kk < SCHUR_TL; kk += 8) {
        const float s0 = S[kk + tig], s1 = S[kk + tig + 4];
        const float* p0 = P + (kk + tig) * LDP;
        const float* p1 = P + (kk + tig + 4) * LDP;
        float af[4] = {s0 * p0[row0 + gid], s0 * p0[row0 + gid + 8], s1 * p1[row0 + gid], s1 * p1[row0 + gid + 8]};
        unsigned ah[4], al[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ah[q] = f2tf32(af[q]);
          al[q] = f2tf32(af[q] - __uint_as_float(ah[q]));
        }
#pragma unroll
        for (int nt8 = 0; nt8 < 4; ++nt8) {
          const float bf[2] = {p0[col0 + 8 * nt8 + gid], p1[col0 + 8 * nt8 + gid]};
          unsigned bh[2], bl[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            bh[q] = f2tf32(bf[q]);
            bl[q] = f2tf32(bf[q] - __uint_as_float(bh[q]));
          }
          mma_tf32(acc[a][nt8], al, bh);
          mma_tf32(acc[a][nt8], ah, bl);
          mma_tf32(acc[a][nt8], ah, bh);
        }
      }
    }
    if (tid < D) {
      const float* B = Bs + stage * SCHUR_TL;
      float b = 0.f;
#pragma unroll 8
      for (int ls = 0; ls < SCHUR_TL; ++ls) b += B[ls] * P[ls * LDP + tid];
      bacc += b;
    }
    __syncthreads();
    stage ^= 1;
  }
#pragma unroll
  for (int a = 0; a < SCHUR_MAXWT; ++a) {
    const int wt = warp + 8 * a;
    if (wt >= nwt) break;
    const int row0 = 16 * (wt / nct), col0 = 32 * (wt % nct);
    if (col0 + 31 < row0) continue;
#pragma unroll
    for (int nt8 = 0; nt8 < 4; ++nt8)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int r = row0 + gid + ((q & 2) ? 8 : 0);
        const int c = col0 + 8 * nt8 + 2 * tig + (q & 1);
        if (r <= c && c < D) Hs[(size_t)blockIdx.x * D * D + (size_t)r * D + c] = (double)acc[a][nt8][q];
      }
  }
  if (tid < D) bs[(size_t)blockIdx.x * D + tid] = (double)bacc;
}

// ------------------------------------------------------------------------------------------------
// assembly of H_pp / b_p from the per-pair cores, in fp64 -- two kernels, no atomics, no zero-fill:
//   k_core_reduce   : second stage of the per-pair core reduction (sum over the host frame's chunk CTAs)
//   k_assemble      : GATHER form of the reference's scatter + symmetrisation
//                     (hessian_block_evaluation.hpp:118-163, quirk Q3): one CTA per 8x8 block (bi, bj), bj >= bi
//     H[r,r] = sum_t ( B_rt^T C_rt B_rt + C_tr ),  lower triangle mirrored (selfadjointView<Lower>)
//     H[i,j] = -B_ij^T C_ij - (B_ji^T C_ji)^T  for i < j,  H[j,i] = H[i,j]^T
//     b[r]   = sum_t ( B_rt^T q_rt - q_tr )
//   with C_rt / q_rt the core of the ordered pair (reference r -> target t) and B = blockdiag(Adj, 1, s').
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_core_reduce(const __grid_constant__ WindowDev w,
                                                     const float* __restrict__ core_part, int lpb, int chunk_stride,
                                                     double* __restrict__ core, const LmCtl* __restrict__ ctl) {
  KStamp kstamp_(1);
  if (lm_skip(ctl, 2)) return;
  __shared__ double red[4][PBA_CORE];
  const int N = w.n_frames;
  const int r = blockIdx.x / (N - 1);
  const int tw = blockIdx.x % (N - 1);  // the target's warp index in k_linearize_fused
  const int t = tw + (tw >= r);
  const int g = threadIdx.x >> 6, k = threadIdx.x & 63;
  double acc = 0;
  if (k < PBA_CORE) {
    const int nchunk = (w.n_lm[r] + lpb - 1) / lpb;
    const float* p = core_part + ((size_t)r * chunk_stride * (N - 1) + tw) * PBA_CORE + k;
    for (int ch = g; ch < nchunk; ch += 4) acc += (double)p[(size_t)ch * (N - 1) * PBA_CORE];
    red[g][k] = acc;
  }
  __syncthreads();
  if (g == 0 && k < PBA_CORE) core[(size_t)(r * PBA_MAXF + t) * PBA_CORE + k] = red[0][k] + red[1][k] + red[2][k] + red[3][k];
}

__device__ __forceinline__ double core_at(const double* c, int i, int j) {  // symmetric 8x8 from its upper triangle
  const int a = min(i, j), b = max(i, j);
  return c[a * 8 - (a * (a - 1)) / 2 + (b - a)];
}
__device__ __forceinline__ double bm_at(const PairAssemble& pa, int fej, int i, int j) {  // B = blockdiag(Adj, 1, s')
  if (i < 6 && j < 6) return (fej ? pa.adj_fej : pa.adj_cur)[i * 6 + j];
  if (i == 6 && j == 6) return 1.0;
  if (i == 7 && j == 7) return fej ? pa.s0 : pa.s;
  return 0.0;
}

// One warp per contributing ordered pair: all its loads (core from L2, B from the pair table) are issued at once and
// the two 8x8x8 fp64 products run warp-locally (lane = row i, two columns); the warps' terms meet in shared memory.
// A serial loop over the targets costs one L2 round trip and four barriers per target instead.
constexpr int ASM_W = 264;  // doubles of shared memory per warp: C, B, T1, out (64 each) + b (8)
__global__ void __launch_bounds__(512) k_assemble(const __grid_constant__ WindowDev w, int fej,
                                                  const double* __restrict__ core, double* __restrict__ Hp,
                                                  double* __restrict__ bp, const LmCtl* __restrict__ ctl, int peer_push) {
  KStamp kstamp_(4);
  if (lm_skip(ctl, 2)) return;
  const unsigned epoch = peer_push ? peer_epoch() : 0u;
  extern __shared__ double asm_sm[];
  const int N = w.n_frames, D = 8 * N;
  // blockIdx.x enumerates the upper-triangular block pairs (bi <= bj)
  int bi = 0, rem = blockIdx.x;
  while (rem >= N - bi) {
    rem -= N - bi;
    ++bi;
  }
  const int bj = bi + rem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = lane >> 2, j0 = (lane & 3) * 2;
  double* Cw = asm_sm + warp * ASM_W;
  double* Bw = Cw + 64;
  double* Tw = Bw + 64;
  double* Ow = Tw + 64;
  double* bw = Ow + 64;
  const bool diag = bi == bj;
  const int nterms = diag ? N - 1 : 2;
  if (warp < nterms) {
    // diagonal block r: term t gives B_rt^T C_rt B_rt + C_tr.  Off-diagonal block (bi, bj): the ordered pair r -> t
    // contributes H_rt = -B^T C to block (r, t); side 1 is transposed when the terms are added.
    const int r = diag ? bi : (warp ? bj : bi);
    const int t = diag ? warp + (warp >= bi) : (warp ? bi : bj);
    const double* c_rt = core + (size_t)(r * PBA_MAXF + t) * PBA_CORE;
    const double* c_tr = core + (size_t)(t * PBA_MAXF + r) * PBA_CORE;
    const PairAssemble& pa = w.pairs_asm[r * PBA_MAXF + t];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      Cw[i * 8 + j0 + q] = core_at(c_rt, i, j0 + q);
      Bw[i * 8 + j0 + q] = bm_at(pa, fej, i, j0 + q);
    }
    const double ctr0 = diag ? core_at(c_tr, i, j0) : 0.0, ctr1 = diag ? core_at(c_tr, i, j0 + 1) : 0.0;
    const double q_rt = (diag && lane < 8) ? c_rt[36 + lane] : 0.0, q_tr = (diag && lane < 8) ? c_tr[36 + lane] : 0.0;
    if (diag && lane < 8) bw[lane] = q_rt;  // staged for the B^T q product below
    __syncwarp();
    double s0 = 0, s1 = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {  // (B^T C)[i][j]
      s0 += Bw[k * 8 + i] * Cw[k * 8 + j0];
      s1 += Bw[k * 8 + i] * Cw[k * 8 + j0 + 1];
    }
    if (!diag) {
      Ow[i * 8 + j0] = -s0;
      Ow[i * 8 + j0 + 1] = -s1;
    } else {
      Tw[i * 8 + j0] = s0;
      Tw[i * 8 + j0 + 1] = s1;
      double br = 0;
      if (lane < 8) {
#pragma unroll
        for (int k = 0; k < 8; ++k) br += Bw[k * 8 + lane] * bw[k];  // B^T q_rt
      }
      __syncwarp();
      double h0 = ctr0, h1 = ctr1;  // + H_tt of the pair t -> r
#pragma unroll
      for (int k = 0; k < 8; ++k) {  // B^T C B
        h0 += Tw[i * 8 + k] * Bw[k * 8 + j0];
        h1 += Tw[i * 8 + k] * Bw[k * 8 + j0 + 1];
      }
      Ow[i * 8 + j0] = h0;
      Ow[i * 8 + j0 + 1] = h1;
      if (lane < 8) bw[lane] = br - q_tr;  // b_t of the pair t -> r is -q_tr
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 64; o += blockDim.x) {
    const int oi = o >> 3, oj = o & 7;
    if (!diag) {
      const double acc = asm_sm[oi * 8 + oj + 192] + asm_sm[ASM_W + oj * 8 + oi + 192];
      red_store(&Hp[(size_t)(8 * bi + oi) * D + 8 * bj + oj], acc, peer_push, epoch);
      red_store(&Hp[(size_t)(8 * bj + oj) * D + 8 * bi + oi], acc, peer_push, epoch);
    } else {
      const int ii = max(oi, oj), jj = min(oi, oj);  // selfadjointView<Lower>
      double acc = 0;
      for (int wi = 0; wi < nterms; ++wi) acc += asm_sm[wi * ASM_W + 192 + ii * 8 + jj];
      red_store(&Hp[(size_t)(8 * bi + oi) * D + 8 * bi + oj], acc, peer_push, epoch);
      if (oj == 0) {
        double bacc = 0;
        for (int wi = 0; wi < nterms; ++wi) bacc += asm_sm[wi * ASM_W + 256 + oi];
        red_store(&bp[8 * bi + oi], bacc, peer_push, epoch);
      }
    }
  }
  if (peer_push) peer_arrive(1, epoch, threadIdx.x < 64);
}

// second stage of the fused path's Schur reduction: out[o] = sum over the chunk CTAs of part[cta][o], o over the
// 4x4 upper-triangle tiles and b.  16 groups of 64 outputs per CTA; each group strides over the chunk CTAs.
__global__ void __launch_bounds__(1024) k_finish_fused(const __grid_constant__ WindowDev w, int lpb, int chunks,
                                                       const float* __restrict__ part, double* __restrict__ Hs,
                                                       double* __restrict__ bs, const LmCtl* __restrict__ ctl, int peer_push) {
  KStamp kstamp_(3);
  if (lm_skip(ctl, 2)) return;
  const unsigned epoch = peer_push ? peer_epoch() : 0u;
  __shared__ double red[32][33];
  const int N = w.n_frames, D = 8 * N;
  const int T4 = D / 4, ntri = T4 * (T4 + 1) / 2, nout = ntri * 16 + D;
  const int g = threadIdx.x >> 5, oo = threadIdx.x & 31;  // 32 groups stride over the chunk CTAs, 32 outputs per CTA
  const int o = blockIdx.x * 32 + oo;
  // rows of `part` = (host frame, chunk CTA) flattened: q = f * chunks + c; every thread keeps four independent loads
  // in flight (CTAs past a frame's last landmark never ran: their rows are skipped)
  __shared__ int s_nch[PBA_MAXF];
  if (threadIdx.x < PBA_MAXF) s_nch[threadIdx.x] = threadIdx.x < N ? (w.n_lm[threadIdx.x] + lpb - 1) / lpb : 0;
  __syncthreads();
  double acc = 0;
  if (o < nout) {
    const int rows = N * chunks;
    const float* p = part + o;
    auto P = [&](int q) -> float {
      const int f = q / chunks;
      return (q - f * chunks) < s_nch[f] ? p[(size_t)q * nout] : 0.f;
    };
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int q = g;
    for (; q + 96 < rows; q += 128) {
      a0 += P(q);
      a1 += P(q + 32);
      a2 += P(q + 64);
      a3 += P(q + 96);
      if (((q - g) & 511) == 384) {  // bound the fp32 partial sums: flush to fp64 every 16 rows
        acc += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
        a0 = a1 = a2 = a3 = 0.f;
      }
    }
    for (; q < rows; q += 32) a0 += P(q);
    acc += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
  }
  red[g][oo] = acc;
  __syncthreads();
  if (g == 0 && o < nout) {
    double s = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) s += red[k][oo];
    if (o >= ntri * 16) {
      red_store(&bs[o - ntri * 16], s, peer_push, epoch);
    } else {
      const int tile = o >> 4, e = o & 15;
      int ty = 0, rem = tile;
      while (rem >= T4 - ty) {
        rem -= T4 - ty;
        ++ty;
      }
      const int r = 4 * ty + (e >> 2), cc = 4 * (ty + rem) + (e & 3);
      if (r <= cc) {  // diagonal tiles also carry r > cc: dropped, the mirror keeps H_s exactly symmetric
        red_store(&Hs[(size_t)r * D + cc], s, peer_push, epoch);
        if (r != cc) red_store(&Hs[(size_t)cc * D + r], s, peer_push, epoch);
      }
    }
  }
  if (peer_push) peer_arrive(1, epoch, g == 0);
}

// ------------------------------------------------------------------------------------------------
// k_reduce_system (round 2): the whole second stage of a linearisation in ONE launch -- what k_core_reduce, k_assemble and
// k_finish_fused do in three.  Blocks [0, N (N + 1) / 2) assemble one upper-triangular 8x8 block pair of H_pp each and
// reduce the per-chunk core records they need on the fly (a pair's 48 sums over <= a few dozen chunk CTAs: cheaper than a
// launch boundary); the diagonal blocks also publish the reduced cores (slots 44 / 45 carry the pair energies the LM
// decision reads).  The remaining blocks sum the per-chunk Schur partials.  with_system = 0 (the last evaluation of a
// solve): only the cores are reduced.
// ------------------------------------------------------------------------------------------------
constexpr int RSYS_W = ASM_W + 96;  // per assemble warp: the k_assemble staging + the two reduced cores (48 doubles each)
__global__ void __launch_bounds__(1024) k_reduce_system(const __grid_constant__ WindowDev w, int fej,
                                                        const float* __restrict__ core_part, int lpb, int chunks,
                                                        double* __restrict__ core, double* __restrict__ Hp,
                                                        double* __restrict__ bp, const float* __restrict__ fpart,
                                                        double* __restrict__ Hs, double* __restrict__ bs,
                                                        const LmCtl* __restrict__ ctl, int with_system) {
  cudaGridDependencySynchronize();
  if (lm_skip(ctl, 1)) return;
  extern __shared__ double rs_sm[];
  const int N = w.n_frames, D = 8 * N;
  const int NB = N * (N + 1) / 2;
  const bool st0 = g_stamps_on && threadIdx.x == 0 && blockIdx.x == 0;
  const bool st1 = g_stamps_on && threadIdx.x == 0 && (int)blockIdx.x == NB;
  if (st0) g_stamps[54] = clock64();
  if (st1) g_stamps[59] = clock64();
  if ((int)blockIdx.x < NB) {
    // ---- assemble block (bi <= bj), cf. k_assemble ---------------------------------------------------------------------
    int bi = 0, rem = blockIdx.x;
    while (rem >= N - bi) {
      rem -= N - bi;
      ++bi;
    }
    const int bj = bi + rem;
    const bool diag = bi == bj;
    if (!with_system && !diag) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = lane >> 2, j0 = (lane & 3) * 2;
    const int nterms = diag ? N - 1 : 2;
    double* Cw = rs_sm + (size_t)min(warp, max(2, N - 1) - 1) * RSYS_W;
    double* Bw = Cw + 64;
    double* Tw = Bw + 64;
    double* Ow = Tw + 64;
    double* bw = Ow + 64;
    double* c_rt = Cw + ASM_W;   // reduced core of the pair r -> t
    double* c_tr = c_rt + 48;    // ... of the pair t -> r
    if (warp < nterms) {
      const int r = diag ? bi : (warp ? bj : bi);
      const int t = diag ? warp + (warp >= bi) : (warp ? bi : bj);
      // sum the chunk CTAs' records: [host frame][chunk][target warp][48] floats, fixed order, fp64
      for (int side = 0; side < 2; ++side) {
        const int hf = side ? t : r, tf = side ? r : t;     // host frame and target frame of this record
        const int tw = tf - (tf > hf);                       // the target's warp index in the sweep CTA of hf
        const int nchunk = (w.n_lm[hf] + lpb - 1) / lpb;
        const float* p = core_part + ((size_t)hf * chunks * (N - 1) + tw) * PBA_CORE;
        double a0 = 0, a1 = 0;
        const size_t cs = (size_t)(N - 1) * PBA_CORE;
        const bool hi = lane < PBA_CORE - 32;
        int ch = 0;
        for (; ch + 8 <= nchunk; ch += 8) {  // 16 independent loads in flight per lane, then the fixed-order fp64 sums
          float x[8], y[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float* q = p + (size_t)(ch + u) * cs;
            x[u] = q[lane];
            y[u] = hi ? q[32 + lane] : 0.f;
          }
          // fp32 -> fp64 conversions run on B200's thin fp64 pipe (they were the top stall of this kernel): eight chunk
          // records are summed pairwise in fp32 first (relative rounding 2e-7 on records that are themselves fp32 sums
          // of ~500 products), one conversion per eight
          a0 += (double)(((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7])));
          a1 += (double)(((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7])));
        }
        for (; ch < nchunk; ++ch) {
          const float* q = p + (size_t)ch * cs;
          a0 += (double)q[lane];
          if (hi) a1 += (double)q[32 + lane];
        }
        double* dst = side ? c_tr : c_rt;
        dst[lane] = a0;
        if (lane < PBA_CORE - 32) dst[32 + lane] = a1;
        if (diag && side == 0) {  // publish: every ordered pair (r, t) belongs to exactly one diagonal block's warp
          double* g = core + (size_t)(r * PBA_MAXF + t) * PBA_CORE;
          g[lane] = a0;
          if (lane < PBA_CORE - 32) g[32 + lane] = a1;
        }
      }
      __syncwarp();
      if (st0) g_stamps[55] = clock64();
      if (with_system) {
        const PairAssemble& pa = w.pairs_asm[r * PBA_MAXF + t];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          Cw[i * 8 + j0 + q] = core_at(c_rt, i, j0 + q);
          Bw[i * 8 + j0 + q] = bm_at(pa, fej, i, j0 + q);
        }
        const double ctr0 = diag ? core_at(c_tr, i, j0) : 0.0, ctr1 = diag ? core_at(c_tr, i, j0 + 1) : 0.0;
        const double q_rt = (diag && lane < 8) ? c_rt[36 + lane] : 0.0, q_tr = (diag && lane < 8) ? c_tr[36 + lane] : 0.0;
        if (diag && lane < 8) bw[lane] = q_rt;
        __syncwarp();
        double s0 = 0, s1 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // (B^T C)[i][j]
          s0 += Bw[k * 8 + i] * Cw[k * 8 + j0];
          s1 += Bw[k * 8 + i] * Cw[k * 8 + j0 + 1];
        }
        if (!diag) {
          Ow[i * 8 + j0] = -s0;
          Ow[i * 8 + j0 + 1] = -s1;
        } else {
          Tw[i * 8 + j0] = s0;
          Tw[i * 8 + j0 + 1] = s1;
          double br = 0;
          if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) br += Bw[k * 8 + lane] * bw[k];  // B^T q_rt
          }
          __syncwarp();
          double h0 = ctr0, h1 = ctr1;  // + H_tt of the pair t -> r
#pragma unroll
          for (int k = 0; k < 8; ++k) {  // B^T C B
            h0 += Tw[i * 8 + k] * Bw[k * 8 + j0];
            h1 += Tw[i * 8 + k] * Bw[k * 8 + j0 + 1];
          }
          Ow[i * 8 + j0] = h0;
          Ow[i * 8 + j0 + 1] = h1;
          if (lane < 8) bw[lane] = br - q_tr;  // b_t of the pair t -> r is -q_tr
        }
      }
    }
    if (st0) g_stamps[56] = clock64();
    __syncthreads();
    if (st0) g_stamps[57] = clock64();
    if (!with_system) return;
    for (int o = threadIdx.x; o < 64; o += blockDim.x) {
      const int oi = o >> 3, oj = o & 7;
      if (!diag) {
        const double acc = rs_sm[oi * 8 + oj + 192] + rs_sm[RSYS_W + oj * 8 + oi + 192];
        Hp[(size_t)(8 * bi + oi) * D + 8 * bj + oj] = acc;
        Hp[(size_t)(8 * bj + oj) * D + 8 * bi + oi] = acc;
      } else {
        const int ii = max(oi, oj), jj = min(oi, oj);  // selfadjointView<Lower>
        double acc = 0;
        for (int wi = 0; wi < nterms; ++wi) acc += rs_sm[wi * RSYS_W + 192 + ii * 8 + jj];
        Hp[(size_t)(8 * bi + oi) * D + 8 * bi + oj] = acc;
        if (oj == 0) {
          double bacc = 0;
          for (int wi = 0; wi < nterms; ++wi) bacc += rs_sm[wi * RSYS_W + 256 + oi];
          bp[8 * bi + oi] = bacc;
        }
      }
    }
    if (st0) g_stamps[58] = clock64();
    return;
  }
  if (!with_system) return;
  // ---- Schur partial reduction, cf. k_finish_fused ------------------------------------------------------------------------
  __shared__ double red[32][33];
  __shared__ int s_nch[PBA_MAXF];
  const int T4 = D / 4, ntri = T4 * (T4 + 1) / 2, nout = ntri * 16 + D;
  const int g = threadIdx.x >> 5, oo = threadIdx.x & 31;
  const int o = ((int)blockIdx.x - NB) * 32 + oo;
  if (threadIdx.x < PBA_MAXF) s_nch[threadIdx.x] = threadIdx.x < N ? (w.n_lm[threadIdx.x] + lpb - 1) / lpb : 0;
  __syncthreads();
  double acc = 0;
  if (o < nout) {
    const int rows = N * chunks;
    const float* p = fpart + o;
    auto P = [&](int q) -> float {
      const int f = q / chunks;
      return (q - f * chunks) < s_nch[f] ? p[(size_t)q * nout] : 0.f;
    };
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int q = g;
    for (; q + 96 < rows; q += 128) {
      a0 += P(q);
      a1 += P(q + 32);
      a2 += P(q + 64);
      a3 += P(q + 96);
      if (((q - g) & 511) == 384) {  // bound the fp32 partial sums: flush to fp64 every 16 rows
        acc += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
        a0 = a1 = a2 = a3 = 0.f;
      }
    }
    for (; q < rows; q += 32) a0 += P(q);
    acc += ((double)a0 + (double)a1) + ((double)a2 + (double)a3);
  }
  if (st1) g_stamps[60] = clock64();
  red[g][oo] = acc;
  __syncthreads();
  if (st1) g_stamps[61] = clock64();
  if (g == 0 && o < nout) {
    double sum = 0;
#pragma unroll
    for (int k = 0; k < 32; ++k) sum += red[k][oo];
    if (o >= ntri * 16) {
      bs[o - ntri * 16] = sum;
    } else {
      const int tile = o >> 4, e = o & 15;
      int ty = 0, rem2 = tile;
      while (rem2 >= T4 - ty) {
        rem2 -= T4 - ty;
        ++ty;
      }
      const int r = 4 * ty + (e >> 2), cc = 4 * (ty + rem2) + (e & 3);
      if (r <= cc) {
        Hs[(size_t)r * D + cc] = sum;
        Hs[(size_t)cc * D + r] = sum;
      }
    }
  }
  if (st1) g_stamps[62] = clock64();
}

// second stage of the Schur reduction + symmetrisation of both systems (hessian_block_evaluation.hpp:147-163):
// H_s[i][j] = H_s[j][i] = sum over the SYRK CTAs of their upper-triangle partials; H_p mirrored as the reference does
__global__ void k_finish_system(int D, double* __restrict__ Hp, double* __restrict__ Hs, double* __restrict__ bs,
                                const double* __restrict__ schur_part, const double* __restrict__ bs_part, int nsb,
                                const LmCtl* __restrict__ ctl) {
  if (lm_skip(ctl, 2)) return;
  const int a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= D * D) return;
  const int i = a / D, j = a % D;
  if (Hs && i <= j) {
    double s = 0;
    for (int b = 0; b < nsb; ++b) s += schur_part[(size_t)b * D * D + a];
    Hs[(size_t)i * D + j] = s;
    Hs[(size_t)j * D + i] = s;
    if (i == 0) {
      double t = 0;
      for (int b = 0; b < nsb; ++b) t += bs_part[(size_t)b * D + j];
      bs[j] = t;
    }
  }
  if (!Hp) return;
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
                                                         const LmCtl* __restrict__ ctl, double* __restrict__ norms /* per-CTA (state, step) partials or null */) {
  KStamp kstamp_(6);
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
      const float id = w.lmk[gl].z;
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
      reinterpret_cast<double2*>(norms)[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = make_double2(a, b);
    }
  }
}

// acceptStep / rejectStep over landmarks (problem.hpp:377-384,395-399); norms in fp64
__global__ void __launch_bounds__(256) k_accept_landmarks(const __grid_constant__ WindowDev w, int accept,
                                                          double* __restrict__ scal, const LmCtl* __restrict__ ctl,
                                                          int with_statuses) {
  KStamp kstamp_(8);
  if (ctl) {  // device LM: k_lm_energy decided (and already closed the iteration's bookkeeping)
    if (!ctl->apply) return;
    accept = ctl->accept;
  }
  const int f = blockIdx.y;
  const int M = w.n_lm[f];
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  double st = 0, sp = 0;
  if (l < M && with_statuses) {  // changeResidualStatuses folded in (problem.hpp:20-35)
    for (int t = 0; t < w.n_frames; ++t) {
      if (t == f) continue;
      const size_t res = res_index(w, f, t, l);
      if (accept > 0) w.status[res] = w.cand[res];
      else w.cand[res] = w.status[res];
    }
  }
  if (l < M) {
    const int gl = lm_index(w, f, l);
    const float s = w.idepth_step[gl];
    if (accept > 0) {
      const float id = w.lmk[gl].z;
      st = (double)id * id;
      sp = (double)s * s;
      w.lmk[gl].z = id + s;
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
  const float id = w.lmk[gl].z;
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
// Immature-landmark activation refine: LandmarkActivationProblem + optimizeImmatureLandmark,
// tracker/landmarks_activator/src/landmarks_activator.cpp:122-316 -- per candidate a 1-D Levenberg-Marquardt on the
// inverse depth over all other frames of the window (8-pixel pattern, Huber weight, inlier energy cap 8 * 12^2,
// lambda0 = 0.1, <= 3 iterations, /2 on accept, x5 on reject, ptol 1e-8, ftol 0).  Thousands of independent tiny
// problems: one WARP per candidate, 8 lanes per target frame (4 targets per pass), the whole LM in registers.
// One evaluation yields the energy AND (H, b) at the same inverse depth: linearize() after an accepted step is at the
// trial point the energy was just evaluated at, so the reference's two passes are one here.
// ------------------------------------------------------------------------------------------------
struct ActEval {
  float energy, H, b;
  int n;
};
__device__ __forceinline__ ActEval act_eval(const WindowDev& w, const PairConst* __restrict__ pcs, int r, float u0, float v0,
                                            float patch, float rho, float sigma, int lane) {
  const int N = w.n_frames;
  const int px = lane & 7, grp = lane >> 3;
  const float xmax = (float)(w.W - 5), ymax = (float)(w.H - 5);
  const float ur = u0 + pat_x(px), vr = v0 + pat_y(px);
  ActEval out{0.f, 0.f, 0.f, 0};
  for (int base = 0; base < N - 1; base += 4) {  // uniform trip count: the shuffles below need the whole warp
    const int ti = base + grp;
    const bool have = ti < N - 1;
    const int t = have ? ti + (ti >= r) : (r == 0 ? 1 : 0);
    const PairConst& pc = pcs[r * PBA_MAXF + t];
    bool ok = have && valid_idepth(rho) && in_roi(ur, vr, xmax, ymax);
    const float X = row_apply4(ldf4(pc.A + 0), ur, vr, rho);
    const float Y = row_apply4(ldf4(pc.A + 4), ur, vr, rho);
    const float Z = row_apply4(ldf4(pc.A + 8), ur, vr, rho);
    ok = ok && Z > 0.f;
    const float rz = __frcp_rn(Z);
    const float tu = __fmul_rn(X, rz), tv = __fmul_rn(Y, rz);
    ok = ok && in_roi(tu, tv, xmax, ymax);
    ok = group_all(ok, lane);
    {  // target_mask.valid(target_pattern), bounds are implied by the ROI test; branch-free (groups differ in t)
      const bool need = !w.mask_all[t];
      const int midx = (ok && need) ? (int)roundf(tv) * w.W + (int)roundf(tu) : 0;
      const bool mv = need ? w.mask[t][midx] != 0 : true;
      ok = group_all(ok && mv, lane);
    }
    const float su = ok ? tu : 8.f, sv = ok ? tv : 8.f;
    const int ix = (int)su, iy = (int)sv;
    const float dx = su - (float)ix, dy = sv - (float)iy, dxdy = dx * dy;
    const float w11 = dxdy, w10 = dy - dxdy, w01 = dx - dxdy, w00 = 1.f - dx - dy + dxdy;
    const float4* p = w.img[t] + ((size_t)iy * w.W + ix) * 2;
    float4 t00, t01, t10, t11;
    ldg256_nc(p, t00, t01);
    ldg256_nc(p + 2 * (size_t)w.W, t10, t11);
    const float I = w11 * t11.x + w10 * t10.x + w01 * t01.x + w00 * t00.x;
    const float dIu = w11 * t11.y + w10 * t10.y + w01 * t01.y + w00 * t00.y;
    const float dIv = w11 * t11.z + w10 * t10.z + w01 * t01.z + w00 * t00.z;
    const float res = ok ? (I - pc.b_t) - pc.s * (patch - pc.b_r) : 0.f;
    const float n2 = group_sum(res * res);
    const float wgt = n2 > sigma * sigma ? sigma * rsqrt_approx(n2) : 1.f;  // :176
    // d r / d idepth (camera_reproject.hpp:339-344)
    const float qx = row_apply4(ldf4(pc.M + 0), ur, vr, rho);
    const float qy = row_apply4(ldf4(pc.M + 4), ur, vr, rho);
    const float qz = row_apply4(ldf4(pc.M + 8), ur, vr, rho);
    const float sI = rcp_approx(ok ? qz : 1.f);
    const float b0 = qx * sI, b1 = qy * sI;
    const float du = pc.fx_t * (pc.tr[0] * sI - pc.tr[2] * sI * b0), dv = pc.fy_t * (pc.tr[1] * sI - pc.tr[2] * sI * b1);
    const float d = ok ? dIu * du + dIv * dv : 0.f;
    const float hd = group_sum(d * d), bd = group_sum(d * res);
    if (px == 0 && ok) {  // one lane per target adds the target's terms
      if (n2 < 8.f * 144.f) {  // kMaxEnergyForInliers, :124,178-183
        out.energy += wgt * n2;
        out.n += 1;
      } else {
        out.energy += 8.f * 144.f;
      }
      out.H += wgt * hd;  // linearize() has no inlier cap, :243-244
      out.b += wgt * bd;
    }
  }
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {  // across the 4 groups (only lane px == 0 of each holds data)
    out.energy += __shfl_xor_sync(FULL, out.energy, o);
    out.H += __shfl_xor_sync(FULL, out.H, o);
    out.b += __shfl_xor_sync(FULL, out.b, o);
    out.n += __shfl_xor_sync(FULL, out.n, o);
  }
  // broadcast lane 0's totals
  out.energy = __shfl_sync(FULL, out.energy, 0);
  out.H = __shfl_sync(FULL, out.H, 0);
  out.b = __shfl_sync(FULL, out.b, 0);
  out.n = __shfl_sync(FULL, out.n, 0);
  return out;
}

__global__ void __launch_bounds__(256) k_refine_immature(const __grid_constant__ WindowDev w, int r, int n,
                                                         const float2* __restrict__ proj, const float* __restrict__ idepth_in,
                                                         const float* __restrict__ patch8, int min_inliers, float sigma,
                                                         float* __restrict__ idepth_out, uint8_t* __restrict__ activate,
                                                         int* __restrict__ n_valid_out) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n) return;  // whole warps leave together
  const float2 uv = proj[c];
  const float patch = patch8[(size_t)c * 8 + (lane & 7)];
  float rho = idepth_in[c];
  // levenberg_marquardt_algorithm::solve (lm.hpp:77-128) with the options of :294-300
  ActEval cur = act_eval(w, w.pairs, r, uv.x, uv.y, patch, rho, sigma, lane);
  float energy = cur.energy, H = cur.H, b = cur.b, lambda = 0.1f;
  int nvalid = cur.n;
  bool stop = nvalid == 0, converged = false;  // calculateEnergy(): no valid residual -> idepth = -1, stop (:192-195)
  for (int it = 0; it < 3 && !converged && nvalid > 0; ++it) {
    if (H == 0.f) stop = true;  // linearize(): hessian_ == 0 (:249)
    const float step = b / (H + H * lambda);  // calculateStep (:252-256)
    const float old = rho;
    rho -= step;
    if (stop) {  // the trial calculateEnergy() returns {0, 0}: rejectStep(); break (lm.hpp:95-98)
      rho = old;
      break;
    }
    const ActEval tr = act_eval(w, w.pairs, r, uv.x, uv.y, patch, rho, sigma, lane);
    if (tr.n == 0) {
      stop = true;
      rho = old;
      break;
    }
    if (tr.energy < energy) {  // acceptStep(): (idepth^2, step^2) (:258); function_tolerance = 0 never fires
      if (step * step < 1e-8f * (rho * rho + 1e-8f)) converged = true;
      energy = tr.energy;
      nvalid = tr.n;
      H = tr.H;
      b = tr.b;
      lambda *= 0.5f;
    } else {
      rho = old;  // rejectStep(): the previous (H, b) stay valid
      lambda *= 5.f;
    }
  }
  if (stop) rho = -1.f;  // the trailing calculateEnergy() with stop_ set (:150-153)
  if (lane == 0) {
    const bool act = !(nvalid < min_inliers || rho < 0.f);  // :308-314
    idepth_out[c] = rho;
    activate[c] = act ? 1 : 0;
    n_valid_out[c] = nvalid;
  }
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

__global__ void __launch_bounds__(128) k_pair_setup(const FrameParams* __restrict__ fr, int N,
                                                     PairConst* __restrict__ pairs, PairAssemble* __restrict__ pasm) {
  KStamp kstamp_(7);
  // one CTA per reference frame r, FOUR warps = four roles, lane = frame.  The kernel sits on the critical path of an LM
  // iteration (after the LM step, before the sweep) and is a chain of dependent fp64 operations: 1.1 k instructions per warp
  // at ~10 cycles each when one thread does everything for its pair (profiles/r02c_k_pair_setup.md).  The work of a pair is
  // therefore split by OUTPUT over the warps -- each recomputes the cheap products it needs -- so that the roles run on four
  // schedulers at once; every output is computed by the same sequence of operations as before.
  //   phase 1  warp 0: exp(-eps) of every frame;  warp 1: T_lin^-1 and the affine state of every frame;
  //            warp 2: exp(+eps) and T_lin of r
  //   phase 2  warp 0: t_t_r -> transform_unproject_ / reproject_ / translation_ at the current state;
  //            warp 1: t_t_r -> Adj at the current state;  warp 2: t_t_r0 -> the same at the linearisation point, scalars;
  //            warp 3: t_t_r0 -> Adj at the linearisation point
  //   phase 3  coalesced write of the staged records
  __shared__ SE3d s_et[PBA_MAXF], s_ti[PBA_MAXF];
  __shared__ SE3d s_er, s_tl;
  __shared__ double s_a[PBA_MAXF], s_b[PBA_MAXF];
  __shared__ PairConst s_pc[PBA_MAXF];
  __shared__ PairAssemble s_pa[PBA_MAXF];
  const int tid = threadIdx.x, role = tid >> 5, lane = tid & 31;
  const int r = blockIdx.x;
  if (lane < N) {
    const FrameParams& F = fr[lane];
    if (role == 0 || (role == 2 && lane == r)) {
      double e[6];
      for (int k = 0; k < 6; ++k) e[k] = F.eps[k] + F.step[k];
      if (role == 0) se3_exp(e, -1.0, s_et[lane]);
      else se3_exp(e, 1.0, s_er);
    }
    if (role == 1 || (role == 2 && lane == r)) {
      SE3d T;
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T.R[i * 3 + j] = F.T_lin[i * 4 + j];
        T.t[i] = F.T_lin[i * 4 + 3];
      }
      if (role == 1) {
        se3_inv(T, s_ti[lane]);
        s_a[lane] = F.ab0[0] + F.eps[6] + F.step[6];
        s_b[lane] = F.ab0[1] + F.eps[7] + F.step[7];
      } else {
        s_tl = T;
      }
    }
  }
  __syncthreads();
  const int t = lane;
  if (t < N && t != r) {
    const FrameParams& R = fr[r];
    const FrameParams& T = fr[t];
    PairConst& pc = s_pc[t];
    PairAssemble& pa = s_pa[t];
    SE3d T0;
    se3_mul(s_ti[t], s_tl, T0);  // t_t_r0 (evaluate_jacobians.hpp:47-48)
    if (role < 2) {
      SE3d tmp, Tc;
      se3_mul(T0, s_er, tmp);
      se3_mul(s_et[t], tmp, Tc);   // t_t_r = exp(-eps_t) T0 exp(eps_r)  (:49)
      if (role == 0) {
        make_proj(Tc, R.intr, T.intr, pc.M, pc.A);
        for (int i = 0; i < 3; ++i) pc.tr[i] = (float)Tc.t[i];
        pc.tr[3] = 0.f;
      } else {
        se3_adj(Tc, pa.adj_cur);
        for (int i = 0; i < 36; ++i) pc.adj[i] = (float)pa.adj_cur[i];
      }
    } else if (role == 2) {
      make_proj(T0, R.intr, T.intr, pc.M0, nullptr);
      for (int i = 0; i < 3; ++i) pc.t0[i] = (float)T0.t[i];
      pc.t0[3] = 0.f;
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
      pc.pad0 = pc.pad1 = 0.f;
    } else {
      se3_adj(T0, pa.adj_fej);
      for (int i = 0; i < 36; ++i) pc.adj0[i] = (float)pa.adj_fej[i];
    }
  }
  __syncthreads();
  for (int tt = 0; tt < N; ++tt) {
    if (tt == r) continue;
    const float4* src = reinterpret_cast<const float4*>(&s_pc[tt]);
    float4* dst = reinterpret_cast<float4*>(&pairs[r * PBA_MAXF + tt]);
    for (int i = tid; i < (int)(sizeof(PairConst) / 16); i += blockDim.x) dst[i] = src[i];
    const double* ps = reinterpret_cast<const double*>(&s_pa[tt]);
    double* pd = reinterpret_cast<double*>(&pasm[r * PBA_MAXF + tt]);
    for (int i = tid; i < (int)(sizeof(PairAssemble) / 8); i += blockDim.x) pd[i] = ps[i];
  }
}

// residual vectors of a freshly pushed frame: rows (phys -> p) and (p -> phys) of status / cand / jac_valid / energy := 0
// (ResidualPoint ctor, local_frame.hpp:212-219).  grid = (ceil(mp / 256), max_frames, 2)
__global__ void k_clear_frame_rows(uint8_t* __restrict__ status, uint8_t* __restrict__ cand,
                                   uint8_t* __restrict__ jac_valid, float* __restrict__ energy, int phys, int mp) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= mp) return;
  const int p = blockIdx.y;
  const size_t row = blockIdx.z ? (size_t)(p * PBA_MAXF + phys) : (size_t)(phys * PBA_MAXF + p);
  const size_t i = row * mp + l;
  status[i] = 0;
  cand[i] = 0;
  jac_valid[i] = 0;
  energy[i] = 0.f;
}

// Device image layout: per pixel x a 32-byte record {texel(x), texel(x + 1)}, texel = {I, dx, dy, 0} (the right
// neighbour is clamped at the last column, which the 4-pixel ROI border keeps from ever being sampled).
// {I,dx,dy} float3 -> records
__global__ void k_pack_image(const float* __restrict__ src, float4* __restrict__ dst, int n, int W) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = ((i % W) == W - 1) ? i : i + 1;
  dst[2 * (size_t)i] = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
  dst[2 * (size_t)i + 1] = make_float4(src[3 * j], src[3 * j + 1], src[3 * j + 2], 0.f);
}

// gradient packing from the intensity plane, features/src/calculate_pixelinfo.cpp:340-374
__device__ __forceinline__ float4 pixelinfo_at(const float* __restrict__ I, int x, int y, int W, int H) {
  const float c = I[y * W + x];
  float dx, dy;
  if (x == 0) dx = 1.0f * (I[y * W + 1] - c);
  else if (x == W - 1) dx = 1.0f * (c - I[y * W + x - 1]);
  else dx = 0.5f * (I[y * W + x + 1] - I[y * W + x - 1]);
  const int yu = y == 0 ? y : y - 1, yb = y == H - 1 ? y : y + 1;
  dy = ((y == 0 || y == H - 1) ? 1.0f : 0.5f) * (I[yb * W + x] - I[yu * W + x]);
  return make_float4(c, dx, dy, 0.f);
}
// photometricallyCorrectedImage, features/src/photometrically_corrected_image.cpp:9-29:
// I = lut[gray] * (max_vignetting / (vignetting + 1)); lut == null: identity, vignetting == null: no second factor
__global__ void k_photometric(const uint8_t* __restrict__ gray, const float* __restrict__ lut,
                              const uint8_t* __restrict__ vignetting, float max_v, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t g = gray[i];
  float v = lut ? lut[g] : (float)g;
  if (vignetting) v = __fmul_rn(v, __fdiv_rn(max_v, __fadd_rn((float)vignetting[i], 1.f)));
  out[i] = v;
}

// downscaleImage, features/internal/features/camera/downscale_image.hpp:16-33: 0.25 * (((A + B) + C) + D) with
// A = (even, even), B = (odd, odd), C = (even, odd), D = (odd, even) -- the reference's summation order
__global__ void k_downscale(const float* __restrict__ src, float* __restrict__ dst, int W, int H) {
  const int W2 = W / 2, H2 = H / 2;
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W2 || y >= H2) return;
  const float a = src[(2 * y) * W + 2 * x], b = src[(2 * y + 1) * W + 2 * x + 1];
  const float c = src[(2 * y) * W + 2 * x + 1], d = src[(2 * y + 1) * W + 2 * x];
  dst[y * W2 + x] = __fmul_rn(0.25f, __fadd_rn(__fadd_rn(__fadd_rn(a, b), c), d));
}

// {I,dx,dy} interleaved float3 (the PixelMap<1> storage) from the intensity plane, for host consumers
__global__ void k_pixelinfo3(const float* __restrict__ I, float* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  const float4 t = pixelinfo_at(I, x, y, W, H);
  float* o = dst + 3 * ((size_t)y * W + x);
  o[0] = t.x;
  o[1] = t.y;
  o[2] = t.z;
}

__global__ void k_pixelinfo(const float* __restrict__ I, float4* __restrict__ dst, int W, int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= W || y >= H) return;
  dst[2 * ((size_t)y * W + x)] = pixelinfo_at(I, x, y, W, H);
  dst[2 * ((size_t)y * W + x) + 1] = pixelinfo_at(I, min(x + 1, W - 1), y, W, H);
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
  ctl->apply = 0;
  ctl->relin = 0;
  ctl->iteration = 0;
  ctl->iterations_executed = 0;
}

// calculateEnergy() tail (problem.hpp:293-316), the accept decision (levenberg_marquardt_algorithm.hpp:95-104) and,
// for a trial energy, acceptStep / rejectStep of the frame state plus the loop bookkeeping (problem.hpp:366-402,
// lm.hpp:104-122) -- one single-CTA kernel per energy evaluation.  k_accept_landmarks runs right after it and applies
// the decision to the landmarks when ctl->apply is set.
__device__ __forceinline__ double block_sum_256(double v, double* red /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  __syncthreads();  // red may still be read from the previous call
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += red[k];
  return s;
}

__device__ __forceinline__ void lm_energy_body(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr,
                                                   int N, double* scal, const double* Hmarg,
                                                   const double* bmarg, int kind, const double2* __restrict__ e_part,
                                                   int n_e, const double2* __restrict__ n_part, int n_n, int from_core) {
  // from_core: e_part is the per-pair core array (k_core_reduce): entry (r, t) holds (energy, n_valid) of the pair in
  // slots 44 / 45 of its 48-double record; n_e = N (N - 1) ordered pairs
  if (kind == pba::LM_ENERGY_TRIAL && ctl->done) {
    if (threadIdx.x == 0) ctl->apply = 0;
    return;
  }
  __shared__ double s[PBA_MAXF * 8];
  __shared__ double red[8];
  __shared__ int s_accept;
  const int D = 8 * N, i = threadIdx.x;
  stamp(50);
  // Thread 0 takes the decision.  It reads the loop state and the options ONCE, up front and in one batch of independent loads
  // that overlaps the reductions below, keeps everything in registers, and only stores from then on: the decision used to be a
  // chain of dependent L2 round trips (ctl->energy, ctl->iteration, scal[] just written by the same thread, ...).
  LmCtl c;
  LmOptionsDev o;
  if (i == 0) {
    c = *ctl;
    o = *opt;
  }
  double sa = 0, sb = 0, sc = 0, sd = 0;  // (energy, n_valid, landmark state norm, landmark step norm) summed over everything
  if (e_part) {  // single-GPU: the second stage of the (energy, n) / norm reductions happens here, no extra launch
    double a = 0, b = 0, cc = 0, d = 0;
    for (int k = i; k < n_e; k += 256) {
      size_t idx = k;
      if (from_core) {
        const int r = k / (N - 1);
        int t = k % (N - 1);
        t += (t >= r);
        idx = ((size_t)(r * PBA_MAXF + t) * PBA_CORE + 44) / 2;
      }
      const double2 v = e_part[idx];
      a += v.x;
      b += v.y;
    }
    for (int k = i; k < n_n; k += 256) {
      const double2 v = n_part[k];
      cc += v.x;
      d += v.y;
    }
    sa = block_sum_256(a, red);
    sb = block_sum_256(b, red);
    if (n_part) {
      sc = block_sum_256(cc, red);
      sd = block_sum_256(d, red);
    }
    if (i == 0) {
      scal[0] = sa;
      scal[1] = sb;
      if (n_part) {
        scal[2] = sc;
        scal[3] = sd;
      }
    }
  } else if (i == 0) {  // several GPUs: the exchanged sums
    sa = scal[0];
    sb = scal[1];
  }
  if (i == 0 && !(e_part && n_part)) {  // norms that an earlier kernel (or the exchange) left in the scalar slots
    sc = scal[2];
    sd = scal[3];
  }
  stamp(51);
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
  const double prior_e = block_sum_256(acc, red);
  if (i == 0) {
    const double E = o.energy_marg + prior_e + sa;
    const int n = (int)llrint(sb);
    if (kind == pba::LM_ENERGY_INITIAL) {
      ctl->energy = E;
      ctl->n_valid = n;
      if (n <= 0 || o.max_it <= 0) ctl->done = 1;
    } else if (kind == pba::LM_ENERGY_TRIAL) {
      c.next_energy = E;
      c.next_n = n;
      ctl->next_energy = E;
      ctl->next_n = n;
      ctl->iterations_executed = c.iterations_executed + 1;
      // landmark parts of the norms, accumulated by k_back_substitute (problem.hpp:377-382)
      c.state_sq = sc;
      c.step_sq = sd;
      ctl->state_sq = sc;
      ctl->step_sq = sd;
      if (n == 0) {
        c.accept = -1;  // rejectStep(); break
      } else {
        if (fabs(c.energy - E) / c.energy < o.ftol) {  // before the accept test (Q7)
          c.converged = 1;
          ctl->converged = 1;
        }
        c.accept = (E < c.energy || (o.force_accept && c.iteration < o.min_it)) ? 1 : 0;
      }
      ctl->accept = c.accept;
      if (c.accept <= 0) ctl->relin = 1;  // speculative mode: the landmark fields now belong to a rejected state
      ctl->apply = 1;
      s_accept = c.accept;
    }
  }
  stamp(52);
  if (kind != pba::LM_ENERGY_TRIAL) return;
  __syncthreads();
  // acceptStep / rejectStep for the frame state (problem.hpp:366-376,392-402); one thread per state entry
  const int accept = s_accept;
  double st = 0, sp = 0;
  if (i < D) {
    const int f = i / 8, k = i % 8;
    const double e = fr[f].eps[k], stp = fr[f].step[k];
    if (accept > 0) {
      st = e * e;
      if (k >= 6) st += fr[f].ab0[k - 6] * fr[f].ab0[k - 6];
      sp = stp * stp;
      fr[f].eps[k] = e + stp;
    }
    fr[f].step[k] = 0;
  }
  st = block_sum_256(st, red);
  sp = block_sum_256(sp, red);
  if (i) return;
  if (accept > 0) {
    const double stt = c.state_sq + st, spp = c.step_sq + sp;
    ctl->state_sq = stt;
    ctl->step_sq = spp;
    if (spp < o.ptol * (stt + o.ptol)) {
      c.converged = 1;
      ctl->converged = 1;
    }
    ctl->energy = c.next_energy;
    ctl->n_valid = c.next_n;
    c.n_valid = c.next_n;
    ctl->lambda = c.lambda / o.dec;
    ctl->system_valid = 0;
  } else {
    if (accept < 0 || o.force_accept) {
      ctl->done = 1;
      return;
    }
    ctl->lambda = c.lambda * o.inc;
    ctl->system_valid = 1;
  }
  c.iteration += 1;
  ctl->iteration = c.iteration;
  if (c.iteration >= o.max_it || c.converged || c.n_valid <= 0) ctl->done = 1;
}

__global__ void __launch_bounds__(256) k_lm_energy(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr,
                                                   int N, double* scal, const double* Hmarg,
                                                   const double* bmarg, int kind, const double2* __restrict__ e_part,
                                                   int n_e, const double2* __restrict__ n_part, int n_n, int from_core,
                                                   int peer_collect_mode, int peer_scal_off) {
  KStamp kstamp_(2);
  if (peer_collect_mode && !(kind == pba::LM_ENERGY_TRIAL && ctl->done)) {
    // fused exchange: the scalars of every rank (one arrival each), summed in rank order
    const unsigned epoch = peer_epoch();
    peer_wait_arrivals(0, epoch, 1u);
    peer_collect(scal, (size_t)peer_scal_off, 8, epoch);
    __syncthreads();
    if (peer_collect_mode == 2 && threadIdx.x == 0) peer_close(epoch);
    __syncthreads();
  }
  lm_energy_body(ctl, opt, fr, N, scal, Hmarg, bmarg, kind, e_part, n_e, n_part, n_n, from_core);
}

// The energy decision straight from the sweep's per-chunk records (slots 44 / 45 of every [frame][chunk][target] record
// carry the pair energies and valid counts): k_core_reduce leaves the critical path of an iteration -- the block
// assembly reduces the cores it needs itself, on a side branch (k_reduce_system).
__global__ void __launch_bounds__(256) k_lm_energy_cp(const __grid_constant__ WindowDev w, LmCtl* ctl, const LmOptionsDev* opt,
                                                      FrameParams* fr, double* scal, const double* Hmarg, const double* bmarg,
                                                      int kind, const float* __restrict__ core_part, int lpb, int chunks,
                                                      const double2* __restrict__ n_part, int n_n) {
  if (!(kind == pba::LM_ENERGY_TRIAL && ctl->done)) {
    __shared__ double red4[4][8];
    const int N = w.n_frames;
    double a = 0, b = 0, c = 0, d = 0;
    const int recs = N * chunks * (N - 1);  // [frame][chunk][target warp]
    for (int i = threadIdx.x; i < recs; i += 256) {
      const int f = i / (chunks * (N - 1)), ch = (i / (N - 1)) % chunks;
      if (ch * lpb < w.n_lm[f]) {  // chunk CTAs past a frame's last landmark never ran
        const float2 v = *reinterpret_cast<const float2*>(core_part + (size_t)i * PBA_CORE + 44);
        a += (double)v.x;
        b += (double)v.y;
      }
    }
    for (int i = threadIdx.x; i < n_n; i += 256) {
      const double2 v = n_part[i];
      c += v.x;
      d += v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(FULL, a, o);
      b += __shfl_xor_sync(FULL, b, o);
      c += __shfl_xor_sync(FULL, c, o);
      d += __shfl_xor_sync(FULL, d, o);
    }
    if ((threadIdx.x & 31) == 0) {
      red4[0][threadIdx.x >> 5] = a;
      red4[1][threadIdx.x >> 5] = b;
      red4[2][threadIdx.x >> 5] = c;
      red4[3][threadIdx.x >> 5] = d;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double t = 0;
      for (int k = 0; k < 8; ++k) t += red4[threadIdx.x][k];
      if (threadIdx.x < 2 || n_part) scal[threadIdx.x] = t;
    }
    __syncthreads();
  }
  lm_energy_body(ctl, opt, fr, w.n_frames, scal, Hmarg, bmarg, kind, nullptr, 0, nullptr, 0, 0);
}

// calculateStep (problem.hpp:342-357): priors (problem.hpp:37-77), full system, Jacobi preconditioner + LDL^T
// (normal_linear_system.cpp:10-59), all fp64 in one CTA.
//
// Blocked right-looking LDL^T, block size 8, on the lower triangle held in shared memory.  The right-hand side rides
// along as row D of the matrix, so z = L^-1 b falls out of the factorisation.  Per block step (2 barriers):
//   panel : thread i owns row kb+i (rows m0 .. m0 + 7, m0 = kb + 8, are the look-ahead warp's).  It reads the FACTORED 8x8 diagonal block (packed triangle + reciprocals, 44 broadcast
//           loads from Gf) and runs the block's recurrence on its own 8 panel entries
//             a[c] -= a[j] * l_cj   (j < c),   l_cj = x_cj / d_j,   x = L D  ("raw" columns)
//   update: A22 -= X L21^T over the trailing lower triangle, 16x16 thread tiling with 2x2 register tiles, K = 8.
//   look-ahead: the eight dependent reciprocals of a diagonal block (~90 cycles each) used to open every block step.
//           Warp 7 now runs AHEAD of the step.  After the top barrier it does the panel work of rows m0 .. m0 + 7 (the
//           next diagonal block's own rows), ARRIVES on the panel barrier without waiting (named barrier 1: the other
//           seven warps bar.sync on it), then updates the 36 entries of the next diagonal block out of the trailing
//           update, factors them and leaves the result in the other half of Gf -- all while the other warps do their
//           panel rows and the rest of the trailing update.  Same sums in the same order: the factorisation is
//           bit-identical to the plain blocked form.
// A column-by-column factorisation needs 8N barrier-separated steps whose critical path (publish column, barrier,
// reciprocal, update) measured ~1500 cycles each on B200 (profiles/r01d_k_lm_step.md); this form has N of them.
__device__ __forceinline__ double rcp64(double d) {
  // IEEE division: measured 67 cycles on B200 against 131 for an fp32 seed + two Newton steps (the f32<->f64
  // conversions dominate), tools/fp64_probe.cu
  return d == 0.0 ? 0.0 : 1.0 / d;
}

// LDL^T of an 8x8 block held as a packed lower triangle g[r (r + 1) / 2 + c] in registers: afterwards the diagonal holds
// d_j, the entries below it the raw columns x_rj = l_rj d_j, inv[j] = 1 / d_j
__device__ __forceinline__ void ldlt8_packed(double (&g)[36], double (&inv)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    inv[j] = rcp64(g[j * (j + 1) / 2 + j]);
#pragma unroll
    for (int r = j + 1; r < 8; ++r) {
      const double l = g[r * (r + 1) / 2 + j] * inv[j];
#pragma unroll
      for (int c = j + 1; c <= r; ++c) g[r * (r + 1) / 2 + c] -= l * g[c * (c + 1) / 2 + j];
    }
  }
}

template <int DP>
__device__ __forceinline__ void lm_step_body(const LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr,
                                             const int* fixed, int N, const double* __restrict__ Hp,
                                             const double* __restrict__ bp, const double* __restrict__ Hs,
                                             const double* __restrict__ bs, const double* __restrict__ Hmarg,
                                             const double* __restrict__ bmarg, double* __restrict__ step_dev) {
  const int done = ctl->done;  // consulted after the loads below have been issued: its round trip overlaps theirs
  constexpr int LD = DP + 1;   // row stride of S (odd: conflict-free column walks)
  constexpr int LPS = 9;       // row stride of Lp
  extern __shared__ double sh[];
  const int D = 8 * N, tid = threadIdx.x;
  double* S = sh;                       // [DP + 1][LD] lower triangle of the preconditioned system, row D = rhs
  double* Lp = S + (DP + 1) * LD;       // [DP + 1][LPS] l_cj of the current panel
  double* dinv = Lp + (DP + 1) * LPS;   // [DP] 1 / d
  double* pre = dinv + DP;              // [DP] Jacobi preconditioner
  double* st = pre + DP;                // [DP] state eps, later the solution
  double* hm = st + DP;                 // [DP] H_marg * state
  double* Gf = hm + DP;                 // [2][48] factored diagonal blocks (36 packed + 8 reciprocals) of this / the next step
  const double lambda = ctl->lambda, ks = -1.0 / (1.0 + lambda);
  // Every global input of the fill is requested UP FRONT, in one batch of independent loads: the first 16 x 256 entries of
  // (H_pose, H_schur[, H_marg]) -- the whole 64 x 64 system of an 8-keyframe window --, the diagonal, the right-hand sides,
  // the state and the options.  The fill used to walk five dependent L2 round trips (state, diagonal, two trips of the
  // triangle, right-hand side) and took 10.5 k of the kernel's 40 k cycles (tools/lm_stamps.py).
  constexpr int PRE = 16;
  double hp0[PRE], hs0[PRE], hg0[PRE];
#pragma unroll
  for (int q = 0; q < PRE; ++q) {
    const int idx = q * 256 + tid;
    const bool in = idx < D * D;
    hp0[q] = in ? Hp[idx] : 0.0;
    hs0[q] = in ? Hs[idx] : 0.0;
    hg0[q] = (in && Hmarg) ? Hmarg[idx] : 0.0;
  }
  const double stv = (tid < D) ? fr[tid / 8].eps[tid % 8] : 0.0;
  double dg_hp = 0, dg_hs = 0, dg_hm = 0, v_bp = 0, v_bs = 0, v_bm = 0, v_ab0 = 0;
  int v_fixed = 0;
  if (tid < D) {
    const size_t idx = (size_t)tid * D + tid;
    dg_hp = Hp[idx];
    dg_hs = Hs[idx];
    if (Hmarg) {
      dg_hm = Hmarg[idx];
      v_bm = bmarg[tid];
    }
    v_bp = bp[tid];
    v_bs = bs[tid];
    v_fixed = fixed[tid / 8];
    if (tid % 8 >= 6) v_ab0 = fr[tid / 8].ab0[tid % 8 - 6];
  }
  const double fixed_reg = opt->fixed_reg, ab_reg0 = opt->ab_reg[0], ab_reg1 = opt->ab_reg[1];
  if (done) return;
  if (tid < DP) st[tid] = stv;
  __syncthreads();
  if (Hmarg) {  // warp per row, lanes along the row: coalesced
    const int warp = tid >> 5, lane = tid & 31;
    for (int i = warp; i < D; i += 8) {
      double t = 0;
      for (int j = lane; j < D; j += 32) t += Hmarg[(size_t)i * D + j] * st[j];
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) t += __shfl_xor_sync(FULL, t, s);
      if (lane == 0) hm[i] = t;
    }
  }
  // unscaled system:  H_pose(+priors, +lambda on the diagonal) + H_marg - H_schur / (1 + lambda)
  if (tid < D) {
    const int k = tid % 8;
    double hp = dg_hp;
    if (v_fixed) hp += fixed_reg;
    else if (k >= 6) hp += (k == 6 ? ab_reg0 : ab_reg1);
    hp += hp * lambda;  // H.diagonal() += system_pose.H.diagonal() * lambda (prior included)
    const double v = hp + dg_hm + ks * dg_hs;
    dinv[tid] = v;                        // parked here until the fill below
    pre[tid] = 1.0 / sqrt(v + 10.0);      // jacobiPreconditioner, +10 floor
  }
  __syncthreads();
  // fill the lower triangle from the registers ...
#pragma unroll
  for (int q = 0; q < PRE; ++q) {
    const int idx = q * 256 + tid;
    const int i = idx / D, j = idx - i * D;
    if (idx < D * D && j <= i) S[i * LD + j] = (i == j ? dinv[i] : hp0[q] + hg0[q] + ks * hs0[q]) * pre[i] * pre[j];
  }
  // ... and, for windows of more than 8 keyframes, the rest of it in trips of 8 independent loads per thread
  for (int base = PRE * 256; base < D * D; base += 256 * 8) {
    double hp[8], hs[8], hg[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = base + q * 256 + tid;
      const bool in = idx < D * D;
      hp[q] = in ? Hp[idx] : 0.0;
      hs[q] = in ? Hs[idx] : 0.0;
      hg[q] = (in && Hmarg) ? Hmarg[idx] : 0.0;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = base + q * 256 + tid;
      const int i = idx / D, j = idx - i * D;
      if (idx < D * D && j <= i) S[i * LD + j] = (i == j ? dinv[i] : hp[q] + hg[q] + ks * hs[q]) * pre[i] * pre[j];
    }
  }
  if (tid < D) {
    const int k = tid % 8;
    double b = v_bp + ks * v_bs;
    if (v_fixed) b += fixed_reg * stv;
    else if (k >= 6) b += (k == 6 ? ab_reg0 : ab_reg1) * (v_ab0 + stv);
    if (Hmarg) b += v_bm + hm[tid];
    S[D * LD + tid] = b * pre[tid];
  }
  const int ty = tid >> 4, tx = tid & 15;
  const int warp_id = tid >> 5, lane_id = tid & 31;
  constexpr int LOOKAHEAD_WARP = 7;
  stamp(2);
  __syncthreads();
  if (warp_id == LOOKAHEAD_WARP) {  // the first diagonal block
    double g[36], inv[8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int c = 0; c <= r; ++c) g[r * (r + 1) / 2 + c] = S[r * LD + c];
    ldlt8_packed(g, inv);
    if (lane_id == 0) {
#pragma unroll
      for (int e = 0; e < 36; ++e) Gf[e] = g[e];
#pragma unroll
      for (int j = 0; j < 8; ++j) Gf[36 + j] = inv[j];
    }
  }
  for (int kb = 0; kb < D; kb += 8) {
    const double* Gc = Gf + 48 * ((kb >> 3) & 1);  // this step's factored diagonal block
    double* Gn = Gf + 48 * (((kb >> 3) & 1) ^ 1);  // the next step's, written by the look-ahead warp meanwhile
    __syncthreads();  // the trailing matrix is up to date and Gc is complete
    if (kb < 64) stamp(3 + kb / 8);
    const int m0 = kb + 8;
    double g[36], inv[8], a[8];
    if (warp_id == LOOKAHEAD_WARP) {
      // The look-ahead warp runs AHEAD of the block step: it does the panel work of the next diagonal block's own rows
      // (m0 .. m0 + 7; lanes 0 .. 7), hands them to the others (arrive on the panel barrier, no wait), then updates and
      // factors the next diagonal block while the other warps are still in their panel phase and trailing update.
      // (keeping the factored block in this warp's registers across the steps instead of re-reading it measured slower)
#pragma unroll
      for (int e = 0; e < 36; ++e) g[e] = Gc[e];
#pragma unroll
      for (int j = 0; j < 8; ++j) inv[j] = Gc[36 + j];
      const int i = m0 + lane_id;
      if (lane_id < 8 && i <= D) {
#pragma unroll
        for (int c = 0; c < 8; ++c) a[c] = S[i * LD + kb + c];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int c = j + 1; c < 8; ++c) a[c] -= a[j] * (g[c * (c + 1) / 2 + j] * inv[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          S[i * LD + kb + j] = a[j];
          Lp[i * LPS + j] = a[j] * inv[j];
        }
      }
      __threadfence_block();
      __syncwarp();
      asm volatile("bar.arrive 1, 256;" ::: "memory");
      if (m0 < D) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {  // the 8 x 8 square over 32 lanes x 2; the upper triangle idles
          const int r = (lane_id >> 3) + 4 * q, c = lane_id & 7;
          if (c <= r) {
            double acc = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += S[(m0 + r) * LD + kb + j] * Lp[(m0 + c) * LPS + j];
            const double v = S[(m0 + r) * LD + m0 + c] - acc;
            S[(m0 + r) * LD + m0 + c] = v;  // what the next step's panel threads read as their own row
            Gn[r * (r + 1) / 2 + c] = v;
          }
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 36; ++e) g[e] = Gn[e];
        ldlt8_packed(g, inv);
        __syncwarp();
        if (lane_id == 0) {
#pragma unroll
          for (int e = 0; e < 36; ++e) Gn[e] = g[e];
#pragma unroll
          for (int j = 0; j < 8; ++j) Gn[36 + j] = inv[j];
        }
      }
      continue;
    }
    const int i = kb + tid;  // this thread's row (row D is the right-hand side); rows m0 .. m0 + 7 are the look-ahead warp's
    if (i <= D && (tid < 8 || tid >= 16)) {
#pragma unroll
      for (int e = 0; e < 36; ++e) g[e] = Gc[e];
#pragma unroll
      for (int j = 0; j < 8; ++j) inv[j] = Gc[36 + j];
#pragma unroll
      for (int c = 0; c < 8; ++c) a[c] = (i >= kb + c) ? S[i * LD + kb + c] : 0.0;
      // own row: the block's recurrence (for a row of the diagonal block the entries right of the diagonal are unused)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int c = j + 1; c < 8; ++c) a[c] -= a[j] * (g[c * (c + 1) / 2 + j] * inv[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (i >= kb + j) {
          S[i * LD + kb + j] = a[j];        // raw column entry x_ij = l_ij d_j
          Lp[i * LPS + j] = a[j] * inv[j];  // l_ij
        }
      }
      if (tid == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) dinv[kb + j] = inv[j];
      }
    }
    if (kb == 0) stamp(13);
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the panel is complete (the look-ahead warp only arrives)
    if (kb == 0) stamp(14);
    // trailing update of the lower triangle (and the rhs row): S[r][c] -= sum_j x_rj l_cj, rows below the next diagonal block.
    // 2 x 2 register tile per thread -- rows {r, r + 14} x columns {c, c + 16}, so that the 16 lanes of a half-warp keep
    // walking Lp and S with the conflict-free strides -- : a panel entry l_cj fetched from shared memory serves two rows
    // and a row's x_rj two columns; every element is still the same 8-term sum in the same order.
    const int own_lo = min(m0 + 8, D);  // rows [m0, own_lo) belong to the look-ahead warp, which takes no other share
    for (int r = m0 + ty; r <= D; r += 28) {
      const int r1 = r + 14;
      const bool has0 = r >= own_lo, has1 = r1 <= D;  // (r1 >= m0 + 14 is never a look-ahead row)
      double x0[8], x1[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x0[j] = has0 ? S[r * LD + kb + j] : 0.0;
        x1[j] = has1 ? S[r1 * LD + kb + j] : 0.0;
      }
      const int cmax0 = has0 ? min(r, D - 1) : -1, cmax1 = has1 ? min(r1, D - 1) : -1;
      const int cend = max(cmax0, cmax1);
      for (int c = m0 + tx; c <= cend; c += 32) {
        const int c1 = c + 16;
        const bool hc1 = c1 <= cend;
        double a00 = 0, a01 = 0, a10 = 0, a11 = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double l0 = Lp[c * LPS + j], l1 = hc1 ? Lp[c1 * LPS + j] : 0.0;
          a00 += x0[j] * l0;
          a01 += x0[j] * l1;
          a10 += x1[j] * l0;
          a11 += x1[j] * l1;
        }
        if (c <= cmax0) S[r * LD + c] -= a00;
        if (c1 <= cmax0) S[r * LD + c1] -= a01;
        if (c <= cmax1) S[r1 * LD + c] -= a10;
        if (c1 <= cmax1) S[r1 * LD + c1] -= a11;
      }
    }
  }
  __syncthreads();
  stamp(18);
  if (tid < D) st[tid] = S[D * LD + tid] * dinv[tid];  // z = D^-1 L^-1 b
  // blocked back substitution L^T x = z, l_ki = x_ki / d_i: every thread solves the 8x8 triangle of the block
  // redundantly in registers, then row i < kb subtracts the block's contribution -- one barrier per 8 unknowns
  for (int kb = D - 8; kb >= 0; kb -= 8) {
    __syncthreads();
    // warps without a row above the block (and that do not publish the solution) skip the redundant 8 x 8 solve: its 44
    // broadcast loads per warp were half of a block step's time on the CTA's one load / store unit
    if ((tid & ~31) >= kb && tid >= 32) continue;
    double x[8];
#pragma unroll
    for (int j = 7; j >= 0; --j) {
      double v = st[kb + j];
      const double dj = dinv[kb + j];
#pragma unroll
      for (int m = 7; m > j; --m) v -= S[(kb + m) * LD + kb + j] * dj * x[m];
      x[j] = v;
    }
    if (tid == 0) {  // the solution is collected in hm (dead since the right-hand side was formed): no second barrier
#pragma unroll
      for (int j = 0; j < 8; ++j) hm[kb + j] = x[j];
    }
    if (tid < kb) {
      double acc = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc += S[(kb + j) * LD + tid] * x[j];
      st[tid] -= dinv[tid] * acc;
    }
  }
  __syncthreads();
  if (tid < D) {
    const double x = hm[tid] * pre[tid];
    step_dev[tid] = x;
    fr[tid / 8].step[tid % 8] = -x;  // frame.state_eps_step = -frame_step (problem.hpp:353-357)
  }
}

template <int DP>
__global__ void __launch_bounds__(256, 1) k_lm_step(const LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr,
                                                 const int* fixed, int N, const double* __restrict__ Hp,
                                                 const double* __restrict__ bp, const double* __restrict__ Hs,
                                                 const double* __restrict__ bs, const double* __restrict__ Hmarg,
                                                 const double* __restrict__ bmarg, double* __restrict__ step_dev,
                                                 int peer_expected, double* sys_out, int n_sys) {
  KStamp kstamp_(5);
  stamp(22);
  if (peer_expected && !ctl->done) {
    // fused exchange: [H_pp | b_p | H_s | b_s] of every rank, summed in rank order into the block lm_step_body reads
    const unsigned epoch = peer_epoch();
    peer_wait_arrivals(1, epoch, (unsigned)peer_expected);
    peer_collect(sys_out, 0, n_sys, epoch);
    __syncthreads();
    if (threadIdx.x == 0) peer_close(epoch);
    __syncthreads();
  }
  lm_step_body<DP>(ctl, opt, fr, fixed, N, Hp, bp, Hs, bs, Hmarg, bmarg, step_dev);
  stamp(23);
}

// Per-pair constants by ONE CTA: thread per frame for the exponentials, then thread per ORDERED PAIR (same arithmetic,
// same results as k_pair_setup, which uses a CTA per reference frame) -- the tail of k_lm_solve.
__device__ __forceinline__ void pair_setup_body(const FrameParams* __restrict__ fr, int N, PairConst* __restrict__ pairs,
                                                PairAssemble* __restrict__ pasm) {
  __shared__ SE3d s_et[PBA_MAXF], s_ti[PBA_MAXF], s_ep[PBA_MAXF], s_tl[PBA_MAXF];
  __shared__ double s_a[PBA_MAXF], s_b[PBA_MAXF];
  const int tid = threadIdx.x;
  if (tid < N) {
    const FrameParams& F = fr[tid];
    double e[6];
    for (int k = 0; k < 6; ++k) e[k] = F.eps[k] + F.step[k];
    se3_exp(e, -1.0, s_et[tid]);
    se3_exp(e, 1.0, s_ep[tid]);
    SE3d T;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T.R[i * 3 + j] = F.T_lin[i * 4 + j];
      T.t[i] = F.T_lin[i * 4 + 3];
    }
    se3_inv(T, s_ti[tid]);
    s_tl[tid] = T;
    s_a[tid] = F.ab0[0] + F.eps[6] + F.step[6];
    s_b[tid] = F.ab0[1] + F.eps[7] + F.step[7];
  }
  __syncthreads();
  for (int p = tid; p < N * (N - 1); p += blockDim.x) {
    const int r = p / (N - 1);
    int t = p % (N - 1);
    t += (t >= r);
    const FrameParams& R = fr[r];
    const FrameParams& T = fr[t];
    SE3d T0, tmp, Tc;
    se3_mul(s_ti[t], s_tl[r], T0);  // t_t_r0 (evaluate_jacobians.hpp:47-48)
    se3_mul(T0, s_ep[r], tmp);
    se3_mul(s_et[t], tmp, Tc);      // t_t_r = exp(-eps_t) T0 exp(eps_r)  (:49)
    PairConst pc;
    PairAssemble& pa = pasm[r * PBA_MAXF + t];
    make_proj(Tc, R.intr, T.intr, pc.M, pc.A);
    make_proj(T0, R.intr, T.intr, pc.M0, nullptr);
    for (int i = 0; i < 3; ++i) {
      pc.tr[i] = (float)Tc.t[i];
      pc.t0[i] = (float)T0.t[i];
    }
    pc.tr[3] = pc.t0[3] = 0.f;
    double adj_cur[36], adj_fej[36];
    se3_adj(Tc, adj_cur);
    se3_adj(T0, adj_fej);
    for (int i = 0; i < 36; ++i) {
      pc.adj[i] = (float)adj_cur[i];
      pc.adj0[i] = (float)adj_fej[i];
      pa.adj_cur[i] = adj_cur[i];
      pa.adj_fej[i] = adj_fej[i];
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
    pc.pad0 = pc.pad1 = 0.f;
    const float4* src = reinterpret_cast<const float4*>(&pc);
    float4* dst = reinterpret_cast<float4*>(&pairs[r * PBA_MAXF + t]);
#pragma unroll
    for (int i = 0; i < (int)(sizeof(PairConst) / 16); ++i) dst[i] = src[i];
  }
}

// One launch for the serial middle of a device-LM iteration: the energy decision (k_lm_energy), calculateStep
// (k_lm_step) and the per-pair constants of the new trial state (k_pair_setup) -- three single-CTA kernels and their
// launch boundaries become one.  with_step = 0: decision only (the last evaluation of a solve).
template <int DP>
__global__ void __launch_bounds__(256, 1) k_lm_solve(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, const int* fixed,
                                                  int N, double* scal, const double* __restrict__ Hp,
                                                  const double* __restrict__ bp, const double* __restrict__ Hs,
                                                  const double* __restrict__ bs, const double* __restrict__ Hmarg,
                                                  const double* __restrict__ bmarg, double* __restrict__ step_dev, int kind,
                                                  const double2* __restrict__ e_part, int n_e,
                                                  const double2* __restrict__ n_part, int n_n, int from_core, int with_step,
                                                  PairConst* __restrict__ pairs, PairAssemble* __restrict__ pasm) {
  cudaGridDependencySynchronize();
  stamp(0);
  lm_energy_body(ctl, opt, fr, N, scal, Hmarg, bmarg, kind, e_part, n_e, n_part, n_n, from_core);
  stamp(1);
  if (!with_step) return;
  __syncthreads();  // ctl / frame state written by thread 0 and the state threads above
  if (ctl->done) return;
  lm_step_body<DP>(ctl, opt, fr, fixed, N, Hp, bp, Hs, bs, Hmarg, bmarg, step_dev);
  __syncthreads();  // fr[].step
  stamp(20);
  pair_setup_body(fr, N, pairs, pasm);
  stamp(21);
}


// Programmatic dependent launch: the kernel may be SCHEDULED while its predecessor in the stream drains (its CTAs are
// resident and their prologue runs); it calls cudaGridDependencySynchronize() before touching anything the predecessor
// wrote.  Removes most of the launch gap between the small kernels of a device-LM iteration; captured into the graph as a
// programmatic edge.
bool g_pdl = false;  // measured: no effect on the captured LM graph (profiles/r02_ab.md); option "pdl"
template <typename... KArgs, typename... Args>
void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

int max_landmarks(const WindowDev& w) {
  int m = 0;
  for (int f = 0; f < w.n_frames; ++f) m = w.n_lm[f] > m ? w.n_lm[f] : m;
  return m;
}

}  // namespace

namespace pba {

std::atomic<long long> g_launches{0};
bool g_schur_mma = false;  // true: Schur SYRK as 3xTF32 mma.sync (faster inner product, ~10x larger rounding error in the reduced system); default: fp32 FFMA kernel
void set_schur_mma(bool on) { g_schur_mma = on; }
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

void launch_lm_energy(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, int N, double* scal,
                      const double* Hmarg, const double* bmarg, int kind, cudaStream_t s, const double* e_part, int n_e,
                      const double* n_part, int n_n, int from_core, int peer_collect) {
  ++g_launches;
  const int D = 8 * N;
  k_lm_energy<<<1, 256, 0, s>>>(ctl, opt, fr, N, scal, Hmarg, bmarg, kind, reinterpret_cast<const double2*>(e_part), n_e,
                                reinterpret_cast<const double2*>(n_part), n_n, from_core, peer_collect, 2 * (D * D + D));
}

void launch_lm_energy_from_records(const WindowDev& w, LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, double* scal,
                                   const double* Hmarg, const double* bmarg, int kind, ReduceBuf rb, FusedShape shape,
                                   const double* n_part, int n_n, cudaStream_t s) {
  ++g_launches;
  k_lm_energy_cp<<<1, 256, 0, s>>>(w, ctl, opt, fr, scal, Hmarg, bmarg, kind, rb.core_part, shape.lpb, shape.chunks,
                                   reinterpret_cast<const double2*>(n_part), n_n);
}

void launch_lm_step(const LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, const int* fixed, int N, ReduceBuf rb,
                    const double* Hmarg, const double* bmarg, double* step_dev, cudaStream_t s, int peer_expected) {
  const int D = 8 * N;
  const int n_sys = 2 * (D * D + D);
  ++g_launches;
  auto smem_of = [](int DP) { return (size_t)((DP + 1) * (DP + 1) + (DP + 1) * 9 + 4 * DP + 96) * sizeof(double); };
  if (D <= 64) {
    k_lm_step<64><<<1, 256, smem_of(64), s>>>(ctl, opt, fr, fixed, N, rb.Hp, rb.bp, rb.Hs, rb.bs, Hmarg, bmarg, step_dev,
                                              peer_expected, rb.Hp, n_sys);
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(k_lm_step<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(128));
      attr_set = true;
    }
    k_lm_step<128><<<1, 256, smem_of(128), s>>>(ctl, opt, fr, fixed, N, rb.Hp, rb.bp, rb.Hs, rb.bs, Hmarg, bmarg, step_dev,
                                                peer_expected, rb.Hp, n_sys);
  }
}

void launch_lm_solve(LmCtl* ctl, const LmOptionsDev* opt, FrameParams* fr, const int* fixed, int N, double* scal,
                     ReduceBuf rb, const double* Hmarg, const double* bmarg, double* step_dev, int kind,
                     const double* e_part, int n_e, const double* n_part, int n_n, int from_core, int with_step,
                     PairConst* pairs, PairAssemble* pasm, cudaStream_t s) {
  const int D = 8 * N;
  ++g_launches;
  auto smem_of = [](int DP) { return (size_t)((DP + 1) * (DP + 1) + (DP + 1) * 9 + 4 * DP + 96) * sizeof(double); };
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_lm_solve<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(64));
    cudaFuncSetAttribute(k_lm_solve<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of(128));
    attr_set = true;
  }
  if (D <= 64)
    launch_pdl(k_lm_solve<64>, dim3(1), dim3(256), smem_of(64), s, ctl, opt, fr, fixed, N, scal, (const double*)rb.Hp,
               (const double*)rb.bp, (const double*)rb.Hs, (const double*)rb.bs, Hmarg, bmarg, step_dev, kind,
               (const double2*)e_part, n_e, (const double2*)n_part, n_n, from_core, with_step, pairs, pasm);
  else
    launch_pdl(k_lm_solve<128>, dim3(1), dim3(256), smem_of(128), s, ctl, opt, fr, fixed, N, scal, (const double*)rb.Hp,
               (const double*)rb.bp, (const double*)rb.Hs, (const double*)rb.bs, Hmarg, bmarg, step_dev, kind,
               (const double2*)e_part, n_e, (const double2*)n_part, n_n, from_core, with_step, pairs, pasm);
}

void launch_reduce_system(const WindowDev& w, int fej, ReduceBuf rb, FusedShape shape, int with_system, cudaStream_t s,
                          const LmCtl* ctl, int with_schur) {
  const int N = w.n_frames, D = 8 * N;
  if (N < 2 || shape.lpb == 0) return;
  const int NB = N * (N + 1) / 2;
  const int T4 = D / 4, nout = T4 * (T4 + 1) / 2 * 16 + D;
  const int NF = (with_system && with_schur) ? (nout + 31) / 32 : 0;
  const size_t smem = (size_t)std::max(2, N - 1) * RSYS_W * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(k_reduce_system, cudaFuncAttributeMaxDynamicSharedMemorySize, 15 * RSYS_W * (int)sizeof(double));
    attr_set = true;
  }
  ++g_launches;
  launch_pdl(k_reduce_system, dim3(NB + NF), dim3(1024), smem, s, w, fej, (const float*)rb.core_part, shape.lpb, shape.chunks,
             rb.core, rb.Hp, rb.bp, (const float*)rb.fschur_part, rb.Hs, rb.bs, ctl, with_system);
}

void launch_pair_setup(const FrameParams* frames, int n_frames, PairConst* pairs, PairAssemble* pasm, cudaStream_t s) {
  ++g_launches;
  k_pair_setup<<<n_frames, 128, 0, s>>>(frames, n_frames, pairs, pasm);
}

void launch_clear_frame_rows(uint8_t* status, uint8_t* cand, uint8_t* jac_valid, float* energy, int phys, int mp,
                             int max_frames, cudaStream_t s) {
  ++g_launches;
  k_clear_frame_rows<<<dim3((mp + 255) / 256, max_frames, 2), 256, 0, s>>>(status, cand, jac_valid, energy, phys, mp);
}

void launch_pack_image(const float* src3, float4* dst, int n_px, int W, cudaStream_t s) {
  ++g_launches;
  k_pack_image<<<(n_px + 255) / 256, 256, 0, s>>>(src3, dst, n_px, W);
}

void launch_photometric(const uint8_t* gray, const float* lut, const uint8_t* vignetting, float max_v, float* out, int n,
                        cudaStream_t s) {
  ++g_launches;
  k_photometric<<<(n + 255) / 256, 256, 0, s>>>(gray, lut, vignetting, max_v, out, n);
}
void launch_downscale(const float* src, float* dst, int W, int H, cudaStream_t s) {
  dim3 b(32, 8), g((W / 2 + 31) / 32, (H / 2 + 7) / 8);
  ++g_launches;
  k_downscale<<<g, b, 0, s>>>(src, dst, W, H);
}
void launch_pixelinfo3(const float* I, float* dst, int W, int H, cudaStream_t s) {
  dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
  ++g_launches;
  k_pixelinfo3<<<g, b, 0, s>>>(I, dst, W, H);
}

// Measured on B200 (dpba_debug_pixelinfo_ab, profiles/r02_ab.md): the TMA-staged variant (image_tma.cu) produces bit-identical
// records in the same time -- 6.16 against 6.16 us per launch at 640x480, 4.2 against 3.9 us at 160x120, 20.5 against 21.4 us at
// 1920x1080.  Every input pixel is touched by at most five neighbouring threads of one CTA, which L1 serves, and the kernel
// is bound by its 32-byte-per-pixel output stream (3.2 TB/s at 1080p) and, below VGA, by the launch itself.  The direct
// kernel stays the default; option "pixelinfo_tma" selects the TMA variant.
static bool g_pixelinfo_tma = false;
void set_pixelinfo_tma(bool on) { g_pixelinfo_tma = on; }
bool get_pixelinfo_tma() { return g_pixelinfo_tma; }
void launch_pixelinfo(const float* I, float4* dst, int W, int H, cudaStream_t s) {
  if (g_pixelinfo_tma && launch_pixelinfo_tma(I, dst, W, H, s)) return;
  dim3 b(32, 8), g((W + 31) / 32, (H + 7) / 8);
  ++g_launches;
  k_pixelinfo<<<g, b, 0, s>>>(I, dst, W, H);
}

int launch_residual_sweep(const WindowDev& w, float sigma, int huber, int fej, double* part, cudaStream_t s,
                          const LmCtl* ctl, int ctl_mode) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return 0;
  const int N = w.n_frames;
  // landmarks per CTA: fill the machine once (threads of N - 1 warps, up to 2048 resident per SM)
  const int threads = 32 * (N - 1);
  const int resident = sm_count() * std::max(1, 2048 / threads);
  int lpb = 8;
  while (lpb < 256 && (long)((m + lpb - 1) / lpb) * N > resident) lpb += 4;
  dim3 g((m + lpb - 1) / lpb, N);
  const size_t smem = (size_t)(N - 1) * sizeof(PairConst) + (size_t)lpb * sizeof(LandmarkRec) + (size_t)(N - 1) * 8;
  ++g_launches;
  if (fej) k_residual_sweep<true><<<g, threads, smem, s>>>(w, sigma, huber, lpb, reinterpret_cast<double2*>(part), ctl, ctl_mode);
  else k_residual_sweep<false><<<g, threads, smem, s>>>(w, sigma, huber, lpb, reinterpret_cast<double2*>(part), ctl, ctl_mode);
  return (int)(g.x * g.y);  // partial slots written
}

void launch_reduce_scal(const LmCtl* ctl, int ctl_mode, const double* e_part, int n_e, const double* n_part, int n_n,
                        double* scal, cudaStream_t s, int core_frames, int peer_push) {
  ++g_launches;
  k_reduce_scal<<<1, 1024, 0, s>>>(ctl, ctl_mode, reinterpret_cast<const double2*>(e_part), n_e,
                                   reinterpret_cast<const double2*>(n_part), n_n, scal, core_frames, peer_push);
}

void set_peer_context(const PeerDev& pd) { cudaMemcpyToSymbol(g_peer, &pd, sizeof(PeerDev)); }

int system_producer_ctas(int n_frames) {
  const int D = 8 * n_frames, T4 = D / 4, nout = T4 * (T4 + 1) / 2 * 16 + D;
  return n_frames * (n_frames + 1) / 2 + (nout + 31) / 32;
}

void launch_materialise_sweep(const WindowDev& w, float sigma, int huber, int fej, cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  // groups of 32 landmarks per CTA: 1 for small windows (more CTAs than one wave needs), up to 8 for large ones
  const int subs = std::max(1, std::min(8, m / 2048));
  dim3 g((m + 32 * subs - 1) / (32 * subs), w.n_frames * (w.n_frames - 1));
  ++g_launches;
  // TARGET-major: one target image (9.8 MB) is the L2 working set while ~12 MB of Jacobians per pair stream out.
  // Measured at 8 x 20000 landmarks (DRAM bytes read per launch): reference-major 502 MB, tiles of 4 targets 296 MB,
  // target-major (tile = 1) 150 MB -- little more than one image survives next to 600 MB of streaming stores.
  constexpr int TILE = 1;
  PairOrder order;
  int k = 0;
  const int N = w.n_frames;
  for (int tile = 0; tile < N; tile += TILE)
    for (int r = 0; r < N; ++r)
      for (int t = tile; t < std::min(tile + TILE, N); ++t)
        if (r != t) {
          order.r[k] = (uint8_t)r;
          order.t[k] = (uint8_t)t;
          ++k;
        }
  if (fej) k_materialise_sweep<true><<<g, 256, 0, s>>>(w, order, sigma, huber, subs);
  else k_materialise_sweep<false><<<g, 256, 0, s>>>(w, order, sigma, huber, subs);
}

int g_fused_version = 2;  // 2: one thread per patch-residual (k_linearize_fused2); 1: 8 lanes per patch-residual
void set_fused_version(int v) { g_fused_version = v == 1 ? 1 : 2; }
void set_pdl(bool on) { g_pdl = on; }
int g_fused_minb = 3;  // resident CTAs per SM the fused linearise is compiled for (N <= 9): 4 -> 64 registers, 3 -> 80
void set_fused_min_blocks(int b) { g_fused_minb = b <= 3 ? 3 : 4; }

int g_fused2_minb = 2;  // A/B (option "fused2_min_blocks"): 3 = the second-generation sweep compiled for three CTAs per SM (<= 96 registers)
void set_fused2_min_blocks(int b) { g_fused2_minb = b == 3 ? 3 : 2; }
int g_fused_lpb_max = 256;  // largest landmark chunk per CTA the launcher may pick (option "fused_lpb_max")
void set_fused_lpb_max(int v) { g_fused_lpb_max = v < 32 ? 32 : (v > 256 ? 256 : (v / 32) * 32); }
int g_fused_reduce_once = 0;  // A/B (option "fused_reduce_once"): the per-lane sums of the landmark groups meet in shared memory and are reduced once per CTA -- measured slower (32.6 against 30.6 us; 268 against 213 us at 1.12 M units): 41 KB more shared memory per CTA comes out of L1
void set_fused_reduce_once(int v) { g_fused_reduce_once = v != 0; }
int g_fused_epilogue = 1;  // second-generation epilogue of k_linearize_fused2 (option "fused_epilogue", 0 = the first generation's)
void set_fused_epilogue(int v) { g_fused_epilogue = v != 0; }
bool g_fused_prefetch = false;  // L1 prefetch of the next group's image taps: measured 42.6 us against 39.1 us without (issue-bound kernel), kept as an A/B switch (option "fused_prefetch")
void set_fused_prefetch(bool on) { g_fused_prefetch = on; }

template <int NWMAX, int MINB, bool PF>
static void launch_fused_tp(const WindowDev& w, float sigma, int huber, int fej, int for_marg, int lpb, dim3 g,
                           int threads, size_t smem, float* core, float* schur, cudaStream_t s, const LmCtl* ctl,
                           int ctl_mode) {
  static bool attr[2] = {false, false};
  if (smem > 48 * 1024 && !attr[fej ? 1 : 0]) {
    if (fej) cudaFuncSetAttribute(k_linearize_fused<true, NWMAX, MINB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    else cudaFuncSetAttribute(k_linearize_fused<false, NWMAX, MINB, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr[fej ? 1 : 0] = true;
  }
  ++g_launches;
  if (fej) k_linearize_fused<true, NWMAX, MINB, PF><<<g, threads, smem, s>>>(w, sigma, huber, for_marg, lpb, core, schur, ctl, ctl_mode);
  else k_linearize_fused<false, NWMAX, MINB, PF><<<g, threads, smem, s>>>(w, sigma, huber, for_marg, lpb, core, schur, ctl, ctl_mode);
}

template <int NWMAX, int MINB>
static void launch_fused_t(const WindowDev& w, float sigma, int huber, int fej, int for_marg, int lpb, dim3 g,
                           int threads, size_t smem, float* core, float* schur, cudaStream_t s, const LmCtl* ctl,
                           int ctl_mode) {
  if (g_fused_prefetch) launch_fused_tp<NWMAX, MINB, true>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, core, schur, s, ctl, ctl_mode);
  else launch_fused_tp<NWMAX, MINB, false>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, core, schur, s, ctl, ctl_mode);
}

template <int NWMAX, int MINB>
static void launch_fused2_t(const WindowDev& w, float sigma, int huber, int fej, int for_marg, int lpb, dim3 g, int threads,
                            size_t smem, float* core_part, float* schur_part, cudaStream_t s, const LmCtl* ctl, int ctl_mode,
                            int fold, const double* step_pose, double2* norms, int epi) {
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(k_linearize_fused2<true, NWMAX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_linearize_fused2<false, NWMAX, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr = true;
  }
  ++g_launches;
  if (fej) launch_pdl(k_linearize_fused2<true, NWMAX, MINB>, g, dim3(threads), smem, s, w, sigma, huber, for_marg, lpb, core_part, schur_part, ctl, ctl_mode, fold, step_pose, norms, epi);
  else launch_pdl(k_linearize_fused2<false, NWMAX, MINB>, g, dim3(threads), smem, s, w, sigma, huber, for_marg, lpb, core_part, schur_part, ctl, ctl_mode, fold, step_pose, norms, epi);
}

// second generation: `lpb` is a multiple of 32 (a lane owns a landmark); 2 CTAs per SM are resident (<= 144 registers)
static FusedShape launch_linearize_fused2(const WindowDev& w, float sigma, int huber, int fej, int for_marg, ReduceBuf rb,
                                          cudaStream_t s, const LmCtl* ctl, int ctl_mode, const double* fold_step) {
  FusedShape shape{0, 0};
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return shape;
  const int N = w.n_frames, D = 8 * N;
  const long resident = (long)sm_count() * ((N <= 8 && g_fused2_minb == 3) ? 3 : 2);
  int lpb = 32;
  double best = 1e300;
  for (int cand = 32; cand <= g_fused_lpb_max; cand += 32) {
    // two CTAs must stay resident per SM: 2 x (dynamic + 1 KB static + 1 KB reserved) within the 227 KB an SM can carve out,
    // with room to spare for the carve-out steps (measured: 106 KB per CTA -- lpb 192 with the second-generation epilogue --
    // ran one CTA per SM and the 1.12 M-unit sweep took 1.6x as long)
    const size_t need = (size_t)(N - 1) * sizeof(PairConst) + ((size_t)cand * (D + 2 + 2 * (N - 1)) + 3) * sizeof(float) +
                        (g_fused_epilogue ? (size_t)cand * (8 * (N - 1) + 1) * sizeof(float) : 0) +
                        ((g_fused_epilogue && g_fused_reduce_once) ? (size_t)(N - 1) * 46 * 32 * sizeof(float) : 0);
    if (cand > 32 && need > 96 * 1024) break;
    const long blocks = (long)((m + cand - 1) / cand) * N;
    const double waves = (double)((blocks + resident - 1) / resident);
    const double cost = waves * (cand / 32 + 0.5);  // per CTA: cand / 32 landmark groups + prologue / epilogue
    if (cost < best - 1e-9) {
      best = cost;
      lpb = cand;
    }
  }
  const int epi = g_fused_epilogue ? (1 | (g_fused_reduce_once ? 2 : 0)) : 0;
  const size_t smem = (size_t)(N - 1) * sizeof(PairConst) + ((size_t)lpb * (D + 2 + 2 * (N - 1)) + 3) * sizeof(float) +
                      (epi ? (size_t)lpb * (8 * (N - 1) + 1) * sizeof(float) : 0) +
                      ((epi & 2) ? (size_t)(N - 1) * 46 * 32 * sizeof(float) : 0);
  dim3 g((m + lpb - 1) / lpb, N);
  const int threads = 32 * (N - 1);
  const int fold = fold_step != nullptr && ctl != nullptr;
  double2* norms = reinterpret_cast<double2*>(rb.n_part);
  if (N <= 8 && g_fused2_minb == 3) launch_fused2_t<7, 3>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode, fold, fold_step, norms, epi);
  else if (N <= 8) launch_fused2_t<7, 2>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode, fold, fold_step, norms, epi);
  else if (N <= 9) launch_fused2_t<8, 2>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode, fold, fold_step, norms, epi);
  else launch_fused2_t<15, 1>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode, fold, fold_step, norms, epi);
  shape.lpb = lpb;
  shape.chunks = (int)g.x;
  return shape;
}

int fused_version() { return g_fused_version; }
void debug_cta_times(long long* out, int n) { cudaMemcpyFromSymbol(out, g_cta_times, sizeof(long long) * (size_t)(n < 4096 ? n : 4096)); }
void stamps_off_async(cudaStream_t s) {  // a memset node when captured: the stamps keep what they hold from here on
  void* p = nullptr;
  cudaGetSymbolAddress(&p, g_stamps_on);
  cudaMemsetAsync(p, 0, sizeof(int), s);
}
void debug_kernel_times(long long out[32]) { cudaMemcpyFromSymbol(out, g_kst, sizeof(long long) * 32); }
void debug_stamps(int enable, long long out[64]) {
  cudaMemcpyToSymbol(g_stamps_on, &enable, sizeof(int));
  if (out) cudaMemcpyFromSymbol(out, g_stamps, 64 * sizeof(long long));
}

FusedShape launch_linearize_fused(const WindowDev& w, float sigma, int huber, int fej, int for_marg, ReduceBuf rb,
                                  cudaStream_t s, const LmCtl* ctl, int ctl_mode, const double* fold_step) {
  if (g_fused_version == 2) return launch_linearize_fused2(w, sigma, huber, fej, for_marg, rb, s, ctl, ctl_mode, fold_step);
  FusedShape shape{0, 0};
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return shape;
  const int N = w.n_frames, D = 8 * N;
  // landmarks per CTA: the candidate whose wave count times per-CTA work (lpb/4 loop trips + the fixed prologue /
  // epilogue, worth ~1.5 trips) is smallest -- small windows want many small CTAs, large ones amortise the epilogue
  const int minb = N <= 9 ? g_fused_minb : 2;
  const int resident = sm_count() * minb;
  int lpb = 16;
  double best = 1e30;
  for (int cand = 8; cand <= 64; cand += 4) {
    const long blocks = (long)((m + cand - 1) / cand) * N;
    const double waves = (double)((blocks + resident - 1) / resident);
    const double cost = waves * (cand / 4 + 1.5);
    if (cost < best - 1e-9) {
      best = cost;
      lpb = cand;
    }
  }
  const size_t smem = (size_t)(N - 1) * sizeof(PairConst) + ((size_t)lpb * (D + 2 + 2 * (N - 1)) + 3) * sizeof(float) +
                      (size_t)lpb * sizeof(LandmarkRec);
  dim3 g((m + lpb - 1) / lpb, N);
  const int threads = 32 * (N - 1);
  if (N <= 9 && minb == 4) launch_fused_t<8, 4>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode);
  else if (N <= 9) launch_fused_t<8, 3>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode);
  else launch_fused_t<15, 2>(w, sigma, huber, fej, for_marg, lpb, g, threads, smem, rb.core_part, rb.fschur_part, s, ctl, ctl_mode);
  shape.lpb = lpb;
  shape.chunks = (int)g.x;
  return shape;
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

int launch_schur(const WindowDev& w, int for_marg, ReduceBuf rb, cudaStream_t s, const LmCtl* ctl) {
  const int N = w.n_frames, D = 8 * N;
  int tiles = 0;
  for (int f = 0; f < N; ++f) tiles += (w.n_lm[f] + SCHUR_TL - 1) / SCHUR_TL;
  if (tiles == 0) return 0;
  const int T4 = D / 4;
  const int ntri = T4 * (T4 + 1) / 2;
  const int gthreads = ((std::max(ntri, D) + 31) / 32) * 32;
  const int G = std::max(1, std::min(4, 1024 / gthreads));
  const size_t smem = (size_t)(2 * SCHUR_TL * D + 4 * SCHUR_TL) * sizeof(float) + (size_t)(ntri * 16 + D) * sizeof(double);
  const int grid = std::min(sm_count(), tiles);
  static bool attr = false;
  if (smem > 48 * 1024 && !attr) {
    cudaFuncSetAttribute(k_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    attr = true;
  }
  ++g_launches;
  if (g_schur_mma) {
    const int DP = (D + 31) & ~31;
    const size_t smem2 = (size_t)(2 * SCHUR_TL * (DP + 8) + 4 * SCHUR_TL) * sizeof(float);
    k_schur_mma<<<grid, 256, smem2, s>>>(w, for_marg, rb.schur_part, rb.bs_part, ctl);
  } else {
    k_schur<<<grid, G * gthreads, smem, s>>>(w, for_marg, G, gthreads, rb.schur_part, rb.bs_part, ctl);
  }
  return grid;  // CTAs that wrote a partial
}

void launch_core_reduce(const WindowDev& w, ReduceBuf rb, FusedShape shape, cudaStream_t s, const LmCtl* ctl) {
  const int N = w.n_frames;
  if (N < 2 || shape.lpb == 0) return;
  ++g_launches;
  k_core_reduce<<<N * (N - 1), 256, 0, s>>>(w, rb.core_part, shape.lpb, shape.chunks, rb.core, ctl);
}

void launch_assemble_blocks(const WindowDev& w, int fej, ReduceBuf rb, FusedShape shape, cudaStream_t s,
                            const LmCtl* ctl, int peer_push) {
  const int N = w.n_frames;
  if (N < 2 || shape.lpb == 0) return;
  ++g_launches;
  const int threads = 32 * std::max(2, N - 1);
  k_assemble<<<N * (N + 1) / 2, threads, (size_t)(threads / 32) * ASM_W * sizeof(double), s>>>(w, fej, rb.core, rb.Hp, rb.bp, ctl,
                                                                                                peer_push);
}

void launch_assemble(const WindowDev& w, int fej, ReduceBuf rb, FusedShape shape, cudaStream_t s, const LmCtl* ctl) {
  launch_core_reduce(w, rb, shape, s, ctl);
  launch_assemble_blocks(w, fej, rb, shape, s, ctl);
}

// fused path: sums the per-chunk Schur partials written by k_linearize_fused and mirrors H_s (k_assemble already
// wrote a symmetric Hp)
void launch_finish_fused(const WindowDev& w, ReduceBuf rb, FusedShape shape, cudaStream_t s, const LmCtl* ctl, int peer_push) {
  const int N = w.n_frames, D = 8 * N;
  if (N < 2 || shape.lpb == 0) return;
  const int T4 = D / 4, nout = T4 * (T4 + 1) / 2 * 16 + D;
  ++g_launches;
  k_finish_fused<<<(nout + 31) / 32, 1024, 0, s>>>(w, shape.lpb, shape.chunks, rb.fschur_part, rb.Hs, rb.bs, ctl, peer_push);
}

// sums the Schur partials of `nsb` CTAs into Hs / bs (skipped when nsb == 0 and Hs == nullptr) and symmetrises Hp
void launch_finish_system(int D, ReduceBuf rb, int nsb, cudaStream_t s, const LmCtl* ctl) {
  ++g_launches;
  k_finish_system<<<(D * D + 255) / 256, 256, 0, s>>>(D, rb.Hp, rb.Hs, rb.bs, rb.schur_part, rb.bs_part, nsb, ctl);
}

void launch_back_substitute(const WindowDev& w, const double* step_pose_dev, double lambda, cudaStream_t s,
                            const LmCtl* ctl, double* norms) {
  const int m = max_landmarks(w);
  if (m == 0) return;
  dim3 g((m + 31) / 32, w.n_frames);
  ++g_launches;
  k_back_substitute<<<g, 256, 0, s>>>(w, step_pose_dev, (float)(1.0 / (1.0 + lambda)), ctl, norms);
}

void launch_accept(const WindowDev& w, int accept, double* scal, cudaStream_t s, const LmCtl* ctl, int with_statuses) {
  const int m = max_landmarks(w);
  if (m == 0) return;
  dim3 g((m + 255) / 256, w.n_frames);
  ++g_launches;
  k_accept_landmarks<<<g, 256, 0, s>>>(w, accept, scal, ctl, with_statuses);
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

void launch_refine_immature(const WindowDev& w, int r, int n, const float* proj, const float* idepth_in, const float* patch8,
                            int min_inliers, float sigma, float* idepth_out, uint8_t* activate, int* n_valid_out,
                            cudaStream_t s) {
  if (n <= 0) return;
  ++g_launches;
  k_refine_immature<<<(n + 7) / 8, 256, 0, s>>>(w, r, n, reinterpret_cast<const float2*>(proj), idepth_in, patch8, min_inliers,
                                              sigma, idepth_out, activate, n_valid_out);
}

void launch_first_estimate(const WindowDev& w, cudaStream_t s) {
  const int m = max_landmarks(w);
  if (m == 0 || w.n_frames < 2) return;
  dim3 g((m + 31) / 32, w.n_frames * (w.n_frames - 1));
  ++g_launches;
  k_first_estimate<<<g, 256, 0, s>>>(w);
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
