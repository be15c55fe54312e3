// Reference depth maps of the coarse tracker, built on the device from the window resident in the BA handle.
//
// Reference: createReferenceDepthMaps, src/tracker/tracker/src/create_depth_maps.cpp:122-146, called by the tracker right
// after every BA solve (src/tracker/tracker/src/monocular_tracker.cpp:465,509):
//   fillFineDepthMap    :19-58   every kOk, non-outlier, non-marginalised landmark of the OLDER keyframes is reprojected
//                                into the newest keyframe (scalar pinhole reproject, camera_reproject.hpp:270-293), and
//                                {idepth / depth_scale * w, w}, w = sqrt(1e-3 / (variance + 1e-12)), is added at the
//                                rounded pixel; depth_scale = z of the transformed point (camera_model_base.hpp:103-107)
//   fillCoarseDepthMaps :70-88   level l = 2x2 sums of level l-1 (both accumulators)
//   dilateDepthMaps     :90-120  empty interior pixels take the mean of their non-empty diagonal (levels 0-1) or axis
//                                (levels >= 2) neighbours
// Everything the function reads is already on the device after a solve (landmarks, statuses towards the newest
// keyframe, per-pair reprojection constants, inv_hessian_idepth_idepth), so nothing is uploaded; the maps are the input
// of dpa_set_reference_depth_map (csrc/pose_alignment.cu).  fp32 with float atomics for the rare collisions of two
// landmarks on one pixel (order of two or three additions: last-bit differences, stated in the test).
//
// STATUS: written in round 1 after the GPU budget was spent -- compiled for sm_100a, NOT yet run on hardware; new entry
// point, nothing else calls it.  Oracle: oracle/depth_map_oracle.py; GPU comparison: tests/test_zz_gpu_experimental.py.
#include <cuda_runtime.h>

#include <cstdint>

#include "pba_internal.h"

namespace pba {

namespace {

constexpr int DM_K_OK = 0;
constexpr int DM_LM_MARG = 1, DM_LM_OUTLIER = 4;

__device__ __forceinline__ float dm_row(const float* a, float u, float v, float rho) {
  // same rounding as the sweeps' reprojection (explicitly rounded, never contracted): the ROI predicate below decides
  // which pixel a landmark lands on
  return __fadd_rn(__fadd_rn(__fmul_rn(a[0], u), __fmul_rn(a[1], v)), __fadd_rn(a[2], __fmul_rn(a[3], rho)));
}

__global__ void __launch_bounds__(256) k_depth_splat(const __grid_constant__ WindowDev w, float const_var,
                                                     float* __restrict__ idw, float* __restrict__ wgt) {
  const int t = w.n_frames - 1;
  const int r = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= w.n_lm[r]) return;
  const size_t gl = (size_t)w.phys[r] * w.max_pts + l;
  if (w.flags[gl] & (DM_LM_MARG | DM_LM_OUTLIER)) return;                                   // :40
  const size_t res = ((size_t)(w.phys[r] * PBA_MAXF + w.phys[t])) * w.max_pts + l;
  if (w.status[res] != DM_K_OK) return;                                                     // :38
  const float4 lm = w.lmk[gl];
  const float u = lm.x, v = lm.y, rho = lm.z;
  const float xmax = (float)(w.W - 5), ymax = (float)(w.H - 5);
  if (!(rho > -1e-4f && rho < 1010.f)) return;                                              // validIdepth
  if (!(u >= 4.f && v >= 4.f && u <= xmax && v <= ymax)) return;                            // insideCameraROI(reference)
  const PairConst& pc = w.pairs[r * PBA_MAXF + t];
  const float X = dm_row(pc.A + 0, u, v, rho), Y = dm_row(pc.A + 4, u, v, rho), Z = dm_row(pc.A + 8, u, v, rho);
  if (!(Z > 0.f)) return;
  const float rz = __frcp_rn(Z);
  const float tu = __fmul_rn(X, rz), tv = __fmul_rn(Y, rz);
  if (!(tu >= 4.f && tv >= 4.f && tu <= xmax && tv <= ymax)) return;                        // insideCameraROI(target)
  const int ix = (int)floorf(tu + 0.5f), iy = (int)floorf(tv + 0.5f);                       // round(), positive operands
  const float qz = dm_row(pc.M + 8, u, v, rho);                                             // getDepthScale
  const float var = const_var >= 0.f ? const_var : w.inv_hdd[gl];
  const float wt = sqrtf(1e-3f / (var + 1e-12f));                                           // :52
  atomicAdd(&idw[(size_t)iy * w.W + ix], rho / qz * wt);
  atomicAdd(&wgt[(size_t)iy * w.W + ix], wt);
}

__global__ void __launch_bounds__(256) k_depth_coarse(const float* __restrict__ idw_up, const float* __restrict__ wgt_up,
                                                      int W_up, float* __restrict__ idw, float* __restrict__ wgt, int W,
                                                      int H) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t a = (size_t)(2 * y) * W_up + 2 * x, b = a + W_up;
  // order of the reference's sum: (2x,2y) + (2x+1,2y) + (2x,2y+1) + (2x+1,2y+1)
  idw[(size_t)y * W + x] = ((idw_up[a] + idw_up[a + 1]) + idw_up[b]) + idw_up[b + 1];
  wgt[(size_t)y * W + x] = ((wgt_up[a] + wgt_up[a + 1]) + wgt_up[b]) + wgt_up[b + 1];
}

__global__ void __launch_bounds__(256) k_depth_dilate(const float* __restrict__ idw_in, const float* __restrict__ wgt_in,
                                                      float* __restrict__ idw_out, float* __restrict__ wgt_out, int W,
                                                      int H, int axis_neighbours) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t i = (size_t)y * W + x;
  float id = idw_in[i], wt = wgt_in[i];
  if (wt <= 0.f && x >= 1 && y >= 1 && x < W - 1 && y < H - 1) {
    // offsets in the reference's order (:103-107): axis (1,0) (-1,0) (0,1) (0,-1); diagonal (1,1) (-1,-1) (1,-1) (-1,1)
    const int dxa[4] = {1, -1, 0, 0}, dya[4] = {0, 0, 1, -1};
    const int dxd[4] = {1, -1, 1, -1}, dyd[4] = {1, -1, -1, 1};
    float sum = 0.f, num = 0.f, numn = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int dx = axis_neighbours ? dxa[k] : dxd[k], dy = axis_neighbours ? dya[k] : dyd[k];
      const size_t j = (size_t)(y + dy) * W + (x + dx);
      const float nw = wgt_in[j];
      if (nw > 0.f) {
        sum += idw_in[j];
        num += nw;
        numn += 1.f;
      }
    }
    if (numn > 0.f) {
      id = sum / numn;
      wt = num / numn;
    }
  }
  idw_out[i] = id;
  wgt_out[i] = wt;
}

}  // namespace

// buf: 4 * sum_l (W>>l)(H>>l) floats: per level [idw_raw | wgt_raw | idw | wgt]; the dilated maps of level l start at
// dm_level_offset(...) + 2 * n_l.  The caller has the per-pair constants at the accepted state (k_pair_setup).
size_t dm_level_offset(int W, int H, int level) {
  size_t off = 0;
  for (int l = 0; l < level; ++l) off += 4 * (size_t)(W >> l) * (size_t)(H >> l);
  return off;
}

void launch_reference_depth_maps(const WindowDev& w, int n_levels, float const_var, float* buf, cudaStream_t s) {
  const int W = w.W, H = w.H;
  float* l0 = buf;
  const size_t n0 = (size_t)W * H;
  cudaMemsetAsync(l0, 0, 2 * n0 * sizeof(float), s);
  int m = 0;
  for (int f = 0; f + 1 < w.n_frames; ++f) m = w.n_lm[f] > m ? w.n_lm[f] : m;
  int launches = 0;
  if (m > 0 && w.n_frames >= 2) {
    k_depth_splat<<<dim3((m + 255) / 256, w.n_frames - 1), 256, 0, s>>>(w, const_var, l0, l0 + n0);
    ++launches;
  }
  for (int l = 0; l < n_levels; ++l) {
    const int Wl = W >> l, Hl = H >> l;
    const size_t nl = (size_t)Wl * Hl;
    float* base = buf + dm_level_offset(W, H, l);
    if (l > 0) {
      const float* up = buf + dm_level_offset(W, H, l - 1);  // the un-dilated sums of the finer level
      const size_t nu = (size_t)(W >> (l - 1)) * (H >> (l - 1));
      k_depth_coarse<<<dim3((Wl + 255) / 256, Hl), 256, 0, s>>>(up, up + nu, W >> (l - 1), base, base + nl, Wl, Hl);
      ++launches;
    }
    k_depth_dilate<<<dim3((Wl + 255) / 256, Hl), 256, 0, s>>>(base, base + nl, base + 2 * nl, base + 3 * nl, Wl, Hl,
                                                              l > 1 ? 1 : 0);
    ++launches;
  }
  add_launches(launches);
}

}  // namespace pba
