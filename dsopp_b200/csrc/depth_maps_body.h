// Per-thread bodies of the reference-depth-map kernels (depth_maps.cu), written as __host__ __device__ functions so that
// tests/emu/ can run exactly this code on the CPU (one call per CUDA thread) against oracle/depth_map_oracle.py.
// On the device the accumulation is a float atomicAdd and the reprojection uses the explicitly rounded intrinsics of the
// sweeps; on the host the same operations are plain float arithmetic (tests/emu is compiled with -ffp-contract=off).
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>

#include "pba_internal.h"

namespace pba {

#ifdef __CUDA_ARCH__
#define DM_MUL(a, b) __fmul_rn((a), (b))
#define DM_ADD(a, b) __fadd_rn((a), (b))
#define DM_RCP(a) __frcp_rn(a)
#define DM_ACCUMULATE(p, v) atomicAdd((p), (v))
#else
#define DM_MUL(a, b) ((a) * (b))
#define DM_ADD(a, b) ((a) + (b))
#define DM_RCP(a) (1.0f / (a))
#define DM_ACCUMULATE(p, v) (*(p) += (v))
#endif

constexpr int DM_K_OK = 0;
constexpr int DM_LM_MARG = 1, DM_LM_OUTLIER = 4;

__host__ __device__ inline float dm_row(const float* a, float u, float v, float rho) {
  // same rounding as the sweeps' reprojection (explicitly rounded, never contracted): the ROI predicate below decides
  // which pixel a landmark lands on
  return DM_ADD(DM_ADD(DM_MUL(a[0], u), DM_MUL(a[1], v)), DM_ADD(a[2], DM_MUL(a[3], rho)));
}

// fillFineDepthMap (create_depth_maps.cpp:19-58) for landmark l of older keyframe r
__host__ __device__ inline void dm_splat_thread(const WindowDev& w, float const_var, float* idw, float* wgt, int r, int l) {
  const int t = w.n_frames - 1;
  if (l >= w.n_lm[r]) return;
  const size_t gl = (size_t)w.phys[r] * w.max_pts + l;
  if (w.flags[gl] & (DM_LM_MARG | DM_LM_OUTLIER)) return;                                   // :40
  const size_t res = ((size_t)(w.phys[r] * PBA_MAXF + w.phys[t])) * w.max_pts + l;
  if (w.status[res] != DM_K_OK) return;                                                     // :38
  const float4 lm = w.lmk[gl];
  const float u = lm.x, v = lm.y;
  float rho = lm.z;
  // the reference splats the TRACK's landmarks, i.e. the solver's after updateFrame's post-processing
  // (photometric_bundle_adjustment.cpp:232-238): |idepth| < 1e-8 becomes 0, any other negative idepth marks an outlier
  if (fabsf(rho) < 1e-8f) rho = 0.f;
  else if (rho < 0.f) return;
  const float xmax = (float)(w.W - 5), ymax = (float)(w.H - 5);
  if (!(rho > -1e-4f && rho < 1010.f)) return;                                              // validIdepth
  if (!(u >= 4.f && v >= 4.f && u <= xmax && v <= ymax)) return;                            // insideCameraROI(reference)
  const PairConst& pc = w.pairs[r * PBA_MAXF + t];
  const float X = dm_row(pc.A + 0, u, v, rho), Y = dm_row(pc.A + 4, u, v, rho), Z = dm_row(pc.A + 8, u, v, rho);
  if (!(Z > 0.f)) return;
  const float rz = DM_RCP(Z);
  const float tu = DM_MUL(X, rz), tv = DM_MUL(Y, rz);
  if (!(tu >= 4.f && tv >= 4.f && tu <= xmax && tv <= ymax)) return;                        // insideCameraROI(target)
  const int ix = (int)floorf(tu + 0.5f), iy = (int)floorf(tv + 0.5f);                       // round(), positive operands
  const float qz = dm_row(pc.M + 8, u, v, rho);                                             // getDepthScale
  const float var = const_var >= 0.f ? const_var : w.inv_hdd[gl];
  const float wt = sqrtf(1e-3f / (var + 1e-12f));                                           // :52
  DM_ACCUMULATE(&idw[(size_t)iy * w.W + ix], rho / qz * wt);
  DM_ACCUMULATE(&wgt[(size_t)iy * w.W + ix], wt);
}

// fillCoarseDepthMaps (:70-88) for pixel (x, y) of the coarser level
__host__ __device__ inline void dm_coarse_pixel(const float* idw_up, const float* wgt_up, int W_up, float* idw, float* wgt,
                                                int W, int H, int x, int y) {
  if (x >= W || y >= H) return;
  const size_t a = (size_t)(2 * y) * W_up + 2 * x, b = a + W_up;
  // order of the reference's sum: (2x,2y) + (2x+1,2y) + (2x,2y+1) + (2x+1,2y+1)
  idw[(size_t)y * W + x] = ((idw_up[a] + idw_up[a + 1]) + idw_up[b]) + idw_up[b + 1];
  wgt[(size_t)y * W + x] = ((wgt_up[a] + wgt_up[a + 1]) + wgt_up[b]) + wgt_up[b + 1];
}

// dilateDepthMaps (:90-120) for pixel (x, y), out of place
__host__ __device__ inline void dm_dilate_pixel(const float* idw_in, const float* wgt_in, float* idw_out, float* wgt_out,
                                                int W, int H, int axis_neighbours, int x, int y) {
  if (x >= W || y >= H) return;
  const size_t i = (size_t)y * W + x;
  float id = idw_in[i], wt = wgt_in[i];
  if (wt <= 0.f && x >= 1 && y >= 1 && x < W - 1 && y < H - 1) {
    // offsets in the reference's order (:103-107): axis (1,0) (-1,0) (0,1) (0,-1); diagonal (1,1) (-1,-1) (1,-1) (-1,1)
    const int dxa[4] = {1, -1, 0, 0}, dya[4] = {0, 0, 1, -1};
    const int dxd[4] = {1, -1, 1, -1}, dyd[4] = {1, -1, -1, 1};
    float sum = 0.f, num = 0.f, numn = 0.f;
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

// buf: 4 * sum_l (W>>l)(H>>l) floats: per level [idw_raw | wgt_raw | idw | wgt]
__host__ __device__ inline size_t dm_level_offset_of(int W, int H, int level) {
  size_t off = 0;
  for (int l = 0; l < level; ++l) off += 4 * (size_t)(W >> l) * (size_t)(H >> l);
  return off;
}

}  // namespace pba
