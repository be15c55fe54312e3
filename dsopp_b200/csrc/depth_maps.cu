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
// of dpa_set_reference_depth_map (csrc/pose_alignment.cu).  The per-thread bodies live in depth_maps_body.h so that
// tests/emu can run them on the CPU.  fp32 with float atomics for the rare collisions of two
// landmarks on one pixel (order of two or three additions: last-bit differences, stated in the test).
//
// Oracle: oracle/depth_map_oracle.py, pinned against the reference's own create_depth_maps.cpp (tests/
// test_reference_tracker.py).  GPU comparison: tests/test_gpu_device_paths.py (oracle) and tests/test_reference_tracker.py
// (the reference's golden maps); the per-thread bodies also run on the CPU in tests/test_kernel_emulation.py.
#include <cuda_runtime.h>

#include <cstdint>

#include "depth_maps_body.h"
#include "pba_internal.h"

namespace pba {

namespace {

__global__ void __launch_bounds__(256) k_depth_splat(const __grid_constant__ WindowDev w, float const_var,
                                                     float* __restrict__ idw, float* __restrict__ wgt) {
  dm_splat_thread(w, const_var, idw, wgt, blockIdx.y, blockIdx.x * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(256) k_depth_coarse(const float* __restrict__ idw_up, const float* __restrict__ wgt_up,
                                                      int W_up, float* __restrict__ idw, float* __restrict__ wgt, int W,
                                                      int H) {
  dm_coarse_pixel(idw_up, wgt_up, W_up, idw, wgt, W, H, blockIdx.x * blockDim.x + threadIdx.x, blockIdx.y);
}

__global__ void __launch_bounds__(256) k_depth_dilate(const float* __restrict__ idw_in, const float* __restrict__ wgt_in,
                                                      float* __restrict__ idw_out, float* __restrict__ wgt_out, int W,
                                                      int H, int axis_neighbours) {
  dm_dilate_pixel(idw_in, wgt_in, idw_out, wgt_out, W, H, axis_neighbours, blockIdx.x * blockDim.x + threadIdx.x,
                  blockIdx.y);
}

}  // namespace

// buf: 4 * sum_l (W>>l)(H>>l) floats: per level [idw_raw | wgt_raw | idw | wgt]; the dilated maps of level l start at
// dm_level_offset(...) + 2 * n_l.  The caller has the per-pair constants at the accepted state (k_pair_setup).
size_t dm_level_offset(int W, int H, int level) { return dm_level_offset_of(W, H, level); }

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
