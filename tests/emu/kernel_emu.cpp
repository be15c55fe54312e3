// CPU emulation of kernels whose per-thread bodies are __host__ __device__ functions (test infrastructure only).
// Every function below walks the same grid the CUDA launcher uses and calls the SAME body once per CUDA thread, so the
// kernel logic (indexing, predicates, rounding, layout) is checked against the oracles without a GPU.  What this cannot
// show: anything that depends on real concurrency (atomics order, warp intrinsics) -- that stays with the -m gpu tests.
//
// Built by tests/test_kernel_emulation.py:  g++ -O1 -ffp-contract=off -I/usr/local/cuda/include -I dsopp_b200/csrc
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "depth_maps_body.h"
#include "energy_quantile_body.h"
#include "optical_flow_body.h"

using namespace pba;

namespace {

// WindowDev over plain host arrays; logical slot f is stored at physical slot phys[f] (given by the caller)
struct HostWindow {
  WindowDev w;
  std::vector<PairConst> pairs;
};

void make_window(HostWindow& hw, int n_frames, int W, int H, int max_pts, const int* n_lm, const int* phys,
                 const int* frame_marg, const float* lmk4, const uint8_t* flags, const uint8_t* status,
                 const float* energy, const float* inv_hdd, const float* A, const float* M) {
  memset(&hw.w, 0, sizeof(hw.w));
  hw.w.n_frames = n_frames;
  hw.w.W = W;
  hw.w.H = H;
  hw.w.max_pts = max_pts;
  for (int f = 0; f < n_frames; ++f) {
    hw.w.n_lm[f] = n_lm[f];
    hw.w.phys[f] = phys[f];
    hw.w.frame_marg[f] = frame_marg ? frame_marg[f] : 0;
  }
  hw.w.lmk = reinterpret_cast<float4*>(const_cast<float*>(lmk4));
  hw.w.flags = const_cast<uint8_t*>(flags);
  hw.w.status = const_cast<uint8_t*>(status);
  hw.w.energy = const_cast<float*>(energy);
  hw.w.inv_hdd = const_cast<float*>(inv_hdd);
  hw.pairs.assign((size_t)PBA_MAXF * PBA_MAXF, PairConst{});
  if (A && M)
    for (int r = 0; r < n_frames; ++r)
      for (int t = 0; t < n_frames; ++t) {
        PairConst& pc = hw.pairs[(size_t)r * PBA_MAXF + t];
        memcpy(pc.A, A + ((size_t)r * n_frames + t) * 12, 12 * sizeof(float));
        memcpy(pc.M, M + ((size_t)r * n_frames + t) * 12, 12 * sizeof(float));
      }
  hw.w.pairs = hw.pairs.data();
}

}  // namespace

extern "C" {

// launch_reference_depth_maps (depth_maps.cu) thread by thread; buf as there: per level [idw_raw | wgt_raw | idw | wgt]
size_t emu_depth_maps_floats(int W, int H, int n_levels) { return dm_level_offset_of(W, H, n_levels); }

void emu_reference_depth_maps(int n_frames, int W, int H, int max_pts, const int* n_lm, const int* phys,
                              const float* lmk4, const uint8_t* flags, const uint8_t* status, const float* inv_hdd,
                              const float* A, const float* M, int n_levels, float const_var, float* buf) {
  HostWindow hw;
  make_window(hw, n_frames, W, H, max_pts, n_lm, phys, nullptr, lmk4, flags, status, nullptr, inv_hdd, A, M);
  const WindowDev& w = hw.w;
  const size_t n0 = (size_t)W * H;
  memset(buf, 0, 2 * n0 * sizeof(float));
  int m = 0;
  for (int f = 0; f + 1 < n_frames; ++f) m = n_lm[f] > m ? n_lm[f] : m;
  const int bx = (m + 255) / 256;
  for (int r = 0; r + 1 < n_frames; ++r)        // blockIdx.y
    for (int b = 0; b < bx; ++b)                // blockIdx.x
      for (int t = 0; t < 256; ++t)             // threadIdx.x
        dm_splat_thread(w, const_var, buf, buf + n0, r, b * 256 + t);
  for (int l = 0; l < n_levels; ++l) {
    const int Wl = W >> l, Hl = H >> l;
    const size_t nl = (size_t)Wl * Hl;
    float* base = buf + dm_level_offset_of(W, H, l);
    const int gx = (Wl + 255) / 256;
    if (l > 0) {
      const float* up = buf + dm_level_offset_of(W, H, l - 1);
      const size_t nu = (size_t)(W >> (l - 1)) * (H >> (l - 1));
      for (int y = 0; y < Hl; ++y)
        for (int x = 0; x < gx * 256; ++x) dm_coarse_pixel(up, up + nu, W >> (l - 1), base, base + nl, Wl, Hl, x, y);
    }
    for (int y = 0; y < Hl; ++y)
      for (int x = 0; x < gx * 256; ++x)
        dm_dilate_pixel(base, base + nl, base + 2 * nl, base + 3 * nl, Wl, Hl, l > 1 ? 1 : 0, x, y);
  }
}

// launch_energy_quantile (energy_quantile.cu): four histogram + pick passes; returns count, *value = k-th smallest
unsigned emu_energy_quantile(int n_frames, int max_pts, const int* n_lm, const int* phys, const int* frame_marg,
                             const uint8_t* flags, const uint8_t* status, const float* energy, double frac,
                             float* value) {
  HostWindow hw;
  make_window(hw, n_frames, 0, 0, max_pts, n_lm, phys, frame_marg, nullptr, flags, status, energy, nullptr, nullptr,
              nullptr);
  const WindowDev& w = hw.w;
  int nmax = 0;
  for (int f = 0; f < n_frames; ++f) nmax = n_lm[f] > nmax ? n_lm[f] : nmax;
  SelectState st;
  select_init(&st);
  if (nmax > 0) {
    const long long total = (long long)n_frames * n_frames * nmax;
    for (int shift = 24; shift >= 0; shift -= 8) {
      for (long long idx = 0; idx < total; ++idx) {  // every CUDA thread of k_select_hist, grid-stride flattened
        unsigned key = 0;
        if (residual_key(w, idx, nmax, key) && (key & st.mask) == st.prefix) ++st.hist[(key >> shift) & 255u];
      }
      unsigned h[256];
      memcpy(h, st.hist, sizeof(h));
      memset(st.hist, 0, sizeof(st.hist));
      select_pick(&st, h, shift, frac);
    }
  }
  *value = st.value;
  return st.count;
}

// k_optical_flow (pose_alignment.cu): the per-landmark term over all landmarks, fp64 sums; returns the count
int emu_optical_flow(int n, const float* lm4, const double* T_target_reference, const double* intr, int W, int H,
                     double* flow) {
  const FlowConst c = make_flow_const(T_target_reference, intr, W, H);
  double sum = 0.0, cnt = 0.0;
  for (int i = 0; i < n; ++i) {
    float sq;
    if (flow_term(c, reinterpret_cast<const float4*>(lm4)[i], sq)) {
      sum += (double)sq;
      cnt += 1.0;
    }
  }
  *flow = std::sqrt(sum / cnt);
  return (int)cnt;
}

}  // extern "C"
