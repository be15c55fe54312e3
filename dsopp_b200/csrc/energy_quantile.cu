// Device-side 75 % energy quantile of updatePointStatuses (option "device_quantile", default ON since round 2).
//
// Reference: PhotometricBundleAdjustment::updatePointStatuses, first half
// (src/energy/problems/src/photometric_bundle_adjustment/photometric_bundle_adjustment.cpp:325-361): the energies of all
// kOk residuals of non-marginalised landmarks towards non-marginalised targets are collected into one vector and
// std::nth_element picks element k = size_t(n * 0.75); the outlier threshold is that energy + sigma^2 / 2.
// The host version (option device_quantile = 0) reads N(N-1) status / energy rows back and calls nth_element; this one
// never moves the rows: an EXACT radix select on the order-preserving integer image of the floats, most significant
// byte first -- per pass one histogram kernel over all residuals (per-CTA shared-memory histograms with
// warp-aggregated atomics, then at most 256 global atomics per CTA) and one tiny kernel that picks the byte and narrows
// (prefix, k).  Four passes give the k-th smallest bit pattern, i.e. the very float nth_element returns.  With landmarks
// sharded over several GPUs the 256-bin histogram of every pass is summed over the ranks before the pick, so every rank
// selects the same element of the union.
//
// Validated on B200 against the host path (tests/test_gpu_device_paths.py, tests/test_gpu_baseline_sizes.py: identical
// threshold bits, statuses, flags and inlier counts).  The enumeration and selection code lives in
// energy_quantile_body.h and also runs on the CPU in tests/test_kernel_emulation.py.
#include <cuda_runtime.h>

#include <cstdint>

#include "energy_quantile_body.h"
#include "pba_internal.h"

namespace pba {

namespace {

// one radix pass: histogram of byte (key >> shift) over the residuals whose higher bytes equal st->prefix
__global__ void __launch_bounds__(256) k_select_hist(const __grid_constant__ WindowDev w, SelectState* __restrict__ st,
                                                     int shift, int nmax) {
  __shared__ unsigned hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const unsigned prefix = st->prefix, mask = st->mask;
  const long long total = (long long)w.n_frames * w.n_frames * nmax;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  // uniform trip count per warp: the match below is a full-warp operation
  for (long long base = (long long)blockIdx.x * blockDim.x; base < total; base += stride) {
    const long long idx = base + threadIdx.x;
    unsigned key = 0;
    const bool in = idx < total && residual_key(w, idx, nmax, key) && (key & mask) == prefix;
    const unsigned bin = in ? ((key >> shift) & 255u) : 256u;
    const unsigned peers = __match_any_sync(0xffffffffu, bin);
    if (in && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
  }
  __syncthreads();
  const unsigned c = hist[threadIdx.x];
  if (c) atomicAdd(&st->hist[threadIdx.x], c);
}

// picks the byte that holds element k, narrows the search, clears the histogram for the next pass
__global__ void __launch_bounds__(256) k_select_pick(SelectState* __restrict__ st, int shift, double frac) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = st->hist[threadIdx.x];
  st->hist[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x == 0) select_pick(st, h, shift, frac);
}

__global__ void k_select_init(SelectState* st) {
  if (threadIdx.x == 0) select_init(st);
}

}  // namespace

// leaves {count, value = k-th smallest energy (k = size_t(count * frac))} in *st; `nmax` = max landmarks of a frame
void launch_energy_quantile(const WindowDev& w, int nmax, double frac, SelectState* st, cudaStream_t s,
                            int (*after_hist)(void*), void* after_hist_arg) {
  k_select_init<<<1, 32, 0, s>>>(st);
  if (nmax > 0) {
    const long long total = (long long)w.n_frames * w.n_frames * nmax;
    long long ctas = (total + 255) / 256;
    const long long cap = 2LL * sm_count();  // enough to hide the load latency, few enough to keep the flush cheap
    if (ctas > cap) ctas = cap;
    for (int shift = 24; shift >= 0; shift -= 8) {
      k_select_hist<<<(unsigned)ctas, 256, 0, s>>>(w, st, shift, nmax);
      if (after_hist && after_hist(after_hist_arg)) return;
      k_select_pick<<<1, 256, 0, s>>>(st, shift, frac);
    }
    add_launches(8);
  }
  add_launches(1);
}

}  // namespace pba
