// Thread-level pieces of the device-side energy quantile (energy_quantile.cu) as __host__ __device__ functions, so that
// tests/emu can run the same enumeration and selection code on the CPU against np.partition.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstring>

#include "pba_internal.h"

namespace pba {

constexpr int SEL_K_OK = 0;     // ResidualStatus kOk
constexpr int SEL_LM_MARG = 1;  // landmark is_marginalized

// order-preserving map float -> unsigned (negative floats reversed below the positive ones)
__host__ __device__ inline unsigned ordered_key(float x) {
  unsigned u;
#ifdef __CUDA_ARCH__
  u = __float_as_uint(x);
#else
  memcpy(&u, &x, 4);
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float key_to_float(unsigned k) {
  const unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  float x;
#ifdef __CUDA_ARCH__
  x = __uint_as_float(u);
#else
  memcpy(&x, &u, 4);
#endif
  return x;
}

// key of flattened residual `idx` (pair-major, `nmax` landmark slots per ordered pair); false when it does not take part
// (photometric_bundle_adjustment.cpp:325-356: kOk residuals of non-marginalised landmarks towards non-marginalised frames)
__host__ __device__ inline bool residual_key(const WindowDev& w, long long idx, int nmax, unsigned& key) {
  const int N = w.n_frames;
  const int p = (int)(idx / nmax), l = (int)(idx - (long long)p * nmax);
  const int r = p / N, t = p - r * N;
  if (p >= N * N || r == t || w.frame_marg[t] || l >= w.n_lm[r]) return false;
  if (w.flags[(size_t)w.phys[r] * w.max_pts + l] & SEL_LM_MARG) return false;
  const size_t res = ((size_t)(w.phys[r] * PBA_MAXF + w.phys[t])) * w.max_pts + l;
  if (w.status[res] != SEL_K_OK) return false;
  key = ordered_key(w.energy[res]);
  return true;
}

__host__ __device__ inline void select_init(SelectState* st) {
  for (int b = 0; b < 256; ++b) st->hist[b] = 0;
  st->prefix = 0;
  st->mask = 0;
  st->k = 0;
  st->count = 0;
  st->value = 0.f;
}

// one thread: picks the byte that holds element k and narrows (prefix, k); h = the histogram of this pass
__host__ __device__ inline void select_pick(SelectState* st, const unsigned* h, int shift, double frac) {
  unsigned long long k = st->k;
  if (shift == 24) {  // first pass: the histogram covers everything -> n and k = size_t(n * 0.75)  (:358)
    unsigned long long n = 0;
    for (int b = 0; b < 256; ++b) n += h[b];
    st->count = (unsigned)n;
    k = (unsigned long long)((double)n * frac);
    if (n == 0) {
      st->value = 0.f;
      return;
    }
  } else if (st->count == 0) {
    return;
  }
  unsigned long long below = 0;
  int b = 0;
  for (; b < 255; ++b) {
    if (below + h[b] > k) break;
    below += h[b];
  }
  st->k = k - below;
  st->prefix |= (unsigned)b << shift;
  st->mask |= 255u << shift;
  if (shift == 0) st->value = key_to_float(st->prefix);
}

}  // namespace pba
