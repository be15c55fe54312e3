// The steps of the NVLink mailbox all-reduce (peer_exchange.cu) as __host__ __device__ functions.  On the device the
// synchronising accesses are system-scope PTX (st.release.sys / ld.acquire.sys / fence.sc.sys), on the host they are the
// corresponding C++11 atomics, so tests/emu/peer_emu.cpp can run THIS code with one host thread per CTA and per rank --
// real concurrency, arbitrary interleavings, ThreadSanitizer-compatible -- before it ever meets an NVLink.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "pba_internal.h"

#ifndef __CUDA_ARCH__
#include <sched.h>

#include <chrono>
#endif

namespace pba {

// ~15 s at 1.9 GHz (device cycles) / 30 s (host nanoseconds).  Ranks legitimately reach an exchange seconds apart
// (first-call graph instantiation, a rank that does extra host work), so the bound is generous; a peer that has not
// arrived by then has failed -- give up loudly (DPBA_E_COMM on the host) instead of hanging the GPU until an outer limit
// kills the process
constexpr long long PEER_TIMEOUT_TICKS = 30000000000LL;

#ifdef __CUDA_ARCH__
__device__ __forceinline__ void peer_store_release(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned peer_load_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void peer_fence_system() { __threadfence_system(); }
__device__ __forceinline__ double2 peer_load_data(const double2* p) { return __ldcg(p); }  // L2: peer stores land there
__device__ __forceinline__ unsigned peer_load_word(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ void peer_store_word(unsigned* p, unsigned v) { *reinterpret_cast<volatile unsigned*>(p) = v; }
__device__ __forceinline__ unsigned peer_fetch_inc(unsigned* p) { return atomicAdd(p, 1u); }
__device__ __forceinline__ void peer_fence_device() { __threadfence(); }
__device__ __forceinline__ void peer_backoff() { __nanosleep(32); }
__device__ __forceinline__ long long peer_ticks() { return clock64(); }
#else
inline void peer_store_release(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned peer_load_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void peer_fence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline double2 peer_load_data(const double2* p) { return *p; }
inline unsigned peer_load_word(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
inline void peer_store_word(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
inline unsigned peer_fetch_inc(unsigned* p) { return __atomic_fetch_add(p, 1u, __ATOMIC_ACQ_REL); }
inline void peer_fence_device() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void peer_backoff() { sched_yield(); }
inline long long peer_ticks() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
#endif

// CTAs of one exchange of n doubles: one per 256 double2 (4 KB), at most PEER_MAXC -- 17 for the 66.6 KB block of an
// 8-keyframe window, 1 for the 8 scalars
__host__ __device__ inline int peer_grid(size_t n) {
  const size_t n2 = n / 2;
  size_t ctas = (n2 + 255) / 256;
  if (ctas < 1) ctas = 1;
  if (ctas > (size_t)PEER_MAXC) ctas = PEER_MAXC;
  return (int)ctas;
}

struct PeerSlice {
  size_t lo, hi;  // double2 indices of this CTA's slice of the block
};
__host__ __device__ inline PeerSlice peer_slice(size_t n2, int c, int C) {
  const size_t per = (n2 + C - 1) / C;
  PeerSlice s;
  s.lo = (size_t)c * per < n2 ? (size_t)c * per : n2;
  s.hi = s.lo + per < n2 ? s.lo + per : n2;
  return s;
}

// epoch of the exchange that is starting: one more than the number of completed exchanges of this rank
__host__ __device__ inline unsigned peer_begin(const PeerDev& pd) { return peer_load_word(pd.seq) + 1u; }

// one thread: element i of the local block goes into slot [parity][rank] of EVERY rank's mailbox (posted stores)
__host__ __device__ inline void peer_push_elem(const PeerDev& pd, const double* in, size_t off, unsigned epoch, size_t i) {
  const size_t par = epoch & 1u;
  const size_t my_slot = (par * PEER_MAXW + (size_t)pd.rank) * pd.slot + off;
  const double2 v = reinterpret_cast<const double2*>(in + off)[i];
#pragma unroll
  for (int r = 0; r < PEER_MAXW; ++r)  // unrolled: on the device the pointer table stays in the constant bank
    if (r < pd.world) reinterpret_cast<double2*>(pd.data[r] + my_slot)[i] = v;
}

// thread s (< world) after the CTA's pushes are complete (barrier): tell rank s that slice c of this rank has landed ...
__host__ __device__ inline void peer_signal(const PeerDev& pd, int c, unsigned epoch, int s) {
  unsigned* peer_flags = nullptr;
#pragma unroll
  for (int r = 0; r < PEER_MAXW; ++r)
    if (r == s) peer_flags = pd.flag[r];
  // no second fence here: every pushing thread fenced at system scope before the CTA barrier that precedes this call, and
  // the release store orders this thread's own earlier accesses
  peer_store_release(peer_flags + pd.rank * PEER_MAXC + c, epoch);
}

// ... then wait until slice c of rank s has landed here.  After one time-out nobody waits again.
__host__ __device__ inline void peer_wait(const PeerDev& pd, int c, unsigned epoch, int s) {
  unsigned* my_flags = nullptr;
#pragma unroll
  for (int r = 0; r < PEER_MAXW; ++r)
    if (r == pd.rank) my_flags = pd.flag[r];
  const unsigned* f = my_flags + s * PEER_MAXC + c;
  const long long t0 = peer_ticks();
  while (peer_load_word(reinterpret_cast<const unsigned*>(pd.error)) == 0u && (int)(peer_load_acquire(f) - epoch) < 0) {
    peer_backoff();
    if (peer_ticks() - t0 > PEER_TIMEOUT_TICKS) {
      peer_store_word(reinterpret_cast<unsigned*>(pd.error), 1u);
      peer_store_word(reinterpret_cast<unsigned*>(pd.error_host), 1u);
      break;
    }
  }
}

// one thread: element i summed over the ranks IN RANK ORDER from this rank's own mailbox
__host__ __device__ inline void peer_sum_elem(const PeerDev& pd, double* out, size_t off, unsigned epoch, size_t i) {
  const size_t par = epoch & 1u;
  const double* box = nullptr;
#pragma unroll
  for (int r = 0; r < PEER_MAXW; ++r)
    if (r == pd.rank) box = pd.data[r];
  box += par * PEER_MAXW * pd.slot + off;
  double2 acc = peer_load_data(reinterpret_cast<const double2*>(box) + i);
  for (int r = 1; r < pd.world; ++r) {
    const double2 v = peer_load_data(reinterpret_cast<const double2*>(box + (size_t)r * pd.slot) + i);
    acc.x += v.x;
    acc.y += v.y;
  }
  reinterpret_cast<double2*>(out + off)[i] = acc;
}

// thread 0 of a CTA that is done: the LAST CTA of the call to finish publishes the epoch (every CTA has read pd.seq by
// then, whatever the grid size of the call)
__host__ __device__ inline void peer_finish(const PeerDev& pd, int C, unsigned epoch) {
  peer_fence_device();
  if (peer_fetch_inc(pd.done) == (unsigned)C - 1u) {
    peer_store_word(pd.done, 0u);
    peer_fence_device();
    peer_store_word(pd.seq, epoch);
  }
}

}  // namespace pba
