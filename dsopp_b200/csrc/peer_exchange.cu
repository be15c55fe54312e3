// One-shot all-reduce of the exchange block over NVLink peer memory (option "peer_exchange", default OFF).
//
// the only data-path exchange of the sharded bundle adjustment is the sum over ranks of the
// packed [H_pp | b_p | H_s | b_s | scal] block (66.6 KB at 8 keyframes).  At that size an allreduce is pure latency:
// NCCL's kernel costs ~25 us per call on two B200s (profiles/r01h_bench_2gpu.json: 122 us per iteration against 97 us
// on one GPU).  This kernel replaces the library call with stores into the peers' mailboxes:
//
//   * every rank owns ONE mailbox allocation (cudaMalloc, exported with cudaIpcGetMemHandle and opened by the peers
//     with cudaIpcOpenMemHandle): data[2][PEER_MAXW][slot] doubles + flag[PEER_MAXW][PEER_MAXC] words;
//   * CTA c of rank s PUSHES its slice of the local block into slot [parity][s] of EVERY rank's mailbox (its own
//     included) with plain 16-byte stores -- NVLink stores are posted, loads are round trips --, fences at system
//     scope and raises flag[s][c] = epoch in every mailbox (st.release.sys);
//   * CTA c of rank d waits for flag[s][c] >= epoch of all s (ld.acquire.sys, with a time-out instead of a hang),
//     then sums slot [parity][0..W) in RANK ORDER from its own mailbox (L2 loads) into the out buffer.
//
// Every rank adds the same W values in the same order, so all ranks hold bitwise identical sums -- the replicated
// LM step (k_lm_step) relies on that, as it does with NCCL.  CTAs only ever wait for the remote CTA of the same index
// and the grid is at most PEER_MAXC CTAs (all co-resident), so there is no scheduling dependency between CTAs.
// The exchange counter ("epoch") lives in device memory: every CTA reads it when it starts and the LAST CTA to finish
// bumps it, so the kernel can be captured in the LM graph and replayed, and all CTAs of one call -- whatever the grid
// size of the call -- agree on the epoch and on the parity (epoch & 1) of the mailbox half they use.  Two parities are
// enough: all exchange kernels of a rank run in stream order, so rank A can only reach exchange k+2 after rank B
// pushed k+1, which B does after its kernel of exchange k (the reader of parity k&1) has completed.
//
// STATUS: runs on B200 since round 2 (2, 4 and 8 GPUs: tests/test_gpu_multi.py, profiles/r02_bench_g*.json); bench.py
// uses it by default for world > 1.  It is off unless dpba_peer_attach() succeeded AND the option "peer_exchange" is
// set; tools/multigpu_check.py (DPBA_PEER_EXCHANGE=1) compares it with the NCCL path on the same handle.  The steps live in peer_exchange_body.h and
// run on the CPU with real threads (one per CTA and rank, C++11 atomics for the PTX accesses) in
// tests/test_kernel_emulation.py.
#include <cuda_runtime.h>

#include <cstdint>

#include "pba_internal.h"
#include "peer_exchange_body.h"

namespace pba {

namespace {

// diagnostics (dpba_debug_kernel_times slots 10 / 11): entry of CTA 0, start of its wait for the peers, latest exit
__device__ unsigned long long g_peer_kst[4];
__device__ int g_peer_stamps_on = 0;
__device__ __forceinline__ unsigned long long peer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerDev pd, const double* __restrict__ in,
                                                        double* __restrict__ out, size_t off, size_t n2, int fence_all) {
  const int c = blockIdx.x;
  if (g_peer_stamps_on && c == 0 && threadIdx.x == 0) g_peer_kst[0] = peer_ns();
  const int C = gridDim.x;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = peer_begin(pd);
  __syncthreads();
  const unsigned epoch = s_epoch;
  const PeerSlice sl = peer_slice(n2, c, C);

  // ---- push this CTA's slice to every mailbox ----------------------------------------------------------------------
  for (size_t i = sl.lo + threadIdx.x; i < sl.hi; i += blockDim.x) peer_push_elem(pd, in, off, epoch, i);
  // Publication: the pushes of ALL threads of the CTA are ordered before the flag by the block barrier (CTA scope) followed by
  // ONE system-scope fence + release store per destination in the signalling threads -- the PTX memory model's causality order
  // is cumulative across the barrier (the pattern of cooperative-groups grid sync and of NCCL's postPeer).  The round-1 form, a
  // system-scope fence in every pushing thread (4352 of them for the 66 KB block), measured 17.8 us per exchange of which 4.5 us
  // were spent waiting for the peer; fence_all = 1 brings it back for the A/B.
  if (fence_all) __threadfence_system();
  __syncthreads();
  // ---- raise our flags, wait for the same slice of every rank ---------------------------------------------------------
  if (g_peer_stamps_on && c == 0 && threadIdx.x == 0) g_peer_kst[2] = peer_ns();
  if ((int)threadIdx.x < pd.world) {
    if (!fence_all) peer_fence_system();
    peer_signal(pd, c, epoch, (int)threadIdx.x);
    peer_wait(pd, c, epoch, (int)threadIdx.x);
  }
  __syncthreads();
  if (g_peer_stamps_on && c == 0 && threadIdx.x == 0) g_peer_kst[3] = peer_ns();
  // ---- sum in rank order from the local mailbox ------------------------------------------------------------------------
  for (size_t i = sl.lo + threadIdx.x; i < sl.hi; i += blockDim.x) peer_sum_elem(pd, out, off, epoch, i);
  if (threadIdx.x == 0) peer_finish(pd, C, epoch);
  if (g_peer_stamps_on && threadIdx.x == 0) atomicMax(&g_peer_kst[1], peer_ns());
}

}  // namespace

static int g_peer_fence_all = 0;  // option "peer_fence_all"
void set_peer_fence_all(int v) { g_peer_fence_all = v != 0; }
void peer_stamps_off_async(cudaStream_t s) {
  void* p = nullptr;
  cudaGetSymbolAddress(&p, g_peer_stamps_on);
  cudaMemsetAsync(p, 0, sizeof(int), s);
}
void debug_peer_times(int enable, long long out[4]) {
  cudaMemcpyToSymbol(g_peer_stamps_on, &enable, sizeof(int));
  if (out) cudaMemcpyFromSymbol(out, g_peer_kst, 4 * sizeof(long long));
}

// n doubles starting at `off` (both even: every block boundary of RedLayout is) of `in`, summed over ranks into `out`
void launch_peer_allreduce(const PeerDev& pd, const double* in, double* out, size_t off, size_t n, cudaStream_t s) {
  k_peer_allreduce<<<peer_grid(n), 256, 0, s>>>(pd, in, out, off, n / 2, g_peer_fence_all);
  add_launches(1);
}

}  // namespace pba
