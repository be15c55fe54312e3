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
// STATUS: written in round 1 after the GPU budget was spent -- compiled for sm_100a, NOT yet run on hardware.  It is
// off unless dpba_peer_attach() succeeded AND the option "peer_exchange" is set; tools/multigpu_check.py
// (DPBA_PEER_EXCHANGE=1) compares it with the NCCL path on the same handle.
#include <cuda_runtime.h>

#include <cstdint>

#include "pba_internal.h"

namespace pba {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ~15 s at 1.9 GHz.  Ranks legitimately reach an exchange seconds apart (first-call graph instantiation, a rank that
// does extra host work), so the bound is generous; a peer that has not arrived by then has failed -- give up loudly
// (DPBA_E_COMM on the host) instead of hanging the GPU until an outer limit kills the process
constexpr long long PEER_TIMEOUT_CYCLES = 30000000000LL;

__global__ void __launch_bounds__(256) k_peer_allreduce(PeerDev pd, const double* __restrict__ in,
                                                        double* __restrict__ out, size_t off, size_t n2) {
  const int c = blockIdx.x;
  const int C = gridDim.x;
  __shared__ unsigned s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile unsigned*>(pd.seq) + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  const size_t par = epoch & 1u;

  const size_t per = (n2 + C - 1) / C;
  const size_t lo = (size_t)c * per < n2 ? (size_t)c * per : n2;
  const size_t hi = lo + per < n2 ? lo + per : n2;
  const double2* src = reinterpret_cast<const double2*>(in + off);
  const size_t my_slot = (par * PEER_MAXW + (size_t)pd.rank) * pd.slot + off;

  // ---- push this CTA's slice to every mailbox ----------------------------------------------------------------------
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const double2 v = src[i];
#pragma unroll
    for (int r = 0; r < PEER_MAXW; ++r)  // unrolled: the pointer table stays in the constant bank
      if (r < pd.world) reinterpret_cast<double2*>(pd.data[r] + my_slot)[i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pd.world) {
    unsigned* peer_flags = nullptr;
    unsigned* my_flags = nullptr;
#pragma unroll
    for (int r = 0; r < PEER_MAXW; ++r) {
      if (r == (int)threadIdx.x) peer_flags = pd.flag[r];
      if (r == pd.rank) my_flags = pd.flag[r];
    }
    __threadfence_system();
    st_release_sys(peer_flags + pd.rank * PEER_MAXC + c, epoch);
    // ---- wait for the same slice of rank threadIdx.x ---------------------------------------------------------------
    const unsigned* f = my_flags + threadIdx.x * PEER_MAXC + c;
    const long long t0 = clock64();
    // after one time-out the run is lost anyway: later exchanges do not wait again (the host reports DPBA_E_COMM)
    while (*reinterpret_cast<volatile int*>(pd.error) == 0 && (int)(ld_acquire_sys(f) - epoch) < 0) {
      __nanosleep(32);
      if (clock64() - t0 > PEER_TIMEOUT_CYCLES) {
        *reinterpret_cast<volatile int*>(pd.error) = 1;
        *reinterpret_cast<volatile int*>(pd.error_host) = 1;
        break;
      }
    }
  }
  __syncthreads();

  // ---- sum in rank order from the local mailbox (peer stores land in this GPU's L2: bypass L1) -------------------
  const double* box = nullptr;
#pragma unroll
  for (int r = 0; r < PEER_MAXW; ++r)
    if (r == pd.rank) box = pd.data[r];
  box += par * PEER_MAXW * pd.slot + off;
  double2* dst = reinterpret_cast<double2*>(out + off);
  for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    double2 acc = __ldcg(reinterpret_cast<const double2*>(box) + i);
    for (int r = 1; r < pd.world; ++r) {
      const double2 v = __ldcg(reinterpret_cast<const double2*>(box + (size_t)r * pd.slot) + i);
      acc.x += v.x;
      acc.y += v.y;
    }
    dst[i] = acc;
  }
  // the last CTA to finish publishes the epoch (every CTA of this call has read pd.seq by then)
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(pd.done, 1u) == (unsigned)C - 1u) {
      *pd.done = 0u;
      __threadfence();
      *reinterpret_cast<volatile unsigned*>(pd.seq) = epoch;
    }
  }
}

}  // namespace

// n doubles starting at `off` (both even: every block boundary of RedLayout is) of `in`, summed over ranks into `out`
void launch_peer_allreduce(const PeerDev& pd, const double* in, double* out, size_t off, size_t n, cudaStream_t s) {
  const size_t n2 = n / 2;
  // one CTA per 256 double2 (4 KB), at most PEER_MAXC: 17 CTAs for the 66.6 KB block, 1 for the 8 scalars
  int ctas = (int)((n2 + 255) / 256);
  if (ctas < 1) ctas = 1;
  if (ctas > PEER_MAXC) ctas = PEER_MAXC;
  k_peer_allreduce<<<ctas, 256, 0, s>>>(pd, in, out, off, n2);
  add_launches(1);
}

}  // namespace pba
