// The NVLink mailbox all-reduce (dsopp_b200/csrc/peer_exchange_body.h) run on the CPU with REAL concurrency: one host
// thread per CTA, one "stream" thread per rank (a rank's exchange kernels run in order, its CTAs concurrently), shared
// host memory in place of the peers' HBM, C++11 atomics in place of the system-scope PTX accesses.  The calls below are
// the kernel's own steps in the kernel's order; a CTA's 256 threads are walked sequentially between the points where
// the kernel has a __syncthreads().  Test infrastructure only (tests/test_kernel_emulation.py); build with -pthread,
// optionally -fsanitize=thread.
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

#include "peer_exchange_body.h"

using namespace pba;

namespace {

struct RankMem {
  std::vector<double> data;     // [2][PEER_MAXW][slot]
  std::vector<unsigned> flags;  // [PEER_MAXW][PEER_MAXC]
  unsigned ctr[4] = {0, 0, 0, 0};  // seq, done, error, pad
  unsigned error_host = 0;
  std::vector<double> in, out;
};

double contribution(int rank, int call, size_t i) { return 1e-3 * (double)(i % 977) + 17.0 * rank + 1000.0 * call + 0.125 * rank * call; }

void run_cta(const PeerDev& pd, const double* in, double* out, size_t off, size_t n2, int c, int C) {
  const unsigned epoch = peer_begin(pd);                     // thread 0, then __syncthreads()
  const PeerSlice sl = peer_slice(n2, c, C);
  for (size_t i = sl.lo; i < sl.hi; ++i) peer_push_elem(pd, in, off, epoch, i);
  peer_fence_system();                                        // __threadfence_system(); __syncthreads()
  for (int s = 0; s < pd.world; ++s) peer_signal(pd, c, epoch, s);   // threads 0..world-1 run concurrently on the device:
  for (int s = 0; s < pd.world; ++s) peer_wait(pd, c, epoch, s);     // every signal is issued before any of them waits
  for (size_t i = sl.lo; i < sl.hi; ++i) peer_sum_elem(pd, out, off, epoch, i);   // after __syncthreads()
  peer_finish(pd, C, epoch);
}

}  // namespace

extern "C" {

// world ranks run n_calls exchanges of (offs[k], ns[k]) doubles with per-rank random pauses of up to jitter_us between
// kernels; returns the number of wrong output elements (over all ranks and calls), -1 after a time-out flag.
long long emu_peer_run(int world, int n_calls, const long long* offs, const long long* ns, long long slot, unsigned seed,
                       int jitter_us) {
  std::vector<RankMem> mem(world);
  for (auto& m : mem) {
    m.data.assign(2 * (size_t)PEER_MAXW * slot, -1.0);
    m.flags.assign((size_t)PEER_MAXW * PEER_MAXC, 0u);
    m.in.assign(slot, 0.0);
    m.out.assign(slot, 0.0);
  }
  std::vector<PeerDev> pds(world);
  for (int r = 0; r < world; ++r) {
    PeerDev& pd = pds[r];
    memset(&pd, 0, sizeof(pd));
    for (int q = 0; q < world; ++q) {
      pd.data[q] = mem[q].data.data();
      pd.flag[q] = mem[q].flags.data();
    }
    pd.seq = &mem[r].ctr[0];
    pd.done = &mem[r].ctr[1];
    pd.error = reinterpret_cast<int*>(&mem[r].ctr[2]);
    pd.error_host = reinterpret_cast<int*>(&mem[r].error_host);
    pd.rank = r;
    pd.world = world;
    pd.slot = (size_t)slot;
  }
  std::atomic<long long> wrong{0};
  std::vector<std::thread> streams;
  for (int r = 0; r < world; ++r) {
    streams.emplace_back([&, r]() {
      std::mt19937 rng(seed * 7919u + (unsigned)r);
      RankMem& me = mem[r];
      for (int k = 0; k < n_calls; ++k) {
        if (jitter_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % (unsigned)(jitter_us + 1)));
        const size_t off = (size_t)offs[k], n = (size_t)ns[k];
        for (size_t i = 0; i < n; ++i) me.in[off + i] = contribution(r, k, i);
        const int C = peer_grid(n);
        std::vector<std::thread> ctas;
        for (int c = 0; c < C; ++c)
          ctas.emplace_back(run_cta, std::cref(pds[r]), me.in.data(), me.out.data(), off, n / 2, c, C);
        for (auto& t : ctas) t.join();  // kernel boundary
        for (size_t i = 0; i < n; ++i) {
          double expect = contribution(0, k, i);
          for (int q = 1; q < world; ++q) expect += contribution(q, k, i);
          if (me.out[off + i] != expect) wrong.fetch_add(1);
        }
      }
    });
  }
  for (auto& t : streams) t.join();
  for (auto& m : mem)
    if (m.ctr[2] || m.error_host) return -1;
  return wrong.load();
}

}  // extern "C"
