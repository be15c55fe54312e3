// Stand-alone driver of tests/emu/peer_emu.cpp for a ThreadSanitizer build (tests/test_kernel_emulation.py): the mailbox
// all-reduce with real threads must be free of data races -- pushes are ordered before the sums by the release / acquire
// flags alone, and the reuse of a mailbox half two exchanges later by the stream order of the kernels.
#include <cstdio>
#include <vector>
extern "C" long long emu_peer_run(int, int, const long long*, const long long*, long long, unsigned, int);
int main() {
  const long long D = 64, sysn = 2 * (D * D + D), slot = sysn + 8;
  std::vector<long long> offs, ns;
  auto add = [&](long long o, long long n) { offs.push_back(o); ns.push_back(n); };
  for (int i = 0; i < 3; ++i) add(0, sysn + 8);
  add(sysn, 8); add(0, sysn); add(sysn, 8); add(sysn, 8);
  for (int i = 0; i < 4; ++i) add(0, sysn + 8);
  add(sysn, 8);
  long long bad = 0;
  for (int world : {2, 4})
    for (unsigned seed = 0; seed < 3; ++seed) bad += emu_peer_run(world, (int)offs.size(), offs.data(), ns.data(), slot, seed, 100);
  printf("wrong elements: %lld\n", bad);
  return bad != 0;
}
