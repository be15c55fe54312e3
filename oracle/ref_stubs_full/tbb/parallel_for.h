// STAND-IN -- this is NOT oneTBB.  Test infrastructure only: the reference's loops run serially, in index order.
#pragma once
#include <cstddef>
namespace tbb {
template <class T>
class blocked_range {
 public:
  blocked_range(T b, T e) : b_(b), e_(e) {}
  T begin() const { return b_; }
  T end() const { return e_; }

 private:
  T b_, e_;
};
template <class Range, class F>
void parallel_for(const Range& r, const F& f) {
  Range copy = r;
  f(copy);
}
}  // namespace tbb
