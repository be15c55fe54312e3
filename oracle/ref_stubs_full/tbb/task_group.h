// STAND-IN -- this is NOT oneTBB.  Test infrastructure only: tasks run at once on the calling thread.
#pragma once
namespace tbb {
class task_group {
 public:
  template <class F>
  void run(const F& f) {
    f();
  }
  void wait() {}
};
}  // namespace tbb
