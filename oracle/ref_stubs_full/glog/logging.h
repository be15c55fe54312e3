// STAND-IN -- this is NOT glog.  Test infrastructure only: CHECK* abort with a message, LOG goes to a null stream.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <unordered_map>
namespace ref_stub_glog {
struct Fatal {
  std::ostringstream s;
  Fatal(const char* file, int line, const char* what) { s << file << ":" << line << " CHECK failed: " << what << " "; }
  [[noreturn]] ~Fatal() {
    std::cerr << s.str() << std::endl;
    std::abort();
  }
  template <class T>
  Fatal& operator<<(const T& v) {
    s << v;
    return *this;
  }
};
struct Null {
  template <class T>
  Null& operator<<(const T&) {
    return *this;
  }
  Null& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
struct Voidify {
  void operator&(const Fatal&) {}
  void operator&(const Null&) {}
};
}  // namespace ref_stub_glog
#define CHECK(cond) (cond) ? (void)0 : ref_stub_glog::Voidify() & ref_stub_glog::Fatal(__FILE__, __LINE__, #cond)
#define CHECK_EQ(a, b) CHECK(static_cast<long long>(a) == static_cast<long long>(b))
#define CHECK_NE(a, b) CHECK((a) != (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_NOTNULL(p) (p)
#define LOG(severity) ref_stub_glog::Null()
#define VLOG(n) ref_stub_glog::Null()
#define DLOG(severity) ref_stub_glog::Null()
#define LOG_IF(severity, c) ref_stub_glog::Null()
