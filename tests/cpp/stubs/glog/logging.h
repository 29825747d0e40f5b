// TEST STUB (not glog): just enough of the glog macros for the reference's deploy_core headers to compile in the
// boundary compile-check (tests/test_cpp_adaptor.py); messages go to stderr.
#pragma once
#include <cstdlib>
#include <iostream>
#include <sstream>
namespace fpb_test_glog {
struct Line {
  std::ostringstream os;
  bool fatal;
  explicit Line(const char* sev, bool f = false) : fatal(f) { os << "[" << sev << "] "; }
  ~Line() {
    std::cerr << os.str() << std::endl;
    if (fatal) std::abort();
  }
  template <typename T>
  Line& operator<<(const T& v) {
    os << v;
    return *this;
  }
};
}  // namespace fpb_test_glog
#define LOG(sev) ::fpb_test_glog::Line(#sev)
#define CHECK(cond) \
  if (!(cond)) ::fpb_test_glog::Line("FATAL", true) << "Check failed: " #cond " "
