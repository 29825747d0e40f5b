// TEST STUB (not OpenCV): the few members of cv::Mat that the reference's public headers and this repository's C++
// boundary touch, for the compile-check of tests/test_cpp_adaptor.py.
#pragma once
#include <cstdint>
#define CV_8U 0
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() = default;
  Mat(int r, int c, int type, void* data) : rows(r), cols(c), type_(type), data_(static_cast<unsigned char*>(data)) {}
  bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return true; }
  int channels() const { return (type_ >> 3) + 1; }
  int depth() const { return type_ & 7; }
  template <typename T>
  T* ptr(int r = 0) {
    return reinterpret_cast<T*>(data_) + static_cast<long long>(r) * cols * channels();
  }
  template <typename T>
  const T* ptr(int r = 0) const {
    return reinterpret_cast<const T*>(data_) + static_cast<long long>(r) * cols * channels();
  }

 private:
  int type_ = 0;
  unsigned char* data_ = nullptr;
};
}  // namespace cv
