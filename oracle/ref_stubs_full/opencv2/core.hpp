// STAND-IN -- this is NOT OpenCV.  Test infrastructure only: cv::Mat as a plain 8-bit / float raster, enough for the
// reference's CameraMask (sensors/camera_calibration/mask/camera_mask.hpp) and for the declarations that name cv::Mat.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5
#define CV_64F 6
typedef unsigned char uchar;
namespace cv {
enum InterpolationFlags { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2 };
enum MorphShapes { MORPH_RECT = 0 };
struct Size {
  int width = 0, height = 0;
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
};
struct Point {
  int x = 0, y = 0;
  Point() = default;
  Point(int xx, int yy) : x(xx), y(yy) {}
};
class Mat {
 public:
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;  // first byte of the raster (shared by copies, like OpenCV's header copies)
  Mat() = default;
  Mat(int r, int c, int type, double fill = 0) : rows(r), cols(c), type_(type) {
    d_ = std::make_shared<std::vector<unsigned char>>(static_cast<size_t>(r) * static_cast<size_t>(c) * elem(), 0);
    if (type == CV_8UC1) std::memset(d_->data(), static_cast<int>(fill), d_->size());
    data = d_->data();
  }
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  bool empty() const { return rows == 0 || cols == 0; }
  Mat clone() const {
    Mat m;
    m.rows = rows, m.cols = cols, m.type_ = type_;
    if (d_) m.d_ = std::make_shared<std::vector<unsigned char>>(*d_);
    m.data = m.d_ ? m.d_->data() : nullptr;
    return m;
  }
  template <class T>
  T& at(int y, int x) {
    return reinterpret_cast<T*>(d_->data())[static_cast<size_t>(y) * static_cast<size_t>(cols) + static_cast<size_t>(x)];
  }
  template <class T>
  const T& at(int y, int x) const {
    return reinterpret_cast<const T*>(d_->data())[static_cast<size_t>(y) * static_cast<size_t>(cols) + static_cast<size_t>(x)];
  }
  template <class T, class F>
  void forEach(const F& f) {
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols; ++x) {
        const int pos[2] = {y, x};
        f(at<T>(y, x), pos);
      }
  }
  unsigned char* ptr() { return d_ ? d_->data() : nullptr; }

 private:
  size_t elem() const { return type_ == CV_8UC1 ? 1 : (type_ == CV_32F ? 4 : 8); }
  int type_ = CV_8UC1;
  std::shared_ptr<std::vector<unsigned char>> d_;
};
// extrema of an 8-bit raster (photometricallyCorrectedImage: the vignette's maximum)
inline void minMaxLoc(const Mat& m, double* min_value, double* max_value) {
  double lo = 255, hi = 0;
  for (int y = 0; y < m.rows; ++y)
    for (int x = 0; x < m.cols; ++x) {
      const double v = m.at<unsigned char>(y, x);
      lo = v < lo ? v : lo, hi = v > hi ? v : hi;
    }
  if (min_value) *min_value = lo;
  if (max_value) *max_value = hi;
}
}  // namespace cv
