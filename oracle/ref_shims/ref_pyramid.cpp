// C entry points over the REFERENCE'S OWN image pyramid (test infrastructure; built by oracle/build_ref_pba.py):
//   photometricallyCorrectedImage   src/features/src/photometrically_corrected_image.cpp:9-29
//   downscaleImage                  src/features/internal/features/camera/downscale_image.hpp:16-33
//   PixelDataFrame::PixelDataFrame  src/features/src/pixel_data_frame.cpp:12-31  (correction, then levels-1 halvings, each
//                                   level packed into {I, dx, dy} by PixelMap<1>'s constructor / calculate_pixelinfo)
// The reference's code runs unchanged; this file only moves plain arrays in and out.
#include <array>
#include <cstring>
#include <vector>

#include <opencv2/opencv.hpp>

#include "common/settings.hpp"
#include "features/camera/downscale_image.hpp"
#include "features/camera/photometrically_corrected_image.hpp"
#include "features/camera/pixel_data_frame.hpp"
#include "features/camera/pixel_map.hpp"

namespace {
using dsopp::Precision;

cv::Mat raster(const unsigned char* p, int h, int w) {
  if (!p) return cv::Mat();
  cv::Mat m(h, w, CV_8UC1);
  std::memcpy(m.data, p, static_cast<size_t>(h) * static_cast<size_t>(w));
  return m;
}
std::array<Precision, 256> table(const double* lut) {
  std::array<Precision, 256> t;
  for (size_t i = 0; i < 256; ++i) t[i] = static_cast<Precision>(lut[i]);
  return t;
}
}  // namespace

extern "C" {

int refpyr_sizeof_precision() { return static_cast<int>(sizeof(Precision)); }

// out: h * w values
void refpyr_photometric_correction(const unsigned char* gray, int h, int w, const double* lut, const unsigned char* vignetting,
                                   double* out) {
  const auto r = dsopp::features::photometricallyCorrectedImage(raster(gray, h, w), table(lut), raster(vignetting, h, w));
  for (size_t i = 0; i < r.size(); ++i) out[i] = static_cast<double>(r[i]);
}

// out: (h / 2) * (w / 2) values
void refpyr_downscale(const double* image, int h, int w, double* out) {
  std::vector<Precision, dsopp::PrecisionAllocator> in(static_cast<size_t>(h) * static_cast<size_t>(w));
  for (size_t i = 0; i < in.size(); ++i) in[i] = static_cast<Precision>(image[i]);
  const auto r = dsopp::features::downscaleImage(in, h, w);
  for (size_t i = 0; i < r.size(); ++i) out[i] = static_cast<double>(r[i]);
}

// out: the levels one after the other, level l holding (h >> l) * (w >> l) records {I, dx, dy}; returns the number of levels
// the reference built (it clamps to PixelDataFrame::kMaxPyramidDepth)
int refpyr_pixel_data_frame(const unsigned char* gray, int h, int w, const double* lut, const unsigned char* vignetting,
                            int levels, double* out) {
  dsopp::features::PixelDataFrame frame(raster(gray, h, w), table(lut), raster(vignetting, h, w), static_cast<size_t>(levels));
  size_t k = 0;
  for (size_t l = 0; l < frame.size(); ++l) {
    const auto& level = frame.getLevel(l);
    for (long y = 0; y < level.height(); ++y)
      for (long x = 0; x < level.width(); ++x) {
        const auto& px = level(static_cast<int>(x), static_cast<int>(y));
        out[k++] = static_cast<double>(px.intensity());
        out[k++] = static_cast<double>(px.jacobian()(0));
        out[k++] = static_cast<double>(px.jacobian()(1));
      }
  }
  return static_cast<int>(frame.size());
}
}
