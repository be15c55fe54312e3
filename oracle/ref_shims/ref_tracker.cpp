// C entry points over the REFERENCE'S OWN tracker arithmetic that sits either side of the bundle adjustment (test
// infrastructure; built by oracle/build_ref_tracker.py into oracle/_ref/libdsopp_ref_tracker.so):
//   createReferenceDepthMaps  src/tracker/tracker/src/create_depth_maps.cpp  -- the WHOLE FILE compiled unchanged
//                             (fillFineDepthMap :19-58, fillCoarseDepthMaps :70-88, dilateDepthMaps :90-120, :122-146)
//   LandmarkActivationProblem, optimizeImmatureLandmark
//                             src/tracker/landmarks_activator/src/landmarks_activator.cpp:122-316 -- the file's own lines,
//                             given to the compiler through REF_ACTIVATOR_PREFIX (the rest of that file drives the track
//                             subsystem), under the reference's LM driver (levenberg_marquardt_algorithm.hpp:77-128)
// The track CONTAINERS those functions read (ActiveKeyframe, the landmark records, FrameConnection, CameraCalibration) are
// the plain records of oracle/ref_stubs_track; everything that computes -- the pinhole model, ArrayReprojector, SE3,
// PixelMap::Evaluate, CameraMask::valid, SimilarityMeasureSSD, the LM driver -- is the reference's code.
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <vector>

#include <glog/logging.h>
#include <opencv2/opencv.hpp>

#include "common/pattern/pattern.hpp"
#include "common/settings.hpp"
#include "energy/camera_model/pinhole/pinhole_camera.hpp"
#include "energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp"
#include "energy/motion/se3_motion.hpp"
#include "energy/projector/camera_reproject.hpp"
#include "features/camera/pattern_patch.hpp"
#include "features/camera/pixel_map.hpp"
#include "measures/similarity_measure_ssd.hpp"
#include "sensors/camera_calibration/camera_calibration.hpp"
#include "sensors/camera_calibration/mask/camera_mask.hpp"
#include "track/frames/active_keyframe.hpp"
#include "tracker/create_depth_maps.hpp"

namespace dsopp {
namespace tracker {
namespace {
#include REF_ACTIVATOR_PREFIX
}  // namespace
}  // namespace tracker
}  // namespace dsopp

namespace {
using dsopp::Precision;
using Motion = dsopp::energy::motion::SE3<Precision>;
using Model = dsopp::energy::model::PinholeCamera<Precision>;
using Keyframe = dsopp::track::ActiveKeyframe<Motion>;
constexpr size_t kSensor = 0;

Motion pose_from(const double* T34) {
  Eigen::Matrix<Precision, 3, 3> R;
  Eigen::Vector<Precision, 3> t;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R(i, j) = static_cast<Precision>(T34[4 * i + j]);
    t(i) = static_cast<Precision>(T34[4 * i + 3]);
  }
  return Motion(R, t);
}
dsopp::features::PixelMap<1> level_from(const double* image, int width, int height) {
  std::vector<Precision, dsopp::PrecisionAllocator> data(static_cast<size_t>(width) * static_cast<size_t>(height));
  for (size_t i = 0; i < data.size(); ++i) data[i] = image ? static_cast<Precision>(image[i]) : Precision(0);
  return dsopp::features::PixelMap<1>(std::move(data), static_cast<long>(width), static_cast<long>(height));
}
}  // namespace

extern "C" {

// frames: n poses (3x4 each) in window order, the last one is the target.  Landmarks of frame f < n - 1 are the records
// lm_offset[f] .. lm_offset[f + 1].  out: per level l, (W >> l) * (H >> l) idepth sums (row-major, [y][x]) followed by as
// many weights.  Returns the number of levels.
int reftrk_create_reference_depth_maps(int n_frames, const double* T_w_agent, const double* intr, int width, int height,
                                       int levels, const int32_t* lm_offset, const double* uv, const double* idepth,
                                       const double* idepth_variance, const uint8_t* outlier, const uint8_t* marginalized,
                                       const uint8_t* status, double* out) {
  Eigen::Vector2<Precision> image_size(static_cast<Precision>(width), static_cast<Precision>(height));
  Eigen::VectorX<Precision> k(4);
  for (int i = 0; i < 4; ++i) k(i) = static_cast<Precision>(intr[i]);
  dsopp::sensors::calibration::CameraCalibration calibration(image_size, k, dsopp::energy::model::ModelType::kPinholeCamera);
  auto model = calibration.cameraModel<Model>();

  std::vector<std::unique_ptr<Keyframe>> owned;
  std::deque<Keyframe*> frames;
  for (int f = 0; f < n_frames; ++f) {
    auto kf = std::make_unique<Keyframe>();
    kf->id_ = static_cast<size_t>(10 + 3 * f), kf->keyframe_id_ = static_cast<size_t>(f);
    kf->t_world_agent_ = pose_from(T_w_agent + 12 * f);
    if (f + 1 < n_frames) {
      auto& lms = kf->active_landmarks_[kSensor];
      auto& st = kf->connections_[static_cast<size_t>(n_frames - 1)].statuses;
      for (int32_t i = lm_offset[f]; i < lm_offset[f + 1]; ++i) {
        dsopp::track::landmarks::ActiveTrackingLandmark lm;
        lm.projection_ = Eigen::Vector2<Precision>(static_cast<Precision>(uv[2 * i]), static_cast<Precision>(uv[2 * i + 1]));
        model->unproject(lm.projection_, lm.direction_);  // what the track stores with a landmark
        lm.idepth_ = static_cast<Precision>(idepth[i]);
        lm.idepth_variance_ = static_cast<Precision>(idepth_variance[i]);
        lm.outlier_ = outlier[i] != 0, lm.marginalized_ = marginalized[i] != 0;
        lms.push_back(lm);
        st.push_back(static_cast<dsopp::track::PointConnectionStatus>(status[i]));
      }
    } else {
      for (int l = 0; l < levels; ++l) kf->pyramids_[kSensor].push_back(level_from(nullptr, width >> l, height >> l));
    }
    frames.push_back(kf.get());
    owned.push_back(std::move(kf));
  }
  auto maps = dsopp::tracker::createReferenceDepthMaps<Motion, Model>(frames, calibration);
  const auto& pyramid = maps.at(kSensor);
  size_t o = 0;
  for (size_t l = 0; l < pyramid.size(); ++l) {
    const auto& m = pyramid[l].map;  // indexed (x, y): rows() is the width
    const long w = m.rows(), h = m.cols();
    for (long y = 0; y < h; ++y)
      for (long x = 0; x < w; ++x) out[o++] = static_cast<double>(m(x, y).idepth);
    for (long y = 0; y < h; ++y)
      for (long x = 0; x < w; ++x) out[o++] = static_cast<double>(m(x, y).weight);
  }
  return static_cast<int>(pyramid.size());
}

// One immature landmark of frame `ref_index` refined against all the frames of the window (optimizeImmatureLandmark).
// images: n_frames level-0 rasters (height x width, raw intensities); masks: n_frames 8-bit rasters or null.
// Returns the activation status (0 activate, 2 delete); idepth_out is the landmark's idepth() afterwards.
int reftrk_optimize_immature_landmark(int n_frames, const double* T_w_agent, const double* exposure, const double* affine,
                                      const double* images, const uint8_t* masks, const double* intr, int width,
                                      int height, int ref_index, const double* projection, const double* patch,
                                      double idepth_min, double idepth_max, int minimum_inliers, double sigma_huber,
                                      double* idepth_out) {
  Eigen::Vector2<Precision> image_size(static_cast<Precision>(width), static_cast<Precision>(height));
  Eigen::Vector<Precision, 4> k(static_cast<Precision>(intr[0]), static_cast<Precision>(intr[1]),
                                static_cast<Precision>(intr[2]), static_cast<Precision>(intr[3]));
  Model model(image_size, k);
  const size_t npx = static_cast<size_t>(width) * static_cast<size_t>(height);
  std::vector<std::unique_ptr<Keyframe>> owned;
  std::deque<Keyframe*> frames;
  for (int f = 0; f < n_frames; ++f) {
    auto kf = std::make_unique<Keyframe>();
    kf->id_ = static_cast<size_t>(10 + 3 * f), kf->keyframe_id_ = static_cast<size_t>(f);
    kf->t_world_agent_ = pose_from(T_w_agent + 12 * f);
    kf->exposure_time_ = static_cast<Precision>(exposure[f]);
    kf->affine_brightness_ = Eigen::Vector2<Precision>(static_cast<Precision>(affine[2 * f]), static_cast<Precision>(affine[2 * f + 1]));
    kf->pyramids_[kSensor].push_back(level_from(images + npx * static_cast<size_t>(f), width, height));
    cv::Mat m(height, width, CV_8UC1, 255);
    if (masks) std::memcpy(m.data, masks + npx * static_cast<size_t>(f), npx);
    kf->masks_[kSensor].emplace_back(m);
    frames.push_back(kf.get());
    owned.push_back(std::move(kf));
  }
  dsopp::track::landmarks::ImmatureTrackingLandmark lm;
  lm.projection_ = Eigen::Vector2<Precision>(static_cast<Precision>(projection[0]), static_cast<Precision>(projection[1]));
  for (int i = 0; i < dsopp::Pattern::kSize; ++i) lm.patch_(i) = static_cast<Precision>(patch[i]);
  lm.idepth_min_ = static_cast<Precision>(idepth_min), lm.idepth_max_ = static_cast<Precision>(idepth_max);
  const Keyframe& ref = *frames[static_cast<size_t>(ref_index)];
  const auto status = dsopp::tracker::optimizeImmatureLandmark<Motion, Model, dsopp::features::PixelMap, 1>(
      lm, ref.id(), ref.exposureTime(), ref.affineBrightness(), kSensor, frames, model, minimum_inliers, ref.tWorldAgent(),
      static_cast<Precision>(sigma_huber));
  *idepth_out = static_cast<double>(lm.idepth());
  return static_cast<int>(status);
}
}
