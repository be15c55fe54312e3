// STAND-IN -- this is NOT the reference's track subsystem.  Test infrastructure only: an active keyframe as a plain record
// with the accessors the tracker's depth-map build (create_depth_maps.cpp) and landmark activation (landmarks_activator.cpp:
// 122-316) read.  The images (features::PixelMap), masks (sensors::calibration::CameraMask) and poses it hands out are the
// reference's own classes.
#ifndef DSOPP_ACTIVE_KEYFRAME_HPP
#define DSOPP_ACTIVE_KEYFRAME_HPP
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <vector>

#include <Eigen/Dense>

#include "common/settings.hpp"
#include "energy/motion/motion.hpp"
#include "features/camera/pixel_data_frame.hpp"
#include "features/camera/pixel_map.hpp"
#include "sensors/camera_calibration/mask/camera_mask.hpp"
#include "track/connections/frame_connection.hpp"
#include "track/landmarks/active_tracking_landmark.hpp"
#include "track/landmarks/immature_tracking_landmark.hpp"

namespace dsopp::track {
template <energy::motion::Motion Motion>
class ActiveKeyframe {
 public:
  enum struct ImmatureLandmarkActivationStatus : uint8_t { kActivate = 0, kSkip = 1, kDelete = 2 };
  using Pyramids = std::map<size_t, features::Pyramid>;

  size_t id_ = 0, keyframe_id_ = 0;
  Motion t_world_agent_;
  Precision exposure_time_ = 1;
  Eigen::Vector<Precision, 2> affine_brightness_ = Eigen::Vector<Precision, 2>::Zero();
  Pyramids pyramids_;
  std::map<size_t, std::vector<sensors::calibration::CameraMask>> masks_;
  std::map<size_t, std::vector<landmarks::ActiveTrackingLandmark>> active_landmarks_;
  mutable std::map<size_t, FrameConnection<typename Motion::Product>> connections_;

  size_t id() const { return id_; }
  size_t keyframeId() const { return keyframe_id_; }
  const Motion& tWorldAgent() const { return t_world_agent_; }
  Precision exposureTime() const { return exposure_time_; }
  const Eigen::Vector<Precision, 2>& affineBrightness() const { return affine_brightness_; }
  const Pyramids& pyramids() const { return pyramids_; }
  const features::PixelMap<1>& getLevel(const size_t sensor, size_t level) const { return pyramids_.at(sensor).at(level); }
  const sensors::calibration::CameraMask& getMask(const size_t sensor, size_t level) const { return masks_.at(sensor).at(level); }
  const std::vector<landmarks::ActiveTrackingLandmark>& activeLandmarks(size_t sensor) const { return active_landmarks_.at(sensor); }
  FrameConnection<typename Motion::Product>& getConnection(size_t target_keyframe_id) const {
    return connections_.at(target_keyframe_id);
  }
};
}  // namespace dsopp::track
#endif
