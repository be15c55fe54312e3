// STAND-IN -- this is NOT the reference's track subsystem.  Test infrastructure only: an immature landmark as a plain
// record with the accessors landmarks_activator.cpp:285-316 uses; idepth() is the reference's midpoint of the search
// interval (track/landmarks/src/immature_tracking_landmark.cpp:23).
#ifndef DSOPP_IMMATURE_TRACKING_LANDMARK_HPP
#define DSOPP_IMMATURE_TRACKING_LANDMARK_HPP
#include <Eigen/Dense>

#include "common/pattern/pattern.hpp"
#include "common/settings.hpp"

namespace dsopp::track::landmarks {
class ImmatureTrackingLandmark {
 public:
  Eigen::Vector2<Precision> projection_;
  Eigen::Vector<Precision, Pattern::kSize> patch_;
  Precision idepth_min_ = 0, idepth_max_ = 0;
  const Eigen::Vector2<Precision>& projection() const { return projection_; }
  const Eigen::Vector<Precision, Pattern::kSize>& patch() const { return patch_; }
  Precision idepth() const { return idepth_max_ * 0.5_p + idepth_min_ * 0.5_p; }
  Precision idepthMin() const { return idepth_min_; }
  Precision idepthMax() const { return idepth_max_; }
  void setIdepthMin(Precision v) { idepth_min_ = v; }
  void setIdepthMax(Precision v) { idepth_max_ = v; }
};
}  // namespace dsopp::track::landmarks
#endif
