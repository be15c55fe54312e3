// STAND-IN -- this is NOT the reference's track subsystem.  Test infrastructure only: an active landmark as a plain record
// with the accessors create_depth_maps.cpp:38-52 reads.
#ifndef DSOPP_ACTIVE_TRACKING_LANDMARK_HPP
#define DSOPP_ACTIVE_TRACKING_LANDMARK_HPP
#include <Eigen/Dense>

#include "common/settings.hpp"

namespace dsopp::track::landmarks {
class ActiveTrackingLandmark {
 public:
  Eigen::Vector2<Precision> projection_;
  Eigen::Vector3<Precision> direction_;
  Precision idepth_ = 0, idepth_variance_ = 0;
  bool outlier_ = false, marginalized_ = false;
  const Eigen::Vector2<Precision>& projection() const { return projection_; }
  const Eigen::Vector3<Precision>& direction() const { return direction_; }
  Precision idepth() const { return idepth_; }
  Precision idepthVariance() const { return idepth_variance_; }
  bool isOutlier() const { return outlier_; }
  bool isMarginalized() const { return marginalized_; }
};
}  // namespace dsopp::track::landmarks
#endif
