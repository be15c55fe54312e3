// STAND-IN -- this is NOT the reference's CameraCalibration.  Test infrastructure only: image size + intrinsics + the
// cameraModel<Model>() factory with the reference's body (sensors/camera_calibration/src/camera_calibration.cpp:67-70); the
// undistorter (OpenCV remap) it also carries is not on this path.
#ifndef DSOPP_CAMERA_CALIBRATION_HPP
#define DSOPP_CAMERA_CALIBRATION_HPP
#include <Eigen/Dense>
#include <memory>

#include "common/settings.hpp"
#include "common/time/time.hpp"
#include "energy/camera_model/camera_model_base.hpp"

namespace dsopp::sensors::calibration {
class CameraCalibration {
 public:
  CameraCalibration(const Eigen::Vector2<Precision>& image_size, const Eigen::VectorX<Precision>& camera_intrinsics,
                    energy::model::ModelType type)
      : image_size_(image_size), camera_intrinsics_(camera_intrinsics), type_(type), shutter_time_(std::chrono::seconds(0)) {}
  template <energy::model::Model Model>
  std::unique_ptr<Model> cameraModel(size_t level_shift = 0) const {
    return std::make_unique<Model>(image_size_, camera_intrinsics_, shutter_time_, 1 << level_shift);
  }

 private:
  Eigen::Vector2<Precision> image_size_;
  Eigen::VectorX<Precision> camera_intrinsics_;
  energy::model::ModelType type_;
  const time::duration shutter_time_;
};
}  // namespace dsopp::sensors::calibration
#endif
