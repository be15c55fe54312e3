// CudaPoseAlignment: host side of the coarse-tracker direct image alignment, mirroring the reference's
//   PoseAlignment<SE3, Pinhole, 1, PixelMap, 1>      src/energy/problems/include/energy/problems/pose_alignment/pose_alignment.hpp:22-63
//   EigenPoseAlignment<..., OPTIMIZE_AFFINE = true>   src/energy/problems/include/energy/problems/pose_alignment/eigen_pose_alignment.hpp:25-70
//                                                     src/energy/problems/src/eigen_pose_alignment.cpp:243-335
// on top of the C ABI in include/dsopp_cuda_pose_alignment.h.  The call sequence is the tracker's
// (src/tracker/tracker/src/monocular_tracker.cpp:199-214): reset(); pushFrame(reference + depth map, kFixed);
// pushFrame(new frame, kFree); solve(); getPose / getAffineBrightness; tTargetReferenceCovariance().
// The reference toolchain (Eigen, Sophus) is absent here, so frames arrive as plain-struct views; INTEGRATION.md shows
// the adapter inside DSOPP.
#pragma once
#include <array>
#include <cstring>
#include <map>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../../include/dsopp_cuda_pba.h"
#include "../../../include/dsopp_cuda_pose_alignment.h"
#include "dense.hpp"

namespace dsopp_b200 {

struct PoseAlignmentOptions {  // createPoseAlignment, src/tracker/tracker/src/fabric.cpp:127-147
  size_t max_iterations = 50;
  double initial_trust_region_radius = 1e2;
  double function_tolerance = 1e-5;
  double parameter_tolerance = 1e-5;
  double affine_brightness_regularizer[2] = {1e12, 1e8};
  double sigma_huber_loss = 20;
};

// one pyramid level of a frame, as the two pushFrame overloads of the tracker pass it
struct AlignmentFrameView {
  long long timestamp = 0;
  double t_world_agent[12];
  double exposure_time = 1;
  double affine_brightness[2] = {0, 0};
  double intrinsics[4];
  int width = 0, height = 0;
  const float* image_I_dx_dy = nullptr;
  const uint8_t* mask = nullptr;        // target only
  const float* depth_idepth_sum = nullptr;  // reference only: DepthMap accumulators (create_depth_maps.cpp)
  const float* depth_weight = nullptr;
};

class CudaPoseAlignment {
 public:
  static constexpr double kZeroCost = -1.0;  // pose_alignment.hpp:29
  CudaPoseAlignment(const PoseAlignmentOptions& o, int max_width, int max_height, int device = 0) : options_(o) {
    dpa_config cfg{max_width * max_height, max_width, max_height, device};
    const int rc = dpa_create(&cfg, &h_);
    if (rc != 0) throw std::runtime_error("dpa_create failed (no CUDA device? there is no CPU fallback): " + std::to_string(rc));
  }
  ~CudaPoseAlignment() {
    if (h_) dpa_destroy(h_);
  }
  CudaPoseAlignment(const CudaPoseAlignment&) = delete;
  CudaPoseAlignment& operator=(const CudaPoseAlignment&) = delete;

  // EigenPoseAlignment::reset (:261-264)
  void reset() {
    prior_rotation_.reset();
    have_reference_ = have_target_ = false;
  }
  // setRotationPrior (:254-258), 3x3 row-major
  void setRotationPrior(const double r_t_r[9]) {
    prior_rotation_.emplace();
    std::memcpy(prior_rotation_->data(), r_t_r, 9 * sizeof(double));
  }
  // pushKnownPose (:268-271)
  void pushKnownPose(long long timestamp, const double t_w_agent[12]) {
    std::vector<double> v(t_w_agent, t_w_agent + 12);
    known_poses_[timestamp] = v;
  }
  // pushFrame(reference keyframe, ..., reference_frame_depth_map, level, model, kFixed)
  int pushReferenceFrame(const AlignmentFrameView& f) {
    reference_ = f;
    const int n = dpa_set_reference_depth_map(h_, f.image_I_dx_dy, f.depth_idepth_sum, f.depth_weight, f.t_world_agent,
                                              f.exposure_time, f.affine_brightness, f.intrinsics, f.width, f.height);
    check(n);
    have_reference_ = true;
    return n;
  }
  // pushFrame(new frame, t_w_t, pyramids, masks, exposure, affine brightness, level, model, kFree)
  void pushTargetFrame(const AlignmentFrameView& f) {
    target_ = f;
    std::memcpy(target_pose_, f.t_world_agent, sizeof(target_pose_));
    target_ab_[0] = f.affine_brightness[0];
    target_ab_[1] = f.affine_brightness[1];
    have_target_ = true;
  }
  // EigenPoseAlignment::solve (:275-329): rmse, or kZeroCost when the pose is known
  double solve(const size_t /*number_of_threads*/) {
    if (!have_reference_ || !have_target_) throw std::runtime_error("CudaPoseAlignment::solve: two frames are needed");
    auto known = known_poses_.find(target_.timestamp);
    if (known != known_poses_.end()) {
      std::memcpy(target_pose_, known->second.data(), sizeof(target_pose_));
      return kZeroCost;
    }
    check(dpa_set_target(h_, target_.image_I_dx_dy, target_.mask, target_.t_world_agent, target_.exposure_time,
                         target_.affine_brightness, target_.intrinsics, target_.width, target_.height));
    dpa_options o;
    o.max_num_iterations = (int32_t)options_.max_iterations;
    o.initial_trust_region_radius = options_.initial_trust_region_radius;
    o.function_tolerance = options_.function_tolerance;
    o.parameter_tolerance = options_.parameter_tolerance;
    o.sigma_huber_loss = options_.sigma_huber_loss;
    o.affine_brightness_regularizer[0] = options_.affine_brightness_regularizer[0];
    o.affine_brightness_regularizer[1] = options_.affine_brightness_regularizer[1];
    o.regularizer_decrease_on_accept = 2.;  // :304-305
    o.regularizer_increase_on_reject = 2.;
    dpa_result r;
    check(dpa_solve(h_, &o, prior_rotation_ ? prior_rotation_->data() : nullptr, &r));
    // covariance_t_t_r_ = pinv(problem.hessian()).topLeftCorner<6, 6>()  (:320-323)
    dense::Mat H(8, 8);
    for (int i = 0; i < 64; ++i) H.a[i] = r.hessian[i];
    const dense::Mat P = dense::sym_pinv(H, -1);
    for (int i = 0; i < 6; ++i)
      for (int j = 0; j < 6; ++j) covariance_[i * 6 + j] = P(i, j);
    std::memcpy(target_pose_, r.T_world_target, sizeof(target_pose_));  // :325
    target_ab_[0] = target_.affine_brightness[0] + r.affine_brightness_eps[0];  // :326
    target_ab_[1] = target_.affine_brightness[1] + r.affine_brightness_eps[1];
    last_ = r;
    return r.rmse;
  }
  // getPose / getAffineBrightness of the target frame after solve()
  const double* targetPose() const { return target_pose_; }
  const double* targetAffineBrightness() const { return target_ab_; }
  const double* tTargetReferenceCovariance() const { return covariance_; }  // 6x6 row-major
  const dpa_result& lastResult() const { return last_; }

  // calculateMeanSquareOpticalFlow (src/tracker/tracker/src/monocular_tracker.cpp:104-133) over the reference landmarks
  // of the last pushed reference frame; the tracker evaluates it with the aligned t_t_r and once more with the rotation
  // set to identity (:474-480).  NaN when no landmark reprojects (0 / 0 in the reference).
  double meanSquareOpticalFlow(const double t_target_reference[12]) const {
    double flow = 0;
    check(dpa_mean_square_optical_flow(h_, t_target_reference, &flow, nullptr));
    return flow;
  }

 private:
  void check(int rc) const {
    if (rc < 0) throw std::runtime_error(std::string("pose alignment: ") + dpa_last_error(h_));
  }
  dpa_handle* h_ = nullptr;
  PoseAlignmentOptions options_;
  std::optional<std::array<double, 9>> prior_rotation_;
  std::map<long long, std::vector<double>> known_poses_;
  AlignmentFrameView reference_{}, target_{};
  bool have_reference_ = false, have_target_ = false;
  double target_pose_[12] = {}, target_ab_[2] = {}, covariance_[36] = {};
  dpa_result last_{};
};

}  // namespace dsopp_b200
