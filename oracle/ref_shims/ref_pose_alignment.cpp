// C entry points around the REFERENCE'S OWN coarse-tracker image alignment, compiled from its source where
// it lies (oracle/build_ref_pba.py) -- test infrastructure only.
//
// The algorithm is the class PoseAlignerProblem, which src/energy/problems/src/eigen_pose_alignment.cpp defines in an anonymous
// namespace (lines 24-242): calculateEnergy, linearize, calculateStep, acceptStep / rejectStep.  The rest of that file --
// the members of EigenPoseAlignment -- derives from PhotometricBundleAdjustment, whose definitions need the track / storage
// subsystem.  The build recipe therefore hands the compiler the file's own lines up to the end of that anonymous namespace
// (a temporary file outside the repository, made at build time and removed again; macro REF_POSE_ALIGNMENT_PREFIX names
// it), and this shim performs the ~20 lines of set-up that EigenPoseAlignment::solve (lines 275-329) does around the
// class: options, t_t_r from the two linearisation points, levenberg_marquardt_algorithm::solve, the final pose.  The
// reference frame's landmarks come from the reference's own depth-map LocalFrame constructor (local_frame.hpp:350-393).
#include <cstdint>
#include <deque>
#include <map>
#include <memory>
#include <vector>

#include REF_POSE_ALIGNMENT_PREFIX

namespace {
using dsopp::Precision;
using Motion = dsopp::energy::motion::SE3<Precision>;
using Model = dsopp::energy::model::PinholeCamera<Precision>;
using Frame = dsopp::energy::problem::LocalFrame<Precision, Motion, Model, 1, dsopp::features::PixelMap, 1>;
using Problem = dsopp::energy::problem::PoseAlignerProblem<Motion, Model, 1, dsopp::features::PixelMap, 1, true>;
namespace prob = dsopp::energy::problem;
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
Frame::Pyramids pyramid_from(const double* image, int width, int height) {
  std::vector<Precision, dsopp::PrecisionAllocator> data(static_cast<size_t>(width) * static_cast<size_t>(height));
  for (size_t i = 0; i < data.size(); ++i) data[i] = static_cast<Precision>(image[i]);
  Frame::Pyramids pyr;
  pyr[kSensor].emplace_back(std::move(data), static_cast<long>(width), static_cast<long>(height));
  return pyr;
}
}  // namespace

extern "C" {

// Aligns the target frame to the reference frame's depth map.  Images are raw intensities (height x width); the depth map
// is given as (idepth sum, weight) rasters of the same size; mask may be null.  Outputs: T_t_r (3x4), affine brightness
// increment (2), the 8x8 Hessian of the last linearisation, the landmarks the constructor made (count, and -- if the
// pointers are not null and capacity allows -- their uv / idepth / patch), and the LM result.  Returns the energy.
double refpa_solve(const double* ref_T34, double ref_exposure, const double* ref_ab, const double* ref_image,
                   const double* tgt_T34, double tgt_exposure, const double* tgt_ab, const double* tgt_image,
                   const uint8_t* tgt_mask, const double* intr, int width, int height, const double* idepth_sum,
                   const double* weight, double sigma_huber, const double* affine_reg, int max_iterations,
                   double initial_trust_region_radius, double function_tolerance, double parameter_tolerance,
                   const double* prior_rotation_3x3, double* T_t_r_34, double* ab_eps, double* H88, int32_t* n_valid,
                   int32_t* converged, int32_t* n_landmarks, int capacity, double* lm_uv, double* lm_idepth,
                   double* lm_patch) {
  Eigen::Vector2<Precision> image_size(static_cast<Precision>(width), static_cast<Precision>(height));
  Eigen::Vector<Precision, 4> k(static_cast<Precision>(intr[0]), static_cast<Precision>(intr[1]),
                                static_cast<Precision>(intr[2]), static_cast<Precision>(intr[3]));
  Model model(image_size, k);

  cv::Mat m(height, width, CV_8UC1, 255);
  if (tgt_mask)
    for (int y = 0; y < height; ++y)
      for (int x = 0; x < width; ++x)
        m.at<uchar>(y, x) = tgt_mask[static_cast<size_t>(y) * static_cast<size_t>(width) + static_cast<size_t>(x)];
  dsopp::sensors::calibration::CameraMask mask(m), ref_mask(height, width);
  std::map<size_t, const dsopp::sensors::calibration::CameraMask&> masks_t, masks_r;
  masks_t.insert({kSensor, mask});
  masks_r.insert({kSensor, ref_mask});

  Frame::Pyramids pyr_r = pyramid_from(ref_image, width, height), pyr_t = pyramid_from(tgt_image, width, height);
  // DepthMap::map is indexed (x, y) (local_frame.hpp:372-377: width = rows())
  std::map<size_t, std::vector<prob::DepthMap>> depth_maps;
  depth_maps[kSensor].emplace_back(static_cast<long>(width), static_cast<long>(height));
  auto& dm = depth_maps[kSensor][0].map;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const size_t i = static_cast<size_t>(y) * static_cast<size_t>(width) + static_cast<size_t>(x);
      dm(x, y).idepth = static_cast<Precision>(idepth_sum[i]);
      dm(x, y).weight = static_cast<Precision>(weight[i]);
    }

  const dsopp::time t0{}, t1{std::chrono::duration_cast<dsopp::time::duration>(std::chrono::milliseconds(50))};
  Eigen::Vector2<Precision> rab(static_cast<Precision>(ref_ab[0]), static_cast<Precision>(ref_ab[1]));
  Eigen::Vector2<Precision> tab(static_cast<Precision>(tgt_ab[0]), static_cast<Precision>(tgt_ab[1]));
  // the tracker's two frames (monocular_tracker.cpp:199-214 through PhotometricBundleAdjustment::pushFrame overloads)
  Frame reference(t0, pose_from(ref_T34), pyr_r, masks_r, depth_maps, static_cast<Precision>(ref_exposure), rab, size_t(0),
                  model, prob::FrameParameterization::kFixed);
  Frame target(t1, pose_from(tgt_T34), pyr_t, masks_t, static_cast<Precision>(tgt_exposure), tab, false, size_t(0), model,
               prob::FrameParameterization::kFree);

  const auto& lms = reference.active_landmarks.at(kSensor);
  *n_landmarks = static_cast<int32_t>(lms.size());
  for (size_t i = 0; i < lms.size() && static_cast<int>(i) < capacity; ++i) {
    if (lm_uv) lm_uv[2 * i] = static_cast<double>(lms[i].projection(0)), lm_uv[2 * i + 1] = static_cast<double>(lms[i].projection(1));
    if (lm_idepth) lm_idepth[i] = static_cast<double>(lms[i].idepth);
    if (lm_patch) lm_patch[i] = static_cast<double>(lms[i].patch(0));
  }

  // EigenPoseAlignment::solve, eigen_pose_alignment.cpp:296-326
  namespace lm = dsopp::energy::levenberg_marquardt_algorithm;
  lm::Options options;
  options.initial_levenberg_marquardt_regularizer = static_cast<Precision>(1. / initial_trust_region_radius);
  options.function_tolerance = static_cast<Precision>(function_tolerance);
  options.parameter_tolerance = static_cast<Precision>(parameter_tolerance);
  options.max_num_iterations = static_cast<size_t>(max_iterations);
  options.levenberg_marquardt_regularizer_decrease_on_accept = 2.;
  options.levenberg_marquardt_regularizer_increase_on_reject = 2.;

  Motion::Product t_t_r = target.T_w_agent_linearization_point.inverse() * reference.T_w_agent_linearization_point;
  if (prior_rotation_3x3) {
    Eigen::Matrix<Precision, 3, 3> R;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) R(i, j) = static_cast<Precision>(prior_rotation_3x3[3 * i + j]);
    t_t_r.setRotationMatrix(R);
  }
  Eigen::Vector2<Precision> affine_brightness_eps = Eigen::Vector2<Precision>::Zero();
  const Eigen::Vector2<Precision> reg(static_cast<Precision>(affine_reg[0]), static_cast<Precision>(affine_reg[1]));
  Problem problem(reference, target, kSensor, target.masks.at(kSensor), static_cast<Precision>(sigma_huber), reg, t_t_r,
                  affine_brightness_eps);
  auto result = lm::solve(problem, options);

  const auto T = t_t_r.matrix3x4();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) T_t_r_34[4 * i + j] = static_cast<double>(T(i, j));
  ab_eps[0] = static_cast<double>(affine_brightness_eps(0));
  ab_eps[1] = static_cast<double>(affine_brightness_eps(1));
  const auto H = problem.hessian();
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) H88[8 * i + j] = static_cast<double>(H(i, j));
  *n_valid = result.number_of_valid_residuals;
  *converged = result.converged ? 1 : 0;
  return static_cast<double>(result.energy);
}

}  // extern "C"
