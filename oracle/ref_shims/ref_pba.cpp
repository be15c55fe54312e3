// C entry points around the REFERENCE'S OWN photometric bundle adjustment, compiled from its sources where they lie under
// /root/reference (oracle/build_ref_pba.py) -- test infrastructure only; nothing under dsopp_b200/ links this.
//
// What is the reference's code here (compiled unchanged, never copied):
//   PBA/local_frame.hpp                    LocalFrame, Landmark, ResidualPoint                      
//   features/camera/pixel_map.hpp + .cpp   PixelMap bilinear sampler over {I,dx,dy}                 (a4, a5)
//   energy/projector/camera_reproject.hpp  ArrayReprojector<T, PinholeCamera, SE3>                  (a6-a8)
//   energy/camera_model/...                PinholeCamera, CameraModelBase (ROI / idepth predicates)
//   energy/motion/se3_motion.hpp           motion::SE3 (increments, log transformers)
//   PBA/evaluate_jacobians.hpp             evaluateJacobians<...>                                   (a9)
//   PBA/first_estimate_jacobians.hpp       firstEstimateJacobians_                                  (a10)
//   PBA/hessian_block_evaluation.hpp       PosePose, Schur complement, calculateIdepths             (a11-a13)
//   PBA/eigen_photometric_bundle_adjustment_problem.hpp   the Problem class, priors, energies, marginalisation  (a14-a16, a20)
//   energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp   the LM driver            (a17)
//   energy/normal_linear_system.hpp + .cpp NormalLinearSystem::solve / reduce_system                (a18)
//   energy/problems/src/photometric_bundle_adjustment.cpp:299-406   updatePointStatuses, relinearizeSystem    (a19)
//   sensors/camera_calibration/mask/camera_mask.hpp + .cpp   CameraMask::valid
// What is NOT the reference's: the third-party libraries under it (Eigen, Sophus, oneTBB, glog, OpenCV, Ceres, the
// protobuf-generated headers) are absent from this image and are replaced by the minimal stand-ins in
// oracle/ref_stubs_full/ (each file says so in its first line).  The frames are filled through LocalFrame's public
// members (its "frontend target" constructor, local_frame.hpp:448-471, then `active_landmarks` / `residuals`), because the
// constructor from track::ActiveKeyframe would drag the whole track / storage subsystem in; the Landmark and ResidualPoint
// constructors that run are the reference's.
#include <cstdint>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <set>
#include <vector>

// src/energy/problems/src/photometric_bundle_adjustment.cpp, lines 1-413: the member-function TEMPLATES of
// PhotometricBundleAdjustment (updatePointStatuses :318-406, relinearizeSystem :307-316) without the explicit instantiations
// that follow them (those would instantiate the members that need the track subsystem).  The build recipe hands these lines
// over as a temporary file outside the repository (oracle/build_ref_pba.py).
#include <glog/logging.h>  // the reference file gets LOG transitively from the real third-party headers

#include REF_PBA_PREFIX

#include "energy/camera_model/pinhole/pinhole_camera.hpp"
#include "energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp"
#include "energy/normal_linear_system.hpp"
#include "energy/problems/photometric_bundle_adjustment/evaluate_jacobians.hpp"
#include "energy/problems/photometric_bundle_adjustment/first_estimate_jacobians.hpp"
#include "energy/problems/photometric_bundle_adjustment/hessian_block_evaluation.hpp"
#include "energy/problems/photometric_bundle_adjustment/local_frame.hpp"
// after evaluate_jacobians.hpp, as in eigen_photometric_bundle_adjustment.cpp:13-16
#include "energy/problems/photometric_bundle_adjustment/eigen_photometric_bundle_adjustment_problem.hpp"
#include "features/camera/pixel_map.hpp"
#include "sensors/camera_calibration/mask/camera_mask.hpp"

namespace dsopp {
// common/time/time.hpp:10 declares it; its definition lives in a file of the reference that is not part of this build
std::ostream& operator<<(std::ostream& os, const time&) { return os; }
}  // namespace dsopp

namespace {

using dsopp::Precision;
using Motion = dsopp::energy::motion::SE3<Precision>;
using Model = dsopp::energy::model::PinholeCamera<Precision>;
constexpr int kP = dsopp::Pattern::kSize;
template <int C>
using Grid = dsopp::features::PixelMap<C>;
using Frame = dsopp::energy::problem::LocalFrame<Precision, Motion, Model, kP, dsopp::features::PixelMap, 1>;
using Residual = dsopp::energy::problem::ResidualPoint<Precision, Motion, kP, 1>;
using Status = dsopp::track::PointConnectionStatus;
using System = dsopp::energy::NormalLinearSystem<Precision>;
namespace prob = dsopp::energy::problem;
constexpr int kBlock = Motion::DoF + 2;
constexpr size_t kSensor = 0;

struct Window {
  std::deque<std::unique_ptr<Frame>> frames;
  // what the frames point to (local_frame.hpp:464-470 keeps pointers into the pyramids and references to the masks)
  std::vector<std::unique_ptr<Frame::Pyramids>> pyramids;
  std::vector<std::unique_ptr<dsopp::sensors::calibration::CameraMask>> masks;
  dsopp::energy::NormalLinearSystem<double> system_marginalized;
  Precision energy_marginalized = 0;
  Eigen::Vector2<Precision> affine_reg;
};

Window* W(void* h) { return static_cast<Window*>(h); }

template <bool FEJ, bool JAC, bool HUBER>
void evaluate(Window* w, double sigma) {
  prob::evaluateJacobians<Precision, Motion, Model, kP, dsopp::features::PixelMap, 1, FEJ, true, JAC, true, HUBER>(
      w->frames, static_cast<Precision>(sigma));
}

template <bool FEJ>
void solve_lm(Window* w, const dsopp::energy::levenberg_marquardt_algorithm::Options& opt, double sigma, double fixed_reg,
              double* energy, int32_t* valid, int32_t* converged) {
  // eigen_photometric_bundle_adjustment.cpp:71-83
  prob::PhotometricBundleAdjustmentProblem<Motion, Model, kP, dsopp::features::PixelMap, true, true, FEJ, 1> problem(
      w->frames, kSensor, static_cast<Precision>(sigma), w->system_marginalized.cast<Precision>(), w->energy_marginalized,
      w->affine_reg, static_cast<Precision>(fixed_reg));
  if constexpr (FEJ) prob::firstEstimateJacobians_<Precision, Motion, Model, kP, dsopp::features::PixelMap, 1>(w->frames);
  auto result = dsopp::energy::levenberg_marquardt_algorithm::solve(problem, opt);
  *energy = static_cast<double>(result.energy);
  *valid = result.number_of_valid_residuals;
  *converged = result.converged ? 1 : 0;
}

}  // namespace

// The reference's header names this class, at global scope, as a friend of PhotometricBundleAdjustment
// (photometric_bundle_adjustment.hpp:15,169 -- its own benchmark uses it to reach the protected members).  Here it runs
// the protected, non-virtual updatePointStatuses / relinearizeSystem on the window's frames.  The solver class is
// abstract and its virtual members need the track subsystem, so no object of it is ever constructed: the two functions read
// nothing but `frames_`, and are called on raw storage in which exactly that member has been constructed.
class LinearSystemBenchmarkData {
 public:
  using PBA = prob::PhotometricBundleAdjustment<Precision, Motion, Model, kP, dsopp::features::PixelMap, true, true, true, 1>;
  static void run(Window* w, int what, size_t min_valid, Precision sigma) {
    alignas(PBA) static unsigned char storage[sizeof(PBA)];
    PBA* p = reinterpret_cast<PBA*>(storage);
    auto* frames = new (&p->frames_) std::deque<std::unique_ptr<Frame>>();
    std::swap(*frames, w->frames);
    if (what == 0)
      p->updatePointStatuses(min_valid, sigma);
    else
      p->relinearizeSystem();
    std::swap(*frames, w->frames);
    frames->~deque();
  }
};

extern "C" {

// PhotometricBundleAdjustment::updatePointStatuses (photometric_bundle_adjustment.cpp:318-406) and ::relinearizeSystem
// (:307-316), which EigenPhotometricBundleAdjustment::solve runs after the LM loop (eigen_photometric_bundle_adjustment.cpp:88,99)
void refpba_update_point_statuses(void* h, int min_valid, double sigma) {
  LinearSystemBenchmarkData::run(W(h), 0, static_cast<size_t>(min_valid), static_cast<Precision>(sigma));
}
void refpba_relinearize_system(void* h) { LinearSystemBenchmarkData::run(W(h), 1, 0, 0); }
void refpba_get_relative_baseline(void* h, int f, double* out) {
  const auto& lms = W(h)->frames[static_cast<size_t>(f)]->active_landmarks.at(kSensor);
  for (size_t i = 0; i < lms.size(); ++i) out[i] = static_cast<double>(lms[i].relative_baseline);
}
void refpba_get_linearization_point(void* h, int f, double* T_34, double* ab0) {
  const Frame& fr = *W(h)->frames[static_cast<size_t>(f)];
  const auto m = fr.T_w_agent_linearization_point.matrix3x4();
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j) T_34[4 * i + j] = static_cast<double>(m(i, j));
  ab0[0] = static_cast<double>(fr.affine_brightness0(0));
  ab0[1] = static_cast<double>(fr.affine_brightness0(1));
}

void* refpba_create() {
  auto* w = new Window;
  w->affine_reg.setZero();
  return w;
}
void refpba_destroy(void* h) { delete W(h); }
int refpba_precision_bytes() { return static_cast<int>(sizeof(Precision)); }

// one keyframe: pose of the linearisation point as a 3x4 row-major [R|t], raw intensities (height x width, row-major),
// mask (height x width, 0 = invalid) or null
int refpba_add_frame(void* h, int id, int64_t timestamp_ns, const double* T_w_agent_3x4, double exposure, const double* ab,
                     const double* intr, const double* image, int width, int height, const uint8_t* mask, int fixed,
                     const double* state_eps) {
  Window* w = W(h);
  Eigen::Matrix<Precision, 3, 3> R;
  Eigen::Vector<Precision, 3> t;
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R(i, j) = static_cast<Precision>(T_w_agent_3x4[4 * i + j]);
    t(i) = static_cast<Precision>(T_w_agent_3x4[4 * i + 3]);
  }
  Motion pose(R, t);

  std::vector<Precision, dsopp::PrecisionAllocator> data(static_cast<size_t>(width) * static_cast<size_t>(height));
  for (size_t i = 0; i < data.size(); ++i) data[i] = static_cast<Precision>(image[i]);
  auto pyr = std::make_unique<Frame::Pyramids>();
  (*pyr)[kSensor].emplace_back(std::move(data), static_cast<long>(width), static_cast<long>(height));

  cv::Mat m(height, width, CV_8UC1, 255);
  if (mask)
    for (int y = 0; y < height; ++y)
      for (int x = 0; x < width; ++x) m.at<uchar>(y, x) = mask[static_cast<size_t>(y) * static_cast<size_t>(width) + static_cast<size_t>(x)];
  auto cm = std::make_unique<dsopp::sensors::calibration::CameraMask>(m);
  std::map<size_t, const dsopp::sensors::calibration::CameraMask&> masks;
  masks.insert({kSensor, *cm});

  Eigen::Vector2<Precision> image_size(static_cast<Precision>(width), static_cast<Precision>(height));
  Eigen::Vector<Precision, 4> k(static_cast<Precision>(intr[0]), static_cast<Precision>(intr[1]),
                                static_cast<Precision>(intr[2]), static_cast<Precision>(intr[3]));
  Model model(image_size, k);
  Eigen::Vector2<Precision> affine(static_cast<Precision>(ab[0]), static_cast<Precision>(ab[1]));
  const dsopp::time stamp{std::chrono::duration_cast<dsopp::time::duration>(std::chrono::nanoseconds(timestamp_ns))};

  auto frame = std::make_unique<Frame>(stamp, pose, *pyr, masks, static_cast<Precision>(exposure), affine, false, size_t(0),
                                       model, fixed ? prob::FrameParameterization::kFixed : prob::FrameParameterization::kFree);
  frame->id = id;
  frame->active_landmarks[kSensor];  // sensors() lists the keys of active_landmarks (local_frame.hpp:537-543)
  if (state_eps)
    for (int i = 0; i < kBlock; ++i) frame->state_eps(i) = static_cast<Precision>(state_eps[i]);
  w->frames.push_back(std::move(frame));
  w->pyramids.push_back(std::move(pyr));
  w->masks.push_back(std::move(cm));
  return static_cast<int>(w->frames.size()) - 1;
}

// flags: bit 0 is_marginalized, bit 1 to_marginalize, bit 2 is_outlier
void refpba_add_landmarks(void* h, int f, int n, const double* uv, const double* idepth, const double* patch,
                          const uint8_t* flags) {
  Frame& fr = *W(h)->frames[static_cast<size_t>(f)];
  auto& lms = fr.active_landmarks[kSensor];
  lms.reserve(lms.size() + static_cast<size_t>(n));
  for (int i = 0; i < n; ++i) {
    Eigen::Vector2<Precision> p(static_cast<Precision>(uv[2 * i]), static_cast<Precision>(uv[2 * i + 1]));
    Eigen::Matrix<Precision, kP, 1> pa;
    for (int k = 0; k < kP; ++k) pa(k) = static_cast<Precision>(patch[kP * i + k]);
    const uint8_t fl = flags ? flags[i] : 0;
    lms.emplace_back(Frame::Landmark{p, static_cast<Precision>(idepth[i]), pa, (fl & 1) != 0, (fl & 4) != 0});
    lms.back().to_marginalize = (fl & 2) != 0;
  }
}

// one ResidualPoint per landmark of frame f towards frame t, constructed from its connection status
void refpba_set_statuses(void* h, int f, int t, int n, const uint8_t* status) {
  Window* w = W(h);
  auto& v = w->frames[static_cast<size_t>(f)]->residuals[{kSensor, kSensor}][w->frames[static_cast<size_t>(t)]->id];
  v.clear();
  for (int i = 0; i < n; ++i) v.push_back(Residual(static_cast<Status>(status[i])));
}

void refpba_set_frame_state(void* h, int f, const double* state_eps, const double* state_eps_step) {
  Frame& fr = *W(h)->frames[static_cast<size_t>(f)];
  for (int i = 0; i < kBlock; ++i) {
    if (state_eps) fr.state_eps(i) = static_cast<Precision>(state_eps[i]);
    if (state_eps_step) fr.state_eps_step(i) = static_cast<Precision>(state_eps_step[i]);
  }
}
void refpba_set_frame_flags(void* h, int f, int to_marginalize, int is_marginalized) {
  Frame& fr = *W(h)->frames[static_cast<size_t>(f)];
  fr.to_marginalize = to_marginalize != 0;
  fr.is_marginalized = is_marginalized != 0;
}
void refpba_set_idepth_steps(void* h, int f, const double* step) {
  auto& lms = W(h)->frames[static_cast<size_t>(f)]->active_landmarks[kSensor];
  for (size_t i = 0; i < lms.size(); ++i) lms[i].idepth_step = static_cast<Precision>(step[i]);
}
void refpba_set_marginalized(void* h, int size, const double* H, const double* b, double energy) {
  Window* w = W(h);
  w->system_marginalized = dsopp::energy::NormalLinearSystem<double>(size);
  for (int i = 0; i < size; ++i) {
    for (int j = 0; j < size; ++j) w->system_marginalized.H(i, j) = H ? H[i * size + j] : 0.0;
    w->system_marginalized.b(i) = b ? b[i] : 0.0;
  }
  w->energy_marginalized = static_cast<Precision>(energy);
}
void refpba_get_marginalized(void* h, double* H, double* b, double* energy, int* size) {
  Window* w = W(h);
  const int n = static_cast<int>(w->system_marginalized.b.size());
  *size = n;
  if (H)
    for (int i = 0; i < n; ++i)
      for (int j = 0; j < n; ++j) H[i * n + j] = w->system_marginalized.H(i, j);
  if (b)
    for (int i = 0; i < n; ++i) b[i] = w->system_marginalized.b(i);
  *energy = static_cast<double>(w->energy_marginalized);
}

int refpba_n_frames(void* h) { return static_cast<int>(W(h)->frames.size()); }
int refpba_n_landmarks(void* h, int f) {
  return static_cast<int>(W(h)->frames[static_cast<size_t>(f)]->active_landmarks[kSensor].size());
}

// firstEstimateJacobians_ (PBA/first_estimate_jacobians.hpp:17-70)
void refpba_first_estimate(void* h) {
  prob::firstEstimateJacobians_<Precision, Motion, Model, kP, dsopp::features::PixelMap, 1>(W(h)->frames);
}

// evaluateJacobians<..., FEJ, OPTIMIZE_IDEPTHS = true, EVALUATE_JACOBIANS, NEW_EVALUATION_POINT = true, HUBER>
// (PBA/evaluate_jacobians.hpp:20-202) -- the three instantiations the Eigen backend makes (problem.hpp:291,323,232)
void refpba_evaluate(void* h, int fej, int jacobians, int huber, double sigma) {
  Window* w = W(h);
  const int key = (fej ? 4 : 0) | (jacobians ? 2 : 0) | (huber ? 1 : 0);
  switch (key) {
    case 0: evaluate<false, false, false>(w, 0); break;
    case 1: evaluate<false, false, true>(w, sigma); break;
    case 2: evaluate<false, true, false>(w, 0); break;
    case 3: evaluate<false, true, true>(w, sigma); break;
    case 4: evaluate<true, false, false>(w, 0); break;
    case 5: evaluate<true, false, true>(w, sigma); break;
    case 6: evaluate<true, true, false>(w, 0); break;
    default: evaluate<true, true, true>(w, sigma); break;
  }
}
void refpba_change_statuses(void* h, int accept) { prob::changeResidualStatuses(W(h)->frames, accept != 0); }

// every field of the ResidualPoints of (f -> t); any output pointer may be null
void refpba_get_residuals(void* h, int f, int t, uint8_t* status, uint8_t* cand, double* residuals, double* du_idepth,
                          double* dv_idepth, double* du_t, double* dv_t, uint8_t* jac_valid, double* J_ref, double* J_tgt,
                          double* d_idepth, double* huber_weight, double* energy, double* bcs) {
  Window* w = W(h);
  const auto& v = w->frames[static_cast<size_t>(f)]->residuals.at({kSensor, kSensor}).at(w->frames[static_cast<size_t>(t)]->id);
  for (size_t i = 0; i < v.size(); ++i) {
    const Residual& r = v[i];
    if (status) status[i] = static_cast<uint8_t>(r.connection_status);
    if (cand) cand[i] = static_cast<uint8_t>(r.connection_status_candidate);
    if (jac_valid) jac_valid[i] = r.reprojection_jacobians_valid ? 1 : 0;
    if (huber_weight) huber_weight[i] = static_cast<double>(r.huber_weight);
    if (energy) energy[i] = static_cast<double>(r.energy);
    if (bcs) bcs[i] = static_cast<double>(r.brightness_change_scale);
    for (int p = 0; p < kP; ++p) {
      if (residuals) residuals[kP * i + p] = static_cast<double>(r.residuals(p));
      if (du_idepth) du_idepth[kP * i + p] = static_cast<double>(r.d_u_idepth(p));
      if (dv_idepth) dv_idepth[kP * i + p] = static_cast<double>(r.d_v_idepth(p));
      if (d_idepth) d_idepth[kP * i + p] = static_cast<double>(r.d_idepth(p));
      for (int k = 0; k < 6; ++k) {
        if (du_t) du_t[(kP * i + p) * 6 + k] = static_cast<double>(r.d_u_tReferenceTarget(p, k));
        if (dv_t) dv_t[(kP * i + p) * 6 + k] = static_cast<double>(r.d_v_tReferenceTarget(p, k));
      }
      for (int k = 0; k < kBlock; ++k) {
        if (J_ref) J_ref[(kP * i + p) * kBlock + k] = static_cast<double>(r.d_reference_state_eps(p, k));
        if (J_tgt) J_tgt[(kP * i + p) * kBlock + k] = static_cast<double>(r.d_target_state_eps(p, k));
      }
    }
  }
}

// per-landmark state of frame f; hpd is (n, 8 * n_frames) row-major (zero rows while the reference's vector is empty)
void refpba_get_landmarks(void* h, int f, double* idepth, double* idepth_step, double* inv_hdd, double* b_d, double* hpd,
                          uint8_t* ill, uint8_t* flags, double* ref_pattern, double* corrected, int64_t* n_inliers) {
  Window* w = W(h);
  const auto& lms = w->frames[static_cast<size_t>(f)]->active_landmarks.at(kSensor);
  const long D = kBlock * static_cast<long>(w->frames.size());
  for (size_t i = 0; i < lms.size(); ++i) {
    const auto& l = lms[i];
    if (idepth) idepth[i] = static_cast<double>(l.idepth);
    if (idepth_step) idepth_step[i] = static_cast<double>(l.idepth_step);
    if (inv_hdd) inv_hdd[i] = static_cast<double>(l.inv_hessian_idepth_idepth);
    if (b_d) b_d[i] = static_cast<double>(l.b_idepth_block);
    if (ill) ill[i] = l.ill_conditioned ? 1 : 0;
    if (flags) flags[i] = static_cast<uint8_t>((l.is_marginalized ? 1 : 0) | (l.to_marginalize ? 2 : 0) | (l.is_outlier ? 4 : 0));
    if (n_inliers) n_inliers[i] = static_cast<int64_t>(l.number_of_inlier_residuals);
    if (hpd)
      for (long k = 0; k < D; ++k)
        hpd[static_cast<long>(i) * D + k] =
            k < l.hessian_poses_idepth_block.size() ? static_cast<double>(l.hessian_poses_idepth_block(k)) : 0.0;
    for (int p = 0; p < kP; ++p) {
      if (ref_pattern) {
        ref_pattern[(kP * i + p) * 2 + 0] = static_cast<double>(l.reference_pattern(0, p));
        ref_pattern[(kP * i + p) * 2 + 1] = static_cast<double>(l.reference_pattern(1, p));
      }
      if (corrected) corrected[kP * i + p] = static_cast<double>(l.corrected_intensities(p));
    }
  }
}
void refpba_get_frame_state(void* h, int f, double* state_eps, double* state_eps_step, double* T_w_agent_3x4) {
  const Frame& fr = *W(h)->frames[static_cast<size_t>(f)];
  for (int i = 0; i < kBlock; ++i) {
    if (state_eps) state_eps[i] = static_cast<double>(fr.state_eps(i));
    if (state_eps_step) state_eps_step[i] = static_cast<double>(fr.state_eps_step(i));
  }
  if (T_w_agent_3x4) {
    const auto m = fr.tWorldAgent().matrix3x4();  // local_frame.hpp:525-527
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) T_w_agent_3x4[4 * i + j] = static_cast<double>(m(i, j));
  }
}

// evaluateLinearSystemPosePose + evaluateLinearSystemPoseDepthSchurComplement (PBA/hessian_block_evaluation.hpp:83-235),
// optionally the priors (problem.hpp:35-82) added to the pose system as linearize() does (problem.hpp:327-331)
void refpba_linear_systems(void* h, int for_marginalized, int with_prior, const double* affine_reg, double fixed_reg,
                           double* H_pose, double* b_pose, double* H_schur, double* b_schur) {
  Window* w = W(h);
  const int n = kBlock * static_cast<int>(w->frames.size());
  System pose(n), schur(n);
  pose.setZero();
  schur.setZero();
  if (for_marginalized) {
    prob::evaluateLinearSystemPosePose<true>(w->frames, kSensor, pose);
    prob::evaluateLinearSystemPoseDepthSchurComplement<true>(w->frames, kSensor, schur);
  } else {
    prob::evaluateLinearSystemPosePose(w->frames, kSensor, pose);
    prob::evaluateLinearSystemPoseDepthSchurComplement(w->frames, kSensor, schur);
  }
  if (with_prior) {
    Eigen::Vector2<Precision> reg(static_cast<Precision>(affine_reg[0]), static_cast<Precision>(affine_reg[1]));
    prob::evaluateLinearSystemPrior(w->frames, pose, reg, static_cast<Precision>(fixed_reg), for_marginalized != 0);
  }
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) {
      H_pose[i * n + j] = static_cast<double>(pose.H(i, j));
      H_schur[i * n + j] = static_cast<double>(schur.H(i, j));
    }
    b_pose[i] = static_cast<double>(pose.b(i));
    b_schur[i] = static_cast<double>(schur.b(i));
  }
}

// calculateIdepths (PBA/hessian_block_evaluation.hpp:237-262) for a given pose step
void refpba_calculate_idepths(void* h, const double* step_poses, double lambda) {
  Window* w = W(h);
  const int n = kBlock * static_cast<int>(w->frames.size());
  Eigen::VectorX<Precision> step(n);
  for (int i = 0; i < n; ++i) step(i) = static_cast<Precision>(step_poses[i]);
  prob::calculateIdepths(w->frames, kSensor, step, static_cast<Precision>(lambda));
}

// calculateLandmarksEnergy (problem.hpp:100-145)
void refpba_landmarks_energy(void* h, int for_marginalized, double* energy, int32_t* n_valid) {
  Window* w = W(h);
  const auto r = for_marginalized ? prob::calculateLandmarksEnergy<true>(w->frames, kSensor)
                                  : prob::calculateLandmarksEnergy(w->frames, kSensor);
  *energy = static_cast<double>(r.first);
  *n_valid = r.second;
}

// NormalLinearSystem<>::solve (normal_linear_system.cpp:52-60)
void refpba_normal_solve(int n, const double* H, const double* b, double* x) {
  System s(n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < n; ++j) s.H(i, j) = static_cast<Precision>(H[i * n + j]);
    s.b(i) = static_cast<Precision>(b[i]);
  }
  const Eigen::VectorX<Precision> r = s.solve();
  for (int i = 0; i < n; ++i) x[i] = static_cast<double>(r(i));
}

// the whole solve as EigenPhotometricBundleAdjustment::solve sets it up (eigen_photometric_bundle_adjustment.cpp:56-84):
// Problem, firstEstimateJacobians, levenberg_marquardt_algorithm::solve
void refpba_solve(void* h, int fej, int max_iterations, double initial_trust_region_radius, double function_tolerance,
                  double parameter_tolerance, int force_accept, double sigma_huber, const double* affine_reg,
                  double fixed_reg, double* energy, int32_t* n_valid, int32_t* converged) {
  Window* w = W(h);
  namespace lm = dsopp::energy::levenberg_marquardt_algorithm;
  lm::Options options;
  options.initial_levenberg_marquardt_regularizer = static_cast<Precision>(1.0 / initial_trust_region_radius);
  options.function_tolerance = static_cast<Precision>(function_tolerance);
  options.parameter_tolerance = static_cast<Precision>(parameter_tolerance);
  options.max_num_iterations = static_cast<size_t>(max_iterations);
  options.min_num_iterations = 3;
  options.force_accept = force_accept != 0;
  options.levenberg_marquardt_regularizer_decrease_on_accept = 1.;
  options.levenberg_marquardt_regularizer_increase_on_reject = 1.;
  w->affine_reg = Eigen::Vector2<Precision>(static_cast<Precision>(affine_reg[0]), static_cast<Precision>(affine_reg[1]));
  if (w->system_marginalized.b.size() != kBlock * static_cast<long>(w->frames.size())) {
    w->system_marginalized = dsopp::energy::NormalLinearSystem<double>(kBlock * static_cast<int>(w->frames.size()));
    w->system_marginalized.setZero();
  }
  if (fej)
    solve_lm<true>(w, options, sigma_huber, fixed_reg, energy, n_valid, converged);
  else
    solve_lm<false>(w, options, sigma_huber, fixed_reg, energy, n_valid, converged);
}

// the frames_.size() > 1 part of EigenPhotometricBundleAdjustment::pushFrame (eigen_photometric_bundle_adjustment.cpp:
// 121-130): FEJ, linearise with Huber, commit the statuses, updateMarginalizedLinearSystem (problem.hpp:146-203, which
// also erases the frames flagged to_marginalize from the deque)
void refpba_marginalize(void* h, int fej, double sigma_huber, const double* affine_reg, double fixed_reg) {
  Window* w = W(h);
  const int n = kBlock * static_cast<int>(w->frames.size());
  if (w->system_marginalized.b.size() != n) {
    w->system_marginalized = dsopp::energy::NormalLinearSystem<double>(n);
    w->system_marginalized.setZero();
  }
  if (fej) {
    prob::firstEstimateJacobians_<Precision, Motion, Model, kP, dsopp::features::PixelMap, 1>(w->frames);
    evaluate<true, true, true>(w, sigma_huber);
  } else {
    evaluate<false, true, true>(w, sigma_huber);
  }
  prob::changeResidualStatuses(w->frames);
  Eigen::Vector2<Precision> reg(static_cast<Precision>(affine_reg[0]), static_cast<Precision>(affine_reg[1]));
  prob::updateMarginalizedLinearSystem(w->frames, kSensor, w->system_marginalized, w->energy_marginalized, reg,
                                       static_cast<Precision>(fixed_reg));
}

// PatternPatch::getIntensities (features/camera/pattern_patch.hpp:52-64) on frame f's image: the landmark patch a tracker
// samples at activation
void refpba_get_intensities(void* h, int f, int n, const double* uv, double* patch) {
  Window* w = W(h);
  const auto& map = (*w->pyramids[static_cast<size_t>(f)])[kSensor][0];
  for (int i = 0; i < n; ++i) {
    Eigen::Vector2<Precision> p(static_cast<Precision>(uv[2 * i]), static_cast<Precision>(uv[2 * i + 1]));
    Eigen::Matrix<Precision, kP, 1, dsopp::PatchStorageOrder<1>> out;
    dsopp::features::PatternPatch::getIntensities(p, map, out);
    for (int k = 0; k < kP; ++k) patch[kP * i + k] = static_cast<double>(out(k));
  }
}

}  // extern "C"
