// CudaPhotometricBundleAdjustment: drop-in for EigenPhotometricBundleAdjustment
// (src/energy/problems/include/energy/problems/photometric_bundle_adjustment/eigen_photometric_bundle_adjustment.hpp:23-76,
//  src/energy/problems/src/eigen_photometric_bundle_adjustment.cpp:47-141) with the same constructor arguments and
// the four calls the tracker makes: pushFrame, updateLocalFrame, solve(n_threads), updateFrame
// (src/tracker/tracker/src/monocular_tracker.cpp:251-262,462-507).
//
// The reference types (track::ActiveKeyframe, Sophus, Eigen) cannot be compiled in this environment, so the
// keyframe crosses this interface as the plain struct KeyframeView; INTEGRATION.md shows the ~40-line subclass of
// PhotometricBundleAdjustment<...> that fills it from an ActiveKeyframe inside the reference tree.
#pragma once
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

#include "cuda_pba_problem.hpp"
#include "lm_driver.hpp"

namespace dsopp_b200 {

// TrustRegionPhotometricBundleAdjustmentOptions<Precision>
// (…/photometric_bundle_adjustment/trust_region_photometric_bundle_adjustment_options.hpp:13-51)
struct TrustRegionPhotometricBundleAdjustmentOptions {
  size_t max_iterations = 7;                               // fabric.cpp:82-99
  Precision initial_trust_region_radius = 1e5;
  Precision function_tolerance = 1e-8;
  Precision parameter_tolerance = 1e-8;
  double affine_brightness_regularizer[2] = {1e12, 1e8};  // fabric.cpp:68
  Precision fixed_state_regularizer = 1e16;               // fabric.cpp:69
  Precision sigma_huber_loss = 20;                        // monocular_tracker.hpp:55
  size_t min_iterations = 3;                              // eigen_photometric_bundle_adjustment.cpp:71
};

enum class FrameParameterization { kFree, kFixed };

// what LocalFrame copies out of track::ActiveKeyframe (local_frame.hpp:309-335) plus its connections
struct KeyframeView {
  int keyframe_id = 0;
  long long timestamp = 0;
  double t_world_agent[12];            // 3x4 row-major
  double exposure_time = 1;
  double affine_brightness[2] = {0, 0};
  double intrinsics[4];                // fx fy cx cy of `model` at `level`
  const float* image_I_dx_dy = nullptr;  // PixelMap<1> storage of pyramid level `level`
  const uint8_t* mask = nullptr;
  bool is_marginalized = false;
  int n_landmarks = 0;
  const float* projections = nullptr;  // [n][2]
  const float* idepths = nullptr;      // [n]
  const float* patches = nullptr;      // [n][8]
  const uint8_t* landmark_flags = nullptr;  // DPBA_LM_MARGINALIZED | DPBA_LM_OUTLIER, [n]
  // connection statuses keyed by the other keyframe's id:
  //   statuses_as_reference[id][l] : landmark l of THIS frame reprojected into frame `id`
  //   statuses_as_target[id][l]    : landmark l of frame `id` reprojected into THIS frame
  std::map<int, std::vector<uint8_t>> statuses_as_reference, statuses_as_target;
};

struct LandmarkResult {  // what updateFrame writes back (photometric_bundle_adjustment.cpp:223-262)
  std::vector<float> idepth, idepth_variance, relative_baseline;
  std::vector<uint8_t> is_outlier;
  std::vector<uint32_t> number_of_inlier_residuals;
};

class CudaPhotometricBundleAdjustment {
 public:
  CudaPhotometricBundleAdjustment(const TrustRegionPhotometricBundleAdjustmentOptions& trust_region_options,
                                  bool estimate_uncertainty, bool force_accept, int width, int height,
                                  int max_frames = 9, int max_points_per_frame = 4096, int device = 0,
                                  bool device_resident_lm = true)
      : options_(trust_region_options),
        estimate_uncertainty_(estimate_uncertainty),
        force_accept_(force_accept),
        device_resident_lm_(device_resident_lm) {
    dpba_config cfg{max_frames, max_points_per_frame, width, height, device, 0, 1};
    const int rc = dpba_create(&cfg, &h_);
    if (rc != 0) throw DpbaFailure(rc, "dpba_create failed (no CUDA device? there is no CPU fallback)");
  }
  ~CudaPhotometricBundleAdjustment() {
    if (h_) dpba_destroy(h_);
  }
  CudaPhotometricBundleAdjustment(const CudaPhotometricBundleAdjustment&) = delete;
  CudaPhotometricBundleAdjustment& operator=(const CudaPhotometricBundleAdjustment&) = delete;

  dpba_handle* handle() { return h_; }
  const std::vector<FrameMeta>& frames() const { return frames_; }
  size_t lastIterations() const { return last_iterations_; }
  const NormalLinearSystem& systemMarginalized() const { return system_marginalized_; }
  Precision energyMarginalized() const { return energy_marginalized_; }
  const std::map<std::pair<int, int>, std::vector<double>>& covariances() const { return covariance_matrices_; }

  // EigenPhotometricBundleAdjustment::pushFrame (eigen_photometric_bundle_adjustment.cpp:116-141)
  void pushFrame(const KeyframeView& frame, size_t /*level*/, FrameParameterization frame_parameterization) {
    if (frames_.size() > 1) marginalize();
    if (!frames_.empty() && !(frames_.back().timestamp < frame.timestamp))
      throw DpbaFailure(DPBA_E_INVALID, "Frames must be processed in ascending order of time");
    const int slot = dpba_push_frame(h_, frame.keyframe_id, frame.image_I_dx_dy, frame.mask, frame.t_world_agent,
                                     frame.exposure_time, frame.affine_brightness, frame.intrinsics,
                                     frame_parameterization == FrameParameterization::kFixed);
    dpba_check(h_, slot);
    FrameMeta m;
    m.id = frame.keyframe_id;
    m.timestamp = frame.timestamp;
    m.fixed = frame_parameterization == FrameParameterization::kFixed;
    m.is_marginalized = frame.is_marginalized;
    std::memcpy(m.T_w_agent_linearization_point, frame.t_world_agent, sizeof(m.T_w_agent_linearization_point));
    m.affine_brightness0[0] = frame.affine_brightness[0];
    m.affine_brightness0[1] = frame.affine_brightness[1];
    m.exposure_time = frame.exposure_time;
    frames_.push_back(m);
    dpba_check(h_, dpba_set_frame_marginalized(h_, slot, m.is_marginalized));
    dpba_check(h_, dpba_set_landmarks(h_, slot, frame.n_landmarks, frame.projections, frame.idepths, frame.patches,
                                      frame.landmark_flags));
    uploadStatuses(slot, frame);
    // system_marginalized_ grows by one zero block (eigen_photometric_bundle_adjustment.cpp:134-140)
    system_marginalized_.resize(kBlockSize * (int)frames_.size());
  }

  // EigenPhotometricBundleAdjustment::updateLocalFrame (eigen_photometric_bundle_adjustment.cpp:103-113) +
  // LocalFrame::update (local_frame.hpp:484-521): flags of known landmarks, freshly matured landmarks, statuses
  void updateLocalFrame(const KeyframeView& frame) {
    const int slot = slotOf(frame.timestamp);
    if (slot < 0) throw DpbaFailure(DPBA_E_INVALID, "Cannot update frame, there is no local copy in the solver");
    const int old_n = dpba_num_landmarks(h_, slot);
    std::vector<uint8_t> flags(old_n);
    if (old_n) {
      dpba_check(h_, dpba_get_landmarks(h_, slot, old_n, nullptr, nullptr, nullptr, nullptr, flags.data(), nullptr, nullptr));
      for (int l = 0; l < old_n; ++l) {
        const bool was_marg = flags[l] & DPBA_LM_MARGINALIZED;
        const bool now_marg = frame.landmark_flags && (frame.landmark_flags[l] & DPBA_LM_MARGINALIZED);
        const bool now_outlier = frame.landmark_flags && (frame.landmark_flags[l] & DPBA_LM_OUTLIER);
        uint8_t f = flags[l] & ~(DPBA_LM_MARGINALIZED | DPBA_LM_TO_MARGINALIZE);
        if (!was_marg && now_marg && !now_outlier) f |= DPBA_LM_TO_MARGINALIZE;  // local_frame.hpp:493-495
        if (now_marg) f |= DPBA_LM_MARGINALIZED;
        flags[l] = f;
      }
      dpba_check(h_, dpba_set_landmark_flags(h_, slot, old_n, flags.data()));
    }
    if (frame.n_landmarks > old_n) {
      const int k = frame.n_landmarks - old_n;
      dpba_check(h_, dpba_append_landmarks(h_, slot, k, frame.projections + 2 * old_n, frame.idepths + old_n,
                                           frame.patches + 8 * old_n,
                                           frame.landmark_flags ? frame.landmark_flags + old_n : nullptr));
    }
    appendStatuses(slot, frame, old_n);
    FrameMeta& m = frames_[slot];
    m.to_marginalize = frame.is_marginalized && !m.is_marginalized;
    m.is_marginalized = frame.is_marginalized;
    dpba_check(h_, dpba_set_frame_flags(h_, slot, m.fixed, m.to_marginalize));
    dpba_check(h_, dpba_set_frame_marginalized(h_, slot, m.is_marginalized));
  }

  // EigenPhotometricBundleAdjustment::solve (eigen_photometric_bundle_adjustment.cpp:59-101)
  Precision solve(const size_t /*number_of_threads*/) {
    namespace lm = levenberg_marquardt_algorithm;
    lm::Options o;
    o.initial_levenberg_marquardt_regularizer = 1.0 / options_.initial_trust_region_radius;
    o.function_tolerance = options_.function_tolerance;
    o.parameter_tolerance = options_.parameter_tolerance;
    o.max_num_iterations = options_.max_iterations;
    o.min_num_iterations = options_.min_iterations;
    o.force_accept = force_accept_;
    o.levenberg_marquardt_regularizer_decrease_on_accept = 1.;
    o.levenberg_marquardt_regularizer_increase_on_reject = 1.;
    dpba_check(h_, dpba_first_estimate(h_));
    lm::Result result;
    if (device_resident_lm_) {
      // the same loop and problem methods, executed on the device with one host synchronisation (dpba_solve_lm)
      dpba_lm_options d;
      d.max_num_iterations = (int32_t)o.max_num_iterations;
      d.min_num_iterations = (int32_t)o.min_num_iterations;
      d.force_accept = o.force_accept;
      d.first_estimate_jacobians = 1;
      d.initial_levenberg_marquardt_regularizer = o.initial_levenberg_marquardt_regularizer;
      d.function_tolerance = o.function_tolerance;
      d.parameter_tolerance = o.parameter_tolerance;
      d.levenberg_marquardt_regularizer_decrease_on_accept = o.levenberg_marquardt_regularizer_decrease_on_accept;
      d.levenberg_marquardt_regularizer_increase_on_reject = o.levenberg_marquardt_regularizer_increase_on_reject;
      d.sigma_huber_loss = options_.sigma_huber_loss;
      d.affine_brightness_regularizer[0] = options_.affine_brightness_regularizer[0];
      d.affine_brightness_regularizer[1] = options_.affine_brightness_regularizer[1];
      d.fixed_state_regularizer = options_.fixed_state_regularizer;
      if (system_marginalized_.size() != kBlockSize * (int)frames_.size())
        system_marginalized_.resize(kBlockSize * (int)frames_.size());
      dpba_lm_result r;
      dpba_check(h_, dpba_solve_lm(h_, &d, system_marginalized_.H.a.data(), system_marginalized_.b.data(),
                                   energy_marginalized_, &r));
      result.energy = r.energy;
      result.number_of_valid_residuals = r.number_of_valid_residuals;
      result.converged = r.converged != 0;
      result.iterations = (size_t)r.iterations;
    } else {
      CudaPhotometricBundleAdjustmentProblem problem(h_, frames_, options_.sigma_huber_loss, system_marginalized_,
                                                     energy_marginalized_, options_.affine_brightness_regularizer,
                                                     options_.fixed_state_regularizer, true);
      result = lm::solve(problem, o);
    }
    last_iterations_ = result.iterations;
    relinearizeSystem();
    if (estimate_uncertainty_) {
      dpba_check(h_, dpba_first_estimate(h_));
      covarianceMatricesOfRelativePoses(covarianceMatrixPosePose());
    }
    double thr = 0;
    dpba_check(h_, dpba_update_point_statuses(h_, 1, options_.sigma_huber_loss, &thr));
    return result.energy;
  }

  // PhotometricBundleAdjustment::updateFrame (photometric_bundle_adjustment.cpp:182-264): pose, affine brightness,
  // idepths / variances / inlier counts / baselines, and the connection statuses of this frame's landmarks
  void updateFrame(long long timestamp, double t_world_agent[12], double affine_brightness[2], LandmarkResult& lms,
                   std::map<int, std::vector<uint8_t>>& statuses_as_reference) {
    const int slot = slotOf(timestamp);
    if (slot < 0) throw DpbaFailure(DPBA_E_INVALID, "Cannot update frame, there is no local copy in the solver");
    const dense::Vec eps = stateEps();
    tWorldAgent(slot, eps, t_world_agent);
    for (int k = 0; k < 2; ++k) affine_brightness[k] = frames_[slot].affine_brightness0[k] + eps[kBlockSize * slot + 6 + k];
    const int n = dpba_num_landmarks(h_, slot);
    std::vector<float> inv_hdd(n);
    std::vector<uint8_t> flags(n);
    lms.idepth.assign(n, 0);
    lms.relative_baseline.assign(n, 0);
    lms.number_of_inlier_residuals.assign(n, 0);
    lms.idepth_variance.assign(n, 0);
    lms.is_outlier.assign(n, 0);
    if (n)
      dpba_check(h_, dpba_get_landmarks(h_, slot, n, lms.idepth.data(), nullptr, inv_hdd.data(), nullptr, flags.data(),
                                        lms.number_of_inlier_residuals.data(), lms.relative_baseline.data()));
    const double kIdepthEps = 1e-8;
    for (int l = 0; l < n; ++l) {
      lms.is_outlier[l] = (flags[l] & DPBA_LM_OUTLIER) != 0;
      if (!(flags[l] & DPBA_LM_MARGINALIZED)) {
        if (std::fabs(lms.idepth[l]) < kIdepthEps) lms.idepth[l] = 0;
        else if (lms.idepth[l] < 0) lms.is_outlier[l] = 1;
        lms.idepth_variance[l] = estimate_uncertainty_ ? inv_hdd[l] : 1e-5f;
      }
    }
    statuses_as_reference.clear();
    uint8_t* rows[DPBA_MAX_FRAMES] = {};
    for (size_t t = 0; t < frames_.size(); ++t) {
      if ((int)t == slot) continue;
      auto& st = statuses_as_reference[frames_[t].id];
      st.assign(n, 0);
      rows[t] = st.data();
    }
    if (n) dpba_check(h_, dpba_get_frame_statuses(h_, slot, n, rows, nullptr));  // one synchronisation per frame
  }

  // The frames_.size() > 1 half of pushFrame on its own (eigen_photometric_bundle_adjustment.cpp:121-130): what the next
  // pushFrame would do before it appends its frame.  Lets a caller (bench.py configs[4], the tests) run and time the
  // marginalisation update without uploading a new keyframe.
  void marginalizeNow() {
    if (frames_.size() > 1) marginalize();
  }

  // createReferenceDepthMaps (src/tracker/tracker/src/create_depth_maps.cpp:122-146), which the tracker calls right after
  // solve / updateSolver (monocular_tracker.cpp:465,509): built on the device from the window this solver holds.
  // Level l: (height >> l) x (width >> l), row-major [y][x] = the reference's map(x, y); `idepth` is the weighted SUM.
  struct DepthMapLevel {
    int width = 0, height = 0;
    std::vector<float> idepth, weight;
  };
  std::vector<DepthMapLevel> createReferenceDepthMaps(int width, int height, int levels) {
    std::vector<DepthMapLevel> maps(levels);
    std::vector<float*> pi(levels), pw(levels);
    for (int l = 0; l < levels; ++l) {
      maps[l].width = width >> l;
      maps[l].height = height >> l;
      maps[l].idepth.assign((size_t)maps[l].width * maps[l].height, 0.f);
      maps[l].weight.assign((size_t)maps[l].width * maps[l].height, 0.f);
      pi[l] = maps[l].idepth.data();
      pw[l] = maps[l].weight.data();
    }
    // idepthVariance of the track: inv_hessian_idepth_idepth with uncertainty estimation, else the constant 1e-5
    // (photometric_bundle_adjustment.cpp:252-254)
    dpba_check(h_, dpba_create_reference_depth_maps(h_, levels, estimate_uncertainty_ ? -1.0 : 1e-5, pi.data(), pw.data()));
    return maps;
  }

 private:
  int slotOf(long long timestamp) const {
    for (size_t i = 0; i < frames_.size(); ++i)
      if (frames_[i].timestamp == timestamp) return (int)i;
    return -1;
  }
  int slotOfId(int id) const {
    for (size_t i = 0; i < frames_.size(); ++i)
      if (frames_[i].id == id) return (int)i;
    return -1;
  }
  dense::Vec stateEps() const {
    dense::Vec eps(kBlockSize * frames_.size()), step(eps.size());
    dpba_check(h_, dpba_get_state(h_, eps.data(), step.data()));
    return eps;
  }

  // pushFrame (photometric_bundle_adjustment.cpp:104-123): the residual vectors between the new frame and every local,
  // not marginalised frame are created in full, in both directions.
  void uploadStatuses(int slot, const KeyframeView& frame) {
    for (const auto& [other_id, st] : frame.statuses_as_reference) {
      const int o = slotOfId(other_id);
      if (o >= 0 && o != slot && !frames_[o].is_marginalized && !st.empty())
        dpba_check(h_, dpba_set_statuses(h_, slot, o, (int)st.size(), st.data()));
    }
    for (const auto& [other_id, st] : frame.statuses_as_target) {
      const int o = slotOfId(other_id);
      if (o >= 0 && o != slot && !frames_[o].is_marginalized && !st.empty())
        dpba_check(h_, dpba_set_statuses(h_, o, slot, (int)st.size(), st.data()));
    }
  }
  // LocalFrame::update (local_frame.hpp:505-518): only the residual vectors this frame is the REFERENCE of grow, and only
  // by the statuses of the freshly matured landmarks [first, size); what the solver decided for the existing residuals
  // (kOutlier from updatePointStatuses, kOOB from changeResidualStatuses) stays.
  void appendStatuses(int slot, const KeyframeView& frame, int first) {
    for (const auto& [other_id, st] : frame.statuses_as_reference) {
      const int o = slotOfId(other_id);
      if (o < 0 || o == slot || (int)st.size() <= first) continue;
      dpba_check(h_, dpba_append_statuses(h_, slot, o, first, (int)st.size() - first, st.data() + first));
    }
  }

  // ---- SE3 helpers (Sophus closed forms; se3_motion.hpp:231-252) ---------------------------------
  static void se3Exp(const double* xi, double R[9], double t[3]) {
    const double* v = xi;
    const double* w = xi + 3;
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
    double a, b, c;
    if (th < 1e-10) a = 1, b = 0.5, c = 1.0 / 6.0;
    else a = std::sin(th) / th, b = (1 - std::cos(th)) / th2, c = (th - std::sin(th)) / (th2 * th);
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double W2[9], V[9];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += W[i * 3 + k] * W[k * 3 + j];
        W2[i * 3 + j] = s;
      }
    for (int i = 0; i < 9; ++i) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * W[i] + b * W2[i];
      V[i] = I + b * W[i] + c * W2[i];
    }
    for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * v[0] + V[i * 3 + 1] * v[1] + V[i * 3 + 2] * v[2];
  }
  // LocalFrame::tWorldAgent = T_lin * exp(eps[0:6])  (local_frame.hpp:525-527)
  void tWorldAgent(int slot, const dense::Vec& eps, double out[12]) const {
    const double* T = frames_[slot].T_w_agent_linearization_point;
    double R[9], t[3];
    se3Exp(&eps[kBlockSize * slot], R, t);
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) {
        double s = 0;
        for (int k = 0; k < 3; ++k) s += T[i * 4 + k] * R[k * 3 + j];
        out[i * 4 + j] = s;
      }
      double s = T[i * 4 + 3];
      for (int k = 0; k < 3; ++k) s += T[i * 4 + k] * t[k];
      out[i * 4 + 3] = s;
    }
  }

  // relinearizeSystem (photometric_bundle_adjustment.cpp:311-316): only the LAST frame (quirk Q9)
  void relinearizeSystem() {
    const int last = (int)frames_.size() - 1;
    const dense::Vec eps = stateEps();
    FrameMeta& m = frames_[last];
    double T[12];
    tWorldAgent(last, eps, T);
    std::memcpy(m.T_w_agent_linearization_point, T, sizeof(T));
    for (int k = 0; k < 2; ++k) m.affine_brightness0[k] += eps[kBlockSize * last + 6 + k];
    dpba_check(h_, dpba_set_frame_linearization(h_, last, m.T_w_agent_linearization_point, m.affine_brightness0));
  }

  // covarianceMatrixPosePose (problem.hpp:204-242): Jacobians WITHOUT the Huber loss
  dense::Mat covarianceMatrixPosePose() {
    const int n = kBlockSize * (int)frames_.size();
    NormalLinearSystem pose(n), schur(n);
    dpba_check(h_, dpba_linearize(h_, 0.0, 0, 1, 0, pose.H.a.data(), pose.b.data(), schur.H.a.data(), schur.b.data()));
    evaluateLinearSystemPrior(frames_, stateEps(), pose, options_.affine_brightness_regularizer,
                              options_.fixed_state_regularizer);
    NormalLinearSystem full = pose - schur + system_marginalized_;
    return dense::sym_pinv(full.H, 1);  // scale nullspace
  }

  // covarianceMatricesOfRelativePoses (covariance_matrices_of_relative_poses.hpp:24-63) with
  // relativeTransformationUncertainty (se3_motion.hpp:151-158)
  void covarianceMatricesOfRelativePoses(const dense::Mat& cov) {
    covariance_matrices_.clear();
    const dense::Vec eps = stateEps();
    const int N = (int)frames_.size();
    std::vector<std::vector<double>> Tw(N, std::vector<double>(12));
    for (int i = 0; i < N; ++i) tWorldAgent(i, eps, Tw[i].data());
    for (int r = 0; r < N; ++r)
      for (int t = 0; t < N; ++t) {
        if (r == t) continue;
        // adj = Adj(T_w_t^-1 T_w_r)
        double R[9], tr[3];
        for (int i = 0; i < 3; ++i) {
          for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Tw[t][k * 4 + i] * Tw[r][k * 4 + j];
            R[i * 3 + j] = s;
          }
          double s = 0;
          for (int k = 0; k < 3; ++k) s += Tw[t][k * 4 + i] * (Tw[r][k * 4 + 3] - Tw[t][k * 4 + 3]);
          tr[i] = s;
        }
        dense::Mat adj(6, 6);
        const double th[9] = {0, -tr[2], tr[1], tr[2], 0, -tr[0], -tr[1], tr[0], 0};
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += th[i * 3 + k] * R[k * 3 + j];
            adj(i, j) = R[i * 3 + j];
            adj(i, 3 + j) = s;
            adj(3 + i, 3 + j) = R[i * 3 + j];
          }
        dense::Mat s11(6, 6), s22(6, 6), s12(6, 6);
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) {
            s11(i, j) = cov(kBlockSize * r + i, kBlockSize * r + j);
            s22(i, j) = cov(kBlockSize * t + i, kBlockSize * t + j);
            s12(i, j) = cov(kBlockSize * r + i, kBlockSize * t + j);
          }
        const dense::Mat adjT = dense::transpose(adj);
        const dense::Mat a = dense::matmul(dense::matmul(adj, s11), adjT);
        const dense::Mat b = dense::matmul(dense::transpose(s12), adjT);
        const dense::Mat c = dense::matmul(adj, s12);
        std::vector<double> out(36);
        for (int i = 0; i < 6; ++i)
          for (int j = 0; j < 6; ++j) out[i * 6 + j] = a(i, j) - b(i, j) - c(i, j) + s22(i, j);
        covariance_matrices_[{frames_[r].id, frames_[t].id}] = out;
      }
  }

  // the frames_.size() > 1 half of pushFrame: firstEstimateJacobians, evaluateJacobians, changeResidualStatuses,
  // updateMarginalizedLinearSystem (eigen_photometric_bundle_adjustment.cpp:121-130; problem.hpp:146-203)
  void marginalize() {
    const int n = kBlockSize * (int)frames_.size();
    dpba_check(h_, dpba_first_estimate(h_));
    NormalLinearSystem pose(n), schur(n);
    // one fused pass: sweep with Jacobians + PosePose<true> + Schur<true>; statuses are committed afterwards, which
    // is equivalent because the sweep only reads the committed status and writes the candidate
    dpba_check(h_, dpba_linearize(h_, options_.sigma_huber_loss, 1, 1, 1, pose.H.a.data(), pose.b.data(),
                                  schur.H.a.data(), schur.b.data()));
    dpba_check(h_, dpba_change_residual_statuses(h_, 1));
    NormalLinearSystem points = pose - schur;
    const dense::Vec state = stateEps();
    double e_marg = 0;
    int32_t nv = 0;
    dpba_check(h_, dpba_landmarks_energy(h_, 1, &e_marg, &nv));
    const dense::Vec Hs = dense::matvec(points.H, state);
    energy_marginalized_ += e_marg + dense::dot(state, Hs) - dense::dot(state, points.b);  // DSO eq 8.15
    for (int i = 0; i < n; ++i) points.b[i] -= Hs[i];
    if (system_marginalized_.size() != n) system_marginalized_.resize(n);
    system_marginalized_ += points;
    // landmark.to_marginalize = false for every landmark (problem.hpp:175-177)
    for (size_t f = 0; f < frames_.size(); ++f) {
      const int m = dpba_num_landmarks(h_, (int)f);
      if (!m) continue;
      std::vector<uint8_t> flags(m);
      dpba_check(h_, dpba_get_landmarks(h_, (int)f, m, nullptr, nullptr, nullptr, nullptr, flags.data(), nullptr, nullptr));
      for (auto& fl : flags) fl &= ~DPBA_LM_TO_MARGINALIZE;
      dpba_check(h_, dpba_set_landmark_flags(h_, (int)f, m, flags.data()));
    }
    std::vector<int> marginalized_part;
    for (size_t i = 0; i < frames_.size(); ++i)
      if (frames_[i].to_marginalize)
        for (int p = 0; p < kBlockSize; ++p) marginalized_part.push_back((int)i * kBlockSize + p);
    if (marginalized_part.empty()) return;
    NormalLinearSystem prior(n);
    evaluateLinearSystemPrior(frames_, state, prior, options_.affine_brightness_regularizer,
                              options_.fixed_state_regularizer, true);
    const dense::Vec Hp = dense::matvec(prior.H, state);
    for (int i = 0; i < n; ++i) prior.b[i] -= Hp[i];
    system_marginalized_ += prior;
    system_marginalized_.reduce_system(marginalized_part);
    for (int i = (int)frames_.size() - 1; i >= 0; --i)
      if (frames_[i].to_marginalize) {
        dpba_check(h_, dpba_remove_frame(h_, i));
        frames_.erase(frames_.begin() + i);
      }
  }

  dpba_handle* h_ = nullptr;
  TrustRegionPhotometricBundleAdjustmentOptions options_;
  bool estimate_uncertainty_;
  bool force_accept_;
  bool device_resident_lm_;
  std::vector<FrameMeta> frames_;
  NormalLinearSystem system_marginalized_{0};
  Precision energy_marginalized_ = 0;
  size_t last_iterations_ = 0;
  std::map<std::pair<int, int>, std::vector<double>> covariance_matrices_;
};

}  // namespace dsopp_b200
