// CudaPhotometricBundleAdjustmentProblem: the LevenbergMarquardtProblem whose data-parallel work runs on the
// B200 behind the C ABI.  Mirrors PhotometricBundleAdjustmentProblem
// (src/energy/problems/internal/energy/problems/photometric_bundle_adjustment/eigen_photometric_bundle_adjustment_problem.hpp:255-429)
// method for method; the host keeps the priors (:37-77), the marginalised prior terms (:293-298,347-351) and the
// 8N x 8N solve (:352), all in double.
#pragma once
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "dsopp_cuda_pba.h"
#include "normal_linear_system.hpp"

namespace dsopp_b200 {

using Precision = double;
constexpr int kBlockSize = DPBA_BLOCK;

struct FrameMeta {  // what the host needs to know per LocalFrame
  int id = 0;
  long long timestamp = 0;
  bool fixed = false;            // FrameParameterization::kFixed
  bool to_marginalize = false;
  bool is_marginalized = false;
  double T_w_agent_linearization_point[12];
  double affine_brightness0[2] = {0, 0};
  double exposure_time = 1;
};

struct DpbaFailure : std::runtime_error {
  int code;
  DpbaFailure(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void dpba_check(dpba_handle* h, int rc) {
  if (rc < 0) throw DpbaFailure(rc, std::string("dpba error ") + std::to_string(rc) + ": " + dpba_last_error(h));
}

// evaluateLinearSystemPrior (problem.hpp:37-77); MotionPrior<SE3> is identically zero (state_priors.hpp:30-73)
inline void evaluateLinearSystemPrior(const std::vector<FrameMeta>& frames, const dense::Vec& state_eps,
                                      NormalLinearSystem& system_prior, const double affine_reg[2],
                                      double fixed_state_reg, bool for_marginalized = false) {
  for (size_t i = 0; i < frames.size(); ++i) {
    if (frames[i].to_marginalize != for_marginalized) continue;
    const int o = kBlockSize * (int)i;
    if (frames[i].fixed) {
      for (int k = 0; k < kBlockSize; ++k) {
        system_prior.H(o + k, o + k) += fixed_state_reg;
        system_prior.b[o + k] += fixed_state_reg * state_eps[o + k];
      }
    } else {
      for (int k = 0; k < 2; ++k) {
        const double ab = frames[i].affine_brightness0[k] + state_eps[o + 6 + k];
        system_prior.H(o + 6 + k, o + 6 + k) += affine_reg[k];
        system_prior.b[o + 6 + k] += affine_reg[k] * ab;
      }
    }
  }
}

class CudaPhotometricBundleAdjustmentProblem {
 public:
  CudaPhotometricBundleAdjustmentProblem(dpba_handle* handle, const std::vector<FrameMeta>& frames,
                                         Precision sigma_huber_loss, const NormalLinearSystem& system_marginalized,
                                         Precision energy_marginalized, const double affine_brightness_regularizer[2],
                                         Precision fixed_pose_regularizer, bool first_estimate_jacobians = true)
      : h_(handle),
        frames_(frames),
        sigma_huber_loss_(sigma_huber_loss),
        system_marginalized_(system_marginalized),
        energy_marginalized_(energy_marginalized),
        fixed_pose_regularizer_(fixed_pose_regularizer),
        fej_(first_estimate_jacobians),
        system_pose_(kBlockSize * (int)frames.size()),
        system_schur_(kBlockSize * (int)frames.size()) {
    affine_brightness_regularizer_[0] = affine_brightness_regularizer[0];
    affine_brightness_regularizer_[1] = affine_brightness_regularizer[1];
    if (system_marginalized_.size() != system_pose_.size()) system_marginalized_.resize(system_pose_.size());
  }

  std::pair<Precision, int> calculateEnergy() {  // problem.hpp:290-317
    double landmarks_energy = 0;
    int32_t n_valid = 0;
    dpba_check(h_, dpba_evaluate(h_, sigma_huber_loss_, 1, fej_, &landmarks_energy, &n_valid));
    const dense::Vec state = stateEpsStacked(true);
    Precision energy = energy_marginalized_ + dense::dot(system_marginalized_.b, state) +
                       dense::dot(state, dense::matvec(system_marginalized_.H, state)) / 2;  // DSO eq 8.19
    for (size_t i = 0; i < frames_.size(); ++i) {  // every frame, the fixed one included (quirk Q8)
      for (int k = 0; k < 2; ++k) {
        const double ab = frames_[i].affine_brightness0[k] + state[kBlockSize * i + 6 + k];
        energy += ab * affine_brightness_regularizer_[k] * ab / 2;  // AffineBrightnessPrior::energyTerm
      }
    }
    return {energy + landmarks_energy, n_valid};
  }

  void linearize() {  // problem.hpp:322-336
    dpba_check(h_, dpba_linearize(h_, sigma_huber_loss_, 1, fej_, 0, system_pose_.H.a.data(), system_pose_.b.data(),
                                  system_schur_.H.a.data(), system_schur_.b.data()));
    evaluateLinearSystemPrior(frames_, stateEpsStacked(false), system_pose_, affine_brightness_regularizer_,
                              fixed_pose_regularizer_);
  }

  void calculateStep(const Precision lambda) {  // problem.hpp:342-361
    const int n = system_pose_.size();
    const dense::Vec state = stateEpsStacked(false);
    NormalLinearSystem system_full = system_pose_ + system_marginalized_;
    for (int i = 0; i < n; ++i) system_full.H(i, i) += system_pose_.H(i, i) * lambda;
    system_full += system_schur_ * (-1.0 / (1.0 + lambda));
    const dense::Vec hs = dense::matvec(system_marginalized_.H, state);
    for (int i = 0; i < n; ++i) system_full.b[i] += hs[i];
    step_ = system_full.solve();
    dense::Vec neg(n);
    for (int i = 0; i < n; ++i) neg[i] = -step_[i];
    dpba_check(h_, dpba_set_state(h_, nullptr, neg.data()));  // frame.state_eps_step = -frame_step
    dpba_check(h_, dpba_back_substitute(h_, step_.data(), lambda));
  }

  std::pair<Precision, Precision> acceptStep() {  // problem.hpp:366-388
    double state_sq = 0, step_sq = 0;
    dpba_check(h_, dpba_accept(h_, &state_sq, &step_sq));
    return {state_sq, step_sq};
  }

  void rejectStep() { dpba_check(h_, dpba_reject(h_)); }  // problem.hpp:392-402

  bool stop() { return false; }  // problem.hpp:407

  const dense::Vec& lastStep() const { return step_; }
  const NormalLinearSystem& systemPose() const { return system_pose_; }
  const NormalLinearSystem& systemSchur() const { return system_schur_; }

  dense::Vec stateEpsStacked(bool with_step) const {  // problem.hpp:79-92
    const int n = kBlockSize * (int)frames_.size();
    dense::Vec eps(n), step(n);
    dpba_check(h_, dpba_get_state(h_, eps.data(), step.data()));
    if (with_step)
      for (int i = 0; i < n; ++i) eps[i] += step[i];
    return eps;
  }

 private:
  dpba_handle* h_;
  const std::vector<FrameMeta>& frames_;
  const Precision sigma_huber_loss_;
  NormalLinearSystem system_marginalized_;
  const Precision energy_marginalized_;
  double affine_brightness_regularizer_[2];
  const Precision fixed_pose_regularizer_;
  const bool fej_;
  NormalLinearSystem system_pose_;
  NormalLinearSystem system_schur_;
  dense::Vec step_;
};

}  // namespace dsopp_b200
