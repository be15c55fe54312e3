// C entry points around the two pieces of the REFERENCE ITSELF that compile here from their own sources (test
// infrastructure only; nothing under dsopp_b200/ links this):
//
//   src/features/src/calculate_pixelinfo.cpp                      calculate_pixelinfo<1>  (the {I,dx,dy}
//                                                                 gradient definition, AVX2 and plain-C paths)
//   src/energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp
//                                                                 levenberg_marquardt_algorithm::solve (row a17)
//   src/common/pattern/include/common/pattern/pattern.hpp         the 8-pixel residual pattern (constants only)
//
// The reference sources are NOT copied: oracle/build_ref.py compiles them where they lie under /root/reference and this
// file only declares / instantiates what it calls.  The LM driver is a template over a problem type; here it is
// instantiated with a SCRIPTED problem whose energies, valid-residual counts and step norms are fed from arrays and
// which records every call, so that the control flow (accept / reject, lambda schedule, force_accept, convergence
// tests, early exits) of our restatements can be compared call by call with the reference's.
#include <cstdint>
#include <vector>

#include "common/pattern/pattern.hpp"
#include "energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp"
#include "features/camera/calculate_pixelinfo.hpp"

namespace {

using dsopp::Precision;

enum Call : int32_t { kEnergy = 0, kLinearize = 1, kStep = 2, kAccept = 3, kReject = 4 };

struct ScriptedProblem {
  const double* energies;
  const int32_t* valid;
  int n_energy;
  const double* norms;  // [n_norms][2] = (state_squared_norm, step_squared_norm)
  int n_norms;
  int i_energy = 0, i_norms = 0;
  std::vector<int32_t> calls;
  std::vector<double> lambdas;

  std::pair<Precision, int> calculateEnergy() {
    calls.push_back(kEnergy);
    const int i = i_energy < n_energy ? i_energy : n_energy - 1;
    ++i_energy;
    return {static_cast<Precision>(energies[i]), valid[i]};
  }
  void linearize() { calls.push_back(kLinearize); }
  void calculateStep(const Precision lambda) {
    calls.push_back(kStep);
    lambdas.push_back(static_cast<double>(lambda));
  }
  std::pair<Precision, Precision> acceptStep() {
    calls.push_back(kAccept);
    const int i = i_norms < n_norms ? i_norms : n_norms - 1;
    ++i_norms;
    return {static_cast<Precision>(norms[2 * i]), static_cast<Precision>(norms[2 * i + 1])};
  }
  void rejectStep() { calls.push_back(kReject); }
  bool stop() { return false; }  // PhotometricBundleAdjustmentProblem::stop, ...problem.hpp:407
};

}  // namespace

extern "C" {

// dsopp::Pattern (src/common/pattern/include/common/pattern/pattern.hpp:17-34): size, centre index, (x_i, y_i) offsets
int ref_pattern(double* xy16, int* center) {
  for (int i = 0; i < 2 * dsopp::Pattern::kSize; ++i) xy16[i] = static_cast<double>(dsopp::Pattern::pattern_data[i]);
  *center = dsopp::Pattern::kCenter;
  return dsopp::Pattern::kSize;
}

void ref_pixelinfo_f64(const double* in, double* out, int width, int height) {
  dsopp::features::calculate_pixelinfo<1>(in, out, width, height);
}
void ref_pixelinfo_f32(const float* in, float* out, int width, int height) {
  dsopp::features::calculate_pixelinfo<1>(in, out, width, height);
}

// returns the number of calls made; calls_out / lambdas_out are filled up to their capacities
int ref_lm_solve(int max_it, double lambda0, double ftol, double ptol, int force_accept, int min_it, double dec, double inc,
                 const double* energies, const int32_t* valid, int n_energy, const double* norms, int n_norms,
                 int32_t* calls_out, int cap_calls, double* lambdas_out, int cap_lambdas, double* energy_out,
                 int32_t* valid_out, int32_t* converged_out) {
  namespace lm = dsopp::energy::levenberg_marquardt_algorithm;
  lm::Options opt;
  opt.max_num_iterations = static_cast<size_t>(max_it);
  opt.initial_levenberg_marquardt_regularizer = static_cast<Precision>(lambda0);
  opt.function_tolerance = static_cast<Precision>(ftol);
  opt.parameter_tolerance = static_cast<Precision>(ptol);
  opt.force_accept = force_accept != 0;
  opt.min_num_iterations = static_cast<size_t>(min_it);
  opt.levenberg_marquardt_regularizer_decrease_on_accept = static_cast<Precision>(dec);
  opt.levenberg_marquardt_regularizer_increase_on_reject = static_cast<Precision>(inc);
  ScriptedProblem p{energies, valid, n_energy, norms, n_norms};
  const lm::Result r = lm::solve(p, opt);
  for (int i = 0; i < (int)p.calls.size() && i < cap_calls; ++i) calls_out[i] = p.calls[i];
  for (int i = 0; i < (int)p.lambdas.size() && i < cap_lambdas; ++i) lambdas_out[i] = p.lambdas[i];
  *energy_out = static_cast<double>(r.energy);
  *valid_out = r.number_of_valid_residuals;
  *converged_out = r.converged ? 1 : 0;
  return (int)p.calls.size();
}

}  // extern "C"
