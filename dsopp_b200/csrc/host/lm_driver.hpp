// Host-side Levenberg-Marquardt driver.  It serves the same five-method problem concept, option names and results as
// the reference's levenberg_marquardt_algorithm::solve
// (src/energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:16-128), so a
// problem written against the reference drops in unchanged; the five methods are where the C ABI is crossed.
// Behaviour is pinned call by call against the reference's own driver (compiled from its source for the tests:
// oracle/build_ref.py, tests/test_reference_parts.py) through dpbah_lm_scripted in host_capi.cpp.
//
// Organisation: every loop pass evaluates ONE trial point and hands (current, trial) to `judge`, which names the outcome;
// the two stopping tests are the small predicates below.  Reference behaviours kept on purpose:
//   * the function-tolerance test is taken on the trial energy BEFORE the accept decision, so a rejected trial can
//     still mark the run converged;
//   * under force_accept a rejected trial ends the run (after rolling back and re-evaluating the energy);
//   * a trial without valid residuals, or a problem asking to stop, rolls back and ends the run.
#pragma once
#include <cmath>
#include <concepts>
#include <cstddef>
#include <limits>
#include <utility>

namespace dsopp_b200::levenberg_marquardt_algorithm {

using Precision = double;  // reference default (src/common/include/common/settings.hpp:10-14)

template <typename P>
concept LevenbergMarquardtProblem = requires(P& p, const Precision damping) {
  { p.calculateEnergy() } -> std::same_as<std::pair<Precision, int>>;
  { p.linearize() } -> std::same_as<void>;
  { p.calculateStep(damping) } -> std::same_as<void>;
  { p.acceptStep() } -> std::same_as<std::pair<Precision, Precision>>;
  { p.rejectStep() } -> std::same_as<void>;
};

// field names and defaults are the reference's (:38-55): callers fill them by name
struct Options {
  size_t max_num_iterations = 50;
  Precision initial_levenberg_marquardt_regularizer = 1e-5;
  Precision function_tolerance = 1e-8;
  Precision parameter_tolerance = 1e-8;
  bool force_accept = false;
  size_t min_num_iterations = 0;
  Precision levenberg_marquardt_regularizer_decrease_on_accept = 2;
  Precision levenberg_marquardt_regularizer_increase_on_reject = 10;
};

struct Result {
  Precision energy = std::numeric_limits<Precision>::max();
  int number_of_valid_residuals = 0;
  bool converged = false;
  size_t iterations = 0;  // trial points evaluated (not in the reference; the benchmark reads it)
};

namespace detail {

enum class Outcome { kAbandon, kTake, kRefuse };

inline Outcome judge(const Options& o, size_t pass, Precision current, Precision trial, int trial_valid, bool stop_asked) {
  if (stop_asked || trial_valid == 0) return Outcome::kAbandon;
  const bool still_forced = o.force_accept && pass < o.min_num_iterations;
  return (trial < current || still_forced) ? Outcome::kTake : Outcome::kRefuse;
}

inline bool energy_stalled(const Options& o, Precision current, Precision trial) {
  return std::abs(current - trial) / current < o.function_tolerance;
}

inline bool step_negligible(const Options& o, const std::pair<Precision, Precision>& state_and_step_sq) {
  return state_and_step_sq.second < o.parameter_tolerance * (state_and_step_sq.first + o.parameter_tolerance);
}

}  // namespace detail

template <LevenbergMarquardtProblem Problem>
Result solve(Problem& problem, const Options& o) {
  using detail::Outcome;
  Result run;
  {
    const auto start = problem.calculateEnergy();
    run.energy = start.first;
    run.number_of_valid_residuals = start.second;
  }
  Precision damping = o.initial_levenberg_marquardt_regularizer;
  bool system_is_current = false;  // a refused trial leaves the linearisation usable for the next, more damped, step

  for (size_t pass = 0; pass < o.max_num_iterations; ++pass) {
    if (run.converged || run.number_of_valid_residuals <= 0) break;
    ++run.iterations;
    if (!system_is_current) problem.linearize();
    problem.calculateStep(damping);
    const std::pair<Precision, int> trial = problem.calculateEnergy();

    const Outcome outcome = detail::judge(o, pass, run.energy, trial.first, trial.second, problem.stop());
    if (outcome == Outcome::kAbandon) {
      problem.rejectStep();
      break;
    }
    if (detail::energy_stalled(o, run.energy, trial.first)) run.converged = true;

    if (outcome == Outcome::kTake) {
      if (detail::step_negligible(o, problem.acceptStep())) run.converged = true;
      run.energy = trial.first;
      run.number_of_valid_residuals = trial.second;
      damping /= o.levenberg_marquardt_regularizer_decrease_on_accept;
      system_is_current = false;
      continue;
    }
    problem.rejectStep();
    if (o.force_accept) break;
    damping *= o.levenberg_marquardt_regularizer_increase_on_reject;
    system_is_current = true;
  }
  problem.calculateEnergy();  // leaves the problem's residual state at the accepted point (:121, :126)
  return run;
}

}  // namespace dsopp_b200::levenberg_marquardt_algorithm
