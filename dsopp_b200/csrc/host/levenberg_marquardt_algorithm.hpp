// Host LM driver with the reference's problem concept, options and control flow
// (src/energy/problems/include/energy/levenberg_marquardt_algorithm/levenberg_marquardt_algorithm.hpp:16-128).
// The loop stays on the host; the five problem methods are where the C ABI is crossed.
#pragma once
#include <concepts>
#include <cstddef>
#include <limits>
#include <utility>

namespace dsopp_b200::levenberg_marquardt_algorithm {

using Precision = double;  // reference default (src/common/include/common/settings.hpp:10-14)

template <typename Problem>
concept LevenbergMarquardtProblem = requires(Problem& problem, const Precision lambda) {
  { problem.calculateEnergy() } -> std::same_as<std::pair<Precision, int>>;
  { problem.linearize() } -> std::same_as<void>;
  { problem.calculateStep(lambda) } -> std::same_as<void>;
  { problem.acceptStep() } -> std::same_as<std::pair<Precision, Precision>>;
  { problem.rejectStep() } -> std::same_as<void>;
};

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
  size_t iterations = 0;  // loop bodies executed (not in the reference; used by the benchmark)
};

template <LevenbergMarquardtProblem Problem>
Result solve(Problem& problem, const Options& options) {
  Result result;
  Precision lambda = options.initial_levenberg_marquardt_regularizer;
  auto e0 = problem.calculateEnergy();
  result.energy = e0.first;
  result.number_of_valid_residuals = e0.second;
  bool linear_system_valid = false;
  for (size_t iteration = 0;
       iteration < options.max_num_iterations && !result.converged && result.number_of_valid_residuals > 0;
       ++iteration) {
    ++result.iterations;
    if (!linear_system_valid) problem.linearize();
    problem.calculateStep(lambda);
    auto [next_energy, number_of_valid_residuals] = problem.calculateEnergy();
    if (problem.stop() || number_of_valid_residuals == 0) {
      problem.rejectStep();
      break;
    }
    const bool function_tolerance_reached =
        std::abs(result.energy - next_energy) / result.energy < options.function_tolerance;
    result.converged |= function_tolerance_reached;  // set before the accept decision (quirk Q7)
    if (next_energy < result.energy || (options.force_accept && iteration < options.min_num_iterations)) {
      auto [state_squared_norm, step_squared_norm] = problem.acceptStep();
      result.converged |= step_squared_norm < options.parameter_tolerance * (state_squared_norm + options.parameter_tolerance);
      result.energy = next_energy;
      result.number_of_valid_residuals = number_of_valid_residuals;
      lambda /= options.levenberg_marquardt_regularizer_decrease_on_accept;
      linear_system_valid = false;
    } else {
      problem.rejectStep();
      if (options.force_accept) {
        problem.calculateEnergy();
        return result;
      }
      lambda *= options.levenberg_marquardt_regularizer_increase_on_reject;
      linear_system_valid = true;
    }
  }
  problem.calculateEnergy();
  return result;
}

}  // namespace dsopp_b200::levenberg_marquardt_algorithm
