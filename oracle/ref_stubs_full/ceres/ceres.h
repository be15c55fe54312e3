// STAND-IN -- this is NOT Ceres (see jet.h next to this file).
#pragma once
#include "jet.h"
namespace ceres {
class EvaluationCallback {
 public:
  virtual ~EvaluationCallback() = default;
  virtual void PrepareForEvaluation(bool, bool) = 0;
};
}  // namespace ceres
