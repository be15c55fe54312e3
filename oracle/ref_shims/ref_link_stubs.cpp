// Link-time stand-ins (test infrastructure, see oracle/build_ref_pba.py): functions of the reference that the compiled
// sources MENTION on paths the pinned bundle adjustment never takes, and whose own source files would pull in protobuf.
// They abort if they are ever reached.
#include <cstdlib>

#include "semantics/semantic_filter.hpp"

namespace dsopp::semantics {
// named by CameraMask::filterSemanticObjects (sensors/camera_calibration/src/camera_mask.cpp:32-40); defined in
// common/semantics/src/semantic_filter.cpp, which needs SemanticLegend and its protobuf message
bool SemanticFilter::filtered(size_t) const { std::abort(); }
}  // namespace dsopp::semantics
