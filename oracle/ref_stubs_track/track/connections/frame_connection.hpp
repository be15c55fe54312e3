// STAND-IN -- this is NOT the reference's track subsystem.  Test infrastructure only (oracle/build_ref_tracker.py): the
// CONTAINER a keyframe keeps per target keyframe, reduced to the accessor the tracker's depth-map build reads
// (create_depth_maps.cpp:30-31).  The enumerators carry the reference's values (track/connections/frame_connection.hpp:19-25).
#ifndef DSOPP_FRAME_CONNECTION_HPP
#define DSOPP_FRAME_CONNECTION_HPP
#include <cstddef>
#include <cstdint>
#include <vector>

#include "energy/motion/motion.hpp"

namespace dsopp::track {
enum struct PointConnectionStatus : uint8_t { kOk = 0, kOutlier = 1, kOccluded = 2, kOOB = 3, kUnknown = 4 };

template <energy::motion::MotionProduct MotionProduct>
class FrameConnection {
 public:
  using ReprojectionStatuses = std::vector<PointConnectionStatus>;
  ReprojectionStatuses statuses;
  const ReprojectionStatuses& referenceReprojectionStatuses(size_t, size_t) const { return statuses; }
};
}  // namespace dsopp::track
#endif
