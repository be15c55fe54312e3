// STAND-IN for the protobuf-generated header (see frame.pb.h).
#pragma once
namespace dsopp::sensors::calibration::proto {
class CameraSettings {};
}  // namespace dsopp::sensors::calibration::proto
