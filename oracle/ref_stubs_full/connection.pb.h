// STAND-IN for the protobuf-generated header (see frame.pb.h).
#pragma once
namespace dsopp::track::proto {
class Connection {};
class Connections {};
}  // namespace dsopp::track::proto
