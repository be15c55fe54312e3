// STAND-IN for the protobuf-generated header (generated code, absent here).  Test infrastructure only: the track
// classes only NAME these message types in declarations of (de)serialisation members the pinned path never calls.
#pragma once
namespace dsopp::track::proto {
class Keyframe {};
class TrackingFrame {};
class Frame {};
}  // namespace dsopp::track::proto
