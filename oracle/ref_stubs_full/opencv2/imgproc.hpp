// STAND-IN -- this is NOT OpenCV (see core.hpp).  The morphology / resize entry points the reference's CameraMask names
// are outside the pinned path and abort if they are ever reached.
#pragma once
#include <cstdlib>
#include "core.hpp"
namespace cv {
inline Mat getStructuringElement(int, Size, Point) { std::abort(); }
inline void erode(const Mat&, Mat&, const Mat&) { std::abort(); }
inline void resize(const Mat&, Mat&, Size, double, double) { std::abort(); }
}  // namespace cv
