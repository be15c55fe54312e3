// STAND-IN -- this is NOT OpenCV (see ../imgproc.hpp).
#pragma once
#include "../imgproc.hpp"
