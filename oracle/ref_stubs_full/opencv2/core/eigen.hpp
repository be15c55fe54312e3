// STAND-IN -- this is NOT OpenCV (see ../core.hpp).
#pragma once
#include "../core.hpp"
