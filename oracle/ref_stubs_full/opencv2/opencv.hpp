// STAND-IN -- this is NOT OpenCV (see core.hpp next to this file).
#pragma once
#include "core.hpp"
