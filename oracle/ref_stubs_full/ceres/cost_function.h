// STAND-IN -- this is NOT Ceres (see jet.h next to this file).
#pragma once
#include "jet.h"
