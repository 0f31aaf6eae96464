// TEST INFRASTRUCTURE ONLY -- see ../opencv.hpp.
#pragma once
#include "../opencv.hpp"
