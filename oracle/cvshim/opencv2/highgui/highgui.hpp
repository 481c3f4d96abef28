// TEST INFRASTRUCTURE: forwards to the single-header OpenCV stand-in (see opencv.hpp).
#pragma once
#include "../opencv.hpp"
