// TEST INFRASTRUCTURE ONLY: forwards to the functional OpenCV stand-in (oracle/refshim/minicv.hpp)
#pragma once
#include "minicv.hpp"
