#pragma once
#include <opencv2/core/core.hpp>
namespace cv { struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; }; }
