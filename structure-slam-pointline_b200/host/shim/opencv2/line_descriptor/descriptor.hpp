#pragma once
#include <opencv2/core/core.hpp>
namespace cv { namespace line_descriptor {
struct KeyLine { float angle; int class_id; int octave; Point2f pt; float response; float size; float startPointX, startPointY, endPointX, endPointY,
                 sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY; float lineLength; int numOfPixels; };
} }
using cv::line_descriptor::KeyLine;
