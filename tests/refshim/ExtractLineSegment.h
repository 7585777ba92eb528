// Stand-in for the reference's include/ExtractLineSegment.h (lines 53-76): the one method host/ExtractLineSegment_b200.cc defines.
// Compile-check use only, see README.md here.
#pragma once
#include <vector>
#include <Eigen/Core>
#include <opencv2/core/core.hpp>
#include <opencv2/line_descriptor/descriptor.hpp>
using namespace std; using namespace cv; using namespace Eigen;
namespace StructureSLAM { class LineSegment { public:
    LineSegment();
    void ExtractLineSegment(const Mat& image, vector<KeyLine>& lines, Mat& descriptors, vector<Vector3d>& equations, int scale = 1.2, int octaves = 1);
  protected: double nn_mad, nn12_mad; }; }
