// stand-in for include/ExtractLineSegment.h:53-76
#pragma once
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/line_descriptor/descriptor.hpp>
#include <Eigen/Core>
using namespace std;
using namespace cv;
using namespace Eigen;
namespace StructureSLAM {
class LineSegment {
public:
    LineSegment();
    void ExtractLineSegment(const Mat &img, vector<KeyLine> &keylines, Mat &ldesc, vector<Vector3d> &keylineFunctions, int scale = 1.2, int numOctaves = 1);
protected:
    double nn_mad, nn12_mad;
};
}
