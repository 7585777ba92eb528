// TEST INFRASTRUCTURE ONLY: stand-in for opencv_contrib line_descriptor (not installed here).  LSDDetector::detect and
// BinaryDescriptor::compute forward to the oracle's LSD / KeyLine / LBD restatement (oracle/line_oracle.cpp; LSD pinned to
// cv2 4.13, KeyLine packaging and LBD restated from memory of the contrib sources: "parity unpinned" for those two).
#pragma once
#include "minicv.hpp"
namespace cv { namespace line_descriptor {
struct KeyLine {
    float angle; int class_id; int octave; Point2f pt; float response; float size;
    float startPointX, startPointY, endPointX, endPointY, sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int numOfPixels;
    Point2f getStartPoint() const { return Point2f(startPointX, startPointY); }
    Point2f getEndPoint() const { return Point2f(endPointX, endPointY); }
    Point2f getStartPointInOctave() const { return Point2f(sPointInOctaveX, sPointInOctaveY); }
    Point2f getEndPointInOctave() const { return Point2f(ePointInOctaveX, ePointInOctaveY); }
    KeyLine() { memset((void*)this, 0, sizeof(*this)); }
};
static_assert(sizeof(KeyLine) == 68, "KeyLine layout");
class BinaryDescriptor {
public:
    static Ptr<BinaryDescriptor> createBinaryDescriptor() { return Ptr<BinaryDescriptor>(new BinaryDescriptor()); }
    void compute(const Mat& image, std::vector<KeyLine>& keylines, Mat& descriptors, bool returnFloatDescr = false) const;
};
class LSDDetector {
public:
    static Ptr<LSDDetector> createLSDDetector() { return Ptr<LSDDetector>(new LSDDetector()); }
    void detect(const Mat& image, std::vector<KeyLine>& keypoints, int scale, int numOctaves, const Mat& mask = Mat());
};
} }
