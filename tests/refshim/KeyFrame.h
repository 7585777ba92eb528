// stand-in for include/KeyFrame.h (members used by the adapters)
#pragma once
#include <vector>
#include "Frame.h"
using namespace cv;                      // include/KeyFrame.h:43
using namespace cv::line_descriptor;     // :45
namespace StructureSLAM {
class KeyFrame {
public:
    void ComputeBoW();
    cv::Mat GetRotation(); cv::Mat GetTranslation(); cv::Mat GetCameraCenter();
    std::vector<MapPoint*> GetMapPointMatches(); MapPoint* GetMapPoint(const size_t &idx);
    std::vector<MapLine*> GetMapLineMatches(); MapLine* GetMapLine(const size_t &idx);
    const float fx, fy, cx, cy;
    const int N;
    const std::vector<cv::KeyPoint> mvKeysUn;
    const cv::Mat mDescriptors; cv::Mat mLineDescriptors;
    DBoW2::BowVector mBowVec; DBoW2::FeatureVector mFeatVec;
    const std::vector<float> mvScaleFactors, mvLevelSigma2;
    ORBVocabulary* mpORBvocabulary;
    KeyFrame();
};
}
