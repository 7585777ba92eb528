// stand-in for include/Frame.h (members used by the adapters; lines refer to the reference header)
#pragma once
#include <vector>
#include <opencv2/core/core.hpp>
#include <opencv2/line_descriptor/descriptor.hpp>
#include "DBoW2shim.h"
#include "MapPoint.h"
#include "MapLine.h"
namespace StructureSLAM {
using namespace std;
class ORBVocabulary;
class Frame {
public:
    void ComputeBoW();                                                 // :75
    ORBVocabulary* mpORBvocabulary;                                    // :104
    static float fx, fy, cx, cy;                                       // :119-124
    float mbf, mb;                                                     // :130-133
    int N, NL;                                                         // :142-143
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;                        // :150-151
    std::vector<float> mvuRight;                                       // :155
    DBoW2::BowVector mBowVec; DBoW2::FeatureVector mFeatVec;           // :159-160
    cv::Mat mDescriptors, mLdesc;                                      // :163, :166
    std::vector<MapPoint*> mvpMapPoints; std::vector<bool> mvbOutlier; // :172-175
    std::vector<MapLine*> mvpMapLines;
    cv::Mat mTcw;                                                      // :204
    std::vector<float> mvScaleFactors;                                 // :213
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;                       // :219-222
};
}
