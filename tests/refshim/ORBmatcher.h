// stand-in for include/ORBmatcher.h:36-101 (the entry points the adapter defines)
#pragma once
#include <vector>
#include <set>
#include "MapPoint.h"
#include "KeyFrame.h"
#include "Frame.h"
namespace StructureSLAM {
class ORBmatcher {
public:
    ORBmatcher(float nnratio=0.6, bool checkOri=true);
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
    int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono);
    int SearchByBoW(KeyFrame *pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches);
    int SearchByBoW(KeyFrame *pKF1, KeyFrame* pKF2, std::vector<MapPoint*> &vpMatches12);
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo);
protected:
    float mfNNratio;
    bool mbCheckOrientation;
};
}
