// stand-in for include/LSDmatcher.h:36-64
#pragma once
#include <vector>
#include "MapLine.h"
#include "KeyFrame.h"
#include "Frame.h"
namespace StructureSLAM {
class LSDmatcher {
public:
    LSDmatcher(float nnratio=0.6, bool checkOri=true);
    int SearchByDescriptor(KeyFrame* pKF, Frame &currentF, std::vector<MapLine*> &vpMapLineMatches);
    int SearchByDescriptor(KeyFrame* pKF, KeyFrame *pKF2, std::vector<MapLine*> &vpMapLineMatches);
    int SearchByProjection(KeyFrame* pKF,Frame &currentF, vector<MapLine*> &vpMapLineMatches);
    int SerachForInitialize(Frame &InitialFrame, Frame &CurrentFrame, std::vector<std::pair<int,int> > &LineMatches);
    int SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<std::pair<size_t, size_t> > &vMatchedPairs);
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b);
protected:
    float mfNNratio;
    bool mbCheckOrientation;
};
}
