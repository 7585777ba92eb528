// Stand-in for the reference's include/ORBmatcher.h (lines 36-101): only the entry points host/matcher_b200.cc defines, with the
// reference's parameter TYPES (names and layout are this file's own).  Compile-check use only, see README.md here.
#pragma once
#include <set>
#include <vector>
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
namespace StructureSLAM { class ORBmatcher { public:
    ORBmatcher(float ratio = 0.6, bool orientation = true);
    static int DescriptorDistance(const cv::Mat& first, const cv::Mat& second);
    int SearchByProjection(Frame& cur, const Frame& last, const float radius, const bool mono);
    int SearchByBoW(KeyFrame* kf, Frame& frame, std::vector<MapPoint*>& out);
    int SearchByBoW(KeyFrame* kf_a, KeyFrame* kf_b, std::vector<MapPoint*>& out);
    int SearchForTriangulation(KeyFrame* kf_a, KeyFrame* kf_b, cv::Mat fundamental, std::vector<pair<size_t, size_t> >& out, const bool stereo_only);
  protected: float mfNNratio; bool mbCheckOrientation; }; }
