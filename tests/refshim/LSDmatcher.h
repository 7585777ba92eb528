// Stand-in for the reference's include/LSDmatcher.h (lines 36-64): only the entry points host/matcher_b200.cc defines, with the
// reference's parameter TYPES (names and layout are this file's own).  Compile-check use only, see README.md here.
#pragma once
#include <vector>
#include "Frame.h"
#include "KeyFrame.h"
#include "MapLine.h"
namespace StructureSLAM { class LSDmatcher { public:
    LSDmatcher(float ratio = 0.6, bool orientation = true);
    static int DescriptorDistance(const cv::Mat& first, const cv::Mat& second);
    int SearchByDescriptor(KeyFrame* kf, Frame& frame, std::vector<MapLine*>& out);
    int SearchByDescriptor(KeyFrame* kf_a, KeyFrame* kf_b, std::vector<MapLine*>& out);
    int SearchByProjection(KeyFrame* kf, Frame& frame, vector<MapLine*>& out);
    int SerachForInitialize(Frame& first, Frame& second, std::vector<std::pair<int, int> >& out);
    int SearchForTriangulation(KeyFrame* kf_a, KeyFrame* kf_b, std::vector<std::pair<size_t, size_t> >& out);
  protected: float mfNNratio; bool mbCheckOrientation; }; }
