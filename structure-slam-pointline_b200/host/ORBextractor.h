// host/ORBextractor.h — drop-in replacement of the reference's include/ORBextractor.h:45-111.
// Same namespace, class name, constructor and operator() signature, accessors and public mvImagePyramid, so that
// Frame.cc / Tracking.cc compile and link unchanged (Tracking.cc:119-120 constructs it, Frame.cc:158 calls it).
// All arithmetic runs in libsslpl_b200.so (include/sslpl.h); this class only marshals cv:: containers.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <opencv2/core/core.hpp>
#include "sslpl.h"

namespace StructureSLAM
{

class ORBextractor
{
public:
    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);

    int inline GetLevels(){ return nlevels; }
    float inline GetScaleFactor(){ return scaleFactor; }
    std::vector<float> inline GetScaleFactors(){ return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors(){ return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares(){ return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares(){ return mvInvLevelSigma2; }

    // Filled lazily: the device keeps the pyramid; call FetchPyramid() to materialise the bordered levels
    // (no code in the reference reads this member outside ORBextractor.cc).
    std::vector<cv::Mat> mvImagePyramid;
    void FetchPyramid();

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    sslpl_orb* mHandle;
    int mMaxW, mMaxH;
    std::vector<sslpl_keypoint> mKpBuf;
    void EnsureHandle(int w, int h);
};

} //namespace StructureSLAM

#endif
