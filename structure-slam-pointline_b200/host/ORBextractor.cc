// host/ORBextractor.cc — adapter from the reference's ORBextractor class surface to the C ABI.
// Replaces src/ORBextractor.cc of the reference in its CMakeLists.txt source list.
#include "ORBextractor.h"
#include <cassert>
#include <cstring>
#include <stdexcept>
#include <string>

namespace StructureSLAM
{

static void Check(int rc, const char* what)
{
    if(rc != SSLPL_OK)
        throw std::runtime_error(std::string(what) + ": " + sslpl_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST):
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels),
    iniThFAST(_iniThFAST), minThFAST(_minThFAST), mHandle(NULL), mMaxW(0), mMaxH(0)
{
    mvImagePyramid.resize(nlevels);
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    // the tables Frame.cc:77-83 reads; the device workspace is created by the first operator() call, for that frame size
    Check(sslpl_orb_tables_host(nfeatures, (float)scaleFactor, nlevels, &mvScaleFactor[0], &mvInvScaleFactor[0], &mvLevelSigma2[0],
                                &mvInvLevelSigma2[0], NULL, NULL), "sslpl_orb_tables_host");
}

ORBextractor::~ORBextractor()
{
    sslpl_orb_destroy(mHandle);
}

void ORBextractor::EnsureHandle(int w, int h)
{
    if(mHandle && w <= mMaxW && h <= mMaxH)
        return;
    if(mHandle)
        sslpl_orb_destroy(mHandle);
    mHandle = NULL;
    sslpl_orb_params p;
    p.nfeatures = nfeatures; p.scaleFactor = (float)scaleFactor; p.nlevels = nlevels;
    p.iniThFAST = iniThFAST; p.minThFAST = minThFAST;
    p.max_width = w > mMaxW ? w : mMaxW; p.max_height = h > mMaxH ? h : mMaxH; p.max_batch = 1; p.device = sslpl_default_device();
    Check(sslpl_orb_create(&p, &mHandle), "sslpl_orb_create");
    mMaxW = p.max_width; mMaxH = p.max_height;
    mKpBuf.resize(sslpl_orb_max_keypoints(mHandle));
}

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                      cv::OutputArray _descriptors)
{
    if(_image.empty())
        return;                                            // ORBextractor.cc:1046

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1 );                      // ORBextractor.cc:1050

    EnsureHandle(image.cols, image.rows);
    const int cap = (int)mKpBuf.size();
    cv::Mat desc(cap, 32, CV_8U);
    int n = 0;
    Check(sslpl_orb_extract(mHandle, image.data, image.cols, image.rows, (int)image.step, &mKpBuf[0], desc.data, cap, &n),
          "sslpl_orb_extract");

    if(n == 0)
        _descriptors.release();                            // ORBextractor.cc:1064-1065
    else
    {
        _descriptors.create(n, 32, CV_8U);                 // ORBextractor.cc:1068
        cv::Mat out = _descriptors.getMat();
        for(int i=0; i<n; i++)
            memcpy(out.ptr(i), desc.ptr(i), 32);
    }

    _keypoints.clear();                                    // ORBextractor.cc:1072
    _keypoints.reserve(n);
    for(int i=0; i<n; i++)
    {
        const sslpl_keypoint &k = mKpBuf[i];
        _keypoints.push_back(cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id));
    }
}

void ORBextractor::FetchPyramid()
{
    for(int level = 0; level < nlevels; ++level)
    {
        int w = 0, h = 0;
        if(sslpl_orb_level_size(mHandle, level, &w, &h) != SSLPL_OK)
            return;
        const int B = 19;                                  // EDGE_THRESHOLD, ORBextractor.cc:74
        cv::Mat temp(h + 2*B, w + 2*B, CV_8UC1);
        Check(sslpl_orb_download_level(mHandle, 0, level, 1, temp.data, (int)temp.step), "sslpl_orb_download_level");
        mvImagePyramid[level] = temp(cv::Rect(B, B, w, h)); // ORBextractor.cc:1115
    }
}

} //namespace StructureSLAM
