// See ORBextractor.h.  Replaces src/ORBextractor.cc in the reference's CMakeLists.txt source list
// (CMakeLists.txt:53-77) and links against libobslam_b200.so.
#include "ORBextractor.h"
#include "obslam_b200.h"

#include <stdexcept>
#include <string>
#include <cstring>

#ifndef OBS_MAX_WIDTH
#define OBS_MAX_WIDTH 2048
#endif
#ifndef OBS_MAX_HEIGHT
#define OBS_MAX_HEIGHT 1536
#endif
// 1 (default): mvImagePyramid holds the level pixels after every call, as in the reference -- an unmodified
// Frame::ComputeStereoMatches can keep reading them.  0: sizes only (saves the device-to-host copy) for trees whose
// ComputeStereoMatches forwards to ComputeStereoMatchesB200; the Mats are then zero-filled, never stale.
#ifndef OBS_DOWNLOAD_PYRAMID
#define OBS_DOWNLOAD_PYRAMID 1
#endif
#ifndef OBS_DEVICE
#define OBS_DEVICE 0
#endif

namespace ORB_SLAM2
{

static void obsCheck(int rc, const char* what)
{
    if(rc != OBS_OK)
        throw std::runtime_error(std::string(what) + ": " + obs_last_error());
}

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST):
    nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels),
    iniThFAST(_iniThFAST), minThFAST(_minThFAST), mpHandle(NULL)
{
    obs_orb_params prm;
    prm.nfeatures = _nfeatures; prm.scale_factor = _scaleFactor; prm.nlevels = _nlevels;
    prm.ini_th_fast = _iniThFAST; prm.min_th_fast = _minThFAST;
    obsCheck(obs_extractor_create(&prm, OBS_MAX_WIDTH, OBS_MAX_HEIGHT, 1, OBS_DEVICE, &mpHandle), "obs_extractor_create");

    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels);
    mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    obsCheck(obs_extractor_tables(mpHandle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                  mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "obs_extractor_tables");
    mvImagePyramid.resize(nlevels);
}

ORBextractor::~ORBextractor()
{
    obs_extractor_destroy(mpHandle);
}

void ORBextractor::operator()( cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                               cv::OutputArray _descriptors)
{
    if(_image.empty())
        return;

    cv::Mat image = _image.getMat();
    assert(image.type() == CV_8UC1 );

    const int cap = obs_extractor_max_keypoints(mpHandle);
    static_assert(sizeof(cv::KeyPoint) == sizeof(obs_keypoint), "cv::KeyPoint layout");
    _keypoints.resize(cap);
    cv::Mat all(cap, 32, CV_8U);
    int n = 0;
    obsCheck(obs_extract(mpHandle, image.data, image.cols, image.rows, image.step,
                         reinterpret_cast<obs_keypoint*>(_keypoints.data()), all.data, cap, &n), "obs_extract");
    _keypoints.resize(n);

    if( n == 0 )
        _descriptors.release();
    else
    {
        _descriptors.create(n, 32, CV_8U);
        cv::Mat descriptors = _descriptors.getMat();
        for(int i=0; i<n; i++)
            memcpy(descriptors.ptr(i), all.ptr(i), 32);
    }

    // mvImagePyramid: sizes always, pixels only on request
    for(int level=0; level<nlevels; ++level)
    {
        int w = 0, h = 0;
        obsCheck(obs_extractor_get_level(mpHandle, 0, level, 0, NULL, 0, &w, &h), "obs_extractor_get_level");
        mvImagePyramid[level].create(h, w, CV_8UC1);
#if OBS_DOWNLOAD_PYRAMID
        obsCheck(obs_extractor_get_level(mpHandle, 0, level, 0, mvImagePyramid[level].data, mvImagePyramid[level].step, &w, &h),
                 "obs_extractor_get_level");
#else
        memset(mvImagePyramid[level].data, 0, (size_t)mvImagePyramid[level].step * h);
#endif
    }
}

void ComputeStereoMatchesB200(ORBextractor* left, ORBextractor* right, float mbf, float minD, float maxD,
                              int nLeft, std::vector<float>& mvuRight, std::vector<float>& mvDepth)
{
    mvuRight = std::vector<float>(nLeft,-1.0f);
    mvDepth = std::vector<float>(nLeft,-1.0f);
    if(nLeft == 0)
        return;
    obsCheck(obs_stereo_match(left->handle(), right->handle(), mbf, minD, maxD, mvuRight.data(), mvDepth.data(), nLeft),
             "obs_stereo_match");
}

} //namespace ORB_SLAM
