// Drop-in replacement for the reference's include/ORBextractor.h (class ORB_SLAM2::ORBextractor,
// include/ORBextractor.h:45-111).  Same namespace, class name, constructor, operator(), getters and
// public mvImagePyramid member, so Frame.cc / Tracking.cc compile against it unchanged; the work is
// forwarded to libobslam_b200.so (include/obslam_b200.h).  There is no CPU implementation here:
// if the library reports an error the call throws std::runtime_error.
//
// Build-time switches:
//   OBS_MAX_WIDTH / OBS_MAX_HEIGHT  upper bound of the image size the device buffers are sized for
//                                   (default 2048 x 1536; buffers grow on demand anyway)
//   OBS_DOWNLOAD_PYRAMID            1 = copy every pyramid level back into mvImagePyramid after each call
//                                   (only needed by code that still reads the pixels on the host, i.e. a
//                                   Frame::ComputeStereoMatches that has not been switched to
//                                   ORB_SLAM2::ComputeStereoMatchesB200); 0 (default) = mvImagePyramid[l]
//                                   carries the right rows/cols but unspecified pixels.
#ifndef OBS_B200_ORBEXTRACTOR_H
#define OBS_B200_ORBEXTRACTOR_H

#include <vector>
#include <list>
#include <opencv/cv.h>

struct obs_extractor;

namespace ORB_SLAM2
{

class ORBextractor
{
public:
    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    void operator()( cv::InputArray image, cv::InputArray mask,
      std::vector<cv::KeyPoint>& keypoints,
      cv::OutputArray descriptors);

    int inline GetLevels(){ return nlevels; }
    float inline GetScaleFactor(){ return scaleFactor; }
    std::vector<float> inline GetScaleFactors(){ return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors(){ return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares(){ return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares(){ return mvInvLevelSigma2; }

    std::vector<cv::Mat> mvImagePyramid;

    // The device handle (for ORB_SLAM2::ComputeStereoMatchesB200 and the matcher entry points).
    obs_extractor* handle() const { return mpHandle; }

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;

    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor;
    std::vector<float> mvInvScaleFactor;
    std::vector<float> mvLevelSigma2;
    std::vector<float> mvInvLevelSigma2;

    obs_extractor* mpHandle;
};

// Body of Frame::ComputeStereoMatches (src/Frame.cc:706-880) on the device: uses the pyramids,
// keypoints and descriptors the two extractors left in HBM by their last operator() calls.
// minD / maxD: the reference computes them from members it has not initialised yet (:736);
// pass 0 and fx (= mbf / mb).  mvuRight / mvDepth are resized to the left keypoint count.
void ComputeStereoMatchesB200(ORBextractor* left, ORBextractor* right, float mbf, float minD, float maxD,
                              int nLeft, std::vector<float>& mvuRight, std::vector<float>& mvDepth);

} //namespace ORB_SLAM

#endif
