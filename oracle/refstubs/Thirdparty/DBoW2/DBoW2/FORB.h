// TEST INFRASTRUCTURE ONLY.  The descriptor traits type named by include/ORBVocabulary.h:31.
#pragma once
#include <opencv2/core/core.hpp>
namespace DBoW2 {
class FORB {
public:
    typedef cv::Mat TDescriptor;
    typedef const TDescriptor* pDescriptor;
    static const int L = 32;
};
}  // namespace DBoW2
