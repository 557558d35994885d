// TEST INFRASTRUCTURE ONLY.  Interface of DBoW2::BowVector as the reference uses it (Thirdparty/DBoW2 is not in the reference
// tree; README.md:51 has users copy it from ORB_SLAM2): a std::map from word id to weight.
#pragma once
#include <map>
#include <vector>
namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
}  // namespace DBoW2
