// TEST INFRASTRUCTURE ONLY.  Interface of DBoW2::FeatureVector as the reference uses it: a std::map from vocabulary node id to the
// indices of the features under that node, iterated in ascending node order (src/ORBmatcher.cc:175-288).
#pragma once
#include "BowVector.h"
namespace DBoW2 {
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {
public:
    void addFeature(NodeId id, unsigned int i_feature) { (*this)[id].push_back(i_feature); }
};
}  // namespace DBoW2
