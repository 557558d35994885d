// TEST INFRASTRUCTURE ONLY.  The vocabulary type named by include/ORBVocabulary.h:31.  Frame::ComputeBoW / KeyFrame::ComputeBoW call
// transform(); the harness fills mFeatVec itself (the DBoW2::FeatureVector is an INPUT of the compared matchers), so the member only
// has to exist.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <list>
#include <string>
#include <vector>
#include "BowVector.h"
#include "FeatureVector.h"

// DBoW2's own TemplatedVocabulary.h opens namespace std for every file that includes it, and the reference's headers rely on that
// (include/KeyFrameDatabase.h:66 spells `list` unqualified).
using namespace std;
namespace DBoW2 {
template <class TDescriptor, class F>
class TemplatedVocabulary {
public:
    void transform(const std::vector<TDescriptor>&, BowVector&, FeatureVector&, int) const {
        fprintf(stderr, "refstubs: DBoW2 vocabulary transform is not part of the compared path\n");
        abort();
    }
    double score(const BowVector&, const BowVector&) const { abort(); }
};
}  // namespace DBoW2
