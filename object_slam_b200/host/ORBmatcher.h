// Host-side mirror of the reference's ORB_SLAM2::ORBmatcher (include/ORBmatcher.h:37-102) over the C ABI
// (include/obslam_b200.h).  Same class name, constructor, constants and method names / argument order; the methods
// are templates over the Frame / MapPoint types because the reference's own Frame.h cannot be included on its own
// (it pulls in DBoW2, g2o, Eigen and the object layer).  Inside the reference tree they instantiate with
// ORB_SLAM2::Frame and ORB_SLAM2::MapPoint: every member they touch carries the reference's name
// (mvKeysUn, mDescriptors, mvuRight, mvpMapPoints, mnMinX..., mbTrackInView, mTrackProjX, GetDescriptor(), ...).
// Each method gathers those members into flat arrays (a few kB), makes one C-ABI call and scatters the result
// back into the caller's containers -- the matching itself runs in csrc/matcher.cu; there is no CPU search here.
//
// The Frame type additionally carries   obs_frame_set* mpDevFrame   (its keypoints, descriptors, mvuRight and
// grid in HBM), created once per frame with UploadFrame() right after Frame::AssignFeaturesToGrid().
#ifndef OBS_B200_ORBMATCHER_H
#define OBS_B200_ORBMATCHER_H

#include <cstring>
#include <stdint.h>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv/cv.h>

#include "obslam_b200.h"

namespace ORB_SLAM2
{

inline void ObsCheck(int rc) { if (rc != OBS_OK) throw std::runtime_error(std::string("obslam_b200: ") + obs_last_error()); }

// One matcher handle (stream + workspace) per thread, like the reference's stack-local ORBmatcher objects; destroyed when
// the thread exits.  A frame uploaded on one thread (Tracking) can be searched from any other (LocalMapping, LoopClosing):
// the library orders a foreign matcher's stream behind the frame set's last build.
struct ThreadMatcherHolder
{
    obs_matcher* m;
    ThreadMatcherHolder(): m(NULL) {}
    ~ThreadMatcherHolder() { if(m) obs_matcher_destroy(m); }
};
inline obs_matcher* ThreadMatcher()
{
    static thread_local ThreadMatcherHolder h;
    if(!h.m) ObsCheck(obs_matcher_create(0, &h.m));
    return h.m;
}

// Frame -> device, Frame.cc:103/164/224 (after AssignFeaturesToGrid; the device rebuilds the grid itself).
template<class FrameT>
void UploadFrame(FrameT &F)
{
    obs_frame_params fp;
    memset(&fp, 0, sizeof(fp));
    fp.min_x = F.mnMinX; fp.max_x = F.mnMaxX; fp.min_y = F.mnMinY; fp.max_y = F.mnMaxY;
    fp.fx = F.fx; fp.fy = F.fy; fp.cx = F.cx; fp.cy = F.cy; fp.mbf = F.mbf; fp.mb = F.mb;
    fp.nlevels = F.mnScaleLevels;
    for(int i=0; i<F.mnScaleLevels; i++) fp.scale_factors[i] = F.mvScaleFactors[i];
    if(!F.mpDevFrame) ObsCheck(obs_frame_set_create(ThreadMatcher(), &fp, 1, F.N > 0 ? F.N : 1, &F.mpDevFrame));
    std::vector<unsigned char> desc((size_t)F.N*32);
    for(int i=0; i<F.N; i++) memcpy(&desc[(size_t)i*32], F.mDescriptors.ptr(i), 32);
    obs_frame_view fv;
    fv.n = F.N;
    fv.keys_un = reinterpret_cast<const obs_keypoint*>(F.mvKeysUn.data());      // sizeof(cv::KeyPoint) == sizeof(obs_keypoint)
    fv.descriptors = desc.data();
    fv.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
    ObsCheck(obs_frame_set_upload(F.mpDevFrame, &fv, 1));
}

class ORBmatcher
{
public:
    ORBmatcher(float nnratio=0.6, bool checkOri=true): mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

    // include/ORBmatcher.h:44
    static int DescriptorDistance(const cv::Mat &a, const cv::Mat &b)
    {
        int32_t d = 0;
        ObsCheck(obs_descriptor_distance(ThreadMatcher(), a.ptr(0), b.ptr(0), 1, &d));
        return d;
    }

    // include/ORBmatcher.h:48, src/ORBmatcher.cc:45-129
    template<class FrameT, class MapPointT>
    int SearchByProjection(FrameT &F, const std::vector<MapPointT*> &vpMapPoints, const float th=3)
    {
        const int M = (int)vpMapPoints.size();
        std::vector<uint8_t> inView(M), desc((size_t)M*32);
        std::vector<float> px(M), py(M), pxr(M), vc(M);
        std::vector<int32_t> lvl(M), obs(M);
        for(int i=0; i<M; i++)
        {
            MapPointT* pMP = vpMapPoints[i];
            inView[i] = pMP->mbTrackInView && !pMP->isBad();
            px[i] = pMP->mTrackProjX; py[i] = pMP->mTrackProjY; pxr[i] = pMP->mTrackProjXR;
            lvl[i] = pMP->mnTrackScaleLevel; vc[i] = pMP->mTrackViewCos; obs[i] = pMP->Observations();
            const cv::Mat d = pMP->GetDescriptor();
            memcpy(&desc[(size_t)i*32], d.ptr(0), 32);
        }
        std::vector<int32_t> kpObs(F.N), kpMatch(F.N);
        for(int k=0; k<F.N; k++) kpObs[k] = F.mvpMapPoints[k] ? F.mvpMapPoints[k]->Observations() : 0;
        obs_mappoint_view v;
        v.n = M; v.per_frame = 0;
        v.in_view = inView.data(); v.proj_x = px.data(); v.proj_y = py.data(); v.proj_xr = pxr.data();
        v.scale_level = lvl.data(); v.view_cos = vc.data(); v.descriptors = desc.data(); v.observations = obs.data();
        int32_t n = 0;
        // the frame set was sized for F.N keypoints: kp arrays are 1 x capacity, capacity >= F.N (rounded up to 32)
        std::vector<int32_t> obsPad(Capacity(F.N), 0), matchPad(Capacity(F.N), -1);
        std::copy(kpObs.begin(), kpObs.end(), obsPad.begin());
        ObsCheck(obs_search_by_projection(ThreadMatcher(), F.mpDevFrame, &v, th, mfNNratio, obsPad.data(), matchPad.data(), &n));
        for(int k=0; k<F.N; k++)
            if(matchPad[k]>=0) F.mvpMapPoints[k] = vpMapPoints[matchPad[k]];
        return n;
    }

    // include/ORBmatcher.h:65, src/ORBmatcher.cc:405-520
    template<class FrameT>
    int SearchForInitialization(FrameT &F1, FrameT &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize=10)
    {
        const int c1 = Capacity(F1.N);
        std::vector<float> prev((size_t)c1*2, 0.f);
        for(int i=0; i<F1.N; i++) { prev[2*i] = vbPrevMatched[i].x; prev[2*i+1] = vbPrevMatched[i].y; }
        std::vector<int32_t> m12(c1, -1);
        int32_t n = 0;
        ObsCheck(obs_search_for_initialization(ThreadMatcher(), F1.mpDevFrame, F2.mpDevFrame, prev.data(), m12.data(), windowSize,
                                               mfNNratio, mbCheckOrientation, &n));
        vnMatches12.assign(m12.begin(), m12.begin()+F1.N);
        for(int i=0; i<F1.N; i++) { vbPrevMatched[i].x = prev[2*i]; vbPrevMatched[i].y = prev[2*i+1]; }
        return n;
    }

    // include/ORBmatcher.h:54, src/ORBmatcher.cc:159-288.  KeyFrameT / FrameT carry mFeatVec (DBoW2::FeatureVector, any
    // std::map<NodeId, std::vector<unsigned int> >-like container), mDescriptors, mvKeysUn, N; the keyframe GetMapPointMatches().
    template<class KeyFrameT, class FrameT, class MapPointT>
    int SearchByBoW(KeyFrameT* pKF, FrameT &F, std::vector<MapPointT*> &vpMapPointMatches)
    {
        const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPointT*>(F.N, static_cast<MapPointT*>(NULL));
        if(pKF->N <= 0 || F.N <= 0)       // nothing to walk: the reference's loops fall through and return 0
            return 0;
        BowCsr a, b;
        FeatVecToCsr(pKF->mFeatVec, a); FeatVecToCsr(F.mFeatVec, b);
        const int32_t nA = pKF->N, nB = F.N;
        std::vector<uint8_t> valid(nA > 0 ? nA : 1), dA((size_t)(nA > 0 ? nA : 1)*32), dB((size_t)(nB > 0 ? nB : 1)*32);
        for(int i=0; i<nA; i++) { valid[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad(); memcpy(&dA[(size_t)i*32], pKF->mDescriptors.ptr(i), 32); }
        for(int i=0; i<nB; i++) memcpy(&dB[(size_t)i*32], F.mDescriptors.ptr(i), 32);
        obs_bow_side s1 = MakeSide(nA, a, dA, pKF->mvKeysUn, valid.data());
        obs_bow_side s2 = MakeSide(nB, b, dB, F.mvKeysUn, nullptr);
        std::vector<int32_t> m12(s1.cap), m21(s2.cap);
        int32_t n = 0;
        ObsCheck(obs_search_by_bow(ThreadMatcher(), &s1, &s2, 1, TH_LOW, 0, mfNNratio, mbCheckOrientation, m12.data(), m21.data(), &n));
        for(int i=0; i<F.N; i++) if(m21[i]>=0) vpMapPointMatches[i] = vpMapPointsKF[m21[i]];
        return n;
    }

    static const int TH_LOW = 50;
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    static int Capacity(int n) { return ((n > 0 ? n : 1) + 31) & ~31; }

    // DBoW2::FeatureVector -> CSR (node ids ascending = the map's iteration order)
    struct BowCsr { std::vector<uint32_t> id; std::vector<int32_t> start, idx; int32_t n, nNodes; };
    template<class FeatVecT>
    static void FeatVecToCsr(const FeatVecT &fv, BowCsr &c)
    {
        c.start.assign(1, 0);
        for(typename FeatVecT::const_iterator it=fv.begin(); it!=fv.end(); ++it)
        {
            c.id.push_back((uint32_t)it->first);
            c.idx.insert(c.idx.end(), it->second.begin(), it->second.end());
            c.start.push_back((int32_t)c.idx.size());
        }
        c.nNodes = (int32_t)c.id.size();
        if(c.id.empty()) { c.id.push_back(0); c.start.push_back(0); }
        if(c.idx.empty()) c.idx.push_back(0);
    }
    static obs_bow_side MakeSide(const int32_t &n, BowCsr &c, const std::vector<uint8_t> &desc, const std::vector<cv::KeyPoint> &keys,
                                 const uint8_t* valid)
    {
        c.n = n;
        c.idx.resize(n > 0 ? n : 1, 0);
        obs_bow_side s;
        s.cap = n > 0 ? n : 1; s.node_cap = c.nNodes > 0 ? c.nNodes : 1;
        s.n = &c.n; s.descriptors = desc.data(); s.keys_un = reinterpret_cast<const obs_keypoint*>(keys.data());
        s.valid = valid; s.u_right = nullptr; s.n_nodes = &c.nNodes;
        s.node_id = c.id.data(); s.node_start = c.start.data(); s.node_idx = c.idx.data();
        return s;
    }

    // src/ORBmatcher.cc:1601-1642 over the bin sizes
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int &ind1, int &ind2, int &ind3)
    {
        std::vector<int32_t> sizes(L), ind(3);
        for(int i=0; i<L; i++) sizes[i] = (int)histo[i].size();
        ObsCheck(obs_compute_three_maxima(ThreadMatcher(), sizes.data(), 1, L, ind.data()));
        ind1 = ind[0]; ind2 = ind[1]; ind3 = ind[2];
    }

    float mfNNratio;
    bool mbCheckOrientation;
};

}// namespace ORB_SLAM

#endif
