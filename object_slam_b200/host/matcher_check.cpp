// Compile-and-run check of the C++ matcher mirror (host/ORBmatcher.h): mock Frame / MapPoint classes that carry the
// members of the reference's classes under the reference's names, filled from a binary case file written by
// tests/test_dropin_cpp.py; runs ORBmatcher::SearchByProjection(F, vpMapPoints, th) and SearchForInitialization and
// writes what they leave in the caller's containers.
#include "ORBmatcher.h"
#include <cstdio>
#include <cstdlib>
#include <map>

struct MapPoint {                              // include/MapPoint.h: the members the matcher reads
    bool mbTrackInView = false, bad = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0;
    int mnTrackScaleLevel = 0, nObs = 0, id = -1;
    cv::Mat mDescriptor;
    bool isBad() const { return bad; }
    int Observations() const { return nObs; }
    cv::Mat GetDescriptor() const { return mDescriptor; }
};

typedef std::map<unsigned int, std::vector<unsigned int> > FeatureVector;      // DBoW2::FeatureVector

struct Frame {                                 // include/Frame.h: the members the matcher reads
    FeatureVector mFeatVec;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    std::vector<float> mvuRight;
    std::vector<MapPoint*> mvpMapPoints;
    float mnMinX = 0, mnMaxX = 0, mnMinY = 0, mnMaxY = 0, fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0, mb = 0;
    int mnScaleLevels = 8;
    std::vector<float> mvScaleFactors;
    obs_frame_set* mpDevFrame = nullptr;
};

struct KeyFrame {                              // include/KeyFrame.h: the members SearchByBoW reads
    int N = 0;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
};

static FILE* g_in;
template <class T> static void rd(T* p, size_t n) { if (fread(p, sizeof(T), n, g_in) != n) { fprintf(stderr, "short read\n"); exit(2); } }

static void readFrame(Frame& F, std::vector<unsigned char>& descStore) {
    int hdr[2];
    rd(hdr, 2);                                // N, has uRight
    F.N = hdr[0];
    float fl[10];
    rd(fl, 10);
    F.mnMinX = fl[0]; F.mnMaxX = fl[1]; F.mnMinY = fl[2]; F.mnMaxY = fl[3]; F.fx = fl[4]; F.fy = fl[5]; F.cx = fl[6]; F.cy = fl[7]; F.mbf = fl[8]; F.mb = fl[9];
    F.mvScaleFactors.resize(8);
    rd(F.mvScaleFactors.data(), 8);
    F.mvKeysUn.resize(F.N);
    rd(reinterpret_cast<unsigned char*>(F.mvKeysUn.data()), (size_t)F.N * sizeof(cv::KeyPoint));
    descStore.resize((size_t)F.N * 32);
    rd(descStore.data(), descStore.size());
    F.mDescriptors = cv::Mat(F.N, 32, CV_8UC1, descStore.data(), 32);
    if (hdr[1]) { F.mvuRight.resize(F.N); rd(F.mvuRight.data(), F.N); }
    F.mvpMapPoints.assign(F.N, nullptr);
}

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s map|init case.bin out.bin\n", argv[0]); return 2; }
    g_in = fopen(argv[2], "rb");
    if (!g_in) return 2;
    FILE* out = fopen(argv[3], "wb");
    std::string mode = argv[1];
    if (mode == "map") {
        Frame F;
        std::vector<unsigned char> ds;
        readFrame(F, ds);
        ORB_SLAM2::UploadFrame(F);
        int M; float th, ratio;
        rd(&M, 1); rd(&th, 1); rd(&ratio, 1);
        std::vector<MapPoint> pts(M);
        std::vector<unsigned char> inView(M), desc((size_t)M * 32);
        std::vector<float> px(M), py(M), pxr(M), vc(M);
        std::vector<int> lvl(M), obs(M);
        rd(inView.data(), M); rd(px.data(), M); rd(py.data(), M); rd(pxr.data(), M); rd(lvl.data(), M); rd(vc.data(), M);
        rd(desc.data(), desc.size()); rd(obs.data(), M);
        std::vector<MapPoint*> vp(M);
        for (int i = 0; i < M; i++) {
            MapPoint& p = pts[i];
            p.mbTrackInView = inView[i]; p.mTrackProjX = px[i]; p.mTrackProjY = py[i]; p.mTrackProjXR = pxr[i];
            p.mnTrackScaleLevel = lvl[i]; p.mTrackViewCos = vc[i]; p.nObs = obs[i]; p.id = i;
            p.mDescriptor = cv::Mat(1, 32, CV_8UC1, &desc[(size_t)i * 32], 32);
            vp[i] = &p;
        }
        ORB_SLAM2::ORBmatcher matcher(ratio);
        const int n = matcher.SearchByProjection(F, vp, th);
        std::vector<int> res(F.N);
        for (int k = 0; k < F.N; k++) res[k] = F.mvpMapPoints[k] ? F.mvpMapPoints[k]->id : -1;
        fwrite(&n, 4, 1, out);
        fwrite(res.data(), 4, res.size(), out);
        printf("SearchByProjection: %d matches, DescriptorDistance(d0,d1)=%d\n", n,
               ORB_SLAM2::ORBmatcher::DescriptorDistance(pts[0].mDescriptor, pts[1].mDescriptor));
    } else if (mode == "bow") {
        // keyframe side: N, keys, descriptors, valid flags, node of every keypoint; then the frame side without the flags
        KeyFrame KF; Frame F;
        std::vector<unsigned char> dk, df;
        std::vector<MapPoint> pts;
        float ratio;
        rd(&ratio, 1);
        for (int side = 0; side < 2; side++) {
            int N; rd(&N, 1);
            std::vector<cv::KeyPoint> keys(N);
            rd(reinterpret_cast<unsigned char*>(keys.data()), (size_t)N * sizeof(cv::KeyPoint));
            std::vector<unsigned char>& ds = side ? df : dk;
            ds.resize((size_t)N * 32); rd(ds.data(), ds.size());
            std::vector<unsigned char> valid(N); rd(valid.data(), N);
            std::vector<int> node(N); rd(node.data(), N);
            FeatureVector fv;
            for (int i = 0; i < N; i++) fv[(unsigned)node[i]].push_back((unsigned)i);
            if (side == 0) {
                KF.N = N; KF.mvKeysUn = keys; KF.mDescriptors = cv::Mat(N, 32, CV_8UC1, ds.data(), 32); KF.mFeatVec = fv;
                pts.resize(N); KF.mvpMapPoints.assign(N, nullptr);
                for (int i = 0; i < N; i++) { pts[i].id = i; if (valid[i]) KF.mvpMapPoints[i] = &pts[i]; }
            } else {
                F.N = N; F.mvKeysUn = keys; F.mDescriptors = cv::Mat(N, 32, CV_8UC1, ds.data(), 32); F.mFeatVec = fv;
            }
        }
        ORB_SLAM2::ORBmatcher matcher(ratio, true);
        std::vector<MapPoint*> vpMatches;
        const int n = matcher.SearchByBoW(&KF, F, vpMatches);
        std::vector<int> res(F.N);
        for (int i = 0; i < F.N; i++) res[i] = vpMatches[i] ? vpMatches[i]->id : -1;
        fwrite(&n, 4, 1, out);
        fwrite(res.data(), 4, res.size(), out);
        printf("SearchByBoW: %d matches\n", n);
    } else {
        Frame F1, F2;
        std::vector<unsigned char> d1, d2;
        readFrame(F1, d1); readFrame(F2, d2);
        ORB_SLAM2::UploadFrame(F1); ORB_SLAM2::UploadFrame(F2);
        int window; float ratio;
        rd(&window, 1); rd(&ratio, 1);
        std::vector<cv::Point2f> prev(F1.N);
        rd(reinterpret_cast<float*>(prev.data()), (size_t)F1.N * 2);
        std::vector<int> m12;
        ORB_SLAM2::ORBmatcher matcher(ratio, true);
        const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window);
        fwrite(&n, 4, 1, out);
        fwrite(m12.data(), 4, m12.size(), out);
        fwrite(prev.data(), 8, prev.size(), out);
        printf("SearchForInitialization: %d matches\n", n);
    }
    fclose(out);
    return 0;
}
