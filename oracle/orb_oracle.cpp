// TEST INFRASTRUCTURE ONLY -- CPU oracle for the ORB front end (see orc_primitives.h).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may build, load or call this file.  The product path (object_slam_b200/csrc) never does.
//
// Restates, over flat arrays and with no OpenCV dependency, the reference's
//   ORBextractor::ORBextractor            /root/reference/src/ORBextractor.cc:410-470
//   ORBextractor::operator()              :1043-1105
//   ORBextractor::ComputePyramid          :1107-1132
//   ORBextractor::ComputeKeyPointsOctTree :765-853
//   ExtractorNode::DivideNode             :481-537
//   ORBextractor::DistributeOctTree       :539-763
//   IC_Angle / computeOrientation         :77-104, :472-479
//   computeOrbDescriptor / computeDescriptors :108-147, :1034-1041
// Parity status: the reference ships no tests or golden vectors for this path ("parity
// unpinned" by upstream).  This restatement is pinned two ways instead:
//   (1) its OpenCV primitives are checked bit-for-bit against cv2 4.13.0 (tests/),
//   (2) its end-to-end output is checked field-for-field against the reference's own
//       unmodified src/ORBextractor.cc compiled from /root/reference into oracle/_ref
//       (oracle/Makefile; bump allocator => node-pointer ties resolve by creation order).
#include "orc_primitives.h"
#include <cstring>
#include <cstdio>
#include <list>
#include <utility>

namespace orc {

static const int kPattern[256 * 4] = {
#include "../object_slam_b200/csrc/orb_pattern.inc"
};

static const int PATCH_SIZE = 31;
static const int HALF_PATCH_SIZE = 15;
static const int EDGE_THRESHOLD = 19;

struct KeyPt {            // layout == cv::KeyPoint (28 bytes)
    float x, y, size, angle, response;
    int octave, class_id;
};

struct Image {
    int w = 0, h = 0;
    std::vector<uint8_t> px;
    const uint8_t* row(int y) const { return px.data() + (size_t)y * w; }
    uint8_t* row(int y) { return px.data() + (size_t)y * w; }
};

struct Extractor {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;          // the reference stores the float ctor arg in a double member (ORBextractor.h:97)
    std::vector<float> scale, invScale, sigma2, invSigma2;
    std::vector<int> featPerLevel;
    std::vector<int> umax;
    // results of the last extract()
    std::vector<Image> pyr, blur;
    std::vector<std::vector<KeyPt>> cand, sel;   // per level: FAST candidates (border-relative), selected (level coords, with angle)
    std::vector<KeyPt> kps;
    std::vector<uint8_t> desc;
};

// :410-470
static Extractor* make_extractor(int nfeatures, float scaleFactorF, int nlevels, int iniTh, int minTh) {
    Extractor* e = new Extractor;
    e->nfeatures = nfeatures; e->nlevels = nlevels; e->iniTh = iniTh; e->minTh = minTh;
    e->scaleFactor = scaleFactorF;
    e->scale.resize(nlevels); e->sigma2.resize(nlevels);
    e->scale[0] = 1.0f; e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);   // float*double -> double -> float
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    e->invScale.resize(nlevels); e->invSigma2.resize(nlevels);
    for (int i = 0; i < nlevels; i++) {
        e->invScale[i] = 1.0f / e->scale[i];
        e->invSigma2[i] = 1.0f / e->sigma2[i];
    }
    e->featPerLevel.resize(nlevels);
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = (float)(nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        e->featPerLevel[l] = cv_round(nDesired);
        sum += e->featPerLevel[l];
        nDesired *= factor;
    }
    e->featPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);

    e->umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH_SIZE * sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

// :1107-1132.  The 19-px reflect border the reference adds is never read by any later
// stage (FAST cells start at x=16 and FAST itself keeps a 3-px margin; patches have radius
// <= 15 around keypoints >= 19 px from the edge; the blur runs on a border-less clone), so
// the oracle stores border-less levels.
static void compute_pyramid(Extractor* e, const uint8_t* img, int w, int h, size_t stride) {
    e->pyr.assign(e->nlevels, Image());
    for (int l = 0; l < e->nlevels; l++) {
        float s = e->invScale[l];
        int lw = cv_round((float)w * s), lh = cv_round((float)h * s);
        Image& L = e->pyr[l];
        L.w = lw; L.h = lh; L.px.resize((size_t)lw * lh);
        if (l == 0) {
            for (int y = 0; y < h; y++) memcpy(L.row(y), img + (size_t)y * stride, w);
        } else {
            const Image& P = e->pyr[l - 1];
            resize_linear_u8(P.px.data(), P.w, P.h, P.w, L.px.data(), lw, lh, lw);
        }
    }
}

// ---- quadtree (:481-763) ------------------------------------------------------------
struct Node {
    std::vector<KeyPt> keys;
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::list<Node>::iterator lit;
    bool noMore = false;
    long seq = 0;     // creation sequence: stands in for the heap address the reference sorts on
};

static void divide_node(const Node& p, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)ceil(static_cast<float>(p.URx - p.ULx) / 2);
    const int halfY = (int)ceil(static_cast<float>(p.BRy - p.ULy) / 2);
    n1.ULx = p.ULx; n1.ULy = p.ULy;
    n1.URx = p.ULx + halfX; n1.URy = p.ULy;
    n1.BLx = p.ULx; n1.BLy = p.ULy + halfY;
    n1.BRx = p.ULx + halfX; n1.BRy = p.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy;
    n2.URx = p.URx; n2.URy = p.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy;
    n2.BRx = p.URx; n2.BRy = p.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy;
    n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = p.BLx; n3.BLy = p.BLy;
    n3.BRx = n1.BRx; n3.BRy = p.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy;
    n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy;
    n4.BRx = p.BRx; n4.BRy = p.BRy;
    for (size_t i = 0; i < p.keys.size(); i++) {
        const KeyPt& kp = p.keys[i];
        if (kp.x < n1.URx) {
            if (kp.y < n1.BRy) n1.keys.push_back(kp); else n3.keys.push_back(kp);
        } else if (kp.y < n1.BRy) n2.keys.push_back(kp);
        else n4.keys.push_back(kp);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

typedef std::pair<int, std::pair<long, Node*>> SizeSeqNode;   // (size, (seq, node)): sorts like (size, address) under a bump allocator

static std::vector<KeyPt> distribute_octtree(const std::vector<KeyPt>& in, int minX, int maxX, int minY, int maxY, int N) {
    const int nIni = (int)round(static_cast<float>(maxX - minX) / (maxY - minY));
    std::vector<KeyPt> result;
    if (nIni < 1) return result;            // reference indexes an empty vector here (UB); the ABI rejects such shapes
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> nodes;
    std::vector<Node*> ini(nIni);
    long seq = 0;
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
        ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        ni.seq = seq++;
        nodes.push_back(ni);
        ini[i] = &nodes.back();
    }
    for (size_t i = 0; i < in.size(); i++) {
        size_t idx = (size_t)(in[i].x / hX);
        if (idx >= (size_t)nIni) idx = nIni - 1;     // unreachable for extractor inputs; reference would be UB
        ini[idx]->keys.push_back(in[i]);
    }
    auto lit = nodes.begin();
    while (lit != nodes.end()) {
        if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
        else if (lit->keys.empty()) lit = nodes.erase(lit);
        else lit++;
    }
    bool finish = false;
    std::vector<SizeSeqNode> vSize;
    auto push_child = [&](Node& n, int* nToExpand) {
        if (n.keys.size() > 0) {
            n.seq = seq++;
            nodes.push_front(n);
            if (n.keys.size() > 1) {
                if (nToExpand) (*nToExpand)++;
                vSize.push_back(std::make_pair((int)n.keys.size(), std::make_pair(nodes.front().seq, &nodes.front())));
                nodes.front().lit = nodes.begin();
            }
        }
    };
    while (!finish) {
        int prevSize = (int)nodes.size();
        lit = nodes.begin();
        int nToExpand = 0;
        vSize.clear();
        while (lit != nodes.end()) {
            if (lit->noMore) { lit++; continue; }
            Node n1, n2, n3, n4;
            divide_node(*lit, n1, n2, n3, n4);
            push_child(n1, &nToExpand); push_child(n2, &nToExpand);
            push_child(n3, &nToExpand); push_child(n4, &nToExpand);
            lit = nodes.erase(lit);
        }
        if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) {
            finish = true;
        } else if (((int)nodes.size() + nToExpand * 3) > N) {
            while (!finish) {
                prevSize = (int)nodes.size();
                std::vector<SizeSeqNode> prev = vSize;
                vSize.clear();
                std::sort(prev.begin(), prev.end());
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    Node* p = prev[j].second.second;
                    divide_node(*p, n1, n2, n3, n4);
                    push_child(n1, nullptr); push_child(n2, nullptr);
                    push_child(n3, nullptr); push_child(n4, nullptr);
                    nodes.erase(p->lit);
                    if ((int)nodes.size() >= N) break;
                }
                if ((int)nodes.size() >= N || (int)nodes.size() == prevSize) finish = true;
            }
        }
    }
    result.reserve(nodes.size());
    for (auto it = nodes.begin(); it != nodes.end(); it++) {
        const std::vector<KeyPt>& k = it->keys;
        const KeyPt* best = &k[0];
        float maxR = best->response;
        for (size_t i = 1; i < k.size(); i++)
            if (k[i].response > maxR) { best = &k[i]; maxR = k[i].response; }
        result.push_back(*best);
    }
    return result;
}

// :77-104
static float ic_angle(const Image& im, float px, float py, const std::vector<int>& umax) {
    int m01 = 0, m10 = 0;
    int cx = cv_round(px), cy = cv_round(py);
    const uint8_t* c = im.row(cy) + cx;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m10 += u * c[u];
    int step = im.w;
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int vsum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int vp = c[u + v * step], vm = c[u - v * step];
            vsum += (vp - vm);
            m10 += u * (vp + vm);
        }
        m01 += v * vsum;
    }
    return fast_atan2((float)m01, (float)m10);
}

// :108-147
static const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
static void orb_descriptor(const KeyPt& kp, const Image& im, uint8_t* desc) {
    float angle = (float)kp.angle * factorPI;
    float a = cosf(angle), b = sinf(angle);
    const uint8_t* c = im.row(cv_round(kp.y)) + cv_round(kp.x);
    const int step = im.w;
    const int* pat = kPattern;
    for (int i = 0; i < 32; i++, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            int t0 = c[cv_round(x0 * b + y0 * a) * step + cv_round(x0 * a - y0 * b)];
            int t1 = c[cv_round(x1 * b + y1 * a) * step + cv_round(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// :765-853
static void compute_keypoints(Extractor* e) {
    e->cand.assign(e->nlevels, {});
    e->sel.assign(e->nlevels, {});
    const float W = 30;
    std::vector<FastKp> cell;
    for (int level = 0; level < e->nlevels; ++level) {
        const Image& im = e->pyr[level];
        const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = im.w - EDGE_THRESHOLD + 3, maxBY = im.h - EDGE_THRESHOLD + 3;
        std::vector<KeyPt>& vToDist = e->cand[level];
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols > 0 && nRows > 0 && maxBX > minBX && maxBY > minBY) {
            const int wCell = (int)ceil(width / nCols), hCell = (int)ceil(height / nRows);
            for (int i = 0; i < nRows; i++) {
                const float iniY = (float)(minBY + i * hCell);
                float maxY = iniY + hCell + 6;
                if (iniY >= maxBY - 3) continue;
                if (maxY > maxBY) maxY = (float)maxBY;
                for (int j = 0; j < nCols; j++) {
                    const float iniX = (float)(minBX + j * wCell);
                    float maxX = iniX + wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = (float)maxBX;
                    int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
                    fast9_16(im.row(y0) + x0, x1 - x0, y1 - y0, im.w, e->iniTh, true, cell);
                    if (cell.empty()) fast9_16(im.row(y0) + x0, x1 - x0, y1 - y0, im.w, e->minTh, true, cell);
                    for (const FastKp& k : cell) {
                        KeyPt kp;
                        kp.x = (float)k.x + j * wCell; kp.y = (float)k.y + i * hCell;
                        kp.size = 7.f; kp.angle = -1.f; kp.response = (float)k.score;
                        kp.octave = 0; kp.class_id = -1;
                        vToDist.push_back(kp);
                    }
                }
            }
        }
        std::vector<KeyPt>& kps = e->sel[level];
        kps = distribute_octtree(vToDist, minBX, maxBX, minBY, maxBY, e->featPerLevel[level]);
        const int scaledPatch = (int)(PATCH_SIZE * e->scale[level]);
        for (KeyPt& k : kps) {
            k.x += minBX; k.y += minBY; k.octave = level; k.size = (float)scaledPatch;
        }
    }
    for (int level = 0; level < e->nlevels; ++level)
        for (KeyPt& k : e->sel[level]) k.angle = ic_angle(e->pyr[level], k.x, k.y, e->umax);
}

// :1043-1105
static int extract(Extractor* e, const uint8_t* img, int w, int h, size_t stride) {
    e->kps.clear(); e->desc.clear();
    if (!img || w <= 0 || h <= 0) return 0;
    compute_pyramid(e, img, w, h, stride);
    compute_keypoints(e);
    int n = 0;
    for (int l = 0; l < e->nlevels; l++) n += (int)e->sel[l].size();
    e->desc.assign((size_t)n * 32, 0);
    e->kps.reserve(n);
    e->blur.assign(e->nlevels, Image());
    int offset = 0;
    for (int l = 0; l < e->nlevels; l++) {
        std::vector<KeyPt> kps = e->sel[l];
        if (kps.empty()) continue;
        Image& B = e->blur[l];
        B.w = e->pyr[l].w; B.h = e->pyr[l].h; B.px.resize((size_t)B.w * B.h);
        gaussian7x7_u8(e->pyr[l].px.data(), B.w, B.h, B.w, B.px.data(), B.w);
        for (size_t i = 0; i < kps.size(); i++) orb_descriptor(kps[i], B, &e->desc[(size_t)(offset + i) * 32]);
        offset += (int)kps.size();
        if (l != 0) {
            float s = e->scale[l];
            for (KeyPt& k : kps) { k.x *= s; k.y *= s; }
        }
        e->kps.insert(e->kps.end(), kps.begin(), kps.end());
    }
    return n;
}

}  // namespace orc

// ------------------------------- C interface for ctypes --------------------------------
using namespace orc;
extern "C" {

void* orc_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    return make_extractor(nfeatures, scaleFactor, nlevels, iniTh, minTh);
}
void orc_extractor_destroy(void* h) { delete (Extractor*)h; }

int orc_extract(void* h, const uint8_t* img, int w, int ht, size_t stride) { return extract((Extractor*)h, img, w, ht, stride); }

int orc_get_keypoints(void* h, KeyPt* kps, uint8_t* desc, int cap) {
    Extractor* e = (Extractor*)h;
    int n = std::min((int)e->kps.size(), cap);
    if (kps) memcpy(kps, e->kps.data(), (size_t)n * sizeof(KeyPt));
    if (desc) memcpy(desc, e->desc.data(), (size_t)n * 32);
    return (int)e->kps.size();
}
void orc_get_tables(void* h, float* scale, float* invScale, float* sigma2, float* invSigma2, int* featPerLevel, int* umax16) {
    Extractor* e = (Extractor*)h;
    for (int i = 0; i < e->nlevels; i++) {
        if (scale) scale[i] = e->scale[i];
        if (invScale) invScale[i] = e->invScale[i];
        if (sigma2) sigma2[i] = e->sigma2[i];
        if (invSigma2) invSigma2[i] = e->invSigma2[i];
        if (featPerLevel) featPerLevel[i] = e->featPerLevel[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}
void orc_get_level_dims(void* h, int level, int* w, int* ht) {
    Extractor* e = (Extractor*)h; *w = e->pyr[level].w; *ht = e->pyr[level].h;
}
// which: 0 = pyramid level, 1 = blurred level (empty -> returns 0)
int orc_get_level(void* h, int which, int level, uint8_t* dst) {
    Extractor* e = (Extractor*)h;
    const Image& im = which ? e->blur[level] : e->pyr[level];
    if (im.px.empty()) return 0;
    memcpy(dst, im.px.data(), im.px.size());
    return (int)im.px.size();
}
// which: 0 = FAST candidates (coords relative to the 16-px border origin, as handed to the
// quadtree), 1 = selected keypoints in level coordinates with angle.
int orc_get_level_keypoints(void* h, int which, int level, KeyPt* out, int cap) {
    Extractor* e = (Extractor*)h;
    const std::vector<KeyPt>& v = which ? e->sel[level] : e->cand[level];
    int n = std::min((int)v.size(), cap);
    if (out) memcpy(out, v.data(), (size_t)n * sizeof(KeyPt));
    return (int)v.size();
}

// primitives, for pinning against cv2 and for per-stage device parity tests
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t ss, uint8_t* dst, int dw, int dh, size_t ds) {
    resize_linear_u8(src, sw, sh, ss, dst, dw, dh, ds);
}
void orc_gaussian7x7_u8(const uint8_t* src, int w, int h, size_t ss, uint8_t* dst, size_t ds) { gaussian7x7_u8(src, w, h, ss, dst, ds); }
int orc_fast9_16(const uint8_t* img, int w, int h, size_t stride, int th, int nms, int* xys, int cap) {
    std::vector<FastKp> v;
    fast9_16(img, w, h, stride, th, nms != 0, v);
    int n = std::min((int)v.size(), cap);
    for (int i = 0; i < n; i++) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
    return (int)v.size();
}
int orc_fast9_16_simple(const uint8_t* img, int w, int h, size_t stride, int th, int nms, int* xys, int cap) {
    std::vector<FastKp> v;
    fast9_16_simple(img, w, h, stride, th, nms != 0, v);
    int n = std::min((int)v.size(), cap);
    for (int i = 0; i < n; i++) { xys[3 * i] = v[i].x; xys[3 * i + 1] = v[i].y; xys[3 * i + 2] = v[i].score; }
    return (int)v.size();
}
// FAST score map: max(contrast-1, 0) per pixel, 0 in the 3-px margin.
void orc_fast_score_map(const uint8_t* img, int w, int h, size_t stride, uint8_t* out) {
    memset(out, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            int c = fast_contrast(img + (size_t)y * stride + x, (ptrdiff_t)stride) - 1;
            out[(size_t)y * w + x] = (uint8_t)(c < 0 ? 0 : c);
        }
}
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
void orc_fast_atan2_n(const float* y, const float* x, float* out, long n) { for (long i = 0; i < n; i++) out[i] = fast_atan2(y[i], x[i]); }
void orc_sincosf_model_n(const float* a, float* s, float* c, long n) { for (long i = 0; i < n; i++) sincosf_model(a[i], &s[i], &c[i]); }
void orc_sincosf_libm_n(const float* a, float* s, float* c, long n) { for (long i = 0; i < n; i++) { s[i] = sinf(a[i]); c[i] = cosf(a[i]); } }
// Exhaustive model-vs-libm sweep over all binary32 values in [lo, hi]; returns mismatch count.
long orc_sincosf_sweep(float lo, float hi, long stride_ulps) {
    long bad = 0;
    uint32_t a = f32_bits(lo), b = f32_bits(hi);
    for (uint64_t u = a; u <= b; u += (uint64_t)stride_ulps) {
        uint32_t uu = (uint32_t)u; float x; memcpy(&x, &uu, 4);
        float s, c; sincosf_model(x, &s, &c);
        if (s != sinf(x) || c != cosf(x)) bad++;
    }
    return bad;
}
const int* orc_pattern() { return kPattern; }

}  // extern "C"
