/*
 * oracle/orb_oracle.cpp — CPU ORACLE for the ORB path.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates /root/reference/src/ORBextractor.cc (file:line cited per function) plus the OpenCV
 * primitives it delegates to (each pinned against cv2 4.13 in tests/).
 * Canonical choices where the reference is not self-consistent:
 *   - no FMA contraction (build with -ffp-contract=off),
 *   - cos/sin of the keypoint angle = correctly rounded f32 (double evaluation, narrowed),
 *   - DistributeOctTree tie-break = (size, creation counter) == reference under a bump allocator,
 *   - GaussianBlur = OpenCV 4.13 fixed-point path, taps [18,34,48,56,48,34,18]/256.
 */
#include "oracle.h"
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

typedef uint8_t uchar;
inline int cvRoundD(double v) { return (int)lrint(v); }    // cvRound: round-half-even
inline int cvRoundF(float v) { return (int)lrintf(v); }
inline int cvFloorF(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeilF(float v) { int i = (int)v; return i + (i < v); }
inline int reflect101(int p, int len) {                    // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
    if (len == 1) return 0;
    while (p < 0 || p >= len) { if (p < 0) p = -p; else p = 2 * len - 2 - p; }
    return p;
}
inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;   // ORBextractor.cc:72-74
const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

}  // namespace

/* cv::resize(INTER_LINEAR) on CV_8UC1 — SURVEY.md A.1 (OpenCV imgproc resize.cpp, 11-bit fixed point) */
extern "C" void orc_resize_linear_u8(const uchar* src, int sw, int sh, int spitch, uchar* dst, int dw, int dh, int dpitch) {
    double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cvFloorF(fx); fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = (short)cvRoundF((1.f - fx) * 2048); alpha[2 * dx + 1] = (short)cvRoundF(fx * 2048);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cvFloorF(fy); fy -= sy;
        if (sy < 0) { fy = 0; sy = 0; }
        if (sy >= sh - 1) { fy = 0; sy = sh - 1; }
        yofs[dy] = sy;
        beta[2 * dy] = (short)cvRoundF((1.f - fy) * 2048); beta[2 * dy + 1] = (short)cvRoundF(fy * 2048);
    }
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; dy++) {
        const uchar* S0 = src + (size_t)yofs[dy] * spitch;
        const uchar* S1 = src + (size_t)std::min(yofs[dy] + 1, sh - 1) * spitch;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = std::min(sx + 1, sw - 1);
            r0[dx] = S0[sx] * alpha[2 * dx] + S0[sx1] * alpha[2 * dx + 1];
            r1[dx] = S1[sx] * alpha[2 * dx] + S1[sx1] * alpha[2 * dx + 1];
        }
        int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uchar* D = dst + (size_t)dy * dpitch;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uchar)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

/* cv::copyMakeBorder(BORDER_REFLECT_101): dst is (w+2b) x (h+2b) */
extern "C" void orc_border_reflect101_u8(const uchar* src, int w, int h, int spitch, uchar* dst, int dpitch, int b) {
    for (int y = -b; y < h + b; y++) {
        const uchar* S = src + (size_t)reflect101(y, h) * spitch;
        uchar* D = dst + (size_t)(y + b) * dpitch;
        for (int x = -b; x < w + b; x++) D[x + b] = S[reflect101(x, w)];
    }
}

/* cv::GaussianBlur on CV_8UC1, OpenCV 4.13 fixed-point separable path — SURVEY.md A.2:
   out = (sum_j k_j * (sum_i k_i * p) + 32768) >> 16, REFLECT_101, taps in 8.8 fixed point (sum 256) */
extern "C" void orc_sepfilter_fixed_u8(const uchar* src, int w, int h, int spitch, uchar* dst, int dpitch,
                                       const int* taps, int ntaps) {
    // same arithmetic as the plain double loop, organised so that the compiler vectorises it (the CPU baseline of
    // bench.py should not be handicapped): padded source row -> u16 row sums (<= 255*256) -> i32 column sums
    const int r = ntaps / 2;
    std::vector<uint16_t> rows((size_t)w * h);
    std::vector<uchar> pad((size_t)w + 2 * r);
    std::vector<int> acc(w);
    for (int y = 0; y < h; y++) {
        const uchar* S = src + (size_t)y * spitch;
        for (int x = -r; x < w + r; x++) pad[x + r] = S[reflect101(x, w)];
        std::fill(acc.begin(), acc.end(), 0);
        for (int i = 0; i < ntaps; i++) {
            const int t = taps[i]; const uchar* P = pad.data() + i; int* A = acc.data();
            for (int x = 0; x < w; x++) A[x] += t * P[x];
        }
        uint16_t* R = &rows[(size_t)y * w];
        for (int x = 0; x < w; x++) R[x] = (uint16_t)acc[x];
    }
    for (int y = 0; y < h; y++) {
        std::fill(acc.begin(), acc.end(), 0);
        for (int j = 0; j < ntaps; j++) {
            const int t = taps[j]; const uint16_t* R = &rows[(size_t)reflect101(y + j - r, h) * w]; int* A = acc.data();
            for (int x = 0; x < w; x++) A[x] += t * R[x];
        }
        uchar* D = dst + (size_t)y * dpitch;
        for (int x = 0; x < w; x++) D[x] = (uchar)((acc[x] + 32768) >> 16);
    }
}
extern "C" void orc_gauss7_sigma2_u8(const uchar* src, int w, int h, int spitch, uchar* dst, int dpitch) {
    static const int taps[7] = {18, 34, 48, 56, 48, 34, 18};   // ORBextractor.cc:1086 GaussianBlur(7x7, sigma 2)
    orc_sepfilter_fixed_u8(src, w, h, spitch, dst, dpitch, taps, 7);
}

/* cv::fastAtan2 — SURVEY.md A.4 (OpenCV core mathfuncs_core, f32 polynomial, degrees) */
extern "C" float orc_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.141592653589793238462643383279502884);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale,
                p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON); c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON); c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

namespace {

/* FAST-9-16 corner score (OpenCV fast_score.cpp cornerScore<16>): max over the 16 arcs of 9
   contiguous ring pixels and both polarities of min |difference|, minus 1.  SURVEY.md A.3 */
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
inline int fast_score(const uchar* p, int pitch, int t) {
    int d[25];
    int v = p[0];
    // quick reject (same idea as OpenCV's table test): a 9-arc contains one pixel of every opposite pair
    {
        int a = v - p[3 * pitch], b = v - p[-3 * pitch];           // k = 0, 8
        bool br = a > t || b > t, dk = a < -t || b < -t;
        if (!br && !dk) return 0;
        a = v - p[3]; b = v - p[-3];                                // k = 4, 12
        br = br && (a > t || b > t); dk = dk && (a < -t || b < -t);
        if (!br && !dk) return 0;
    }
    for (int k = 0; k < 16; k++) d[k] = v - p[kRingDy[k] * pitch + kRingDx[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int k = 0; k < 16; k++) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1 >= t ? best - 1 : 0;   // corner at threshold t <=> score >= t
}

}  // namespace

/* cv::FAST(img, kps, threshold, nonmaxSuppression=true), TYPE_9_16.  Tested pixels: 3<=x<cols-3,
   3<=y<rows-3; corner <=> score >= threshold; NMS keeps strict maxima over the 8 neighbours where
   non-corners / untested pixels count 0; output in raster order.  SURVEY.md A.3 */
extern "C" int orc_fast9_16(const uchar* img, int cols, int rows, int pitch, int threshold,
                            int* xs, int* ys, int* scores, int cap) {
    if (cols < 7 || rows < 7) return 0;
    std::vector<int> sc((size_t)cols * rows, 0);
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            sc[(size_t)y * cols + x] = fast_score(img + (size_t)y * pitch + x, pitch, threshold);
        }
    int n = 0;
    for (int y = 3; y < rows - 3; y++)
        for (int x = 3; x < cols - 3; x++) {
            int s = sc[(size_t)y * cols + x];
            if (s == 0) continue;   // threshold >= 1 in all uses, so 0 <=> not a corner
            const int* c = &sc[(size_t)y * cols + x];
            if (s > c[-1] && s > c[1] && s > c[-cols - 1] && s > c[-cols] && s > c[-cols + 1] &&
                s > c[cols - 1] && s > c[cols] && s > c[cols + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; scores[n] = s; }
                n++;
            }
        }
    return n;
}

namespace {

struct Key { int x, y, resp; };   // integer-valued in the reference (FAST output + cell offset)

/* ExtractorNode (ORBextractor.h:33-43) */
struct Node {
    int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
    std::vector<int> keys;            // indices into the candidate array, order preserved
    std::list<Node>::iterator lit;
    bool noMore = false;
};

/* ExtractorNode::DivideNode — ORBextractor.cc:481-537 */
void divide_node(const Node& p, const std::vector<Key>& K, Node& n1, Node& n2, Node& n3, Node& n4) {
    const int halfX = (int)std::ceil((float)(p.URx - p.ULx) / 2);
    const int halfY = (int)std::ceil((float)(p.BRy - p.ULy) / 2);
    n1.ULx = p.ULx; n1.ULy = p.ULy; n1.URx = p.ULx + halfX; n1.URy = p.ULy;
    n1.BLx = p.ULx; n1.BLy = p.ULy + halfY; n1.BRx = p.ULx + halfX; n1.BRy = p.ULy + halfY;
    n2.ULx = n1.URx; n2.ULy = n1.URy; n2.URx = p.URx; n2.URy = p.URy;
    n2.BLx = n1.BRx; n2.BLy = n1.BRy; n2.BRx = p.URx; n2.BRy = p.ULy + halfY;
    n3.ULx = n1.BLx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.URy = n1.BRy;
    n3.BLx = p.BLx; n3.BLy = p.BLy; n3.BRx = n1.BRx; n3.BRy = p.BLy;
    n4.ULx = n3.URx; n4.ULy = n3.URy; n4.URx = n2.BRx; n4.URy = n2.BRy;
    n4.BLx = n3.BRx; n4.BLy = n3.BRy; n4.BRx = p.BRx; n4.BRy = p.BRy;
    for (int k : p.keys) {
        const Key& kp = K[k];
        if (kp.x < n1.URx) { if (kp.y < n1.BRy) n1.keys.push_back(k); else n3.keys.push_back(k); }
        else if (kp.y < n1.BRy) n2.keys.push_back(k);
        else n4.keys.push_back(k);
    }
    if (n1.keys.size() == 1) n1.noMore = true;
    if (n2.keys.size() == 1) n2.noMore = true;
    if (n3.keys.size() == 1) n3.noMore = true;
    if (n4.keys.size() == 1) n4.noMore = true;
}

/* ORBextractor::DistributeOctTree — ORBextractor.cc:539-763.  Returns indices of the retained
   candidates in final list order.  Sort key (size, pointer) is canonicalised to (size, counter). */
std::vector<int> distribute_octree(const std::vector<Key>& K, int minX, int maxX, int minY, int maxY, int N) {
    std::vector<int> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));           // :542
    if (nIni <= 0) return result;   // degenerate (tall) level: reference divides by zero below (UB)
    const float hX = (float)(maxX - minX) / nIni;                                        // :544
    std::list<Node> L;
    std::vector<Node*> ini(nIni);
    for (int i = 0; i < nIni; i++) {                                                     // :551-562
        Node ni;
        ni.ULx = (int)(hX * (float)i); ni.ULy = 0;
        ni.URx = (int)(hX * (float)(i + 1)); ni.URy = 0;
        ni.BLx = ni.ULx; ni.BLy = maxY - minY;
        ni.BRx = ni.URx; ni.BRy = maxY - minY;
        L.push_back(ni);
        ini[i] = &L.back();
    }
    for (size_t i = 0; i < K.size(); i++) {                                              // :565-569
        int r = (int)((float)K[i].x / hX);
        if (r >= nIni) r = nIni - 1;   // never taken for in-range candidates (x < maxX-minX); guards UB
        ini[r]->keys.push_back((int)i);
    }
    for (auto lit = L.begin(); lit != L.end();) {                                        // :573-584
        if (lit->keys.size() == 1) { lit->noMore = true; ++lit; }
        else if (lit->keys.empty()) lit = L.erase(lit);
        else ++lit;
    }
    typedef std::pair<std::pair<int, long long>, Node*> Entry;   // ((size, creation counter), node)
    long long counter = 0;
    std::vector<Entry> vSize;
    bool finish = false;
    auto push_children = [&](Node* ch[4], int& nToExpand) {
        for (int c = 0; c < 4; c++) {
            if (ch[c]->keys.empty()) continue;
            L.push_front(*ch[c]);
            if (ch[c]->keys.size() > 1) {
                nToExpand++;
                vSize.push_back(Entry(std::make_pair((int)ch[c]->keys.size(), counter++), &L.front()));
                L.front().lit = L.begin();
            }
        }
    };
    while (!finish) {                                                                    // :591
        int prevSize = (int)L.size();
        int nToExpand = 0;
        vSize.clear();
        for (auto lit = L.begin(); lit != L.end();) {                                    // :603-665
            if (lit->noMore) { ++lit; continue; }
            Node n1, n2, n3, n4; Node* ch[4] = {&n1, &n2, &n3, &n4};
            divide_node(*lit, K, n1, n2, n3, n4);
            push_children(ch, nToExpand);
            lit = L.erase(lit);
        }
        if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;              // :669
        else if ((int)L.size() + nToExpand * 3 > N) {                                    // :673
            while (!finish) {
                prevSize = (int)L.size();
                std::vector<Entry> prev = vSize;
                vSize.clear();
                std::sort(prev.begin(), prev.end(),
                          [](const Entry& a, const Entry& b) { return a.first < b.first; });   // :684
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4; Node* ch[4] = {&n1, &n2, &n3, &n4};
                    divide_node(*prev[j].second, K, n1, n2, n3, n4);
                    int dummy = 0;
                    push_children(ch, dummy);
                    L.erase(prev[j].second->lit);
                    if ((int)L.size() >= N) break;                                       // :730
                }
                if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;      // :734
            }
        }
    }
    for (auto& nd : L) {                                                                 // :741-760
        int best = nd.keys[0];
        for (size_t k = 1; k < nd.keys.size(); k++)
            if (K[nd.keys[k]].resp > K[best].resp) best = nd.keys[k];
        result.push_back(best);
    }
    return result;
}

}  // namespace

extern "C" int orc_octree(const int* xs, const int* ys, const int* resp, int n, int minX, int maxX, int minY, int maxY,
                          int N, int* out_idx, int cap) {
    std::vector<Key> K(n);
    for (int i = 0; i < n; i++) K[i] = Key{xs[i], ys[i], resp[i]};
    std::vector<int> r = distribute_octree(K, minX, maxX, minY, maxY, N);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out_idx[i] = r[i];
    return (int)r.size();
}

struct orc_orb {
    int nfeatures, nlevels, iniTh, minTh;
    float scaleFactorF;
    std::vector<float> scale, invscale, sigma2, invsigma2;
    std::vector<int> nfeat, umax;
    // state of the last call
    struct Level { int w, h, pitch; std::vector<uchar> buf; std::vector<uchar> blur; std::vector<Key> cand;
                   std::vector<Key> kp; std::vector<float> angle; };
    std::vector<Level> lv;
    double ms[6];
};

/* ORBextractor::ORBextractor — ORBextractor.cc:410-470 */
extern "C" orc_orb* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniTh, int minTh) {
    orc_orb* o = new orc_orb();
    o->nfeatures = nfeatures; o->nlevels = nlevels; o->iniTh = iniTh; o->minTh = minTh; o->scaleFactorF = scaleFactor;
    const double sfd = (double)scaleFactor;     // member is `double scaleFactor` (ORBextractor.h:96)
    o->scale.resize(nlevels); o->sigma2.resize(nlevels); o->invscale.resize(nlevels); o->invsigma2.resize(nlevels);
    o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        o->scale[i] = (float)(o->scale[i - 1] * sfd);
        o->sigma2[i] = o->scale[i] * o->scale[i];
    }
    for (int i = 0; i < nlevels; i++) { o->invscale[i] = 1.0f / o->scale[i]; o->invsigma2[i] = 1.0f / o->sigma2[i]; }
    o->nfeat.resize(nlevels);
    float factor = (float)(1.0f / sfd);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        o->nfeat[l] = cvRoundF(nDesired);
        sum += o->nfeat[l];
        nDesired *= factor;
    }
    o->nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);
    o->umax.assign(HALF_PATCH_SIZE + 1, 0);
    int v, v0, vmax = cvFloorF(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = cvCeilF(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) o->umax[v] = cvRoundD(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
    o->lv.resize(nlevels);
    return o;
}
extern "C" void orc_orb_destroy(orc_orb* o) { delete o; }
extern "C" void orc_orb_tables(const orc_orb* o, float* scale, float* invscale, float* sigma2, float* invsigma2,
                               int* nfeat, int* umax16) {
    for (int i = 0; i < o->nlevels; i++) {
        if (scale) scale[i] = o->scale[i];
        if (invscale) invscale[i] = o->invscale[i];
        if (sigma2) sigma2[i] = o->sigma2[i];
        if (invsigma2) invsigma2[i] = o->invsigma2[i];
        if (nfeat) nfeat[i] = o->nfeat[i];
    }
    if (umax16) for (int i = 0; i < 16; i++) umax16[i] = o->umax[i];
}

namespace {

/* IC_Angle — ORBextractor.cc:77-104 (on the un-blurred bordered level) */
float ic_angle(const uchar* center, int step, const std::vector<int>& umax) {
    int m_01 = 0, m_10 = 0;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * step], val_minus = center[u - v * step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor — ORBextractor.cc:107-147 */
void orb_descriptor(float kp_angle, const uchar* center, int step, uchar* desc) {
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);
    float angle = kp_angle * factorPI;
    float a = (float)cos((double)angle), b = (float)sin((double)angle);   // canonical: correctly-rounded f32
    const int8_t* pat = kPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; k++) {
            float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1], x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
            int t0 = center[cvRoundF(x0 * b + y0 * a) * step + cvRoundF(x0 * a - y0 * b)];
            int t1 = center[cvRoundF(x1 * b + y1 * a) * step + cvRoundF(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uchar)val;
    }
}

}  // namespace

/* Stage-level entry points so that IC_Angle and the rBRIEF arithmetic can be pinned against cv2.ORB on level 0
   (tests/test_oracle_cpu.py): orientation of given integer pixel positions on an un-blurred image, and descriptors of given
   (position, angle) on an ALREADY blurred image.  Positions must keep 19 px from the image border, as the extractor's do. */
extern "C" void orc_ic_angles(const orc_orb* o, const uchar* img, int w, int h, int pitch, const int* xs, const int* ys, int n, float* angles) {
    (void)w; (void)h;
    for (int i = 0; i < n; i++) angles[i] = ic_angle(img + (ptrdiff_t)ys[i] * pitch + xs[i], pitch, o->umax);
}
extern "C" void orc_brief_descriptors(const uchar* blurred, int w, int h, int pitch, const int* xs, const int* ys, const float* angles,
                                      int n, uchar* desc) {
    (void)w; (void)h;
    for (int i = 0; i < n; i++) orb_descriptor(angles[i], blurred + (ptrdiff_t)ys[i] * pitch + xs[i], pitch, desc + 32 * (size_t)i);
}

/* ORBextractor::operator() — ORBextractor.cc:1043-1105 (+ ComputePyramid :1107-1132,
   ComputeKeyPointsOctTree :765-853) */
extern "C" int orc_orb_extract(orc_orb* o, const uchar* img, int w, int h, int pitch,
                               orc_keypoint* kps, uchar* desc, int cap) {
    if (!img || w <= 0 || h <= 0) return 0;                                   // :1046
    const int L = o->nlevels, B = EDGE_THRESHOLD;
    double t0 = now_ms();
    // ComputePyramid
    for (int l = 0; l < L; l++) {
        orc_orb::Level& lv = o->lv[l];
        float sc = o->invscale[l];
        lv.w = cvRoundF((float)w * sc); lv.h = cvRoundF((float)h * sc);
        lv.pitch = lv.w + 2 * B;
        lv.buf.assign((size_t)lv.pitch * (lv.h + 2 * B), 0);
        std::vector<uchar> tmp((size_t)lv.w * lv.h);
        if (l == 0) { for (int y = 0; y < h; y++) memcpy(&tmp[(size_t)y * w], img + (size_t)y * pitch, w); }
        else {
            orc_orb::Level& pv = o->lv[l - 1];
            orc_resize_linear_u8(&pv.buf[(size_t)B * pv.pitch + B], pv.w, pv.h, pv.pitch, tmp.data(), lv.w, lv.h, lv.w);
        }
        orc_border_reflect101_u8(tmp.data(), lv.w, lv.h, lv.w, lv.buf.data(), lv.pitch, B);
    }
    double t1 = now_ms();
    double tf = 0, to = 0;
    // ComputeKeyPointsOctTree :765-853
    const float W = 30;
    for (int l = 0; l < L; l++) {
        double ta = now_ms();
        orc_orb::Level& lv = o->lv[l];
        const uchar* base = &lv.buf[(size_t)B * lv.pitch + B];
        const int minBX = EDGE_THRESHOLD - 3, minBY = minBX;
        const int maxBX = lv.w - EDGE_THRESHOLD + 3, maxBY = lv.h - EDGE_THRESHOLD + 3;
        lv.cand.clear(); lv.kp.clear(); lv.angle.clear();
        const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols > 0 && nRows > 0) {
            const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
            std::vector<int> xs(4096), ys(4096), ss(4096);
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
                    int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                    const uchar* sub = base + (ptrdiff_t)y0 * lv.pitch + x0;
                    int n = orc_fast9_16(sub, cw, ch, lv.pitch, o->iniTh, xs.data(), ys.data(), ss.data(), 4096);
                    if (n == 0) n = orc_fast9_16(sub, cw, ch, lv.pitch, o->minTh, xs.data(), ys.data(), ss.data(), 4096);
                    for (int k = 0; k < n; k++) lv.cand.push_back(Key{xs[k] + j * wCell, ys[k] + i * hCell, ss[k]});
                }
            }
        }
        double tb = now_ms();
        std::vector<int> sel = distribute_octree(lv.cand, minBX, maxBX, minBY, maxBY, o->nfeat[l]);
        for (int idx : sel) lv.kp.push_back(Key{lv.cand[idx].x + minBX, lv.cand[idx].y + minBY, lv.cand[idx].resp});
        double tc = now_ms();
        tf += tb - ta; to += tc - tb;
    }
    double t2 = now_ms();
    for (int l = 0; l < L; l++) {                                              // computeOrientation :472-479
        orc_orb::Level& lv = o->lv[l];
        const uchar* base = &lv.buf[(size_t)B * lv.pitch + B];
        for (const Key& k : lv.kp) lv.angle.push_back(ic_angle(base + (ptrdiff_t)k.y * lv.pitch + k.x, lv.pitch, o->umax));
    }
    double t3 = now_ms();
    int n = 0;
    double tblur = 0, tbrief = 0;
    for (int l = 0; l < L; l++) {                                              // :1078-1104
        orc_orb::Level& lv = o->lv[l];
        if (lv.kp.empty()) { lv.blur.clear(); continue; }
        double ta = now_ms();
        lv.blur.assign((size_t)lv.w * lv.h, 0);
        orc_gauss7_sigma2_u8(&lv.buf[(size_t)B * lv.pitch + B], lv.w, lv.h, lv.pitch, lv.blur.data(), lv.w);
        double tb = now_ms();
        const int scaledPatchSize = (int)(PATCH_SIZE * o->scale[l]);           // :836
        for (size_t i = 0; i < lv.kp.size(); i++, n++) {
            if (n >= cap) continue;
            const Key& k = lv.kp[i];
            orb_descriptor(lv.angle[i], &lv.blur[(size_t)k.y * lv.w + k.x], lv.w, desc + (size_t)n * 32);
            orc_keypoint& kp = kps[n];
            kp.x = (float)k.x; kp.y = (float)k.y;
            if (l != 0) { kp.x *= o->scale[l]; kp.y *= o->scale[l]; }
            kp.size = (float)scaledPatchSize; kp.angle = lv.angle[i]; kp.response = (float)k.resp;
            kp.octave = l; kp.class_id = -1;
        }
        tblur += tb - ta; tbrief += now_ms() - tb;
    }
    o->ms[0] = t1 - t0; o->ms[1] = tf; o->ms[2] = to; o->ms[3] = t3 - t2; o->ms[4] = tblur; o->ms[5] = tbrief;
    return n;
}

extern "C" void orc_orb_level_size(const orc_orb* o, int l, int* w, int* h) { *w = o->lv[l].w; *h = o->lv[l].h; }
extern "C" void orc_orb_level_copy(const orc_orb* o, int l, int bordered, uchar* dst, int dpitch) {
    const orc_orb::Level& lv = o->lv[l];
    int B = EDGE_THRESHOLD;
    if (bordered) for (int y = 0; y < lv.h + 2 * B; y++) memcpy(dst + (size_t)y * dpitch, &lv.buf[(size_t)y * lv.pitch], lv.w + 2 * B);
    else for (int y = 0; y < lv.h; y++) memcpy(dst + (size_t)y * dpitch, &lv.buf[(size_t)(y + B) * lv.pitch + B], lv.w);
}
extern "C" void orc_orb_blur_copy(const orc_orb* o, int l, uchar* dst, int dpitch) {
    const orc_orb::Level& lv = o->lv[l];
    if (lv.blur.empty()) return;
    for (int y = 0; y < lv.h; y++) memcpy(dst + (size_t)y * dpitch, &lv.blur[(size_t)y * lv.w], lv.w);
}
extern "C" int orc_orb_candidates(const orc_orb* o, int l, int* xs, int* ys, int* resp, int cap) {
    const auto& c = o->lv[l].cand;
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xs[i] = c[i].x; ys[i] = c[i].y; resp[i] = c[i].resp; }
    return (int)c.size();
}
extern "C" int orc_orb_level_keypoints(const orc_orb* o, int l, int* xs, int* ys, int* resp, float* angle, int cap) {
    const auto& c = o->lv[l].kp;
    for (size_t i = 0; i < c.size() && (int)i < cap; i++) { xs[i] = c[i].x; ys[i] = c[i].y; resp[i] = c[i].resp; angle[i] = o->lv[l].angle[i]; }
    return (int)c.size();
}
extern "C" void orc_orb_stage_ms(const orc_orb* o, double* ms6) { for (int i = 0; i < 6; i++) ms6[i] = o->ms[i]; }
