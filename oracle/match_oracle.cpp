/*
 * oracle/match_oracle.cpp — CPU ORACLE for the Hamming matchers.  TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Restates /root/reference/src/ORBmatcher.cc, src/LSDmatcher.cpp (knnMatch-based entry points),
 * src/Frame.cc:190-215 (lineDescriptorMAD) and cv::BFMatcher(NORM_HAMMING).knnMatch(k=2)
 * (pinned against cv2 in tests/).  KeyFrame/Frame/MapPoint pointers of the reference
 * are replaced by indices + validity masks; DBoW2::FeatureVector (std::map<node, vector<idx>>) is
 * passed as CSR with ascending node ids (FeatureVector.cpp:31-45 keeps per-node indices ascending).
 *
 * PINNED: tests/test_ref_parity_cpu.py runs every function of this file against the reference's own
 * ORBmatcher.cc / LSDmatcher.cpp / Frame.cc / MapPoint.cc / DBoW2 compiled unmodified (oracle/_ref, oracle/ref_build.sh).
 */
#include "oracle.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {
const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;   // ORBmatcher.cc:37-39

/* ORBmatcher::ComputeThreeMaxima — ORBmatcher.cc:1604-1645 */
void three_maxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

inline int rot_bin(float a1, float a2) {                     // ORBmatcher.cc:241-246
    const float factor = 1.0f / HISTO_LENGTH;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)roundf(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

/* std::map::lower_bound merge-walk over two CSR feature vectors: calls f(i1, i2) for every common node */
template <class F>
void walk_common_nodes(const int32_t* nodes1, int nn1, const int32_t* nodes2, int nn2, F f) {
    int i1 = 0, i2 = 0;
    while (i1 < nn1 && i2 < nn2) {
        if (nodes1[i1] == nodes2[i2]) { f(i1, i2); i1++; i2++; }
        else if (nodes1[i1] < nodes2[i2]) i1 = (int)(std::lower_bound(nodes1 + i1, nodes1 + nn1, nodes2[i2]) - nodes1);
        else i2 = (int)(std::lower_bound(nodes2 + i2, nodes2 + nn2, nodes1[i1]) - nodes2);
    }
}
}  // namespace

/* ORBmatcher::DescriptorDistance — ORBmatcher.cc:1650-1666 (= LSDmatcher.cpp:364-380) */
extern "C" int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t pa, pb;
        memcpy(&pa, a + 4 * i, 4); memcpy(&pb, b + 4 * i, 4);
        unsigned int v = pa ^ pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

/* cv::BFMatcher(NORM_HAMMING,false).knnMatch(q,t,out,2) — SURVEY.md A.7: ascending distance, ties -> lower trainIdx */
extern "C" void orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out) {
    for (int i = 0; i < nq; i++) {
        int d0 = 1 << 30, i0 = -1, d1 = 1 << 30, i1 = -1;
        for (int j = 0; j < nt; j++) {
            int d = orc_descriptor_distance(q + 32 * (size_t)i, t + 32 * (size_t)j);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        out[4 * i] = i0; out[4 * i + 1] = i0 < 0 ? -1 : d0; out[4 * i + 2] = i1; out[4 * i + 3] = i1 < 0 ? -1 : d1;
    }
}

/* Synthetic stand-in for DBoW2 TemplatedVocabulary::transform (TemplatedVocabulary.h:1218-1259) collapsed
   to one level: nearest centroid by Hamming distance, strict '<' so the first best centroid wins. */
extern "C" void orc_bow_assign(const uint8_t* desc, int n, const uint8_t* centroids, int nc, int32_t* node) {
    for (int i = 0; i < n; i++) {
        int best = 1 << 30, bi = 0;
        for (int c = 0; c < nc; c++) {
            int d = orc_descriptor_distance(desc + 32 * (size_t)i, centroids + 32 * (size_t)c);
            if (d < best) { best = d; bi = c; }
        }
        node[i] = bi;
    }
}

/* DBoW2 TemplatedVocabulary<FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
   (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1195) with the per-feature tree descent of :1218-1259.
   The tree is given as arrays (what loadFromTextFile :1338-1420 builds): node 0 = root, parent[i] < i, children of a node
   in ascending node id, is_leaf[i] as in the file, word ids numbered over the leaves in node order.
   Per feature: word[i], node[i] (the NodeId at level L - levelsup; 0 when the leaf is reached earlier or the level is <= 0 —
   the reference leaves it uninitialised in the first case), weight[i] (the leaf's weight; TF_IDF / IDF: idf, TF / BINARY: 1
   is already folded into the file's weights).  Features with weight <= 0 ("stopped" words) are not part of the
   FeatureVector (:1162-1166): the caller drops them. */
extern "C" int orc_vocab_transform(int L, int nnodes, const int32_t* parent, const uint8_t* ndesc, const double* weight,
                                   const uint8_t* is_leaf, const uint8_t* feats, int n, int levelsup,
                                   int32_t* word, int32_t* node, double* w) {
    std::vector<std::vector<int>> children(nnodes);
    std::vector<int> word_id(nnodes, -1);
    int nwords = 0;
    for (int i = 1; i < nnodes; i++) {
        if (parent[i] < 0 || parent[i] >= i) return -1;
        children[parent[i]].push_back(i);
        if (is_leaf[i]) word_id[i] = nwords++;
    }
    const int nid_level = L - levelsup;
    for (int f = 0; f < n; f++) {
        const uint8_t* feat = feats + 32 * (size_t)f;
        int nid = 0, final_id = 0, current_level = 0;
        do {
            ++current_level;
            const std::vector<int>& nodes = children[final_id];
            if (nodes.empty()) break;                              // (an empty vocabulary: the reference returns early)
            final_id = nodes[0];
            double best_d = orc_descriptor_distance(feat, ndesc + 32 * (size_t)final_id);
            for (size_t c = 1; c < nodes.size(); c++) {
                const double d = orc_descriptor_distance(feat, ndesc + 32 * (size_t)nodes[c]);
                if (d < best_d) { best_d = d; final_id = nodes[c]; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (!children[final_id].empty());
        word[f] = word_id[final_id]; node[f] = nid; w[f] = weight[final_id];
    }
    return nwords;
}

/* ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) — ORBmatcher.cc:159-291 */
extern "C" int orc_search_by_bow(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                                 const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nn1,
                                 const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nn2,
                                 const uint8_t* valid1, const float* angle1, const float* angle2,
                                 float nnratio, int checkOri, int32_t* match2) {
    (void)n1;
    for (int j = 0; j < n2; j++) match2[j] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    walk_common_nodes(nodes1, nn1, nodes2, nn2, [&](int a, int b) {
        for (int iKF = off1[a]; iKF < off1[a + 1]; iKF++) {
            const int realIdxKF = idx1[iKF];
            if (!valid1[realIdxKF]) continue;                                         // :196-200
            int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
            for (int iF = off2[b]; iF < off2[b + 1]; iF++) {
                const int realIdxF = idx2[iF];
                if (match2[realIdxF] >= 0) continue;                                  // :212
                const int dist = orc_descriptor_distance(d1 + 32 * (size_t)realIdxKF, d2 + 32 * (size_t)realIdxF);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            if (bestDist1 <= TH_LOW) {                                                // :231
                if ((float)bestDist1 < nnratio * (float)bestDist2) {                  // :233
                    match2[bestIdxF] = realIdxKF;
                    if (checkOri) rotHist[rot_bin(angle1[realIdxKF], angle2[bestIdxF])].push_back(bestIdxF);
                    nmatches++;
                }
            }
        }
    });
    if (checkOri) {                                                                   // :270-288
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match2[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* ---------------------------------------------------------------------------------------------
 * Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (Frame.cc:133-148, 462-472, 368-421)
 * and ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono) (ORBmatcher.cc:1331-1473).
 * Pointer state is passed as flags: valid1 = (pMP && !mvbOutlier), obs1 = pMP->Observations() > 0, claimed2 = the current
 * frame's feature already holds a MapPoint with Observations() > 0.  assign2[j] = index (into the last frame) of the
 * MapPoint written to CurrentFrame.mvpMapPoints[j], or -1.  Float arithmetic: every expression in float, no contraction;
 * Rcw*x3Dw+tcw as cv::gemm evaluates a plain 3x3 * 3x1 + 3x1 CV_32F product on cv2 4.13 (float products summed left to right,
 * the addend last; tools/probe_cv_gemm.py).  Pinned against the reference itself in tests/test_ref_parity_cpu.py.
 * --------------------------------------------------------------------------------------------- */
namespace {
struct Grid {
    static const int COLS = 64, ROWS = 48;                         /* Frame.h:45-46 */
    float minX, minY, invW, invH;
    std::vector<int> cell[COLS][ROWS];
    void build(int n, const float* x, const float* y) {            /* AssignFeaturesToGrid + PosInGrid */
        for (int i = 0; i < n; i++) {
            const int px = (int)std::round((x[i] - minX) * invW), py = (int)std::round((y[i] - minY) * invH);
            if (px < 0 || px >= COLS || py < 0 || py >= ROWS) continue;
            cell[px][py].push_back(i);
        }
    }
    void area(float x, float y, float r, int minLevel, int maxLevel, const float* kx, const float* ky, const int32_t* oct,
              std::vector<int>& out) const {                       /* GetFeaturesInArea */
        out.clear();
        const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
        if (nMinCellX >= COLS) return;
        const int nMaxCellX = std::min(COLS - 1, (int)std::ceil((x - minX + r) * invW));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
        if (nMinCellY >= ROWS) return;
        const int nMaxCellY = std::min(ROWS - 1, (int)std::ceil((y - minY + r) * invH));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
                for (int j : cell[ix][iy]) {
                    if (bCheckLevels) {
                        if (oct[j] < minLevel) continue;
                        if (maxLevel >= 0 && oct[j] > maxLevel) continue;
                    }
                    const float distx = kx[j] - x, disty = ky[j] - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(j);
                }
    }
};
}  // namespace

extern "C" int orc_features_in_area(int n, const float* kx, const float* ky, const int32_t* oct, float minX, float minY, float invW, float invH,
                                    float x, float y, float r, int minLevel, int maxLevel, int32_t* out, int cap) {
    Grid* g = new Grid(); g->minX = minX; g->minY = minY; g->invW = invW; g->invH = invH;
    g->build(n, kx, ky);
    std::vector<int> v; g->area(x, y, r, minLevel, maxLevel, kx, ky, oct, v);
    delete g;
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) out[i] = v[i];
    return (int)v.size();
}

extern "C" int orc_search_by_projection_frame(
        int n1, const uint8_t* valid1, const uint8_t* obs1, const float* Xw, const uint8_t* dmp, const int32_t* oct1, const float* angle1,
        int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* angle2, const float* uright2,
        const uint8_t* claimed2, const float* Tcw, const float* Tlw, const float* cam /* fx fy cx cy mbf mb */,
        const float* bounds /* minX maxX minY maxY */, const float* scaleFactors, float th, int bMono, int checkOri, int32_t* assign2) {
    const int TH_HIGH = 100;                                                             /* ORBmatcher.cc:37 */
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], mbf = cam[4], mb = cam[5];
    Grid* g = new Grid();
    g->minX = bounds[0]; g->minY = bounds[2];
    g->invW = (float)Grid::COLS / (bounds[1] - bounds[0]); g->invH = (float)Grid::ROWS / (bounds[3] - bounds[2]);   /* Frame.cc:115-116 */
    g->build(n2, x2, y2);
    std::vector<uint8_t> claimed(n2, 0);
    for (int j = 0; j < n2; j++) { assign2[j] = -1; if (claimed2) claimed[j] = claimed2[j]; }
    /* tlc = Rlw * (-Rcw^T tcw) + tlw, only its z against mb (stereo).  cv::gemm as probed on cv2 4.13 (tools/probe_cv_gemm.py,
       oracle/refshim/minicv.cpp): a transposed operand takes the general path (double accumulation, one rounding); a plain
       3x3 * 3x1 (+ 3x1) runs in FLOAT, products summed left to right, the addend last. */
    bool bForward = false, bBackward = false;
    if (!bMono && Tlw) {
        double twc[3];
        for (int r = 0; r < 3; r++) twc[r] = -((double)Tcw[0 * 4 + r] * Tcw[3] + (double)Tcw[1 * 4 + r] * Tcw[7] + (double)Tcw[2 * 4 + r] * Tcw[11]);
        const float twcf[3] = {(float)twc[0], (float)twc[1], (float)twc[2]};
        float s = Tlw[8] * twcf[0]; s = s + Tlw[9] * twcf[1]; s = s + Tlw[10] * twcf[2];
        const float tlcz = s + Tlw[11];
        bForward = tlcz > mb; bBackward = -tlcz > mb;
    }
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> cand;
    for (int i = 0; i < n1; i++) {
        if (!valid1[i]) continue;
        const float* X = Xw + 3 * (size_t)i;
        float xc3[3];
        for (int r = 0; r < 3; r++) {                                       /* x3Dc = Rcw*x3Dw+tcw (:1364): cv::gemm float path */
            float s = Tcw[4 * r] * X[0]; s = s + Tcw[4 * r + 1] * X[1]; s = s + Tcw[4 * r + 2] * X[2];
            xc3[r] = s + Tcw[4 * r + 3];
        }
        const float xc = xc3[0], yc = xc3[1];
        const float invzc = (float)(1.0 / xc3[2]);
        if (invzc < 0) continue;
        const float u = fx * xc * invzc + cx, v = fy * yc * invzc + cy;
        if (u < bounds[0] || u > bounds[1]) continue;
        if (v < bounds[2] || v > bounds[3]) continue;
        const int nLastOctave = oct1[i];
        const float radius = th * scaleFactors[nLastOctave];
        if (bForward) g->area(u, v, radius, nLastOctave, -1, x2, y2, oct2, cand);
        else if (bBackward) g->area(u, v, radius, 0, nLastOctave, x2, y2, oct2, cand);
        else g->area(u, v, radius, nLastOctave - 1, nLastOctave + 1, x2, y2, oct2, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (claimed[i2]) continue;                                                   /* :1400-1402 */
            if (uright2 && uright2[i2] > 0) {
                const float ur = u - mbf * invzc, er = std::fabs(ur - uright2[i2]);
                if (er > radius) continue;
            }
            const int dist = orc_descriptor_distance(dmp + 32 * (size_t)i, d2 + 32 * (size_t)i2);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            assign2[bestIdx2] = i; claimed[bestIdx2] = obs1[i] ? 1 : 0;
            nmatches++;
            if (checkOri) rotHist[rot_bin(angle1[i], angle2[bestIdx2])].push_back(bestIdx2);
        }
    }
    delete g;
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { assign2[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchByProjection(Frame &F, const vector<MapPoint*> &vpMapPoints, th) — ORBmatcher.cc:45-129, the matcher of
 * Tracking::SearchLocalPoints (TrackLocalMap, Tracking.cc:1736), run on every frame.  Per MapPoint (in vector order) the fields
 * Frame::isInFrustum left on it: inview (mbTrackInView), bad (isBad()), projx/projy (mTrackProjX/Y), projxr (mTrackProjXR, may be
 * NULL), level (mnTrackScaleLevel), viewcos (mTrackViewCos), its descriptor, and obs = Observations() > 0 (a frame feature holding
 * such a point is skipped by later points, :86-88).  held2[j]: 0 = the frame feature holds nothing, 1 = a MapPoint WITH
 * observations (skipped), 2 = a MapPoint without (may be replaced).  assign2[j] = index of the MapPoint written to
 * F.mvpMapPoints[j] (the last writer), -1 none (features that only keep what they held are -1 too). */
extern "C" int orc_search_by_projection_mps(
        int nmp, const uint8_t* inview, const uint8_t* bad, const uint8_t* obs, const float* projx, const float* projy, const float* projxr,
        const int32_t* level, const float* viewcos, const uint8_t* dmp,
        int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* uright2, const uint8_t* held2,
        const float* bounds, const float* scaleFactors, float nnratio, float th, int32_t* assign2) {
    Grid* g = new Grid();
    g->minX = bounds[0]; g->minY = bounds[2];
    g->invW = (float)Grid::COLS / (bounds[1] - bounds[0]); g->invH = (float)Grid::ROWS / (bounds[3] - bounds[2]);
    g->build(n2, x2, y2);
    std::vector<uint8_t> claimed(n2, 0);
    for (int j = 0; j < n2; j++) { assign2[j] = -1; claimed[j] = held2 && held2[j] == 1; }
    int nmatches = 0;
    const bool bFactor = th != 1.0;                                                       /* :49 */
    std::vector<int> cand;
    for (int i = 0; i < nmp; i++) {
        if (!inview[i]) continue;                                                         /* :54 */
        if (bad && bad[i]) continue;                                                      /* :57 */
        const int nPredictedLevel = level[i];
        float r = viewcos[i] > 0.998 ? 2.5f : 4.0f;                                       /* RadiusByViewingCos :131-137 (double compare) */
        if (bFactor) r *= th;
        g->area(projx[i], projy[i], r * scaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, x2, y2, oct2, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (claimed[idx]) continue;                                                   /* :86-88 */
            if (uright2 && uright2[idx] > 0) {
                const float er = std::fabs(projxr[i] - uright2[idx]);
                if (er > r * scaleFactors[nPredictedLevel]) continue;
            }
            const int dist = orc_descriptor_distance(dmp + 32 * (size_t)i, d2 + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct2[idx]; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = oct2[idx]; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;     /* :117 (int -> float) */
            assign2[bestIdx] = i; claimed[bestIdx] = obs && obs[i];
            nmatches++;
        }
    }
    delete g;
    return nmatches;
}

/* ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — ORBmatcher.cc:408-523
 * (Tracking::MonocularInitialization, Tracking.cc:366).  Keypoints are mvKeysUn; prev[2*n1] is vbPrevMatched (in/out). */
extern "C" int orc_search_for_initialization(
        int n1, const uint8_t* d1, const int32_t* oct1, const float* angle1,
        int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* angle2,
        float* prev, const float* bounds, float nnratio, int checkOri, int windowSize, int32_t* matches12) {
    Grid* g = new Grid();
    g->minX = bounds[0]; g->minY = bounds[2];
    g->invW = (float)Grid::COLS / (bounds[1] - bounds[0]); g->invH = (float)Grid::ROWS / (bounds[3] - bounds[2]);
    g->build(n2, x2, y2);
    int nmatches = 0;
    for (int i = 0; i < n1; i++) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> matchedDistance(n2, 0x7fffffff), matches21(n2, -1), cand;
    for (int i1 = 0; i1 < n1; i1++) {
        const int level1 = oct1[i1];
        if (level1 > 0) continue;                                                         /* :426 */
        g->area(prev[2 * i1], prev[2 * i1 + 1], (float)windowSize, level1, level1, x2, y2, oct2, cand);
        if (cand.empty()) continue;
        int bestDist = 0x7fffffff, bestDist2 = 0x7fffffff, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = orc_descriptor_distance(d1 + 32 * (size_t)i1, d2 + 32 * (size_t)i2);
            if (matchedDistance[i2] <= dist) continue;                                    /* :448 */
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if ((float)bestDist < (float)bestDist2 * nnratio) {                          /* :464 */
                if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
                matches12[i1] = bestIdx2; matches21[bestIdx2] = i1; matchedDistance[bestIdx2] = bestDist;
                nmatches++;
                if (checkOri) rotHist[rot_bin(angle1[i1], angle2[bestIdx2])].push_back(i1);
            }
        }
    }
    delete g;
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < n1; i1++)                                                       /* :517-520 */
        if (matches12[i1] >= 0) { prev[2 * i1] = x2[matches12[i1]]; prev[2 * i1 + 1] = y2[matches12[i1]]; }
    return nmatches;
}

/* MapPoint::ComputeDistinctiveDescriptors (MapPoint.cc:247-312) and MapLine::ComputeDistinctiveDescriptors
   (MapLine.cpp:246-317): for each group of observed descriptors, the one with the least MEDIAN Hamming distance to the rest
   (median = sorted[int(0.5 * (N - 1))], first minimum wins).  Groups in CSR form: off[g] .. off[g+1]. */
extern "C" void orc_descriptor_medoid(const uint8_t* desc, const int32_t* off, int ngroups, int32_t* best_idx, int32_t* best_median) {
    std::vector<int> row;
    for (int g = 0; g < ngroups; g++) {
        const int b = off[g], N = off[g + 1] - b;
        best_idx[g] = N > 0 ? 0 : -1; best_median[g] = -1;
        if (N <= 0) continue;
        int BestMedian = 0x7fffffff, BestIdx = 0;
        for (int i = 0; i < N; i++) {
            row.assign(N, 0);
            for (int j = 0; j < N; j++) row[j] = i == j ? 0 : orc_descriptor_distance(desc + 32 * (size_t)(b + i), desc + 32 * (size_t)(b + j));
            std::sort(row.begin(), row.end());
            const int median = row[(size_t)(0.5 * (N - 1))];
            if (median < BestMedian) { BestMedian = median; BestIdx = i; }
        }
        best_idx[g] = BestIdx; best_median[g] = BestMedian;
    }
}

/* ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) — ORBmatcher.cc:525-658 */
extern "C" int orc_search_by_bow_kf(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                                    const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nn1,
                                    const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nn2,
                                    const uint8_t* valid1, const uint8_t* valid2, const float* angle1, const float* angle2,
                                    float nnratio, int checkOri, int32_t* match12) {
    for (int i = 0; i < n1; i++) match12[i] = -1;
    std::vector<char> matched2(n2, 0);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    walk_common_nodes(nodes1, nn1, nodes2, nn2, [&](int a, int b) {
        for (int i1 = off1[a]; i1 < off1[a + 1]; i1++) {
            const int id1 = idx1[i1];
            if (!valid1[id1]) continue;
            int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
            for (int i2 = off2[b]; i2 < off2[b + 1]; i2++) {
                const int id2 = idx2[i2];
                if (matched2[id2] || !valid2[id2]) continue;                          // :579-583
                int dist = orc_descriptor_distance(d1 + 32 * (size_t)id1, d2 + 32 * (size_t)id2);
                if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = id2; }
                else if (dist < bestDist2) bestDist2 = dist;
            }
            if (bestDist1 < TH_LOW) {                                                 // :601 (strict)
                if ((float)bestDist1 < nnratio * (float)bestDist2) {
                    match12[id1] = bestIdx2;
                    matched2[bestIdx2] = 1;
                    if (checkOri) rotHist[rot_bin(angle1[id1], angle2[bestIdx2])].push_back(id1);
                    nmatches++;
                }
            }
        }
    });
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match12[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

/* ORBmatcher::SearchForTriangulation — ORBmatcher.cc:660-826 with CheckDistEpipolarLine :140-157.
   Monocular restatement: mvuRight < 0 everywhere (bStereo1 = bStereo2 = false), bOnlyStereo = false. */
extern "C" int orc_search_for_triangulation(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                        const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nn1,
                        const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nn2,
                        const uint8_t* has_mp1, const uint8_t* has_mp2,
                        const float* x1, const float* y1, const float* a1,
                        const float* x2, const float* y2, const float* a2, const int32_t* oct2,
                        const float* F12, float ex, float ey, const float* scale, const float* sigma2,
                        int checkOri, int32_t* pairs) {
    (void)n2;
    std::vector<int> m12(n1, -1);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    walk_common_nodes(nodes1, nn1, nodes2, nn2, [&](int a, int b) {
        for (int i1 = off1[a]; i1 < off1[a + 1]; i1++) {
            const int id1 = idx1[i1];
            if (has_mp1[id1]) continue;                                               // :702-706
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (int i2 = off2[b]; i2 < off2[b + 1]; i2++) {
                const int id2 = idx2[i2];
                if (has_mp2[id2]) continue;                                           // :725-729 (vbMatched2 never set)
                const int dist = orc_descriptor_distance(d1 + 32 * (size_t)id1, d2 + 32 * (size_t)id2);
                if (dist > TH_LOW || dist > bestDist) continue;                       // :741
                {                                                                     // :746-752
                    const float distex = ex - x2[id2], distey = ey - y2[id2];
                    if (distex * distex + distey * distey < 100 * scale[oct2[id2]]) continue;
                }
                // CheckDistEpipolarLine :140-157
                const float la = x1[id1] * F12[0] + y1[id1] * F12[3] + F12[6];
                const float lb = x1[id1] * F12[1] + y1[id1] * F12[4] + F12[7];
                const float lc = x1[id1] * F12[2] + y1[id1] * F12[5] + F12[8];
                const float num = la * x2[id2] + lb * y2[id2] + lc;
                const float den = la * la + lb * lb;
                if (den == 0) continue;
                const float dsqr = num * num / den;
                if (dsqr < 3.84 * sigma2[oct2[id2]]) { bestIdx2 = id2; bestDist = dist; }   // double compare (:156)
            }
            if (bestIdx2 >= 0) {
                m12[id1] = bestIdx2;
                nmatches++;
                if (checkOri) rotHist[rot_bin(a1[id1], a2[bestIdx2])].push_back(id1);
            }
        }
    });
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { m12[j] = -1; nmatches--; }
        }
    }
    int k = 0;
    for (int i = 0; i < n1; i++) if (m12[i] >= 0) { pairs[2 * k] = i; pairs[2 * k + 1] = m12[i]; k++; }   // :818-823
    return nmatches;
}

/* Frame::lineDescriptorMAD — Frame.cc:190-215.  std::sort there is unstable but only the median VALUE
   is read, which does not depend on the order of equal elements. */
extern "C" void orc_line_mad(const int32_t* knn, int nq, double* nn_mad, double* nn12_mad) {
    if (nq <= 0) { *nn_mad = 0; *nn12_mad = 0; return; }
    std::vector<float> a(nq);
    for (int i = 0; i < nq; i++) a[i] = (float)knn[4 * i + 1];
    std::sort(a.begin(), a.end());
    double med = a[nq / 2];
    for (int i = 0; i < nq; i++) a[i] = fabsf((float)((float)knn[4 * i + 1] - med));
    std::sort(a.begin(), a.end());
    *nn_mad = 1.4826 * a[nq / 2];
    std::vector<float> g(nq);
    for (int i = 0; i < nq; i++) g[i] = (float)knn[4 * i + 3] - (float)knn[4 * i + 1];
    std::sort(g.begin(), g.end(), [](float x, float y) { return x > y; });      // descending (:206)
    double med12 = g[nq / 2];
    for (int i = 0; i < nq; i++) a[i] = fabsf((float)((float)knn[4 * i + 3] - (float)knn[4 * i + 1] - med12));
    std::sort(a.begin(), a.end());
    *nn12_mad = 1.4826 * a[nq / 2];
}

/* LSDmatcher knn-based entry points — LSDmatcher.cpp:143-183 / 286-327 (mode 0), 257-284 (mode 1),
   329-362 (mode 2), 382-415 (mode 3).  q = d1, train = d2 in all of them.
   mode 0: out[tdx] = qdx (table of size n2, -1 = none), later queries overwrite; returns nmatches (counts overwrites, as the reference does)
   mode 1: out = pairs (qdx,tdx);   mode 2: out[qdx] = tdx (table of size n1);   mode 3: out = pairs */
extern "C" int orc_line_match(int mode, const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                              const uint8_t* has_ml1, const uint8_t* has_ml2, int32_t* out, int* nout) {
    std::vector<int32_t> knn(4 * (size_t)std::max(n1, 1));
    orc_knn2(d1, n1, d2, n2, knn.data());
    double nn_mad, nn12_mad;
    orc_line_mad(knn.data(), n1, &nn_mad, &nn12_mad);
    int nmatches = 0, k = 0;
    if (mode == 0) for (int j = 0; j < n2; j++) out[j] = -1;
    if (mode == 2) for (int i = 0; i < n1; i++) out[i] = -1;
    const float minRatio = 1.0f / 1.5f;
    for (int i = 0; i < n1; i++) {                  // already in queryIdx order
        const int tdx = knn[4 * i];
        const float dist0 = (float)knn[4 * i + 1], dist1 = (float)knn[4 * i + 3];
        if (mode == 0) {
            double dist_12 = dist0 / dist1;                                     // float division, widened (:167)
            if (dist_12 < minRatio && has_ml1[i]) { out[tdx] = i; nmatches++; }
        } else if (mode == 1) {
            double dist_12 = dist1 - dist0;
            if (dist_12 > nn12_mad * 0.5) { out[2 * k] = i; out[2 * k + 1] = tdx; k++; nmatches++; }
        } else if (mode == 2) {
            double dist_12 = dist1 - dist0;
            if (dist_12 > nn12_mad * 0.5 && has_ml2[tdx]) { out[i] = tdx; nmatches++; }
        } else {
            if (has_ml1[i] || has_ml2[tdx]) continue;                           // :403
            double dist_12 = dist1 - dist0;
            if (dist_12 > nn12_mad * 0.1) { out[2 * k] = i; out[2 * k + 1] = tdx; k++; nmatches++; }
        }
    }
    if (nout) *nout = k;
    return nmatches;
}

/* ================================================================================================================
 * line projection matchers and Fuse.  Each reference function is split where the product splits it:
 * a projection stage (per map element: gates + projected quantities; host arithmetic in the reference's own Mat types)
 * and a search stage (Hamming scan over the frame's features; the device part).  orc_*_project restate the first,
 * orc_line_window_search / orc_fuse_*_search the second; tests/test_ref_parity_cpu.py composes them against the reference.
 * ================================================================================================================ */
namespace {
/* Rcw * X + tcw through cv::gemm's float path (3x3 * 3x1 + 3x1: products summed left to right, addend last) */
inline void rt_apply(const float* T, const float* X, float* out) {
    for (int r = 0; r < 3; r++) {
        float s = T[4 * r] * X[0]; s = s + T[4 * r + 1] * X[1]; s = s + T[4 * r + 2] * X[2];
        out[r] = s + T[4 * r + 3];
    }
}
/* Frame::GetLinesInArea (Frame.cc:423-460) = KeyFrame::GetLinesInArea (KeyFrame.cc:651-684), one line */
inline bool line_in_area(float x1, float y1, float x2, float y2, float r, int minLevel, int maxLevel, float ptx, float pty, float angle, int octave) {
    const bool bCheckLevels = (minLevel > 0) || (maxLevel > 0);
    const float distance = (0.5 * (x1 + x2) - ptx) * (0.5 * (x1 + x2) - ptx) + (0.5 * (y1 + y2) - pty) * (0.5 * (y1 + y2) - pty);
    if (distance > r * r) return false;
    const float slope = (y1 - y2) / (x1 - x2) - angle;
    if (slope > r * 0.01) return false;
    if (bCheckLevels) {
        if (octave < minLevel) return false;
        if (maxLevel >= 0 && octave > maxLevel) return false;
    }
    return true;
}
}  // namespace

/* Projection stage of LSDmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono) — LSDmatcher.cpp:22-96.
 * valid1[i]: pML && !pML->isBad() && !mvbLineOutlier[i]; Pw: GetWorldPos() narrowed to float as the Mat_<float> initialisers do (:48-49);
 * oct1[i] = LastFrame.mvKeys[i].octave (:84 reads the POINT keypoint of the same index).  Out: active, proj (x1 y1 x2 y2),
 * radius, minLevel, maxLevel per line. */
extern "C" void orc_line_project_frame(int nl1, const uint8_t* valid1, const float* Pw, const int32_t* oct1,
                                       const float* Tcw, const float* Tlw, const float* cam /* fx fy cx cy mb */, const float* bounds,
                                       const float* scaleFactors, float th, int bMono,
                                       uint8_t* active, float* proj, float* radius, int32_t* minLevel, int32_t* maxLevel) {
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], mb = cam[4];
    double twc[3];                                              /* twc = -Rcw.t()*tcw: transposed operand, double accumulation */
    for (int r = 0; r < 3; r++) twc[r] = -((double)Tcw[0 * 4 + r] * Tcw[3] + (double)Tcw[1 * 4 + r] * Tcw[7] + (double)Tcw[2 * 4 + r] * Tcw[11]);
    const float twcf[3] = {(float)twc[0], (float)twc[1], (float)twc[2]};
    float tlc[3]; rt_apply(Tlw, twcf, tlc);
    const bool bForward = tlc[2] > mb && !bMono, bBackward = -tlc[2] > mb && !bMono;
    for (int i = 0; i < nl1; i++) {
        active[i] = 0; radius[i] = 0; minLevel[i] = maxLevel[i] = -1;
        for (int k = 0; k < 4; k++) proj[4 * i + k] = 0;
        if (!valid1[i]) continue;
        float SPc[3], EPc[3];
        rt_apply(Tcw, Pw + 6 * (size_t)i, SPc); rt_apply(Tcw, Pw + 6 * (size_t)i + 3, EPc);
        if (SPc[2] < 0.0f || EPc[2] < 0.0f) continue;
        const float invz1 = 1.0f / SPc[2];
        const float u1 = fx * SPc[0] * invz1 + cx, v1 = fy * SPc[1] * invz1 + cy;
        if (u1 < bounds[0] || u1 > bounds[1]) continue;
        if (v1 < bounds[2] || v1 > bounds[3]) continue;
        const float invz2 = 1.0f / EPc[2];
        const float u2 = fx * EPc[0] * invz2 + cx, v2 = fy * EPc[1] * invz2 + cy;
        if (u2 < bounds[0] || u2 > bounds[1]) continue;
        if (v2 < bounds[2] || v2 > bounds[3]) continue;
        const int o = oct1[i];
        active[i] = 1; proj[4 * i] = u1; proj[4 * i + 1] = v1; proj[4 * i + 2] = u2; proj[4 * i + 3] = v2;
        radius[i] = th * scaleFactors[o];
        if (bForward) { minLevel[i] = o; maxLevel[i] = -1; }
        else if (bBackward) { minLevel[i] = 0; maxLevel[i] = o; }
        else { minLevel[i] = o - 1; maxLevel[i] = o + 1; }
    }
}

/* Projection stage of LSDmatcher::SearchByProjection(Frame&, const vector<MapLine*>&, th) — LSDmatcher.cpp:185-208: the fields
 * Frame::isInFrustum left on the MapLine (mbTrackInView, mnTrackScaleLevel, mTrackViewCos) become radius and level range. */
extern "C" void orc_line_project_mls(int nml, const uint8_t* inview, const uint8_t* bad, const int32_t* level, const float* viewcos,
                                     const float* scaleFactors, float th, uint8_t* active, float* radius, int32_t* minLevel, int32_t* maxLevel) {
    const bool bFactor = th != 1.0;
    for (int i = 0; i < nml; i++) {
        active[i] = inview[i] && !(bad && bad[i]);
        float r = viewcos[i] > 0.998 ? 5.0f : 8.0f;             /* RadiusByViewingCos :550-556 */
        if (bFactor) r *= th;
        radius[i] = active[i] ? r * scaleFactors[level[i]] : 0.f;
        minLevel[i] = level[i] - 1; maxLevel[i] = level[i];
    }
}

/* Search stage shared by both line SearchByProjection overloads (LSDmatcher.cpp:98-137 = :210-251).  kl2: pt.x, pt.y, angle of
 * mvKeylinesUn; oct2 their octave; held2[j]: 1 = the frame line holds a MapLine WITH observations (skipped), otherwise free;
 * obs[i]: the MapLine has observations (so a line it is written to is skipped by later MapLines).  assign2[j] = last writer. */
extern "C" int orc_line_window_search(int nml, const uint8_t* active, const uint8_t* obs, const float* proj, const float* radius,
                                      const int32_t* minLevel, const int32_t* maxLevel, const uint8_t* dml,
                                      int nl2, const uint8_t* ld2, const float* kl2, const int32_t* oct2, const uint8_t* held2,
                                      float nnratio, int32_t* assign2) {
    std::vector<uint8_t> claimed(nl2, 0);
    for (int j = 0; j < nl2; j++) { assign2[j] = -1; claimed[j] = held2 && held2[j] == 1; }
    int nmatches = 0;
    for (int i = 0; i < nml; i++) {
        if (!active[i]) continue;
        const float* p = proj + 4 * (size_t)i;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx = 0; idx < nl2; idx++) {
            if (!line_in_area(p[0], p[1], p[2], p[3], radius[i], minLevel[i], maxLevel[i], kl2[3 * idx], kl2[3 * idx + 1], kl2[3 * idx + 2], oct2[idx])) continue;
            if (claimed[idx]) continue;
            const int dist = orc_descriptor_distance(dml + 32 * (size_t)i, ld2 + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = oct2[idx]; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = oct2[idx]; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            assign2[bestIdx] = i; claimed[bestIdx] = obs && obs[i];
            nmatches++;
        }
    }
    return nmatches;
}

/* Projection stage of ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) — ORBmatcher.cc:828-894.  skip[i]: !pMP || isBad() ||
 * IsInKeyFrame(pKF); minInv/maxInv: Get{Min,Max}DistanceInvariance(); maxRaw: mfMaxDistance (PredictScale, MapPoint.cc:390-405,
 * logf through `using namespace std`).  cam: fx fy cx cy bf. */
extern "C" void orc_fuse_project_points(int nmp, const uint8_t* skip, const float* Xw, const float* normal, const float* minInv, const float* maxInv,
                                        const float* maxRaw, const float* Tcw, const float* Ow, const float* cam, const float* bounds,
                                        int nlevels, float logScaleFactor, uint8_t* active, float* u, float* v, float* ur, int32_t* level) {
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], bf = cam[4];
    for (int i = 0; i < nmp; i++) {
        active[i] = 0; u[i] = v[i] = ur[i] = 0; level[i] = 0;
        if (skip[i]) continue;
        const float* X = Xw + 3 * (size_t)i;
        float pc[3]; rt_apply(Tcw, X, pc);
        if (pc[2] < 0.0f) continue;
        const float invz = 1 / pc[2];
        const float x = pc[0] * invz, y = pc[1] * invz;
        const float uu = fx * x + cx, vv = fy * y + cy;
        if (!(uu >= bounds[0] && uu < bounds[1] && vv >= bounds[2] && vv < bounds[3])) continue;      /* KeyFrame::IsInImage KeyFrame.cc:686-689 */
        const float PO[3] = {X[0] - Ow[0], X[1] - Ow[1], X[2] - Ow[2]};
        const float dist3D = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);   /* cv::norm: double accumulation */
        if (dist3D < minInv[i] || dist3D > maxInv[i]) continue;
        const float* Pn = normal + 3 * (size_t)i;
        const double dot = (double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2];          /* Mat::dot: double accumulation */
        if (dot < 0.5 * dist3D) continue;
        const float ratio = maxRaw[i] / dist3D;
        int nScale = (int)std::ceil(logf(ratio) / logScaleFactor);
        if (nScale < 0) nScale = 0; else if (nScale >= nlevels) nScale = nlevels - 1;
        active[i] = 1; u[i] = uu; v[i] = vv; ur[i] = uu - bf * invz; level[i] = nScale;
    }
}

/* Search stage of ORBmatcher::Fuse — ORBmatcher.cc:896-950: best_idx[i] = nearest KeyFrame feature in the window (first in
 * GetFeaturesInArea order on ties) or -1, best_dist[i] its distance (256 when none).  The caller fuses when best_dist <= TH_LOW. */
extern "C" void orc_fuse_points_search(int nmp, const uint8_t* active, const float* u, const float* v, const float* ur, const int32_t* level, const uint8_t* dmp,
                                       int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* uright2,
                                       const float* bounds, const float* scaleFactors, const float* invLevelSigma2, float th,
                                       int32_t* best_idx, int32_t* best_dist) {
    Grid* g = new Grid();
    g->minX = bounds[0]; g->minY = bounds[2];
    g->invW = (float)Grid::COLS / (bounds[1] - bounds[0]); g->invH = (float)Grid::ROWS / (bounds[3] - bounds[2]);
    g->build(n2, x2, y2);
    std::vector<int> cand;
    for (int i = 0; i < nmp; i++) {
        best_idx[i] = -1; best_dist[i] = 256;
        if (!active[i]) continue;
        const int nPredictedLevel = level[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        g->area(u[i], v[i], radius, -1, -1, x2, y2, oct2, cand);                         /* KeyFrame::GetFeaturesInArea KeyFrame.cc:610-649: no level test */
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const int kpLevel = oct2[idx];
            if (kpLevel < nPredictedLevel - 1 || kpLevel > nPredictedLevel) continue;
            if (uright2 && uright2[idx] >= 0) {
                const float ex = u[i] - x2[idx], ey = v[i] - y2[idx], er = ur[i] - uright2[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * invLevelSigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = u[i] - x2[idx], ey = v[i] - y2[idx];
                const float e2 = ex * ex + ey * ey;
                if (e2 * invLevelSigma2[kpLevel] > 5.99) continue;
            }
            const int dist = orc_descriptor_distance(dmp + 32 * (size_t)i, d2 + 32 * (size_t)idx);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        best_idx[i] = bestIdx; best_dist[i] = bestDist;
    }
    delete g;
}

/* Projection stage of LSDmatcher::Fuse — LSDmatcher.cpp:417-497.  skip[i]: !pML || isBad(); Pw as in orc_line_project_frame;
 * normal: GetNormal() narrowed to float (:489); maxRaw: mfMaxDistance of MapLine::PredictScale (MapLine.cpp:386-395, NOT clamped:
 * a level outside [0, nlevels) makes the reference read mvScaleFactors out of bounds, here the line is dropped and level keeps the value). */
extern "C" void orc_fuse_project_lines(int nml, const uint8_t* skip, const float* Pw, const float* normal, const float* minInv, const float* maxInv,
                                       const float* maxRaw, const float* Tcw, const float* Ow, const float* cam, const float* bounds,
                                       int nlevels, float logScaleFactor, uint8_t* active, float* proj, int32_t* level) {
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    for (int i = 0; i < nml; i++) {
        active[i] = 0; level[i] = 0;
        for (int k = 0; k < 4; k++) proj[4 * i + k] = 0;
        if (skip[i]) continue;
        const float* SP = Pw + 6 * (size_t)i; const float* EP = SP + 3;
        float SPc[3], EPc[3];
        rt_apply(Tcw, SP, SPc); rt_apply(Tcw, EP, EPc);
        if (SPc[2] < 0.0f || EPc[2] < 0.0f) continue;
        const float invz1 = 1.0f / SPc[2];
        const float u1 = fx * SPc[0] * invz1 + cx, v1 = fy * SPc[1] * invz1 + cy;
        if (u1 < bounds[0] || u1 > bounds[1]) continue;
        if (v1 < bounds[2] || v1 > bounds[3]) continue;
        const float invz2 = 1.0f / EPc[2];
        const float u2 = fx * EPc[0] * invz2 + cx, v2 = fy * EPc[1] * invz2 + cy;
        if (u2 < bounds[0] || u2 > bounds[1]) continue;
        if (v2 < bounds[2] || v2 > bounds[3]) continue;
        float OM[3];
        for (int k = 0; k < 3; k++) OM[k] = 0.5f * (SP[k] + EP[k]) - Ow[k];               /* 0.5*(SP+EP) - Ow: float, halving is exact */
        const float dist = (float)std::sqrt((double)OM[0] * OM[0] + (double)OM[1] * OM[1] + (double)OM[2] * OM[2]);
        if (dist < minInv[i] || dist > maxInv[i]) continue;
        const float* pn = normal + 3 * (size_t)i;
        const double dot = (double)OM[0] * pn[0] + (double)OM[1] * pn[1] + (double)OM[2] * pn[2];
        if (dot < 0.5 * dist) continue;
        const float ratio = maxRaw[i] / dist;
        const int lvl = (int)std::ceil(logf(ratio) / logScaleFactor);
        level[i] = lvl;
        if (lvl < 0 || lvl >= nlevels) continue;
        active[i] = 1; proj[4 * i] = u1; proj[4 * i + 1] = v1; proj[4 * i + 2] = u2; proj[4 * i + 3] = v2;
    }
}

/* Search stage of LSDmatcher::Fuse — LSDmatcher.cpp:499-523: oct2 = mvKeyLines[idx].octave; best_dist INT_MAX when nothing qualified */
extern "C" void orc_fuse_lines_search(int nml, const uint8_t* active, const float* proj, const int32_t* level, const uint8_t* dml,
                                      int nl2, const uint8_t* ld2, const float* kl2, const int32_t* oct2, const float* scaleFactors, float th,
                                      int32_t* best_idx, int32_t* best_dist) {
    for (int i = 0; i < nml; i++) {
        best_idx[i] = -1; best_dist[i] = 0x7fffffff;
        if (!active[i]) continue;
        const float* p = proj + 4 * (size_t)i;
        const int nPredictedLevel = level[i];
        const float radius = th * scaleFactors[nPredictedLevel];
        for (int idx = 0; idx < nl2; idx++) {
            if (!line_in_area(p[0], p[1], p[2], p[3], radius, -1, -1, kl2[3 * idx], kl2[3 * idx + 1], kl2[3 * idx + 2], oct2[idx])) continue;
            const int klLevel = oct2[idx];
            if (klLevel < nPredictedLevel - 1 || klLevel > nPredictedLevel) continue;
            const int dist = orc_descriptor_distance(dml + 32 * (size_t)i, ld2 + 32 * (size_t)idx);
            if (dist < best_dist[i]) { best_dist[i] = dist; best_idx[i] = idx; }
        }
    }
}
