// tests/integration/ref_link_test.cpp — TEST INFRASTRUCTURE.
// "Tracking.cc links unchanged": this program is linked from
//   * the reference's own, unmodified translation units (src/Frame.cc, KeyFrame.cc, MapPoint.cc, MapLine.cpp, Map.cc,
//     KeyFrameDatabase.cc, ORBmatcher.cc, LSDmatcher.cpp, Thirdparty/DBoW2) compiled against the reference's real headers (with the
//     OpenCV / Eigen stand-ins of oracle/refshim, since OpenCV C++ is not installed) — the symbols of ORBmatcher.cc / LSDmatcher.cpp
//     are made weak with objcopy, which is what "delete the replaced bodies" of INTEGRATION.md amounts to;
//   * the adapters of structure-slam-pointline_b200/host/ (ORBextractor.cc, ExtractLineSegment_b200.cc, matcher_b200.cc, bow_b200.cc);
//   * libsslpl_b200.so.
// It then does what Tracking does with them: Frame::Frame(imGray, ...) (Frame.cc:69 -> the adapters' ORB and LSD extraction on the
// GPU), Frame::ComputeBoW with a DBoW2 vocabulary, KeyFrame construction, ORBmatcher::SearchByBoW / SearchByProjection / Fuse and
// LSDmatcher::SearchByProjection (KeyFrame and MapLine forms) / Fuse through the adapters, and dumps inputs and outputs as .npy files for tests/test_integration_gpu.py,
// which holds them against the fixtures frozen from the reference and against oracle/_ref.
#include <cstdio>
#include <cstdint>
#include <fstream>
#include <map>
#include <string>
#include <vector>
#define private public
#define protected public
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include "LSDmatcher.h"
#include "Frame.h"
#include "KeyFrame.h"
#include "MapPoint.h"
#include "MapLine.h"
#include "Map.h"
#include "KeyFrameDatabase.h"
#include "ORBVocabulary.h"
#include "Converter.h"
#undef private
#undef protected

using namespace StructureSLAM;
namespace StructureSLAM { void SslplSetVocabularyFile(const std::string& strVocFile); }   // host/bow_b200.cc: the one line added after System.cc:65

std::vector<cv::Mat> StructureSLAM::Converter::toDescriptorVector(const cv::Mat& Descriptors) {   // Converter.cc:30-38 (Converter.cc needs g2o)
    std::vector<cv::Mat> v; v.reserve(Descriptors.rows);
    for (int j = 0; j < Descriptors.rows; j++) v.push_back(Descriptors.row(j));
    return v;
}

static void npy(const std::string& dir, const std::string& name, const char* descr, const void* data, size_t elem, std::vector<size_t> shape) {
    std::string sh = "(";
    size_t n = 1;
    for (size_t i = 0; i < shape.size(); i++) { sh += std::to_string(shape[i]) + ","; n *= shape[i]; }
    sh += ")";
    std::string hdr = std::string("{'descr': '") + descr + "', 'fortran_order': False, 'shape': " + sh + ", }";
    while ((10 + hdr.size() + 1) % 64) hdr += ' ';
    hdr += '\n';
    std::ofstream f(dir + "/" + name + ".npy", std::ios::binary);
    const char magic[] = "\x93NUMPY\x01\x00";
    f.write(magic, 8);
    const uint16_t hl = (uint16_t)hdr.size(); f.write((const char*)&hl, 2); f.write(hdr.data(), hdr.size());
    f.write((const char*)data, (std::streamsize)(n * elem));
}

static void dump_frame(const std::string& dir, const std::string& tag, Frame& F) {
    npy(dir, tag + "_keys", "V28", F.mvKeys.data(), 28, {F.mvKeys.size()});
    npy(dir, tag + "_keysun", "V28", F.mvKeysUn.data(), 28, {F.mvKeysUn.size()});
    npy(dir, tag + "_desc", "|u1", F.mDescriptors.data, 1, {(size_t)F.mDescriptors.rows, 32});
    npy(dir, tag + "_keylines", "V68", F.mvKeylinesUn.data(), 68, {F.mvKeylinesUn.size()});
    npy(dir, tag + "_ldesc", "|u1", F.mLdesc.data, 1, {(size_t)F.mLdesc.rows, 32});
    std::vector<double> eq; for (auto& v : F.mvKeyLineFunctions) for (int k = 0; k < 3; k++) eq.push_back(v(k));
    npy(dir, tag + "_lineeq", "<f8", eq.data(), 8, {F.mvKeyLineFunctions.size(), 3});
    std::vector<int32_t> node(F.N, -1);
    for (auto& kv : F.mFeatVec) for (unsigned i : kv.second) node[i] = (int32_t)kv.first;
    npy(dir, tag + "_node", "<i4", node.data(), 4, {node.size()});
}

int main(int argc, char** argv) {
    if (argc < 6) { fprintf(stderr, "usage: ref_link_test gray1.raw gray2.raw w h vocabulary.txt outdir\n"); return 2; }
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    const std::string out = argv[6];
    cv::Mat im[2];
    for (int i = 0; i < 2; i++) {
        im[i] = cv::Mat(h, w, CV_8UC1);
        std::ifstream f(argv[1 + i], std::ios::binary);
        f.read((char*)im[i].data, (std::streamsize)w * h);
        if (!f) { fprintf(stderr, "cannot read %s\n", argv[1 + i]); return 2; }
    }
    ORBVocabulary voc;
    if (!voc.loadFromTextFile(argv[5])) { fprintf(stderr, "vocabulary?\n"); return 2; }
    SslplSetVocabularyFile(argv[5]);                                 // the device copy of the tree for Frame::ComputeBoW (GPU transform)
    ORBextractor ext(1000, 1.2f, 8, 20, 7);                          // Tracking.cc:119 (the adapter class)
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = 481.2f; K.at<float>(1, 1) = 480.0f; K.at<float>(0, 2) = 319.5f; K.at<float>(1, 2) = 239.5f;
    cv::Mat D = cv::Mat::zeros(4, 1, CV_32F);
    Frame::mbInitialComputations = true;
    Frame F1(im[0], 0.0, &ext, &voc, K, D, 0.f, 0.f);                // Frame.cc:69: ExtractORB + ExtractLSD through the adapters
    Frame F2(im[1], 1.0, &ext, &voc, K, D, 0.f, 0.f);
    F1.ComputeBoW(); F2.ComputeBoW();                                // Frame.cc:474 -> host/bow_b200.cc: the DBoW2 transform on the GPU
    cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
    F1.SetPose(T); F2.SetPose(T);
    dump_frame(out, "f1", F1); dump_frame(out, "f2", F2);

    Map map; KeyFrameDatabase db(voc);
    KeyFrame* KF = new KeyFrame(F1, &map, &db);                      // Tracking.cc:1494
    std::vector<uint8_t> state(F1.N, 0);
    std::map<MapPoint*, int> index; std::map<MapLine*, int> lindex;
    for (int i = 0; i < F1.N; i++) if (i % 10 != 3) {                // every feature but one in ten has a MapPoint; a few are bad
        cv::Mat X = (cv::Mat_<float>(3, 1) << (F1.mvKeysUn[i].pt.x - 319.5f) / 481.2f * 2.f, (F1.mvKeysUn[i].pt.y - 239.5f) / 480.f * 2.f, 2.f);
        MapPoint* p = new MapPoint(X, KF, &map);
        F1.mDescriptors.row(i).copyTo(p->mDescriptor);
        KF->AddMapPoint(p, i); p->AddObservation(KF, i);
        F1.mvpMapPoints[i] = p;
        state[i] = 1; index[p] = i;
        if (i % 37 == 5) { p->mbBad = true; state[i] = 2; }
    }
    npy(out, "state1", "|u1", state.data(), 1, {state.size()});
    {   // ORBmatcher::SearchByBoW(KeyFrame*, Frame&) — Tracking::TrackReferenceKeyFrame (Tracking.cc:1020)
        ORBmatcher matcher(0.7f, true);
        std::vector<MapPoint*> m;
        const int n = matcher.SearchByBoW(KF, F2, m);
        std::vector<int32_t> o(F2.N + 1, -1);
        for (int j = 0; j < F2.N; j++) o[j] = m[j] ? index[m[j]] : -1;
        o[F2.N] = n;
        npy(out, "bow_match2", "<i4", o.data(), 4, {o.size()});
    }
    {   // ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) — Tracking::TrackWithMotionModel (Tracking.cc:1227)
        ORBmatcher matcher(0.9f, true);
        cv::Mat Tcw = cv::Mat::eye(4, 4, CV_32F);
        Tcw.at<float>(0, 3) = 0.01f; Tcw.at<float>(1, 3) = -0.004f;
        F2.SetPose(Tcw);
        std::fill(F2.mvpMapPoints.begin(), F2.mvpMapPoints.end(), static_cast<MapPoint*>(NULL));   // Tracking.cc:1226
        const int n = matcher.SearchByProjection(F2, F1, 15.f, true);
        std::vector<int32_t> o(F2.N + 1, -1);
        for (int j = 0; j < F2.N; j++) o[j] = F2.mvpMapPoints[j] ? index[F2.mvpMapPoints[j]] : -1;
        o[F2.N] = n;
        npy(out, "proj_assign2", "<i4", o.data(), 4, {o.size()});
        std::vector<float> tc(12); for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) tc[4 * r + c] = Tcw.at<float>(r, c);
        npy(out, "proj_Tcw", "<f4", tc.data(), 4, {12});
        std::vector<float> xw((size_t)3 * F1.N, 0.f);
        for (auto& kv : index) { cv::Mat X = kv.first->GetWorldPos(); for (int k = 0; k < 3; k++) xw[3 * kv.second + k] = X.at<float>(k); }
        npy(out, "proj_Xw", "<f4", xw.data(), 4, {(size_t)F1.N, 3});
    }
    {   // LSDmatcher::SearchByProjection(KeyFrame*, Frame&, ...) — Tracking.cc:1024, :1234
        std::vector<uint8_t> has(KF->NL, 0);
        Vector6d P; P << 0.0, 0.0, 1.0, 1.0, 0.0, 1.0;
        for (int i = 0; i < KF->NL; i++) if (i % 4 != 1) { MapLine* l = new MapLine(P, KF, &map); KF->AddMapLine(l, i); has[i] = 1; lindex[l] = i; }
        npy(out, "has_ml1", "|u1", has.data(), 1, {has.size()});
        LSDmatcher lm;
        std::vector<MapLine*> m;
        const int n = lm.SearchByProjection(KF, F2, m);
        std::vector<int32_t> o(F2.NL + 1, -1);
        for (int j = 0; j < F2.NL; j++) o[j] = m[j] ? lindex[m[j]] : -1;
        o[F2.NL] = n;
        npy(out, "line_match2", "<i4", o.data(), 4, {o.size()});
    }
    // ---- SURVEY.md 8(f) row 3 through the adapters: Fuse (points, lines) and LSDmatcher::SearchByProjection(Frame&, MapLines) ----
    std::fill(F2.mvpMapPoints.begin(), F2.mvpMapPoints.end(), static_cast<MapPoint*>(NULL));     // forget the tracking matches made above
    std::fill(F2.mvpMapLines.begin(), F2.mvpMapLines.end(), static_cast<MapLine*>(NULL));
    KeyFrame* KF2 = new KeyFrame(F2, &map, &db);                     // pose = the Tcw set above
    {   // ORBmatcher::Fuse(KeyFrame*, vector<MapPoint*>) — LocalMapping::SearchInNeighbors (LocalMapping.cc:1206)
        std::vector<MapPoint*> mps(F1.mvpMapPoints);
        std::vector<float> nrm((size_t)3 * F1.N, 0.f), dmin(F1.N, 0.f), dmax(F1.N, 0.f);
        for (auto& kv : index) {
            MapPoint* p = kv.first;
            if (!p->isBad()) p->UpdateNormalAndDepth();                // MapPoint.cc:314-358: normal, mfMinDistance, mfMaxDistance from its observation in KF
            const int i = kv.second;
            for (int k = 0; k < 3; k++) nrm[3 * i + k] = p->mNormalVector.empty() ? 0.f : p->mNormalVector.at<float>(k);
            dmin[i] = p->mfMinDistance; dmax[i] = p->mfMaxDistance;
        }
        npy(out, "fuse_normal", "<f4", nrm.data(), 4, {(size_t)F1.N, 3});
        npy(out, "fuse_dmin", "<f4", dmin.data(), 4, {dmin.size()}); npy(out, "fuse_dmax", "<f4", dmax.data(), 4, {dmax.size()});
        std::vector<int32_t> kfobs(F2.N, -1);
        for (int j = 0; j < F2.N; j += 7) {                           // some KeyFrame features already hold a MapPoint (the Replace branch)
            cv::Mat X = (cv::Mat_<float>(3, 1) << 0.f, 0.f, 1000.f);
            MapPoint* q = new MapPoint(X, KF2, &map);
            F2.mDescriptors.row(j).copyTo(q->mDescriptor);
            q->AddObservation(KF2, j); KF2->AddMapPoint(q, j); kfobs[j] = 0;
        }
        npy(out, "fuse_kfobs", "<i4", kfobs.data(), 4, {kfobs.size()});
        ORBmatcher matcher(0.6f, true);
        const int n = matcher.Fuse(KF2, mps, 3.0f);
        std::vector<int32_t> o(F1.N + 1, -1);
        for (int i = 0; i < F1.N; i++) {
            MapPoint* p = mps[i];
            if (!p || state[i] != 1) continue;
            int hops = 0;
            while (p && p->isBad() && hops++ < 8) p = p->GetReplaced();
            if (p && !p->isBad() && p->IsInKeyFrame(KF2)) o[i] = p->GetIndexInKeyFrame(KF2);
        }
        o[F1.N] = n;
        npy(out, "fuse_idx", "<i4", o.data(), 4, {o.size()});
    }
    {   // LSDmatcher::Fuse(KeyFrame*, vector<MapLine*>) (LocalMapping.cc:1243) and LSDmatcher::SearchByProjection(Frame&, MapLines, th) (Tracking.cc:1783)
        const int nl = F1.NL;
        std::vector<MapLine*> mls(nl, static_cast<MapLine*>(NULL));
        std::vector<double> pw((size_t)6 * nl, 0.0), nrm((size_t)3 * nl, 0.0);
        std::vector<float> dmin(nl, 0.f), dmax(nl, 0.f), proj((size_t)4 * nl, 0.f);
        std::vector<uint8_t> lstate(nl, 0);
        for (int i = 0; i < nl; i++) if (i % 9 != 4) {
            const KeyLine& k = F1.mvKeylinesUn[i];
            const double z = 2.0 + 0.01 * i;
            Vector6d P;
            P << (double)(float)((k.startPointX - 319.5f) / 481.2f * z), (double)(float)((k.startPointY - 239.5f) / 480.f * z), (double)(float)z,
                 (double)(float)((k.endPointX - 319.5f) / 481.2f * z), (double)(float)((k.endPointY - 239.5f) / 480.f * z), (double)(float)z;
            MapLine* l = new MapLine(P, KF, &map);
            F1.mLdesc.row(i).copyTo(l->mLDescriptor);
            l->AddObservation(KF, i);
            Vector3d mid; mid << 0.5 * (P(0) + P(3)), 0.5 * (P(1) + P(4)), 0.5 * (P(2) + P(5));
            const double d = mid.norm();
            l->mNormalVector << (double)(float)(mid(0) / d), (double)(float)(mid(1) / d), (double)(float)(mid(2) / d);
            l->mfMaxDistance = (float)(d * (i % 3 == 0 ? 1.15 : 1.0)); l->mfMinDistance = l->mfMaxDistance / 3.f;   // predicted level 1 or 0
            l->mbTrackInView = (i % 11 != 7); l->mnTrackScaleLevel = 0; l->mTrackViewCos = (i % 2) ? 0.9999f : 0.9f;
            l->mTrackProjX1 = k.startPointX; l->mTrackProjY1 = k.startPointY; l->mTrackProjX2 = k.endPointX; l->mTrackProjY2 = k.endPointY;
            if (i % 13 == 6) l->mbBad = true;
            mls[i] = l; lstate[i] = l->mbBad ? 2 : 1;
            for (int c = 0; c < 6; c++) pw[6 * i + c] = P(c);
            for (int c = 0; c < 3; c++) nrm[3 * i + c] = l->mNormalVector(c);
            dmin[i] = l->mfMinDistance; dmax[i] = l->mfMaxDistance;
            proj[4 * i] = k.startPointX; proj[4 * i + 1] = k.startPointY; proj[4 * i + 2] = k.endPointX; proj[4 * i + 3] = k.endPointY;
        }
        npy(out, "lfuse_state", "|u1", lstate.data(), 1, {lstate.size()}); npy(out, "lfuse_pw", "<f8", pw.data(), 8, {(size_t)nl, 6});
        npy(out, "lfuse_normal", "<f8", nrm.data(), 8, {(size_t)nl, 3}); npy(out, "lfuse_dmin", "<f4", dmin.data(), 4, {dmin.size()});
        npy(out, "lfuse_dmax", "<f4", dmax.data(), 4, {dmax.size()}); npy(out, "lproj", "<f4", proj.data(), 4, {(size_t)nl, 4});
        {   // SearchByProjection(Frame&, MapLines) first: it only writes F2.mvpMapLines
            LSDmatcher lm(0.8f, true);
            std::fill(F2.mvpMapLines.begin(), F2.mvpMapLines.end(), static_cast<MapLine*>(NULL));
            const int n = lm.SearchByProjection(F2, mls, 3.0f);
            std::map<MapLine*, int> li; for (int i = 0; i < nl; i++) if (mls[i]) li[mls[i]] = i;
            std::vector<int32_t> o(F2.NL + 1, -1);
            for (int j = 0; j < F2.NL; j++) o[j] = F2.mvpMapLines[j] ? li[F2.mvpMapLines[j]] : -1;
            o[F2.NL] = n;
            npy(out, "lproj_assign2", "<i4", o.data(), 4, {o.size()});
        }
        std::vector<int32_t> kfobs(F2.NL, -1);
        Vector6d P0; P0 << 0.0, 0.0, 1.0, 1.0, 0.0, 1.0;
        for (int j = 0; j < F2.NL; j += 5) { MapLine* q = new MapLine(P0, KF2, &map); F2.mLdesc.row(j).copyTo(q->mLDescriptor); q->AddObservation(KF2, j); KF2->AddMapLine(q, j); kfobs[j] = 0; }
        npy(out, "lfuse_kfobs", "<i4", kfobs.data(), 4, {kfobs.size()});
        LSDmatcher lm(0.6f, true);
        const int n = lm.Fuse(KF2, mls, 10.0f);
        std::vector<int32_t> o(nl + 1, -1);
        for (int i = 0; i < nl; i++) {
            MapLine* l = mls[i];
            if (!l || lstate[i] != 1) continue;
            int hops = 0;
            while (l && l->isBad() && hops++ < 8) l = l->GetReplaced();
            if (l && !l->isBad() && l->IsInKeyFrame(KF2)) o[i] = l->GetIndexInKeyFrame(KF2);
        }
        o[nl] = n;
        npy(out, "lfuse_idx", "<i4", o.data(), 4, {o.size()});
    }
    printf("ref_link_test ok: N=%d/%d NL=%d/%d\n", F1.N, F2.N, F1.NL, F2.NL);
    return 0;
}
