/*
 * oracle/oracle.h — C ABI of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  The oracle is a plain CPU restatement of the reference algorithm
 * (yanyan-li/Structure-SLAM-PointLine, files cited per function in the .cpp files).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The
 * product (structure-slam-pointline_b200/, include/sslpl.h) never includes, links or calls it.
 */
#ifndef SSLPL_ORACLE_H
#define SSLPL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* mirrors cv::KeyPoint (28 B) */
typedef struct { float x, y, size, angle, response; int32_t octave, class_id; } orc_keypoint;

/* mirrors cv::line_descriptor::KeyLine (68 B) */
typedef struct {
    float angle; int32_t class_id; int32_t octave; float pt_x, pt_y; float response; float size;
    float startPointX, startPointY, endPointX, endPointY;
    float sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int32_t numOfPixels;
} orc_keyline;

/* ---------------- primitives (pinned against cv2 4.13 in tests/) ---------------- */
void  orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int spitch, uint8_t* dst, int dw, int dh, int dpitch);
void  orc_border_reflect101_u8(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch, int b);
void  orc_sepfilter_fixed_u8(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch,
                             const int* taps, int ntaps);           /* 8.8 fixed-point, REFLECT_101 */
void  orc_gauss7_sigma2_u8(const uint8_t* src, int w, int h, int spitch, uint8_t* dst, int dpitch);
int   orc_fast9_16(const uint8_t* img, int cols, int rows, int pitch, int threshold,
                   int* xs, int* ys, int* scores, int cap);         /* cv::FAST(...,true) */
float orc_fast_atan2(float y, float x);

/* ---------------- ORB extractor (src/ORBextractor.cc) ---------------- */
typedef struct orc_orb orc_orb;
orc_orb* orc_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
void  orc_orb_destroy(orc_orb*);
/* tables computed by the constructor (ORBextractor.cc:410-470) */
void  orc_orb_tables(const orc_orb*, float* scale, float* invscale, float* sigma2, float* invsigma2,
                     int* nfeat_per_level, int* umax16);
/* operator() (ORBextractor.cc:1043-1105); returns number of keypoints (<= cap) */
int   orc_orb_extract(orc_orb*, const uint8_t* img, int w, int h, int pitch,
                      orc_keypoint* kps, uint8_t* desc, int cap);
/* intermediates of the last extract call (for stage-by-stage parity tests) */
void  orc_orb_level_size(const orc_orb*, int level, int* w, int* h);
void  orc_orb_level_copy(const orc_orb*, int level, int bordered, uint8_t* dst, int dpitch);
void  orc_orb_blur_copy(const orc_orb*, int level, uint8_t* dst, int dpitch);
int   orc_orb_candidates(const orc_orb*, int level, int* xs, int* ys, int* resp, int cap);
int   orc_orb_level_keypoints(const orc_orb*, int level, int* xs, int* ys, int* resp, float* angle, int cap);
/* stage-level entry points (IC_Angle :77-104 on an un-blurred image; computeOrbDescriptor :107-147 on an already blurred one) */
void  orc_ic_angles(const orc_orb*, const uint8_t* img, int w, int h, int pitch, const int* xs, const int* ys, int n, float* angles);
void  orc_brief_descriptors(const uint8_t* blurred, int w, int h, int pitch, const int* xs, const int* ys, const float* angles, int n, uint8_t* desc);
/* per-stage wall times (ms) of the last extract call: pyramid, fast, octree, orient, blur, brief */
void  orc_orb_stage_ms(const orc_orb*, double* ms6);
/* DistributeOctTree alone (ORBextractor.cc:539-763) with the canonical (size,counter) tie rule */
int   orc_octree(const int* xs, const int* ys, const int* resp, int n, int minX, int maxX, int minY, int maxY,
                 int N, int* out_idx, int cap);

/* ---------------- matching (src/ORBmatcher.cc, src/LSDmatcher.cpp, src/Frame.cc) ---------------- */
int   orc_descriptor_distance(const uint8_t* a, const uint8_t* b);          /* ORBmatcher.cc:1650 */
/* cv::BFMatcher(NORM_HAMMING).knnMatch(q,t,.,2): out[4*i] = idx0,d0,idx1,d1 (-1/-1 when nt<2) */
void  orc_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* out);
/* synthetic-vocabulary node assignment (stands in for DBoW2 transform): nearest centroid, first wins */
void  orc_bow_assign(const uint8_t* desc, int n, const uint8_t* centroids, int nc, int32_t* node);
/* DBoW2 transform on a vocabulary tree given as arrays (TemplatedVocabulary.h:1127-1259); returns the number of words or -1 */
int   orc_vocab_transform(int L, int nnodes, const int32_t* parent, const uint8_t* ndesc, const double* weight,
                          const uint8_t* is_leaf, const uint8_t* feats, int n, int levelsup,
                          int32_t* word, int32_t* node, double* w);
/* MapPoint / MapLine ::ComputeDistinctiveDescriptors (MapPoint.cc:247-312, MapLine.cpp:246-317) for CSR groups of descriptors */
void  orc_descriptor_medoid(const uint8_t* desc, const int32_t* off, int ngroups, int32_t* best_idx, int32_t* best_median);
/* SearchByBoW(KeyFrame*,Frame&) ORBmatcher.cc:159-291. FeatureVectors in CSR (node ids ascending).
   valid1[i]!=0 <=> KF feature i has a good MapPoint. match2[j] = KF index matched to frame feature j or -1. */
int   orc_search_by_bow(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                        const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nn1,
                        const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nn2,
                        const uint8_t* valid1, const float* angle1, const float* angle2,
                        float nnratio, int checkOri, int32_t* match2);
/* SearchByBoW(KeyFrame*,KeyFrame*) ORBmatcher.cc:525-658: match12[i] = KF2 index or -1 */
int   orc_search_by_bow_kf(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                           const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nn1,
                           const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nn2,
                           const uint8_t* valid1, const uint8_t* valid2, const float* angle1, const float* angle2,
                           float nnratio, int checkOri, int32_t* match12);
/* SearchForTriangulation ORBmatcher.cc:660-826 (monocular: bOnlyStereo=false, no stereo keypoints).
   has_mp[i]!=0 <=> feature already has a MapPoint (skipped). kp arrays: x,y,angle,octave of mvKeysUn. */
int   orc_search_for_triangulation(const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                        const int32_t* nodes1, const int32_t* off1, const int32_t* idx1, int nn1,
                        const int32_t* nodes2, const int32_t* off2, const int32_t* idx2, int nn2,
                        const uint8_t* has_mp1, const uint8_t* has_mp2,
                        const float* x1, const float* y1, const float* a1,
                        const float* x2, const float* y2, const float* a2, const int32_t* oct2,
                        const float* F12, float ex, float ey, const float* scale, const float* sigma2,
                        int checkOri, int32_t* pairs /*2*n1*/);
/* SearchByProjection(Frame&, vector<MapPoint*>&, th) ORBmatcher.cc:45-129 and SearchForInitialization :408-523 (see match_oracle.cpp) */
int   orc_search_by_projection_mps(int nmp, const uint8_t* inview, const uint8_t* bad, const uint8_t* obs, const float* projx, const float* projy,
                                   const float* projxr, const int32_t* level, const float* viewcos, const uint8_t* dmp,
                                   int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* uright2,
                                   const uint8_t* held2, const float* bounds, const float* scaleFactors, float nnratio, float th, int32_t* assign2);
int   orc_search_for_initialization(int n1, const uint8_t* d1, const int32_t* oct1, const float* angle1,
                                    int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* angle2,
                                    float* prev, const float* bounds, float nnratio, int checkOri, int windowSize, int32_t* matches12);
/* line projection matchers (LSDmatcher.cpp:22-141, :185-255) and Fuse (ORBmatcher.cc:828-973, LSDmatcher.cpp:417-548),
   each split into its projection stage and its search stage (see match_oracle.cpp) */
void  orc_line_project_frame(int nl1, const uint8_t* valid1, const float* Pw, const int32_t* oct1, const float* Tcw, const float* Tlw,
                             const float* cam, const float* bounds, const float* scaleFactors, float th, int bMono,
                             uint8_t* active, float* proj, float* radius, int32_t* minLevel, int32_t* maxLevel);
void  orc_line_project_mls(int nml, const uint8_t* inview, const uint8_t* bad, const int32_t* level, const float* viewcos,
                           const float* scaleFactors, float th, uint8_t* active, float* radius, int32_t* minLevel, int32_t* maxLevel);
int   orc_line_window_search(int nml, const uint8_t* active, const uint8_t* obs, const float* proj, const float* radius,
                             const int32_t* minLevel, const int32_t* maxLevel, const uint8_t* dml,
                             int nl2, const uint8_t* ld2, const float* kl2, const int32_t* oct2, const uint8_t* held2, float nnratio, int32_t* assign2);
void  orc_fuse_project_points(int nmp, const uint8_t* skip, const float* Xw, const float* normal, const float* minInv, const float* maxInv,
                              const float* maxRaw, const float* Tcw, const float* Ow, const float* cam, const float* bounds,
                              int nlevels, float logScaleFactor, uint8_t* active, float* u, float* v, float* ur, int32_t* level);
void  orc_fuse_points_search(int nmp, const uint8_t* active, const float* u, const float* v, const float* ur, const int32_t* level, const uint8_t* dmp,
                             int n2, const uint8_t* d2, const float* x2, const float* y2, const int32_t* oct2, const float* uright2,
                             const float* bounds, const float* scaleFactors, const float* invLevelSigma2, float th, int32_t* best_idx, int32_t* best_dist);
void  orc_fuse_project_lines(int nml, const uint8_t* skip, const float* Pw, const float* normal, const float* minInv, const float* maxInv,
                             const float* maxRaw, const float* Tcw, const float* Ow, const float* cam, const float* bounds,
                             int nlevels, float logScaleFactor, uint8_t* active, float* proj, int32_t* level);
void  orc_fuse_lines_search(int nml, const uint8_t* active, const float* proj, const int32_t* level, const uint8_t* dml,
                            int nl2, const uint8_t* ld2, const float* kl2, const int32_t* oct2, const float* scaleFactors, float th,
                            int32_t* best_idx, int32_t* best_dist);
/* Frame::lineDescriptorMAD Frame.cc:190-215 on a knn2 table */
void  orc_line_mad(const int32_t* knn, int nq, double* nn_mad, double* nn12_mad);
/* LSDmatcher knn-based entry points. mode: 0 = SearchByProjection(KF,F)/SearchByDescriptor(KF,F) (ratio),
   1 = SerachForInitialize (pairs), 2 = SearchByDescriptor(KF,KF2), 3 = SearchForTriangulation */
int   orc_line_match(int mode, const uint8_t* d1, int n1, const uint8_t* d2, int n2,
                     const uint8_t* has_ml1, const uint8_t* has_ml2, int32_t* out, int* nout);

/* ---------------- lines (src/ExtractLineSegment.cpp + OpenCV lsd.cpp / line_descriptor) ---------------- */
typedef struct orc_line orc_line;
orc_line* orc_line_create(int lsdNFeatures);
void  orc_line_destroy(orc_line*);
int   orc_line_extract(orc_line*, const uint8_t* img, int w, int h, int pitch,
                       orc_keyline* kl, uint8_t* ldesc, double* lineeq3, int cap);
/* the two OpenCV-contrib steps on their own: LSDDetector::detect (all KeyLines, detection order; returns the count even when
   cap is smaller) and BinaryDescriptor::compute on given KeyLines */
int   orc_lsd_keylines(const uint8_t* img, int w, int h, int pitch, orc_keyline* out, int cap);
void  orc_lbd_compute(const uint8_t* img, int w, int h, int pitch, const orc_keyline* kls, int n, uint8_t* ldesc);
/* cv::clipLine on integer points (returns 0 when nothing is left) and cv::LineIterator(...,8).count */
int   orc_clip_line(int w, int h, long long* x1, long long* y1, long long* x2, long long* y2);
int   orc_line_iterator_count(int w, int h, int ax, int ay, int bx, int by);
/* raw LSD segments of the last call, before the top-N cut: x1,y1,x2,y2 (f32) */
int   orc_line_raw_segments(const orc_line*, float* seg4, int cap);
void  orc_line_scaled_copy(const orc_line*, uint8_t* dst, int dpitch, int* w, int* h);
void  orc_line_stage_ms(const orc_line*, double* ms4);
/* debug trace of the last call: 10 doubles per region that reached region2rect */
int   orc_line_trace(const orc_line*, double* out, int cap_rows);
/* LSD on an already-scaled image (scale=1.0 path) */
int   orc_lsd_detect_scaled(const uint8_t* img, int w, int h, int pitch, float* seg4, int cap);
void  orc_lbd_prep(const uint8_t* img, int w, int h, int pitch, int16_t* dx, int16_t* dy);

#ifdef __cplusplus
}
#endif
#endif
