// host/bow_b200.cc — GPU-backed Frame::ComputeBoW / KeyFrame::ComputeBoW (reference src/Frame.cc:474-481,
// src/KeyFrame.cc:71-80), i.e. ORBVocabulary::transform(features, BowVector&, FeatureVector&, 4)
// (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1127-1259).  SURVEY.md 8(f) row 1.
//
// Compiled inside the reference tree against ITS headers; in src/Frame.cc and src/KeyFrame.cc the two ComputeBoW bodies are
// deleted (or #if 0'ed) and this file is added to the source list.  The tree nodes of
// DBoW2::TemplatedVocabulary are protected, so the device copy is loaded from the same text file: add ONE line after
// System.cc:65 (`mpVocabulary->loadFromTextFile(strVocFile)`):
//     StructureSLAM::SslplSetVocabularyFile(strVocFile);
// There is no CPU fallback: ComputeBoW throws until that call has been made, and for vocabularies that are not L1 / TF-IDF
// (the ORBvoc.txt defaults "10 6 0 0"), whose BowVector arithmetic this adapter does not restate.
//
// The device returns (word, node, weight) per feature; the BowVector / FeatureVector are then filled with DBoW2's OWN
// addWeight / addFeature / normalize in feature order, exactly as transform() does (:1150-1170, :1194), so the two maps
// (and the float sums inside them) are identical to the CPU result.
#include "Frame.h"
#include "KeyFrame.h"
#include "sslpl.h"
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace StructureSLAM
{
namespace {
struct VocabCtx {
    sslpl_vocab* v; int scoring, weighting; std::mutex mu;
    VocabCtx(): v(NULL), scoring(0), weighting(0) {}
    ~VocabCtx() { sslpl_vocab_destroy(v); }
};
VocabCtx& Voc() { static VocabCtx c; return c; }

struct BowCtx {
    sslpl_matcher* h;
    BowCtx(): h(NULL) {
        sslpl_matcher_params p; p.max_features = 8192; p.max_lines = 16; p.max_nodes = 16; p.max_batch = 1; p.device = sslpl_default_device();
        if(sslpl_matcher_create(&p, &h) != SSLPL_OK) throw std::runtime_error(std::string("sslpl_matcher_create: ") + sslpl_last_error());
    }
    ~BowCtx() { sslpl_matcher_destroy(h); }
};
sslpl_matcher* Ctx() { static thread_local BowCtx c; return c.h; }     // Tracking and LocalMapping threads both call ComputeBoW

void TransformOnDevice(const cv::Mat &descriptors, DBoW2::BowVector &bow, DBoW2::FeatureVector &fv, int levelsup)
{
    VocabCtx &V = Voc();
    if(!V.v) throw std::runtime_error("libsslpl_b200: call StructureSLAM::SslplSetVocabularyFile(strVocFile) after loading the vocabulary (no CPU fallback)");
    if(V.scoring != 0 /* L1_NORM */ || V.weighting != 0 /* TF_IDF */) throw std::runtime_error("libsslpl_b200: only L1 / TF-IDF vocabularies (ORBvoc.txt) are supported");
    const int n = descriptors.rows;
    bow.clear(); fv.clear();
    if(n == 0) return;
    cv::Mat d = descriptors.isContinuous() ? descriptors : descriptors.clone();
    std::vector<int32_t> word(n), node(n); std::vector<double> w(n);
    if(sslpl_bow_transform(Ctx(), V.v, d.ptr<uchar>(), n, levelsup, &word[0], &node[0], &w[0]) != SSLPL_OK)
        throw std::runtime_error(std::string("sslpl_bow_transform: ") + sslpl_last_error());
    for(int i=0; i<n; i++)
        if(w[i] > 0) {                                                   // not stopped (TemplatedVocabulary.h:1162)
            bow.addWeight((DBoW2::WordId)word[i], w[i]);
            fv.addFeature((DBoW2::NodeId)node[i], (unsigned int)i);
        }
    bow.normalize(DBoW2::L1);                                            // L1 scoring always normalises (:1194)
}
}

void SslplSetVocabularyFile(const std::string &strVocFile)
{
    VocabCtx &V = Voc();
    std::lock_guard<std::mutex> lock(V.mu);
    sslpl_vocab_destroy(V.v); V.v = NULL;
    if(sslpl_vocab_load_text(sslpl_default_device(), strVocFile.c_str(), &V.v, &V.scoring, &V.weighting) != SSLPL_OK)
        throw std::runtime_error(std::string("sslpl_vocab_load_text: ") + sslpl_last_error());
}

void Frame::ComputeBoW()                                                 // Frame.cc:474-481
{
    if(mBowVec.empty())
    {
        TransformOnDevice(mDescriptors, mBowVec, mFeatVec, 4);
    }
}

void KeyFrame::ComputeBoW()                                              // KeyFrame.cc:71-80
{
    if(mBowVec.empty() || mFeatVec.empty())
    {
        TransformOnDevice(mDescriptors, mBowVec, mFeatVec, 4);
    }
}

} // namespace StructureSLAM
