// ref_bow_capi.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Runs the reference's OWN bag-of-words code (DBoW2, vendored in the reference tree under Thirdparty/DBoW2): FORB.cpp,
// BowVector.cpp and FeatureVector.cpp are compiled verbatim from where they lie; the two transform() members of
// TemplatedVocabulary (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1197, 1217-1259) are extracted verbatim at build
// time (oracle/ref/gen_match_bodies.py -> oracle/_ref/gen/ref_bow_bodies.inc) into the minimal class below, which carries
// exactly the members those bodies read.  The rest of the template (k-means training, YAML / text IO through
// cv::FileStorage) is not on the path and not built; the vocabulary is loaded from flat arrays instead.
#include <cmath>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>
#include <stdint.h>
#include <opencv2/core/core.hpp>
using namespace cv;                                          // the DBoW2 sources use CV_8U / CV_32F unqualified
using namespace std;                                         // TemplatedVocabulary.h has `using namespace std` at file scope

#include "Thirdparty/DBoW2/DBoW2/FORB.cpp"                   // reference sources, compiled where they lie (-I$(REF))
#include "Thirdparty/DBoW2/DBoW2/BowVector.cpp"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.cpp"
#include "Thirdparty/DBoW2/DBoW2/ScoringObject.h"

namespace DBoW2 {

template<class TDescriptor, class F>
class TemplatedVocabulary {                                  // TemplatedVocabulary.h:38-330, the subset transform() touches
public:
    struct Node {                                            // :297-329
        NodeId id; WordValue weight; vector<NodeId> children; NodeId parent; TDescriptor descriptor; WordId word_id;
        Node() : id(0), weight(0), parent(0), word_id(0) {}
        inline bool isLeaf() const { return children.empty(); }
    };
    virtual ~TemplatedVocabulary() {}
    virtual inline bool empty() const { return m_words.empty(); }                     // :118
    virtual void transform(const std::vector<TDescriptor>& features, BowVector &v, FeatureVector &fv, int levelsup) const;
    virtual void transform(const TDescriptor &feature, WordId &word_id, WordValue &weight, NodeId *nid = NULL, int levelsup = 0) const;
    int m_k, m_L; WeightingType m_weighting; ScoringType m_scoring; GeneralScoring* m_scoring_object;
    std::vector<Node> m_nodes; std::vector<Node*> m_words;
};

#include "ref_bow_bodies.inc"                                // GENERATED: verbatim transform() bodies

// mustNormalize() of the scoring classes (ScoringObject.h:74-92); score() itself is not on the path
struct NormOnlyScoring : public GeneralScoring {
    bool must; LNorm norm;
    NormOnlyScoring(ScoringType s) {
        must = (s != DOT_PRODUCT); norm = (s == L2_NORM) ? L2 : L1;                  // L1, L2, CHI_SQUARE, KL, BHATTACHARYYA normalise; DOT_PRODUCT does not
    }
    virtual double score(const BowVector&, const BowVector&) const { return 0; }
    virtual bool mustNormalize(LNorm& n) const { n = norm; return must; }
};

}  // namespace DBoW2

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // include/ORBVocabulary.h:40-41

extern "C" {

// Same node table as the text format ORBVocabulary::loadFromTextFile reads (TemplatedVocabulary.h:1336-1424): node i + 1 has
// (parent, is_leaf, descriptor, weight); ids in file order, children in push_back order, word ids in order of appearance.
void* ref_voc_create(int k, int L, int weighting, int scoring, int n_nodes, const int* parent, const unsigned char* is_leaf, const unsigned char* desc, const double* weight) {
    ORBVocabulary* v = new ORBVocabulary();
    v->m_k = k; v->m_L = L; v->m_weighting = (DBoW2::WeightingType)weighting; v->m_scoring = (DBoW2::ScoringType)scoring;
    v->m_scoring_object = new DBoW2::NormOnlyScoring(v->m_scoring);
    v->m_nodes.resize(n_nodes + 1);
    v->m_nodes[0].id = 0;
    int nwords = 0;
    for (int i = 0; i < n_nodes; ++i) nwords += is_leaf[i] ? 1 : 0;
    v->m_words.reserve(nwords);
    for (int i = 0; i < n_nodes; ++i) {
        const int nid = i + 1;
        v->m_nodes[nid].id = nid; v->m_nodes[nid].parent = parent[i];
        v->m_nodes[parent[i]].children.push_back(nid);
        v->m_nodes[nid].descriptor = cv::Mat(1, 32, CV_8U, (void*)(desc + (size_t)i * 32)).clone();
        v->m_nodes[nid].weight = weight[i];
        if (is_leaf[i]) { v->m_nodes[nid].word_id = (int)v->m_words.size(); v->m_words.push_back(&v->m_nodes[nid]); }
    }
    return v;
}
void ref_voc_destroy(void* p) { ORBVocabulary* v = (ORBVocabulary*)p; if (v) { delete v->m_scoring_object; delete v; } }

// Frame::ComputeBoW / KeyFrame::ComputeBoW (src/Frame.cc:1033-1049): transform(descriptors, mBowVec, mFeatVec, levelsup)
int ref_voc_transform(void* p, const unsigned char* desc, int n, int levelsup, int* word_of, int* node_of, int* bow_ids, double* bow_vals, int* n_bow,
                      int* fv_nodes, int* fv_offsets, int* fv_idx, int* n_fv) {
    const ORBVocabulary* v = (const ORBVocabulary*)p;
    std::vector<cv::Mat> d(n);
    for (int i = 0; i < n; ++i) d[i] = cv::Mat(1, 32, CV_8U, (void*)(desc + (size_t)i * 32));
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->transform(d, bv, fv, levelsup);
    for (int i = 0; i < n; ++i) { DBoW2::WordId w; DBoW2::WordValue wt; DBoW2::NodeId nid = 0; v->transform(d[i], w, wt, &nid, levelsup); word_of[i] = (int)w; node_of[i] = (int)nid; }
    int o = 0;
    for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++o) { bow_ids[o] = (int)it->first; bow_vals[o] = it->second; }
    *n_bow = o;
    int q = 0, e = 0;
    for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++q) {
        fv_nodes[q] = (int)it->first; fv_offsets[q] = e;
        for (size_t j = 0; j < it->second.size(); ++j) fv_idx[e++] = (int)it->second[j];
    }
    fv_offsets[q] = e; *n_fv = q;
    return 0;
}

}  // extern "C"
