// BoW_b200.cc -- B200-native bodies of the bag-of-words methods.  Member definitions of the reference's OWN
// classes, so Tracking / LocalMapping / LoopClosing keep calling them as they do today (src/Tracking.cc:1565-1567, 1740-1752,
// 2595-2633; src/LocalMapping.cc:212; src/LoopClosing.cc); a maintainer compiles this file, removes (or #ifdef's out) the four
// bodies it replaces and registers the vocabulary file once, right after it is loaded (src/System.cc:84):
//     ORB_SLAM2::RegisterDeviceVocabulary(mpVocabulary, strVocFile);
//
//   Frame::ComputeBoW()                                                           src/Frame.cc:1033-1049
//   KeyFrame::ComputeBoW()                                                        src/KeyFrame.cc:79-89
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)                src/ORBmatcher.cc:230-382
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)             src/ORBmatcher.cc:656-799
//   ORBmatcher::SearchForTriangulation(KeyFrame*, KeyFrame*, cv::Mat F12, vector<pair<size_t,size_t>>&, bool)   src/ORBmatcher.cc:810-1010  (8f rank 3)
//
// The tree descent (60 Hamming distances per descriptor for ORBvoc), the BowVector / FeatureVector assembly with DBoW2's exact
// double arithmetic, the per-node best / second-best search, the ratio test and the rotation histogram run on the GPU; the bodies
// only move the results into the std::map containers the rest of the reference reads.
#include "ORBmatcher.h"
#include "../../include/orbx_b200.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ORB_SLAM2
{
namespace
{
void check(int rc, const char* what) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string(what) + ": " + orbx_last_error());
}
int device_ordinal() { const char* e = std::getenv("ORBX_DEVICE"); return e ? std::atoi(e) : 0; }

std::mutex g_voc_mutex;
std::map<const ORBVocabulary*, orbx_vocabulary*> g_vocs;      // host vocabulary object -> its copy in HBM

orbx_vocabulary* device_vocabulary(const ORBVocabulary* voc) {
    std::lock_guard<std::mutex> lock(g_voc_mutex);
    std::map<const ORBVocabulary*, orbx_vocabulary*>::iterator it = g_vocs.find(voc);
    if (it == g_vocs.end()) throw std::runtime_error("vocabulary not registered: call ORB_SLAM2::RegisterDeviceVocabulary(voc, file) after loadFromTextFile");
    return it->second;
}

// orbx_vocabulary handles own scratch buffers and a stream, so calls on one handle are serialised (Tracking and LocalMapping both compute BoW)
std::mutex g_transform_mutex;

void compute_bow(const ORBVocabulary* voc, const cv::Mat& descriptors, DBoW2::BowVector& bow, DBoW2::FeatureVector& fv) {
    const int n = descriptors.rows;
    std::vector<unsigned char> tmp;
    const unsigned char* d = descriptors.ptr();
    if (n && !descriptors.isContinuous()) { tmp.resize((size_t)n * 32); for (int i = 0; i < n; ++i) std::memcpy(&tmp[(size_t)i * 32], descriptors.ptr(i), 32); d = tmp.data(); }
    std::vector<int> ids(n ? n : 1), nodes(n ? n : 1), offs(n + 1), idx(n ? n : 1);
    std::vector<double> vals(n ? n : 1);
    int nb = 0, nf = 0;
    {
        std::lock_guard<std::mutex> lock(g_transform_mutex);
        check(orbx_vocabulary_transform(device_vocabulary(voc), d, n, 4, NULL, NULL, ids.data(), vals.data(), &nb, nodes.data(), offs.data(), idx.data(), &nf), "orbx_vocabulary_transform");
    }
    bow.clear(); fv.clear();
    for (int i = 0; i < nb; ++i) bow.insert(bow.end(), std::make_pair((DBoW2::WordId)ids[i], (DBoW2::WordValue)vals[i]));       // already in map order
    for (int q = 0; q < nf; ++q) {
        std::vector<unsigned int>& v = fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)nodes[q], std::vector<unsigned int>()))->second;
        v.assign(idx.begin() + offs[q], idx.begin() + offs[q + 1]);
    }
}

// flattened FeatureVector
struct FlatFv {
    std::vector<int> nodes, offs, idx;
    explicit FlatFv(const DBoW2::FeatureVector& fv) {
        offs.push_back(0);
        for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it) {
            nodes.push_back((int)it->first);
            idx.insert(idx.end(), it->second.begin(), it->second.end());
            offs.push_back((int)idx.size());
        }
    }
};
struct Side {
    orbx_bow_side s; FlatFv fv; std::vector<unsigned char> valid, desc;
    Side(const std::vector<cv::KeyPoint>& keys, const cv::Mat& descriptors, const DBoW2::FeatureVector& f, const std::vector<MapPoint*>* mps) : fv(f) {
        s.n = (int)keys.size(); s.keys = reinterpret_cast<const orbx_keypoint*>(keys.data());
        if (descriptors.isContinuous()) s.descriptors = descriptors.ptr();
        else { desc.resize((size_t)s.n * 32); for (int i = 0; i < s.n; ++i) std::memcpy(&desc[(size_t)i * 32], descriptors.ptr(i), 32); s.descriptors = desc.data(); }
        s.valid = NULL;
        if (mps) { valid.assign(s.n, 0); for (int i = 0; i < s.n; ++i) { MapPoint* p = (*mps)[i]; if (p && !p->isBad()) valid[i] = 1; } s.valid = valid.data(); }
        s.n_fv = (int)fv.nodes.size(); s.fv_nodes = fv.nodes.data(); s.fv_offsets = fv.offs.data(); s.fv_indices = fv.idx.data();
    }
};

struct BowMatcherCache {                                        // as in ORBmatcher_b200.cc: matcher objects are short-lived locals
    std::map<std::pair<float, bool>, orbx_matcher*> m;
    ~BowMatcherCache() { for (auto& kv : m) orbx_matcher_destroy(kv.second); }
    orbx_matcher* get(float nnratio, bool checkOri) {
        auto key = std::make_pair(nnratio, checkOri);
        auto it = m.find(key);
        if (it != m.end()) return it->second;
        orbx_matcher* h = nullptr;
        check(orbx_matcher_create(nnratio, checkOri ? 1 : 0, device_ordinal(), &h), "orbx_matcher_create");
        m[key] = h;
        return h;
    }
};
thread_local BowMatcherCache t_bow_matchers;
}  // namespace

// call once after ORBVocabulary::loadFromTextFile(file) succeeded (src/System.cc:84-91)
void RegisterDeviceVocabulary(const ORBVocabulary* voc, const std::string& file)
{
    orbx_vocabulary* v = nullptr;
    check(orbx_vocabulary_load_text(device_ordinal(), file.c_str(), &v), "orbx_vocabulary_load_text");
    std::lock_guard<std::mutex> lock(g_voc_mutex);
    std::map<const ORBVocabulary*, orbx_vocabulary*>::iterator it = g_vocs.find(voc);
    if (it != g_vocs.end()) orbx_vocabulary_destroy(it->second);
    g_vocs[voc] = v;
}

void Frame::ComputeBoW()
{
    if (mBowVec.empty()) compute_bow(mpORBvocabulary, mDescriptors, mBowVec, mFeatVec);                                  // src/Frame.cc:1037
}

void KeyFrame::ComputeBoW()
{
    if (mBowVec.empty() || mFeatVec.empty()) compute_bow(mpORBvocabulary, mDescriptors, mBowVec, mFeatVec);              // src/KeyFrame.cc:81
}

int ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame &F, std::vector<MapPoint*> &vpMapPointMatches)
{
    const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
    vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
    Side a(pKF->mvKeysUn, pKF->mDescriptors, pKF->mFeatVec, &vpMapPointsKF), b(F.mvKeys, F.mDescriptors, F.mFeatVec, NULL);
    std::vector<int> m12(a.s.n ? a.s.n : 1), m21(b.s.n ? b.s.n : 1);
    int nmatches = 0;
    check(orbx_search_by_bow(t_bow_matchers.get(mfNNratio, mbCheckOrientation), 0, &a.s, &b.s, m12.data(), m21.data(), &nmatches), "orbx_search_by_bow");
    for (int j = 0; j < F.N; ++j) if (m21[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[m21[j]];                         // :318
    return nmatches;
}

int ORBmatcher::SearchByBoW(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint*> &vpMatches12)
{
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
    const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
    vpMatches12 = std::vector<MapPoint*>(vpMapPoints1.size(), static_cast<MapPoint*>(NULL));
    Side a(pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mFeatVec, &vpMapPoints1), b(pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mFeatVec, &vpMapPoints2);
    std::vector<int> m12(a.s.n ? a.s.n : 1), m21(b.s.n ? b.s.n : 1);
    int nmatches = 0;
    check(orbx_search_by_bow(t_bow_matchers.get(mfNNratio, mbCheckOrientation), 1, &a.s, &b.s, m12.data(), m21.data(), &nmatches), "orbx_search_by_bow");
    for (int i = 0; i < a.s.n; ++i) if (m12[i] >= 0) vpMatches12[i] = vpMapPoints2[m12[i]];                               // :746
    return nmatches;
}

int ORBmatcher::SearchForTriangulation(KeyFrame *pKF1, KeyFrame *pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> > &vMatchedPairs, const bool bOnlyStereo)
{
    // epipole of camera 1 in image 2 (:816-825)
    cv::Mat Cw = pKF1->GetCameraCenter();
    cv::Mat R2w = pKF2->GetRotation();
    cv::Mat t2w = pKF2->GetTranslation();
    cv::Mat C2 = R2w * Cw + t2w;
    const float invz = 1.0f / C2.at<float>(2);
    const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
    const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
    Side a(pKF1->mvKeysUn, pKF1->mDescriptors, pKF1->mFeatVec, NULL), b(pKF2->mvKeysUn, pKF2->mDescriptors, pKF2->mFeatVec, NULL);
    a.valid.assign(a.s.n, 0); b.valid.assign(b.s.n, 0);                                  // only features WITHOUT a map point take part (:843-845, :862)
    for (int i = 0; i < a.s.n; ++i) a.valid[i] = pKF1->GetMapPoint(i) ? 0 : 1;
    for (int j = 0; j < b.s.n; ++j) b.valid[j] = pKF2->GetMapPoint(j) ? 0 : 1;
    a.s.valid = a.valid.data(); b.s.valid = b.valid.data();
    float F[9];
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F[3 * r + c] = F12.at<float>(r, c);
    std::vector<int> m12(a.s.n ? a.s.n : 1, -1);
    int nmatches = 0;
    check(orbx_search_for_triangulation(t_bow_matchers.get(mfNNratio, mbCheckOrientation), &a.s, &b.s, pKF1->mvuRight.data(), pKF2->mvuRight.data(), F, ex, ey,
                                        (int)pKF2->mvScaleFactors.size(), pKF2->mvScaleFactors.data(), pKF2->mvLevelSigma2.data(), bOnlyStereo ? 1 : 0, m12.data(), &nmatches),
          "orbx_search_for_triangulation");
    vMatchedPairs.clear();
    vMatchedPairs.reserve(nmatches);
    for (int i = 0; i < a.s.n; ++i) if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m12[i]));    // :1000-1006
    return nmatches;
}

} //namespace ORB_SLAM2
