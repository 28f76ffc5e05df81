// bow_port.cpp -- TEST INFRASTRUCTURE ONLY (oracle "port"): plain restatement of the reference's bag-of-words path.
//   DBoW2::TemplatedVocabulary::transform x2     /root/reference/Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1197, 1217-1259
//   DBoW2::FORB::distance                         /root/reference/Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
//   DBoW2::BowVector::addWeight / addIfNotExist / normalize   .../BowVector.cpp:29-87
//   DBoW2::FeatureVector::addFeature              .../FeatureVector.cpp:29-43
//   ORBmatcher::SearchByBoW(KeyFrame*, Frame&, ...)      /root/reference/src/ORBmatcher.cc:230-382
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, ...)   /root/reference/src/ORBmatcher.cc:656-799
// Pinned against oracle/_ref (the reference's own DBoW2 sources and matcher bodies) by tests/test_oracle_bow.py and the golden
// vectors in tests/golden/ref_bow.npz.  Same argument layout as the product's C ABI (include/orbx_b200.h).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <map>
#include <vector>

namespace bowport {

enum { TF_IDF = 0, TF = 1, IDF = 2, BINARY = 3 };                                     // BowVector.h:25-31
enum { L1_NORM = 0, L2_NORM = 1, CHI_SQUARE = 2, KL = 3, BHATTACHARYYA = 4, DOT_PRODUCT = 5 };   // BowVector.h:41-49

struct Voc {
    int k, L, weighting, scoring, n;                          // n nodes + root (id 0)
    std::vector<int> parent, word; std::vector<std::vector<int> > children;
    std::vector<unsigned char> desc; std::vector<double> weight;
};

static inline int hamming(const unsigned char* a, const unsigned char* b) {
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

// descend from the root, at every level to the FIRST child with the smallest distance (strict <); remember the node on level L - levelsup
static void transform_one(const Voc& v, const unsigned char* f, int levelsup, int& word, double& weight, int& nid) {
    const int nid_level = v.L - levelsup;
    nid = 0;
    int cur = 0, level = 0;
    do {
        ++level;
        const std::vector<int>& ch = v.children[cur];
        int best = ch[0], bd = hamming(f, &v.desc[(size_t)ch[0] * 32]);
        for (size_t c = 1; c < ch.size(); ++c) { const int d = hamming(f, &v.desc[(size_t)ch[c] * 32]); if (d < bd) { bd = d; best = ch[c]; } }
        cur = best;
        if (level == nid_level) nid = cur;
    } while (!v.children[cur].empty());
    word = v.word[cur]; weight = v.weight[cur];
}

}  // namespace bowport

using namespace bowport;

extern "C" {

void* port_voc_create(int k, int L, int weighting, int scoring, int n_nodes, const int* parent, const unsigned char* is_leaf, const unsigned char* desc, const double* weight) {
    Voc* v = new Voc();
    v->k = k; v->L = L; v->weighting = weighting; v->scoring = scoring; v->n = n_nodes + 1;
    v->parent.assign(v->n, 0); v->word.assign(v->n, -1); v->children.resize(v->n); v->desc.assign((size_t)v->n * 32, 0); v->weight.assign(v->n, 0.0);
    int nw = 0;
    for (int i = 0; i < n_nodes; ++i) {
        const int id = i + 1;
        v->parent[id] = parent[i]; v->children[parent[i]].push_back(id);
        std::memcpy(&v->desc[(size_t)id * 32], desc + (size_t)i * 32, 32); v->weight[id] = weight[i];
        if (is_leaf[i]) v->word[id] = nw++;
    }
    return v;
}
void port_voc_destroy(void* p) { delete (Voc*)p; }

int port_voc_transform(void* p, const unsigned char* desc, int n, int levelsup, int* word_of, int* node_of, int* bow_ids, double* bow_vals, int* n_bow,
                       int* fv_nodes, int* fv_offsets, int* fv_idx, int* n_fv) {
    const Voc& v = *(const Voc*)p;
    std::map<int, double> bow; std::map<int, std::vector<int> > fv;
    *n_bow = 0; *n_fv = 0; fv_offsets[0] = 0;
    for (int i = 0; i < n; ++i) {
        int w, nid; double wt;
        transform_one(v, desc + (size_t)i * 32, levelsup, w, wt, nid);
        word_of[i] = w; node_of[i] = nid;
        if (!(wt > 0)) continue;                                                       // stopped word
        if (v.weighting == TF_IDF || v.weighting == TF) bow[w] += wt;                  // addWeight: running double sum in feature order
        else if (!bow.count(w)) bow[w] = wt;                                           // addIfNotExist
        fv[nid].push_back(i);
    }
    const bool must = v.scoring != DOT_PRODUCT;
    if ((v.weighting == TF_IDF || v.weighting == TF) && !bow.empty() && !must) {
        const double nd = (double)bow.size();
        for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) it->second /= nd;
    }
    if (must) {
        double norm = 0.0;
        if (v.scoring == L2_NORM) { for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) norm += it->second * it->second; norm = std::sqrt(norm); }
        else for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) norm += std::fabs(it->second);
        if (norm > 0.0) for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it) it->second /= norm;
    }
    int o = 0;
    for (std::map<int, double>::iterator it = bow.begin(); it != bow.end(); ++it, ++o) { bow_ids[o] = it->first; bow_vals[o] = it->second; }
    *n_bow = o;
    int q = 0, e = 0;
    for (std::map<int, std::vector<int> >::iterator it = fv.begin(); it != fv.end(); ++it, ++q) {
        fv_nodes[q] = it->first; fv_offsets[q] = e;
        for (size_t j = 0; j < it->second.size(); ++j) fv_idx[e++] = it->second[j];
    }
    fv_offsets[q] = e; *n_fv = q;
    return 0;
}

// both SearchByBoW variants: merge the two node lists, inside a common node every valid side-1 feature (in list order) takes the best
// still-unmatched valid side-2 feature if best <= / < TH_LOW and best < ratio * second; then the rotation-histogram filter.
// keys: 28-byte cv::KeyPoint records (only .angle, offset 12, is read)
int port_search_by_bow(float nnratio, int checkOri, int kf_kf, int n1, const unsigned char* keys1, const unsigned char* desc1, const unsigned char* valid1,
                       int n_fv1, const int* fv1_nodes, const int* fv1_offsets, const int* fv1_idx,
                       int n2, const unsigned char* keys2, const unsigned char* desc2, const unsigned char* valid2,
                       int n_fv2, const int* fv2_nodes, const int* fv2_offsets, const int* fv2_idx, int* match12, int* match21) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    for (int j = 0; j < n2; ++j) match21[j] = -1;
    std::vector<int> hist[30];
    int nmatches = 0, a = 0, b = 0;
    while (a < n_fv1 && b < n_fv2) {
        if (fv1_nodes[a] < fv2_nodes[b]) { ++a; continue; }                            // lower_bound on a sorted list == skip ahead
        if (fv2_nodes[b] < fv1_nodes[a]) { ++b; continue; }
        for (int e1 = fv1_offsets[a]; e1 < fv1_offsets[a + 1]; ++e1) {
            const int i1 = fv1_idx[e1];
            if (!valid1[i1]) continue;
            int best1 = 256, best2 = 256, bidx = -1;
            for (int e2 = fv2_offsets[b]; e2 < fv2_offsets[b + 1]; ++e2) {
                const int i2 = fv2_idx[e2];
                if (match21[i2] >= 0) continue;                                        // vpMapPointMatches[realIdxF] / vbMatched2[idx2]
                if (kf_kf && !valid2[i2]) continue;
                const int d = hamming(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
                if (d < best1) { best2 = best1; best1 = d; bidx = i2; } else if (d < best2) best2 = d;
            }
            const bool close = kf_kf ? (best1 < TH_LOW) : (best1 <= TH_LOW);           // :741 vs :313
            if (close && (float)best1 < nnratio * (float)best2) {
                match12[i1] = bidx; match21[bidx] = i1;
                if (checkOri) {
                    float a1, a2; std::memcpy(&a1, keys1 + (size_t)i1 * 28 + 12, 4); std::memcpy(&a2, keys2 + (size_t)bidx * 28 + 12, 4);
                    float rot = a1 - a2;
                    if (rot < 0.0) rot += 360.0f;
                    int bin = (int)std::round(rot * ((float)HISTO_LENGTH / 360.0f));
                    if (bin == HISTO_LENGTH) bin = 0;
                    hist[bin].push_back(i1);
                }
                ++nmatches;
            }
        }
        ++a; ++b;
    }
    if (checkOri) {
        // ComputeThreeMaxima (ORBmatcher.cc:1866-1908)
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)hist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < hist[i].size(); ++j) { const int i1 = hist[i][j]; match21[match12[i1]] = -1; match12[i1] = -1; --nmatches; }
        }
    }
    return nmatches;
}


// ORBmatcher::SearchForTriangulation (ORBmatcher.cc:810-1010) with CheckDistEpipolarLine (:188-215).  free1 / free2: the feature holds no map point.
int port_search_for_triangulation(int checkOri, int n1, const unsigned char* keys1, const unsigned char* desc1, const unsigned char* free1, const float* ur1,
                                  int n_fv1, const int* fv1_nodes, const int* fv1_offsets, const int* fv1_idx,
                                  int n2, const unsigned char* keys2, const unsigned char* desc2, const unsigned char* free2, const float* ur2,
                                  int n_fv2, const int* fv2_nodes, const int* fv2_offsets, const int* fv2_idx,
                                  const float* F, float ex, float ey, const float* scale2, const float* sigma2_2, int only_stereo, int* match12) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    struct KP { float x, y, size, angle, response; int octave, class_id; };
    const KP* k1 = (const KP*)keys1; const KP* k2 = (const KP*)keys2;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    std::vector<unsigned char> matched2(n2, 0);
    std::vector<int> hist[30];
    int nmatches = 0, a = 0, b = 0;
    while (a < n_fv1 && b < n_fv2) {
        if (fv1_nodes[a] < fv2_nodes[b]) { ++a; continue; }
        if (fv2_nodes[b] < fv1_nodes[a]) { ++b; continue; }
        for (int e1 = fv1_offsets[a]; e1 < fv1_offsets[a + 1]; ++e1) {
            const int i1 = fv1_idx[e1];
            if (!free1[i1]) continue;
            const bool stereo1 = ur1[i1] >= 0;
            if (only_stereo && !stereo1) continue;
            // epipolar line of kp1 in image 2
            volatile float la = k1[i1].x * F[0] + k1[i1].y * F[3] + F[6], lb = k1[i1].x * F[1] + k1[i1].y * F[4] + F[7], lc = k1[i1].x * F[2] + k1[i1].y * F[5] + F[8];
            int bestDist = TH_LOW, bestIdx2 = -1;
            for (int e2 = fv2_offsets[b]; e2 < fv2_offsets[b + 1]; ++e2) {
                const int i2 = fv2_idx[e2];
                if (matched2[i2] || !free2[i2]) continue;
                const bool stereo2 = ur2[i2] >= 0;
                if (only_stereo && !stereo2) continue;
                const int dist = hamming(desc1 + (size_t)i1 * 32, desc2 + (size_t)i2 * 32);
                if (dist > TH_LOW || dist > bestDist) continue;
                if (!stereo1 && !stereo2) {
                    volatile float dx = ex - k2[i2].x, dy = ey - k2[i2].y;
                    volatile float d2 = dx * dx + dy * dy, lim = 100 * scale2[k2[i2].octave];
                    if (d2 < lim) continue;
                }
                volatile float num = la * k2[i2].x + lb * k2[i2].y + lc, den = la * la + lb * lb;
                if (den == 0) continue;
                volatile float dsqr = num * num / den;
                if (!((double)dsqr < 3.84 * (double)sigma2_2[k2[i2].octave])) continue;
                bestIdx2 = i2; bestDist = dist;
            }
            if (bestIdx2 < 0) continue;
            match12[i1] = bestIdx2; matched2[bestIdx2] = 1; ++nmatches;
            if (checkOri) {
                float rot = k1[i1].angle - k2[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = (int)std::round(rot * ((float)HISTO_LENGTH / 360.0f));
                if (bin == HISTO_LENGTH) bin = 0;
                hist[bin].push_back(i1);
            }
        }
        ++a; ++b;
    }
    if (checkOri) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)hist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < hist[i].size(); ++j) { match12[hist[i][j]] = -1; --nmatches; }
        }
    }
    return nmatches;
}


// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-439) for many points: per point the descriptor with the least median distance to the others
void port_distinctive_descriptors(int n_points, const int* offsets, const unsigned char* desc, int* best_idx) {
    for (int p = 0; p < n_points; ++p) {
        const int o = offsets[p], N = offsets[p + 1] - o;
        best_idx[p] = -1;
        if (N <= 0) continue;
        int bestMedian = INT_MAX, best = 0;
        std::vector<int> row(N);
        for (int i = 0; i < N; ++i) {
            for (int j = 0; j < N; ++j) row[j] = i == j ? 0 : hamming(desc + (size_t)(o + i) * 32, desc + (size_t)(o + j) * 32);
            std::sort(row.begin(), row.end());
            const int median = row[(size_t)(0.5 * (N - 1))];
            if (median < bestMedian) { bestMedian = median; best = i; }
        }
        best_idx[p] = best;
    }
}

}  // extern "C"
