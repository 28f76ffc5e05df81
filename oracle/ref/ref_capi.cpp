// ref_capi.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Thin C API around the reference's OWN ORB_SLAM2::ORBextractor, compiled unmodified from
// /root/reference/src/ORBextractor.cc against the OpenCV-free shim (oracle/shim -> oracle/cvlite).
// Used by tests/ as the parity oracle and by bench.py --impl reference / cpu_baseline as the
// CPU reference arm.  Never linked into the product library.
#include "ORBextractor.h"   // the reference's header: /root/reference/include/ORBextractor.h
#include <cstring>

extern "C" void ref_arena_begin();
extern "C" void ref_arena_end();

namespace {
// exposes the protected members we need for stage-level parity checks
struct RefExtractor : public ORB_SLAM2::ORBextractor {
    using ORB_SLAM2::ORBextractor::ORBextractor;
    using ORB_SLAM2::ORBextractor::DistributeOctTree;
    using ORB_SLAM2::ORBextractor::mnFeaturesPerLevel;
    using ORB_SLAM2::ORBextractor::umax;
    using ORB_SLAM2::ORBextractor::nlevels;
    std::vector<std::vector<cv::KeyPoint> > staged;   // per-level keypoints between detect and describe
};
struct ArenaScope { ArenaScope() { ref_arena_begin(); } ~ArenaScope() { ref_arena_end(); } };
}

extern "C" {

void* ref_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new RefExtractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_extractor_destroy(void* h) { delete (RefExtractor*)h; }

int ref_extractor_info(void* h, int* nlevels, float* scale_factors, int* features_per_level, int* umax16) {
    RefExtractor* e = (RefExtractor*)h;
    *nlevels = e->GetLevels();
    std::vector<float> sf = e->GetScaleFactors();
    for (int i = 0; i < e->GetLevels(); ++i) { scale_factors[i] = sf[i]; features_per_level[i] = e->mnFeaturesPerLevel[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
    return 0;
}

// ORBextractor::operator()(image, mask, keypoints, descriptors)   ORBextractor.cc:1544
int ref_extract(void* h, const unsigned char* img, int rows, int cols, int step,
                cv::KeyPoint* kp_out, unsigned char* desc_out, int cap) {
    RefExtractor* e = (RefExtractor*)h;
    ArenaScope scope;
    int n;
    {
        cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step), desc;
        std::vector<cv::KeyPoint> kps;
        (*e)(image, cv::Mat(), kps, desc);
        n = (int)kps.size();
        if (n > cap) return -n;
        if (n) { std::memcpy(kp_out, kps.data(), sizeof(cv::KeyPoint) * n); std::memcpy(desc_out, desc.data, (size_t)n * 32); }
    }
    return n;
}

// ORBextractor::operator()(image, mask, vector<vector<KeyPoint>>&)   ORBextractor.cc:1672
// keypoints are returned level-major in LEVEL coordinates; level_counts[nlevels].
int ref_detect(void* h, const unsigned char* img, int rows, int cols, int step,
               cv::KeyPoint* kp_out, int* level_counts, int cap) {
    RefExtractor* e = (RefExtractor*)h;
    int n = 0;
    {
        ArenaScope scope;
        cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step);
        std::vector<std::vector<cv::KeyPoint> > all;
        (*e)(image, cv::Mat(), all);
        for (size_t l = 0; l < all.size(); ++l) { level_counts[l] = (int)all[l].size(); n += level_counts[l]; }
        if (n > cap) return -n;
        int o = 0;
        for (size_t l = 0; l < all.size(); ++l) { if (!all[l].empty()) std::memcpy(kp_out + o, all[l].data(), sizeof(cv::KeyPoint) * all[l].size()); o += (int)all[l].size(); }
    }
    return n;
}

// mvImagePyramid[level] (ROI only, tightly packed)
int ref_pyramid_level(void* h, int level, unsigned char* out, int* rows, int* cols) {
    RefExtractor* e = (RefExtractor*)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size() || e->mvImagePyramid[level].empty()) return -1;
    const cv::Mat& m = e->mvImagePyramid[level];
    *rows = m.rows; *cols = m.cols;
    if (out) for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
    return 0;
}
// padded parent buffer of a level ((cols+38) x (rows+38)), to check the REFLECT_101 border export
int ref_pyramid_level_padded(void* h, int level, unsigned char* out) {
    RefExtractor* e = (RefExtractor*)h;
    const cv::Mat& m = e->mvImagePyramid[level];
    const int B = 19;
    for (int y = -B; y < m.rows + B; ++y) std::memcpy(out + (size_t)(y + B) * (m.cols + 2 * B), m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - B, (size_t)m.cols + 2 * B);
    return 0;
}

// ORBextractor::DistributeOctTree   ORBextractor.cc:706
int ref_distribute_octtree(void* h, const cv::KeyPoint* cand, int ncand, int minX, int maxX, int minY, int maxY,
                           int N, int level, cv::KeyPoint* out, int cap) {
    RefExtractor* e = (RefExtractor*)h;
    ArenaScope scope;
    int n;
    {
        std::vector<cv::KeyPoint> v(cand, cand + ncand);
        std::vector<cv::KeyPoint> r = e->DistributeOctTree(v, minX, maxX, minY, maxY, N, level);
        n = (int)r.size();
        if (n > cap) return -n;
        if (n) std::memcpy(out, r.data(), sizeof(cv::KeyPoint) * n);
    }
    return n;
}

// ORBextractor::MovingKeyPoints   ORBextractor.cc:1688  (Amos dynamic-mask culling)
// kp_inout: level-major keypoints in level coordinates with level_counts[nlevels]; filtered in place.
// centers_id[ncenters] = centers[i].id ; label = CV_64F H x W ; mask = CV_8U H x W
// returns number of culled keypoints (written to culled_out in cull order).
int ref_moving_keypoints(void* h, const unsigned char* mask, const double* label, int rows, int cols,
                         const int* centers_id, int ncenters, const int* rm_vector, int nrm,
                         cv::KeyPoint* kp_inout, int* level_counts, cv::KeyPoint* culled_out) {
    RefExtractor* e = (RefExtractor*)h;
    ArenaScope scope;
    int nc;
    {
        cv::Mat imS(rows, cols, CV_8UC1, (void*)mask), imLS(rows, cols, CV_64F, (void*)label), gray;
        std::vector<ORB_SLAM2::center> centers(ncenters);
        for (int i = 0; i < ncenters; ++i) { std::memset(&centers[i], 0, sizeof(ORB_SLAM2::center)); centers[i].id = centers_id[i]; }
        std::vector<int> rm(rm_vector, rm_vector + nrm);
        std::vector<bool> flag;
        int nl = e->GetLevels();
        std::vector<std::vector<cv::KeyPoint> > keys(nl);
        int o = 0;
        for (int l = 0; l < nl; ++l) { keys[l].assign(kp_inout + o, kp_inout + o + level_counts[l]); o += level_counts[l]; }
        std::vector<cv::KeyPoint> dyn = e->MovingKeyPoints(gray, imS, imLS, centers, rm, flag, keys);
        o = 0;
        for (int l = 0; l < nl; ++l) { level_counts[l] = (int)keys[l].size(); if (!keys[l].empty()) std::memcpy(kp_inout + o, keys[l].data(), sizeof(cv::KeyPoint) * keys[l].size()); o += level_counts[l]; }
        nc = (int)dyn.size();
        if (nc && culled_out) std::memcpy(culled_out, dyn.data(), sizeof(cv::KeyPoint) * nc);
    }
    return nc;
}

// ORBextractor::ProcessDesp   ORBextractor.cc:1747  (uses the pyramid kept from the last detect)
int ref_process_desp(void* h, const cv::KeyPoint* kp_in, const int* level_counts,
                     cv::KeyPoint* kp_out, unsigned char* desc_out, int cap) {
    RefExtractor* e = (RefExtractor*)h;
    ArenaScope scope;
    int n;
    {
        int nl = e->GetLevels();
        std::vector<std::vector<cv::KeyPoint> > keys(nl);
        int o = 0;
        for (int l = 0; l < nl; ++l) { keys[l].assign(kp_in + o, kp_in + o + level_counts[l]); o += level_counts[l]; }
        std::vector<cv::KeyPoint> flat; cv::Mat desc;
        e->ProcessDesp(cv::Mat(), cv::Mat(), keys, flat, desc);
        n = (int)flat.size();
        if (n > cap) return -n;
        if (n) { std::memcpy(kp_out, flat.data(), sizeof(cv::KeyPoint) * n); std::memcpy(desc_out, desc.data, (size_t)n * 32); }
    }
    return n;
}

}  // extern "C"
