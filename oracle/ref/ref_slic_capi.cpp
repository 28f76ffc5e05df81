// ref_slic_capi.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref).  C entry points around the reference's OWN SLIC / k-means code
// (/root/reference/src/cluster.cc:88-344 and :345-464, bodies extracted verbatim at build time by gen_match_bodies.py into
// oracle/_ref/gen/ref_slic_bodies.inc) compiled against the OpenCV-free shim.
// Contract of this oracle:
//   * cv::cvtColor(BGR2Lab) is NOT restated (OpenCV's softfloat-built trilinear LUT): the Lab image is an INPUT, computed by the real
//     OpenCV where the fixtures are generated (tests/golden/make_slic_golden.py) and handed to the reference's SLIC() through
//     cvlite's hook -- the same boundary the product has (the drop-in host code calls cv::cvtColor itself).
//   * the reference seeds k-means with rand() % rowLen + 1 (src/cluster.cc:343-352, unseeded and one past the end for the last index):
//     the canonical contract passes the k seed indices in explicitly; everything after the seeding is the reference's own code.
//     `center vec;` in kmeans() (:424) accumulates into uninitialised ints in the reference; the oracle build zero-fills automatic
//     storage (-ftrivial-auto-var-init=zero on this TU), which is the value every sane run sees.
#include <cstdint>
#include <cstring>
#define private public
#include "cluster.h"
#undef private

namespace ORB_SLAM2 {
cluster::cluster(const cv::Mat&, const cv::Mat&, vector<center>&, const int& nk) : k(nk) {}   // the reference's constructor runs SLIC + k-means itself; the oracle drives the stages
cluster::~cluster() {}
#include "ref_slic_bodies.inc"
}

static thread_local const uint8_t* g_lab = nullptr;
static void lab_hook(const cv::Mat& bgr, cv::Mat& lab) {
    for (int y = 0; y < bgr.rows; ++y) std::memcpy(lab.ptr(y), g_lab + (size_t)y * bgr.cols * 3, (size_t)bgr.cols * 3);
}

extern "C" {
// lab: rows x cols x 3 (L, a, b as cv2.cvtColor(bgr, COLOR_BGR2Lab) returns them), depth: rows x cols u16.
// labels_out: rows x cols doubles (the reference's CV_64F labelMask); centers_out: n x 7 ints (x, y, L, A, B, D, label).
int ref_slic(const uint8_t* lab, const uint16_t* depth, int rows, int cols, int len, int m, double* labels_out, int* centers_out, int cap, int* n_out) {
    cv::Mat image(rows, cols, CV_8UC3), imD(rows, cols, CV_16UC1, (void*)depth);
    std::vector<ORB_SLAM2::center> centers;
    ORB_SLAM2::cluster c(image, imD, centers, 1);
    g_lab = lab; cv::cvl_bgr2lab_hook() = lab_hook;
    cv::Mat labelMask;
    c.SLIC(image, imD, labelMask, centers, len, m);
    cv::cvl_bgr2lab_hook() = nullptr;
    for (int y = 0; y < rows; ++y) std::memcpy(labels_out + (size_t)y * cols, labelMask.ptr<double>(y), sizeof(double) * cols);
    *n_out = (int)centers.size();
    for (int i = 0; i < (int)centers.size() && i < cap; ++i) {
        const ORB_SLAM2::center& s = centers[i];
        int* o = centers_out + (size_t)i * 7; o[0] = s.x; o[1] = s.y; o[2] = s.L; o[3] = s.A; o[4] = s.B; o[5] = s.D; o[6] = s.label;
    }
    return 0;
}
// the gradient image SLIC() builds (:308-312) through cvlite's Sobel / addWeighted: rows x cols x 3 doubles (pinned against cv2 by tests/test_oracle_slic.py)
int ref_slic_gradient(const uint8_t* lab, int rows, int cols, double* out) {
    cv::Mat imageLAB(rows, cols, CV_8UC3, (void*)lab), sx, sy, grad;
    cv::Sobel(imageLAB, sx, CV_64F, 0, 1, 3);
    cv::Sobel(imageLAB, sy, CV_64F, 1, 0, 3);
    cv::addWeighted(sx, 0.5, sy, 0.5, 0, grad);
    for (int y = 0; y < rows; ++y) std::memcpy(out + (size_t)y * cols * 3, grad.ptr<double>(y), sizeof(double) * cols * 3);
    return 0;
}
// k-means of the super-pixel centres with explicit seeds (indices into centres, each with D > 0 as the reference's loop demands).
// ids_out[i] = cluster index of centre i (what the reference stores into centers[label - 1].id, src/cluster.cc:20-27).
int ref_slic_kmeans(const int* centers_in, int n, const int* seeds, int k, int* ids_out) {
    std::vector<ORB_SLAM2::center> centers(n);
    for (int i = 0; i < n; ++i) { const int* s = centers_in + (size_t)i * 7; ORB_SLAM2::center c; std::memset(&c, 0, sizeof(c)); c.x = s[0]; c.y = s[1]; c.L = s[2]; c.A = s[3]; c.B = s[4]; c.D = s[5]; c.label = s[6]; centers[i] = c; }
    cv::Mat dummy;
    ORB_SLAM2::cluster c(dummy, dummy, centers, k);
    c.loadDataSet(centers);
    for (int i = 0; i < k; ++i) c.centroids.push_back(c.dataSet[seeds[i]]);
    c.kmeans();
    for (int i = 0; i < n; ++i) ids_out[i] = -1;
    for (int j = 0; j < k; ++j) for (size_t q = 0; q < c.label[j].size(); ++q) ids_out[c.label[j][q] - 1] = j;
    return 0;
}
}
