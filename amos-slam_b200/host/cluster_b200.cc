// cluster_b200.cc -- drop-in body of ORB_SLAM2::cluster::SLIC (/root/reference/src/cluster.cc:295-344) on top of the C ABI.
// Build it INSTEAD of the reference's SLIC / clustering / updateCenter / initilizeCenters / fituneCenter bodies;
// include/cluster.h stays untouched: the class, its constructor, evalImage() and the k-means that follows keep calling SLIC() as before.
// The colour conversion stays OpenCV's own call (the reference's line :305) -- its 8-bit Lab path is a LUT built with OpenCV's
// softfloat and is not restated; everything after it runs on the GPU and returns what the reference's loops return, bit for bit
// (tests/test_gpu_slic.py against the reference build).  ~1 ms per 640 x 480 frame instead of ~130 ms on one core.
#include "cluster.h"
#include <opencv2/imgproc/imgproc.hpp>
#include <stdexcept>
#include <string>
#include "orbx_b200.h"

namespace ORB_SLAM2 {

namespace {
struct SlicHandle {                                  // one device handle per host thread (cluster objects are short-lived locals, Frame.cc:633)
    orbx_slic* h = nullptr;
    ~SlicHandle() { if (h) orbx_slic_destroy(h); }
};
orbx_slic* slic_handle() {
    static thread_local SlicHandle s;
    if (!s.h && orbx_slic_create(0, &s.h) != ORBX_OK) throw std::runtime_error(std::string("orbx_slic_create: ") + orbx_last_error());
    return s.h;
}
}

int cluster::SLIC(cv::Mat const &image, cv::Mat const &image_D, cv::Mat &resultLabel, std::vector<center> &centers, int len, int m)
{
    const int height = image.rows, width = image.cols;
    cv::Mat imageLAB;
    cv::cvtColor(image, imageLAB, cv::COLOR_BGR2Lab);                                   // :305
    cv::Mat labelMask(height, width, CV_64FC1);
    const int cap = (height / len + 1) * (width / len + 1);
    std::vector<orbx_slic_center> c((size_t)cap);
    int n = 0;
    const int rc = orbx_slic_run(slic_handle(), imageLAB.data, imageLAB.step, reinterpret_cast<const uint16_t*>(image_D.data), image_D.step, height, width, len, m, 5,
                                 reinterpret_cast<double*>(labelMask.data), labelMask.step, nullptr, 0, c.data(), cap, &n);
    if (rc != ORBX_OK) throw std::runtime_error(std::string("orbx_slic_run: ") + orbx_last_error());
    for (int i = 0; i < n; ++i) {                                                       // initilizeCenters appends to the caller's vector (:232)
        center cent;
        cent.x = c[i].x; cent.y = c[i].y; cent.L = c[i].L; cent.A = c[i].A; cent.B = c[i].B; cent.D = c[i].D; cent.label = c[i].label; cent.id = 0;
        centers.push_back(cent);
    }
    resultLabel = labelMask;                                                            // :341
    return 0;
}

}  // namespace ORB_SLAM2
