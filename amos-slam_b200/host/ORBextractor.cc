// ORBextractor.cc -- host side of the drop-in ORB_SLAM2::ORBextractor: flattens the reference's C++
// arguments (cv::Mat / std::vector<cv::KeyPoint>) and calls the C ABI of liborbx_b200.so.
// Replaces /root/reference/src/ORBextractor.cc in the reference's build.  No image
// arithmetic happens here; there is no CPU fallback.
#include "ORBextractor.h"
#include "../../include/orbx_b200.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI uses");

namespace ORB_SLAM2
{

static int g_default_device = -1;

static int pick_device() {
    if (g_default_device >= 0) return g_default_device;
    const char* e = std::getenv("ORBX_DEVICE");
    return e ? std::atoi(e) : 0;
}

static void check(int rc, const char* what) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string(what) + ": " + orbx_last_error());
}

void ORBextractor::SetDefaultDevice(int device) { g_default_device = device; }

ORBextractor::ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _iniThFAST, int _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST),
      mpHandle(nullptr), mbExportPyramid(true)
{
    check(orbx_create(_nfeatures, _scaleFactor, _nlevels, _iniThFAST, _minThFAST, pick_device(), &mpHandle), "orbx_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
    mnFeaturesPerLevel.resize(nlevels);
    orbx_get_scale_factors(mpHandle, mvScaleFactor.data());
    orbx_get_inverse_scale_factors(mpHandle, mvInvScaleFactor.data());
    orbx_get_scale_sigma_squares(mpHandle, mvLevelSigma2.data());
    orbx_get_inverse_scale_sigma_squares(mpHandle, mvInvLevelSigma2.data());
    orbx_get_features_per_level(mpHandle, mnFeaturesPerLevel.data());
    mvImagePyramid.resize(nlevels);                                               // src/ORBextractor.cc:526
}

ORBextractor::~ORBextractor() { orbx_destroy(mpHandle); }

void ORBextractor::SyncImagePyramid()
{
    const int B = 19;                                                             // EDGE_THRESHOLD
    for (int l = 0; l < nlevels; ++l) {
        int r = 0, c = 0;
        check(orbx_pyramid_level(mpHandle, l, B, nullptr, 0, &r, &c), "orbx_pyramid_level");
        cv::Mat temp(r, c, CV_8UC1);
        check(orbx_pyramid_level(mpHandle, l, B, temp.ptr(), temp.step, &r, &c), "orbx_pyramid_level");
        mvImagePyramid[l] = temp(cv::Rect(B, B, c - 2 * B, r - 2 * B));           // ROI inside the padded buffer, as :1844-1846
    }
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors)
{
    if (_image.empty()) return;                                                   // :1553
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");   // assert at :1559
    const int cap = orbx_max_keypoints(mpHandle, image.rows, image.cols);
    if (cap < 0) check(cap, "orbx_max_keypoints");
    _keypoints.resize(cap);
    mvDescScratch.resize((size_t)cap * 32);
    int n = 0;
    check(orbx_extract(mpHandle, image.ptr(), image.rows, image.cols, image.step, reinterpret_cast<orbx_keypoint*>(_keypoints.data()),
                       mvDescScratch.data(), cap, &n), "orbx_extract");
    _keypoints.resize(n);
    if (n == 0) _descriptors.release();                                           // :1590-1591
    else {
        _descriptors.create(n, 32, CV_8U);                                        // :1594-1596
        cv::Mat d = _descriptors.getMat();
        for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &mvDescScratch[(size_t)i * 32], 32);
    }
    if (mbExportPyramid) SyncImagePyramid();
}

void ORBextractor::operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<std::vector<cv::KeyPoint> >& _keypoints)
{
    if (_image.empty()) return;                                                   // :1677
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor: image must be CV_8UC1");
    const int cap = orbx_max_keypoints(mpHandle, image.rows, image.cols);
    if (cap < 0) check(cap, "orbx_max_keypoints");
    mvKpScratch.resize(cap);
    std::vector<int> counts(nlevels, 0);
    int n = 0;
    check(orbx_detect(mpHandle, image.ptr(), image.rows, image.cols, image.step, reinterpret_cast<orbx_keypoint*>(mvKpScratch.data()),
                      counts.data(), cap, &n), "orbx_detect");
    _keypoints.assign(nlevels, std::vector<cv::KeyPoint>());
    int o = 0;
    for (int l = 0; l < nlevels; ++l) { _keypoints[l].assign(mvKpScratch.begin() + o, mvKpScratch.begin() + o + counts[l]); o += counts[l]; }
    if (mbExportPyramid) SyncImagePyramid();
}

std::vector<cv::KeyPoint> ORBextractor::MovingKeyPoints(const cv::Mat &imGray, const cv::Mat &imS, const cv::Mat &imLS, std::vector<center> centers,
                                                        std::vector<int> rm_vector, std::vector<bool> DynaFlag,
                                                        std::vector<std::vector<cv::KeyPoint> >& mvKeysT)
{
    if (imS.type() != CV_8UC1 || imLS.type() != CV_64F || imS.rows != imLS.rows || imS.cols != imLS.cols)
        throw std::runtime_error("MovingKeyPoints: imS must be CV_8UC1 and imLS CV_64F of the same size");
    std::vector<int> ids(centers.size());
    for (size_t i = 0; i < centers.size(); ++i) ids[i] = centers[i].id;           // rm_vector[centers[label-1].id]  :1730
    std::vector<int> counts(nlevels, 0);
    size_t total = 0;
    for (int l = 0; l < nlevels && l < (int)mvKeysT.size(); ++l) { counts[l] = (int)mvKeysT[l].size(); total += mvKeysT[l].size(); }
    mvKpScratch.resize(total ? total : 1);
    std::vector<cv::KeyPoint> culled(total ? total : 1);
    size_t o = 0;
    for (int l = 0; l < nlevels && l < (int)mvKeysT.size(); ++l) { std::copy(mvKeysT[l].begin(), mvKeysT[l].end(), mvKpScratch.begin() + o); o += mvKeysT[l].size(); }
    int nculled = 0;
    check(orbx_cull(mpHandle, imS.ptr(), imS.step, imLS.ptr<double>(), imLS.step, imS.rows, imS.cols, ids.data(), (int)ids.size(),
                    rm_vector.data(), (int)rm_vector.size(), reinterpret_cast<orbx_keypoint*>(mvKpScratch.data()), counts.data(),
                    reinterpret_cast<orbx_keypoint*>(culled.data()), &nculled), "orbx_cull");
    o = 0;
    for (int l = 0; l < nlevels && l < (int)mvKeysT.size(); ++l) { mvKeysT[l].assign(mvKpScratch.begin() + o, mvKpScratch.begin() + o + counts[l]); o += counts[l]; }
    culled.resize(nculled);
    return culled;
}

void ORBextractor::ProcessDesp(cv::InputArray _image, cv::InputArray _mask, std::vector<std::vector<cv::KeyPoint> >& _allKeypoints,
                               std::vector<cv::KeyPoint>& _mKeypoints, cv::OutputArray _descriptors)
{
    std::vector<int> counts(nlevels, 0);
    size_t total = 0;
    for (int l = 0; l < nlevels && l < (int)_allKeypoints.size(); ++l) { counts[l] = (int)_allKeypoints[l].size(); total += _allKeypoints[l].size(); }
    _mKeypoints.clear();
    if (total == 0) { _descriptors.release(); return; }                           // :1760-1761
    mvKpScratch.resize(total);
    size_t o = 0;
    for (int l = 0; l < nlevels && l < (int)_allKeypoints.size(); ++l) {
        for (size_t i = 0; i < _allKeypoints[l].size(); ++i) { mvKpScratch[o + i] = _allKeypoints[l][i]; mvKpScratch[o + i].octave = l; }
        o += _allKeypoints[l].size();
    }
    _mKeypoints.resize(total);
    mvDescScratch.resize(total * 32);
    int n = 0;
    check(orbx_describe(mpHandle, reinterpret_cast<const orbx_keypoint*>(mvKpScratch.data()), counts.data(),
                        reinterpret_cast<orbx_keypoint*>(_mKeypoints.data()), mvDescScratch.data(), (int)total, &n), "orbx_describe");
    _mKeypoints.resize(n);
    _descriptors.create(n, 32, CV_8U);
    cv::Mat d = _descriptors.getMat();
    for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), &mvDescScratch[(size_t)i * 32], 32);
}

} //namespace ORB_SLAM2
