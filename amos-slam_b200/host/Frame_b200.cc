// Frame_b200.cc -- B200-native bodies of the Frame methods between the extractor and the matchers.
// Member definitions of the reference's OWN class, so the Frame constructors and Frame::CalDyna keep calling them as
// they do today (src/Frame.cc:187, 240, 378, 428, 510-515, 640-645); a maintainer compiles this file, removes (or
// #ifdef's out) the three bodies it replaces in src/Frame.cc, and adds ONE member to include/Frame.h:
//     struct orbx_frame;                                      (forward declaration, global namespace)
//     #define ORBX_FRAME_HAS_DEVICE 1
//     std::shared_ptr<orbx_frame> mpDeviceFrame;              (+ copy it in Frame's copy constructor, :67-111)
//
//   Frame::UndistortKeyPoints()                       src/Frame.cc:1052-1117
//   Frame::ComputeStereoFromRGBD(const cv::Mat&)      src/Frame.cc:1576-1614
//   Frame::AssignFeaturesToGrid()                     src/Frame.cc:431-461
//
// The keypoints and descriptors the extractor just produced are still in HBM; these bodies build the device-resident frame
// from them (undistortion in double as cv::undistortPoints does, RGB-D stereo, 64x48 grid) and copy back exactly the members
// the rest of the reference reads on the CPU (mvKeysUn, mvuRight, mvDepth, mGrid).  The matcher bodies in ORBmatcher_b200.cc
// then use the device frame instead of uploading the frame again.  Nothing is computed on the CPU here except the N depth
// look-ups imDepth.at<float>(v, u), which would otherwise cost a 1.2 MB upload per frame.
#include "Frame.h"
#include "../../include/orbx_b200.h"

#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI uses");

namespace ORB_SLAM2
{
namespace
{
void check(int rc, const char* what) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string(what) + ": " + orbx_last_error());
}

orbx_camera camera_of(const Frame& F) {               // mK / mDistCoef are CV_32F (Tracking.cc camera block)
    orbx_camera c; std::memset(&c, 0, sizeof(c));
    c.fx = F.mK.at<float>(0, 0); c.fy = F.mK.at<float>(1, 1); c.cx = F.mK.at<float>(0, 2); c.cy = F.mK.at<float>(1, 2);
    const int nd = (int)F.mDistCoef.total();
    if (nd > 0) c.k1 = F.mDistCoef.at<float>(0);
    if (nd > 1) c.k2 = F.mDistCoef.at<float>(1);
    if (nd > 2) c.p1 = F.mDistCoef.at<float>(2);
    if (nd > 3) c.p2 = F.mDistCoef.at<float>(3);
    if (nd > 4) c.k3 = F.mDistCoef.at<float>(4);
    c.bf = F.mbf;
    return c;
}

// Copies of a Frame (mLastFrame = Frame(mCurrentFrame), KeyFrame(F, ...)) share its device frame; a Frame object that is
// built again while a copy is alive gets a fresh one.
orbx_frame* device_frame(Frame& F) {
    if (!F.mpDeviceFrame || F.mpDeviceFrame.use_count() > 1) {
        const char* e = std::getenv("ORBX_DEVICE");
        orbx_frame* f = nullptr;
        check(orbx_frame_create(e ? std::atoi(e) : 0, &f), "orbx_frame_create");
        F.mpDeviceFrame.reset(f, orbx_frame_destroy);
    }
    return F.mpDeviceFrame.get();
}
}  // namespace

void Frame::UndistortKeyPoints()
{
    orbx_frame* f = device_frame(*this);
    // mvKeys / mDescriptors are still on the device when the left extractor's last call produced this frame (ExtractORB* ran
    // just before: Frame.cc:165-176, 367, 505, 637); otherwise (N differs) upload the host copies.
    bool taken = mpORBextractorLeft && orbx_frame_take(f, mpORBextractorLeft->handle()) == ORBX_OK && orbx_frame_taken(f) == N;
    if (!taken) {
        std::vector<unsigned char> tmp;
        const unsigned char* d = mDescriptors.ptr();
        if (N && !mDescriptors.isContinuous()) { tmp.resize((size_t)N * 32); for (int i = 0; i < N; ++i) std::memcpy(&tmp[(size_t)i * 32], mDescriptors.ptr(i), 32); d = tmp.data(); }
        check(orbx_frame_take_host(f, reinterpret_cast<const orbx_keypoint*>(mvKeys.data()), d, N, (int)mvScaleFactors.size(), mvScaleFactors.data()), "orbx_frame_take_host");
    }
    const orbx_camera cam = camera_of(*this);
    mvKeysUn.resize(N);
    check(orbx_frame_undistort_keypoints(f, &cam, N ? reinterpret_cast<orbx_keypoint*>(mvKeysUn.data()) : NULL), "orbx_frame_undistort_keypoints");
}

void Frame::ComputeStereoFromRGBD(const cv::Mat &imDepth)
{
    mvuRight = std::vector<float>(N, -1);                                                      // :1581-1582
    mvDepth = std::vector<float>(N, -1);
    if (N == 0) return;
    std::vector<float> d(N);
    for (int i = 0; i < N; ++i) d[i] = imDepth.at<float>(mvKeys[i].pt.y, mvKeys[i].pt.x);      // :1590-1595 (raw keypoint, float -> int truncation)
    check(orbx_frame_compute_stereo_from_rgbd(mpDeviceFrame.get(), mbf, d.data(), 0, 0, 0, mvuRight.data(), mvDepth.data()), "orbx_frame_compute_stereo_from_rgbd");
}

void Frame::AssignFeaturesToGrid()
{
    orbx_frame* f = mpDeviceFrame.get();
    if (!f) throw std::runtime_error("Frame::AssignFeaturesToGrid before UndistortKeyPoints");
    // stereo constructor: ComputeStereoMatches filled mvuRight / mvDepth on the host after the undistortion step (:187-196)
    if (N && (int)mvuRight.size() == N && (int)mvDepth.size() == N) check(orbx_frame_set_stereo(f, mvuRight.data(), mvDepth.data()), "orbx_frame_set_stereo");
    const float b[6] = {mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv};
    std::vector<int> cs(64 * 48 + 1), en(N ? N : 1);
    check(orbx_frame_assign_features_to_grid(f, b, cs.data(), en.data()), "orbx_frame_assign_features_to_grid");
    for (int x = 0; x < 64; ++x)
        for (int y = 0; y < 48; ++y) {
            const int c = x * 48 + y;
            mGrid[x][y].assign(en.begin() + cs[c], en.begin() + cs[c + 1]);
        }
}

} //namespace ORB_SLAM2
