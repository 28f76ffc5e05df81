// tests/host/MapPoint.h -- TEST STAND-IN for the reference's include/MapPoint.h: just the members the on-path
// matchers read (/root/reference/include/MapPoint.h:86,108,154,211,240-246), backed by plain fields.
#ifndef MAPPOINT_H
#define MAPPOINT_H
#include <opencv2/core/core.hpp>
namespace ORB_SLAM2 {
class MapPoint {
public:
    cv::Mat GetWorldPos() { return mWorldPos; }
    cv::Mat GetDescriptor() { return mDescriptor; }
    cv::Mat GetNormal() { return mNormalVector; }
    cv::Mat mNormalVector;
    int Observations() { return nObs; }
    bool isBad() { return mbBad; }
    float GetMinDistanceInvariance() { return mfMinDistance; }
    float GetMaxDistanceInvariance() { return mfMaxDistance; }
    template <class FrameT> int PredictScale(const float&, FrameT*) { return nPredictedLevel; }
    template <class KF> bool IsInKeyFrame(KF* kf) { return (const void*)kf == pInKF; }
    template <class KF> void AddObservation(KF* kf, size_t idx) { pInKF = kf; nIdxInKF = (int)idx; ++nObs; }
    void Replace(MapPoint* p) { pReplaced = p; mbBad = true; }
    MapPoint* pReplaced = nullptr;
    template <class KF> int GetIndexInKeyFrame(KF* kf) { return (const void*)kf == pInKF ? nIdxInKF : -1; }
    const void* pInKF = nullptr; int nIdxInKF = -1;
    float mfMinDistance = 0.f, mfMaxDistance = 1e30f; int nPredictedLevel = 0;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0;
    bool mbTrackInView = false;
    int mnTrackScaleLevel = 0;
    float mTrackViewCos = 0;
    // stand-in storage
    cv::Mat mWorldPos, mDescriptor; int nObs = 0; bool mbBad = false;
};
}
#endif
