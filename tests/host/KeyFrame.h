// tests/host/KeyFrame.h -- TEST STAND-IN for the reference's include/KeyFrame.h: the members ComputeBoW and the two SearchByBoW
// forms touch (/root/reference/include/KeyFrame.h:84-92, 214, 337-364, 401), no behaviour.
#ifndef KEYFRAME_H
#define KEYFRAME_H
#include <set>
#include <vector>
#include <opencv2/core/core.hpp>
#include "MapPoint.h"
#include "DBoW2_standin.h"
namespace ORB_SLAM2 {
class KeyFrame {
public:
    void ComputeBoW();
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    ORBVocabulary* mpORBvocabulary = nullptr;
    bool IsInImage(const float &x, const float &y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }   // src/KeyFrame.cc:799-802
    MapPoint* GetMapPoint(const size_t &idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* pMP, const size_t &idx) { mvpMapPoints[idx] = pMP; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (size_t i = 0; i < mvpMapPoints.size(); ++i) if (mvpMapPoints[i]) s.insert(mvpMapPoints[i]); return s; }
    float mbf = 0;
    std::vector<float> mvInvLevelSigma2;
    cv::Mat GetCameraCenter() { return mOw; }
    cv::Mat mOw;
    std::vector<float> mvuRight, mvLevelSigma2;
    cv::Mat GetRotation() { return mRcw; }
    cv::Mat GetTranslation() { return mtcw; }
    cv::Mat mRcw, mtcw;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    int mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<float> mvScaleFactors;
};
}
#endif
