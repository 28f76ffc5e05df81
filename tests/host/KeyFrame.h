// tests/host/KeyFrame.h -- TEST STAND-IN for the reference's include/KeyFrame.h: the members ComputeBoW and the two SearchByBoW
// forms touch (/root/reference/include/KeyFrame.h:84-92, 214, 337-364, 401), no behaviour.
#ifndef KEYFRAME_H
#define KEYFRAME_H
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
};
}
#endif
