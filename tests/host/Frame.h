// tests/host/Frame.h -- TEST STAND-IN for the reference's include/Frame.h: the data members and statics the
// on-path matchers / ComputeStereoMatches touch (/root/reference/include/Frame.h:304-489), no behaviour.
#ifndef FRAME_H
#define FRAME_H
#include <vector>
#include <opencv2/core/core.hpp>
#include "MapPoint.h"
#include "ORBextractor.h"
namespace ORB_SLAM2 {
class Frame {
public:
    void ComputeStereoMatches();
    ORBextractor* mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    float fx = 0, fy = 0, cx = 0, cy = 0, mbf = 0, mb = 0;
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    cv::Mat mTcw;
    int mnScaleLevels = 0;
    std::vector<float> mvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
};
}
#endif
