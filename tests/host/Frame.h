// tests/host/Frame.h -- TEST STAND-IN for the reference's include/Frame.h: the data members and statics the
// on-path matchers / ComputeStereoMatches touch (/root/reference/include/Frame.h:304-489), no behaviour.
#ifndef FRAME_H
#define FRAME_H
#define ORBX_FRAME_HAS_DEVICE 1                             // defined next to the mpDeviceFrame member
#include <memory>
#include <vector>
#include <opencv2/core/core.hpp>
#include "MapPoint.h"
#include "DBoW2_standin.h"
#include "ORBextractor.h"
struct orbx_frame;                                          // include/orbx_b200.h (global namespace)
namespace ORB_SLAM2 {
class Frame {
public:
    void ComputeStereoMatches();
    void UndistortKeyPoints();
    void ComputeStereoFromRGBD(const cv::Mat& imDepth);
    void AssignFeaturesToGrid();
    void ComputeBoW();
    ORBVocabulary* mpORBvocabulary = nullptr;
    DBoW2::BowVector mBowVec;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mK, mDistCoef;
    std::vector<std::size_t> mGrid[64][48];                 // FRAME_GRID_COLS x FRAME_GRID_ROWS (include/Frame.h:56-61, 459)
    std::shared_ptr<orbx_frame> mpDeviceFrame;             // the ONE member a maintainer adds to include/Frame.h
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
