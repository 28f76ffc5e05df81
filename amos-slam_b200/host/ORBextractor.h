// ORBextractor.h -- drop-in replacement for the reference's include/ORBextractor.h (B200-native).
//
// Same class name, namespace, public signatures and public members as ORB_SLAM2::ORBextractor
// (/root/reference/include/ORBextractor.h:93-168); every method marshals into the C ABI of
// liborbx_b200.so (include/orbx_b200.h).  Nothing is computed on the CPU here: if the CUDA library or a
// device is missing the constructor throws std::runtime_error (the reference has no error channel --
// void returns and asserts -- so failures surface as exceptions with orbx_last_error() as the text).
//
// Drop-in use: put this directory before the reference's include/ on the include path, compile
// ORBextractor.cc instead of the reference's src/ORBextractor.cc, link -lorbx_b200.  Frame / Tracking
// compile unchanged.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <vector>
#include <list>
#include <opencv/cv.h>
#ifndef ORBX_HOST_NO_CLUSTER_H
#include <cluster.h>          // ORB_SLAM2::center (/root/reference/include/cluster.h:22-31)
#endif

struct orbx_extractor;

namespace ORB_SLAM2
{

class ORBextractor
{
public:
    enum {HARRIS_SCORE=0, FAST_SCORE=1 };

    // include/ORBextractor.h:93 ; the CUDA device is taken from ORBX_DEVICE (default 0) or SetDefaultDevice()
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
    ~ORBextractor();
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // src/ORBextractor.cc:1544-1668.  The mask is ignored, as in the reference.
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors);
    // :1672-1686  (Amos stage 1: per-level keypoints in level coordinates, no descriptors)
    void operator()(cv::InputArray image, cv::InputArray mask, std::vector<std::vector<cv::KeyPoint> >& _keypoints);
    // :1747-1820  (Amos stage 2)
    void ProcessDesp(cv::InputArray image, cv::InputArray mask, std::vector<std::vector<cv::KeyPoint> >& _allKeypoints,
                     std::vector<cv::KeyPoint>& _mKeypoints, cv::OutputArray descriptors);
    // :1688-1745
    std::vector<cv::KeyPoint> MovingKeyPoints(const cv::Mat &imGray, const cv::Mat &imS, const cv::Mat &imLS, std::vector<center> centers,
                                              std::vector<int> rm_vector, std::vector<bool> DynaFlag,
                                              std::vector<std::vector<cv::KeyPoint> >& mvKeysT);

    int inline GetLevels(){ return nlevels; }
    float inline GetScaleFactor(){ return scaleFactor; }
    std::vector<float> inline GetScaleFactors(){ return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors(){ return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares(){ return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares(){ return mvInvLevelSigma2; }

    // The pyramid lives in HBM.  With export enabled (the default, strict drop-in) every call that builds a
    // pyramid also copies it back so that mvImagePyramid[l] is, as in the reference, an ROI view inside a
    // buffer with a 19-px BORDER_REFLECT_101 frame (:1859-1882).  Callers that consume the pyramid on the
    // device (orbx_compute_stereo_matches) switch the export off and save the D2H traffic.
    std::vector<cv::Mat> mvImagePyramid;
    void SetExportPyramid(bool on) { mbExportPyramid = on; }
    void SyncImagePyramid();                       // explicit D2H of the resident pyramid into mvImagePyramid

    // ---- extensions (not in the reference) ----
    orbx_extractor* handle() { return mpHandle; }  // for the device-side matcher calls
    static void SetDefaultDevice(int device);      // device ordinal used by subsequently constructed extractors

protected:
    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;

    orbx_extractor* mpHandle;
    bool mbExportPyramid;
    std::vector<cv::KeyPoint> mvKpScratch;         // flat marshalling buffers, reused between frames
    std::vector<unsigned char> mvDescScratch;
};

} //namespace ORB_SLAM2

#endif
