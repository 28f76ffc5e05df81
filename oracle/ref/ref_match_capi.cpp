// ref_match_capi.cpp -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Runs the reference's OWN matcher code: the bodies of ORBmatcher::{SearchByProjection x2, SearchByBoW x2, SearchForInitialization,
// ComputeThreeMaxima, DescriptorDistance, RadiusByViewingCos} and Frame::{AssignFeaturesToGrid, GetFeaturesInArea,
// PosInGrid, ComputeStereoMatches, UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD} are taken verbatim from /root/reference/src/{ORBmatcher,Frame}.cc at build time
// (oracle/ref/gen_match_bodies.py -> oracle/_ref/gen/ref_match_bodies.inc) and compiled against the minimal
// Frame / MapPoint / ORBmatcher declarations below, which carry exactly the members those bodies touch
// (/root/reference/include/Frame.h, MapPoint.h, ORBmatcher.h).  The whole object graph of the reference
// (KeyFrame, Map, DBoW2, g2o ...) is not needed by these functions and is not built.
#include "ORBextractor.h"   // reference header (for mvImagePyramid in ComputeStereoMatches)
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"   // reference header (mFeatVec in SearchByBoW); FeatureVector.cpp is built in ref_bow_capi.cpp
#include <climits>
#include <cstring>
#include <map>
#include <mutex>
#include <set>
#include <vector>
#include <list>
#include <cmath>
#include <stdint.h>

extern "C" void ref_arena_begin();
extern "C" void ref_arena_end();

#define FRAME_GRID_ROWS 48   // include/Frame.h:56
#define FRAME_GRID_COLS 64   // include/Frame.h:61

namespace ORB_SLAM2 {

class Frame;
class KeyFrame;

class MapPoint {               // the members ORBmatcher.cc:70-175 and :1569-1728 read (include/MapPoint.h)
public:
    bool mbTrackInView; int mnTrackScaleLevel; float mTrackViewCos, mTrackProjX, mTrackProjY, mTrackProjXR;
    bool bad; int nobs; cv::Mat desc, pos; float maxd, mind; int plevel;
    MapPoint() : mbTrackInView(true), mnTrackScaleLevel(0), mTrackViewCos(1.f), mTrackProjX(0), mTrackProjY(0), mTrackProjXR(0), bad(false), nobs(0), maxd(1e30f), mind(0.f), plevel(0), in_kf(NULL), in_idx(-1), mbBad(false), fused_idx(-1), replaced_with(NULL) {}
    float GetMaxDistanceInvariance() { return maxd; }
    float GetMinDistanceInvariance() { return mind; }
    int PredictScale(const float&, Frame*) { return plevel; }          // the level is an input of the C-ABI call: the harness supplies it
    int PredictScale(const float&, KeyFrame*) { return plevel; }
    cv::Mat normal;
    cv::Mat GetNormal() { return normal.clone(); }
    KeyFrame* in_kf; int in_idx;                                       // one observation is enough for SearchBySim3's vbAlreadyMatched2 (:1350)
    int GetIndexInKeyFrame(KeyFrame* kf) { return kf == in_kf ? in_idx : -1; }
    bool IsInKeyFrame(KeyFrame* kf) { return kf == in_kf; }
    // Fuse's map surgery is recorded, not performed: the harness reads which KeyFrame feature each point was fused with
    // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:359-439) reads these (include/MapPoint.h:228-262)
    void ComputeDistinctiveDescriptors();
    std::mutex mMutexFeatures; bool mbBad; std::map<KeyFrame*, size_t> mObservations; cv::Mat mDescriptor;
    int fused_idx; MapPoint* replaced_with;
    void Replace(MapPoint* p);                                         // logged (below)
    void AddObservation(KeyFrame*, size_t idx) { fused_idx = (int)idx; }
    bool isBad() { return bad; }
    int Observations() { return nobs; }
    cv::Mat GetDescriptor() { return desc.clone(); }
    cv::Mat GetWorldPos() { return pos.clone(); }
};

static std::vector<std::pair<MapPoint*, MapPoint*> > g_replace_log;  // (this, argument) of every MapPoint::Replace call, in order
inline void MapPoint::Replace(MapPoint* p) { replaced_with = p; g_replace_log.push_back(std::make_pair(this, p)); }

class Frame {                  // include/Frame.h, only what the extracted bodies use
public:
    Frame() : mpORBextractorLeft(NULL), mpORBextractorRight(NULL), mbf(0), mb(0), N(0) {}
    void AssignFeaturesToGrid();
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const;
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    void ComputeStereoMatches();
    void UndistortKeyPoints();
    void ComputeImageBounds(const cv::Mat& imLeft);
    void ComputeStereoFromRGBD(const cv::Mat& imDepth);
    cv::Mat mK, mDistCoef;
    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    static float fx, fy, cx, cy;
    float mbf, mb;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors, mDescriptorsRight;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    static float mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    cv::Mat mTcw;
    std::vector<float> mvScaleFactors, mvInvScaleFactors;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY;
    DBoW2::FeatureVector mFeatVec;
};

class KeyFrame {               // include/KeyFrame.h, the members SearchByBoW / SearchByProjection(KeyFrame*, Scw, ...) read
public:
    KeyFrame() : kf_bad(false), N(0), fx(0), fy(0), cx(0), cy(0), mnGridCols(FRAME_GRID_COLS), mnGridRows(FRAME_GRID_ROWS), mfGridElementWidthInv(0), mfGridElementHeightInv(0),
                 mnMinX(0), mnMinY(0), mnMaxX(0), mnMaxY(0) {}
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
    bool kf_bad; bool isBad() { return kf_bad; }
    cv::Mat Rcw, tcw, Ow;                                              // empty = KeyFrame at the world origin
    cv::Mat GetRotation() { return Rcw.empty() ? cv::Mat(cv::Mat::eye(3, 3, CV_32F)) : Rcw.clone(); }
    cv::Mat GetTranslation() { return tcw.empty() ? cv::Mat(cv::Mat::zeros(3, 1, CV_32F)) : tcw.clone(); }
    cv::Mat GetCameraCenter() { return Ow.empty() ? cv::Mat(cv::Mat::zeros(3, 1, CV_32F)) : Ow.clone(); }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    void AddMapPoint(MapPoint* p, const size_t& idx) { mvpMapPoints[idx] = p; }
    std::set<MapPoint*> GetMapPoints() { std::set<MapPoint*> s; for (size_t i = 0; i < mvpMapPoints.size(); ++i) if (mvpMapPoints[i]) s.insert(mvpMapPoints[i]); return s; }
    float mbf;
    std::vector<float> mvuRight, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r) const;
    bool IsInImage(const float& x, const float& y) const;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    DBoW2::FeatureVector mFeatVec;
    int N; float fx, fy, cx, cy;
    const int mnGridCols, mnGridRows; float mfGridElementWidthInv, mfGridElementHeightInv;
    int mnMinX, mnMinY, mnMaxX, mnMaxY;                      // include/KeyFrame.h:408-411 (ints there)
    std::vector<float> mvScaleFactors;
    std::vector<std::vector<std::vector<size_t> > > mGrid;   // include/KeyFrame.h:383
};
float Frame::fx, Frame::fy, Frame::cx, Frame::cy, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;

class ORBmatcher {             // include/ORBmatcher.h:57-215, the subset on the hot path
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true);
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    int SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched, int th);
    int SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float& s12, const cv::Mat& R12, const cv::Mat& t12, const float th);
    int Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th = 3.0);
    int Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, cv::Mat F12, std::vector<std::pair<size_t, size_t> >& vMatchedPairs, const bool bOnlyStereo);
    bool CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
protected:
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio; bool mbCheckOrientation;
};

#include "ref_match_bodies.inc"   // GENERATED: verbatim reference function bodies

}  // namespace ORB_SLAM2

using namespace ORB_SLAM2;

namespace {
struct ArenaScope { ArenaScope() { ref_arena_begin(); } ~ArenaScope() { ref_arena_end(); } };

// mirrors orbx_frame_view (include/orbx_b200.h)
struct FrameView {
    int n; const cv::KeyPoint* keys_un; const unsigned char* descriptors; const float* u_right;
    float min_x, min_y, max_x, max_y, gw_inv, gh_inv; int nlevels; const float* scale_factors;
};
void fill_frame(Frame& F, const FrameView* v) {
    F.N = v->n;
    F.mvKeysUn.assign(v->keys_un, v->keys_un + v->n);
    F.mvKeys = F.mvKeysUn;
    F.mDescriptors = cv::Mat(v->n, 32, CV_8U, (void*)v->descriptors).clone();
    if (v->u_right) F.mvuRight.assign(v->u_right, v->u_right + v->n); else F.mvuRight.assign(v->n, -1.f);
    F.mvpMapPoints.assign(v->n, (MapPoint*)NULL);
    F.mvbOutlier.assign(v->n, false);
    F.mvScaleFactors.assign(v->scale_factors, v->scale_factors + v->nlevels);
    Frame::mnMinX = v->min_x; Frame::mnMinY = v->min_y; Frame::mnMaxX = v->max_x; Frame::mnMaxY = v->max_y;
    Frame::mfGridElementWidthInv = v->gw_inv; Frame::mfGridElementHeightInv = v->gh_inv;
    F.AssignFeaturesToGrid();
}
}  // namespace

static void fill_fv(DBoW2::FeatureVector& fv, int n_fv, const int* nodes, const int* offsets, const int* idx) {
    for (int q = 0; q < n_fv; ++q) for (int e = offsets[q]; e < offsets[q + 1]; ++e) fv.addFeature(nodes[q], idx[e]);
}

extern "C" {

// ORBmatcher::DescriptorDistance   ORBmatcher.cc:1913
void ref_descriptor_distance(const unsigned char* a, const unsigned char* b, int n, int* out) {
    for (int i = 0; i < n; ++i) {
        cv::Mat ma(1, 32, CV_8U, (void*)(a + (size_t)i * 32)), mb(1, 32, CV_8U, (void*)(b + (size_t)i * 32));
        out[i] = ORBmatcher::DescriptorDistance(ma, mb);
    }
}

// Frame::GetFeaturesInArea   Frame.cc:894  (result order matters to the matchers' tie-breaks)
int ref_get_features_in_area(const FrameView* v, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    ArenaScope scope;
    int n;
    { Frame F; fill_frame(F, v); std::vector<size_t> idx = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel); n = (int)idx.size(); for (int i = 0; i < n && i < cap; ++i) out[i] = (int)idx[i]; }
    return n;
}

// ORBmatcher::SearchForInitialization   ORBmatcher.cc:515
int ref_search_for_initialization(float nnratio, int checkOri, const FrameView* v1, const FrameView* v2, float* prev_matched_xy, int* matches12, int windowSize) {
    ArenaScope scope;
    int nm;
    {
        Frame F1, F2; fill_frame(F1, v1); fill_frame(F2, v2);
        std::vector<cv::Point2f> prev(v1->n);
        for (int i = 0; i < v1->n; ++i) prev[i] = cv::Point2f(prev_matched_xy[2 * i], prev_matched_xy[2 * i + 1]);
        std::vector<int> m12;
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchForInitialization(F1, F2, prev, m12, windowSize);
        for (int i = 0; i < v1->n; ++i) { matches12[i] = m12[i]; prev_matched_xy[2 * i] = prev[i].x; prev_matched_xy[2 * i + 1] = prev[i].y; }
    }
    return nm;
}

// ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)   ORBmatcher.cc:1569
// Poses are identity (Tcw = Tlw = I), so x3Dc == x3Dw == cam_xyz exactly; the projection the body computes
// (:1608-1618) is returned in proj_uv / proj_invz so the caller can feed the same numbers to the C-ABI call.
int ref_search_by_projection_frame(float nnratio, int checkOri, const FrameView* cur, int n_last, const float* cam_xyz,
                                   const int* last_octave, const float* last_angle, const unsigned char* mp_desc, const unsigned char* valid,
                                   const unsigned char* mp_observed, const unsigned char* cur_occupied, float th, int bMono, float mb, float mbf, float fx, float fy, float cx, float cy,
                                   float* proj_uv, float* proj_invz, int* cur_match) {
    ArenaScope scope;
    int nm;
    {
        Frame C, L; fill_frame(C, cur);
        Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
        C.mb = mb; C.mbf = mbf;
        C.mTcw = cv::Mat::eye(4, 4, CV_32F); L.mTcw = cv::Mat::eye(4, 4, CV_32F);
        std::vector<MapPoint> pts(n_last), occ(cur->n);
        L.N = n_last; L.mvKeys.resize(n_last); L.mvKeysUn.resize(n_last); L.mvpMapPoints.assign(n_last, (MapPoint*)NULL); L.mvbOutlier.assign(n_last, false);
        for (int i = 0; i < n_last; ++i) {
            L.mvKeys[i].octave = last_octave[i]; L.mvKeysUn[i].octave = last_octave[i]; L.mvKeysUn[i].angle = last_angle[i];
            pts[i].pos = cv::Mat(3, 1, CV_32F); for (int k = 0; k < 3; ++k) pts[i].pos.at<float>(k) = cam_xyz[3 * i + k];
            pts[i].desc = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + (size_t)i * 32)).clone();
            pts[i].nobs = (mp_observed && mp_observed[i]) ? 1 : 0;
            if (valid[i]) L.mvpMapPoints[i] = &pts[i];
            // the body's own projection, restated for the caller (:1608-1618)
            const float xc = cam_xyz[3 * i], yc = cam_xyz[3 * i + 1];
            const float invzc = 1.0 / cam_xyz[3 * i + 2];
            proj_invz[i] = invzc; proj_uv[2 * i] = fx * xc * invzc + cx; proj_uv[2 * i + 1] = fy * yc * invzc + cy;
        }
        for (int j = 0; j < cur->n; ++j) if (cur_occupied && cur_occupied[j]) { occ[j].nobs = 1; C.mvpMapPoints[j] = &occ[j]; }
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchByProjection(C, L, th, bMono != 0);
        for (int j = 0; j < cur->n; ++j) {
            MapPoint* p = C.mvpMapPoints[j];
            cur_match[j] = (p && p >= &pts[0] && p < &pts[0] + n_last) ? (int)(p - &pts[0]) : -1;
        }
    }
    return nm;
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)   ORBmatcher.cc:70
int ref_search_by_projection_points(float nnratio, int checkOri, const FrameView* Fv, int n_points, const float* track_uv, const float* track_ur,
                                    const int* track_level, const float* track_view_cos, const unsigned char* mp_desc, const unsigned char* mp_observed,
                                    const unsigned char* f_occupied, float th, int* f_match) {
    ArenaScope scope;
    int nm;
    {
        Frame F; fill_frame(F, Fv);
        std::vector<MapPoint> pts(n_points), occ(Fv->n);
        std::vector<MapPoint*> vp(n_points);
        for (int i = 0; i < n_points; ++i) {
            pts[i].mTrackProjX = track_uv[2 * i]; pts[i].mTrackProjY = track_uv[2 * i + 1]; pts[i].mTrackProjXR = track_ur[i];
            pts[i].mnTrackScaleLevel = track_level[i]; pts[i].mTrackViewCos = track_view_cos[i];
            pts[i].desc = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + (size_t)i * 32)).clone();
            pts[i].nobs = (mp_observed && mp_observed[i]) ? 1 : 0;
            vp[i] = &pts[i];
        }
        for (int j = 0; j < Fv->n; ++j) if (f_occupied && f_occupied[j]) { occ[j].nobs = 1; F.mvpMapPoints[j] = &occ[j]; }
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchByProjection(F, vp, th);
        for (int j = 0; j < Fv->n; ++j) {
            MapPoint* p = F.mvpMapPoints[j];
            f_match[j] = (p && p >= &pts[0] && p < &pts[0] + n_points) ? (int)(p - &pts[0]) : -1;
        }
    }
    return nm;
}

// The frame steps between extractor and matchers, in the order Frame::CalDyna runs them (Frame.cc:636-645), preceded by the
// constructor's one-time ComputeImageBounds + grid constants (:296-303):
//   UndistortKeyPoints (:1052) -> ComputeStereoFromRGBD (:1576, when depth != NULL) -> AssignFeaturesToGrid (:431).
// cam9 = fx fy cx cy k1 k2 p1 p2 k3 (mK / mDistCoef are CV_32F as in Tracking.cc); ndist = 4 or 5 coefficients.
// Outputs: keys_un[n], u_right[n], depth_out[n], bounds[6] = minX maxX minY maxY gridWInv gridHInv,
// cell_start[64*48+1] / entries[n] = mGrid flattened cell-major (x * 48 + y), push_back order inside a cell.
int ref_frame_build(const cv::KeyPoint* keys, int n, const float* cam9, int ndist, float bf, int rows, int cols, const float* depth_img,
                    cv::KeyPoint* keys_un, float* u_right, float* depth_out, float* bounds, int* cell_start, int* entries) {
    ArenaScope scope;
    {
        Frame F;
        F.N = n; F.mvKeys.assign(keys, keys + n); F.mbf = bf;
        F.mK = cv::Mat::eye(3, 3, CV_32F);
        F.mK.at<float>(0, 0) = cam9[0]; F.mK.at<float>(1, 1) = cam9[1]; F.mK.at<float>(0, 2) = cam9[2]; F.mK.at<float>(1, 2) = cam9[3];
        F.mDistCoef = cv::Mat(ndist, 1, CV_32F);
        for (int i = 0; i < ndist; ++i) F.mDistCoef.at<float>(i) = cam9[4 + i];
        cv::Mat gray(rows, cols, CV_8UC1);                                  // ComputeImageBounds only reads its size
        F.ComputeImageBounds(gray);
        Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);     // Frame.cc:301-302
        Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
        F.UndistortKeyPoints();
        if (depth_img) {
            cv::Mat imD(rows, cols, CV_32F, (void*)depth_img);
            F.ComputeStereoFromRGBD(imD);
        } else { F.mvuRight.assign(n, -1.f); F.mvDepth.assign(n, -1.f); }
        F.AssignFeaturesToGrid();
        for (int i = 0; i < n; ++i) { keys_un[i] = F.mvKeysUn[i]; u_right[i] = F.mvuRight[i]; depth_out[i] = F.mvDepth[i]; }
        bounds[0] = Frame::mnMinX; bounds[1] = Frame::mnMaxX; bounds[2] = Frame::mnMinY; bounds[3] = Frame::mnMaxY;
        bounds[4] = Frame::mfGridElementWidthInv; bounds[5] = Frame::mfGridElementHeightInv;
        int o = 0;
        for (int x = 0; x < FRAME_GRID_COLS; ++x) for (int y = 0; y < FRAME_GRID_ROWS; ++y) {
            cell_start[x * FRAME_GRID_ROWS + y] = o;
            for (size_t q = 0; q < F.mGrid[x][y].size(); ++q) entries[o++] = (int)F.mGrid[x][y][q];
        }
        cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = o;
    }
    return 0;
}

// cv::undistortPoints(pts, pts, K, D, Mat(), K) of the shim alone (pinned against cv2 goldens)
void ref_undistort_points(const float* pts, int n, const float* cam9, int ndist, float* out) {
    cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
    K.at<float>(0, 0) = cam9[0]; K.at<float>(1, 1) = cam9[1]; K.at<float>(0, 2) = cam9[2]; K.at<float>(1, 2) = cam9[3];
    cv::Mat D(ndist, 1, CV_32F);
    for (int i = 0; i < ndist; ++i) D.at<float>(i) = cam9[4 + i];
    cv::Mat m(n, 2, CV_32F);
    std::memcpy(m.data, pts, (size_t)n * 8);
    cv::undistortPoints(m, m, K, D, cv::Mat(), K);
    std::memcpy(out, m.data, (size_t)n * 8);
}

// ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, th, ORBdist)   ORBmatcher.cc:1731.  Identity pose as above:
// x3Dc == x3Dw == cam_xyz; the body's projection is returned in proj_uv.  state[i]: 0 = no map point, 1 = good, 2 = bad, 3 = already found.
int ref_search_by_projection_keyframe(float nnratio, int checkOri, const FrameView* cur, int n_kf, const float* cam_xyz, const int* predicted_level, const float* kf_angle,
                                      const unsigned char* mp_desc, const unsigned char* state, const float* min_dist, const float* max_dist,
                                      const unsigned char* cur_occupied, float th, int orb_dist, float fx, float fy, float cx, float cy, float* proj_uv, int* cur_match) {
    ArenaScope scope;
    int nm;
    {
        Frame C; fill_frame(C, cur);
        Frame::fx = fx; Frame::fy = fy; Frame::cx = cx; Frame::cy = cy;
        C.mTcw = cv::Mat::eye(4, 4, CV_32F);
        std::vector<MapPoint> pts(n_kf), occ(cur->n);
        KeyFrame K; K.mvKeysUn.resize(n_kf); K.mvpMapPoints.assign(n_kf, (MapPoint*)NULL);
        std::set<MapPoint*> found;
        for (int i = 0; i < n_kf; ++i) {
            K.mvKeysUn[i].angle = kf_angle[i];
            pts[i].pos = cv::Mat(3, 1, CV_32F); for (int k = 0; k < 3; ++k) pts[i].pos.at<float>(k) = cam_xyz[3 * i + k];
            pts[i].desc = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + (size_t)i * 32)).clone();
            pts[i].plevel = predicted_level[i]; pts[i].mind = min_dist[i]; pts[i].maxd = max_dist[i];
            if (state[i]) K.mvpMapPoints[i] = &pts[i];
            if (state[i] == 2) pts[i].bad = true;
            if (state[i] == 3) found.insert(&pts[i]);
            const float xc = cam_xyz[3 * i], yc = cam_xyz[3 * i + 1];
            const float invzc = 1.0 / cam_xyz[3 * i + 2];
            proj_uv[2 * i] = fx * xc * invzc + cx; proj_uv[2 * i + 1] = fy * yc * invzc + cy;
        }
        for (int j = 0; j < cur->n; ++j) if (cur_occupied && cur_occupied[j]) C.mvpMapPoints[j] = &occ[j];
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchByProjection(C, &K, found, th, orb_dist);
        for (int j = 0; j < cur->n; ++j) {
            MapPoint* p = C.mvpMapPoints[j];
            cur_match[j] = (p && p >= &pts[0] && p < &pts[0] + n_kf) ? (int)(p - &pts[0]) : -1;
        }
    }
    return nm;
}

// ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)   ORBmatcher.cc:388 (KeyFrame::GetFeaturesInArea / IsInImage from
// src/KeyFrame.cc:752-804).  Scw = identity, so p3Dc == p3Dw == cam_xyz.  state[p]: 1 good, 2 bad, 3 already in vpMatched (placed at
// KeyFrame feature found_at[p]); facing[p] != 0: the point's normal looks at the camera (passes the 60-degree gate), else away from it.
int ref_search_by_projection_keyframe_points(float nnratio, int checkOri, const FrameView* kfv, int n_points, const float* cam_xyz, const int* predicted_level,
                                             const unsigned char* mp_desc, const unsigned char* state, const int* found_at, const unsigned char* facing,
                                             const float* min_dist, const float* max_dist, const unsigned char* kf_matched, int th,
                                             float fx, float fy, float cx, float cy, float* proj_uv, int* kf_match) {
    ArenaScope scope;
    int nm;
    {
        Frame G; fill_frame(G, kfv);                                   // builds the 64 x 48 grid the KeyFrame constructor copies (src/KeyFrame.cc:58-66)
        KeyFrame K; K.N = kfv->n; K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy;
        K.mvKeysUn = G.mvKeysUn; K.mDescriptors = G.mDescriptors; K.mvScaleFactors = G.mvScaleFactors;
        K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
        K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
        K.mGrid.resize(FRAME_GRID_COLS);
        for (int i = 0; i < FRAME_GRID_COLS; ++i) { K.mGrid[i].resize(FRAME_GRID_ROWS); for (int j = 0; j < FRAME_GRID_ROWS; ++j) K.mGrid[i][j] = G.mGrid[i][j]; }
        std::vector<MapPoint> pts(n_points), old(kfv->n);
        std::vector<MapPoint*> vp(n_points), matched(kfv->n, (MapPoint*)NULL);
        for (int j = 0; j < kfv->n; ++j) if (kf_matched && kf_matched[j]) matched[j] = &old[j];
        for (int p = 0; p < n_points; ++p) {
            pts[p].pos = cv::Mat(3, 1, CV_32F); pts[p].normal = cv::Mat(3, 1, CV_32F);
            const float x = cam_xyz[3 * p], y = cam_xyz[3 * p + 1], z = cam_xyz[3 * p + 2];
            const float len = std::sqrt(x * x + y * y + z * z), sgn = facing[p] ? 1.f : -1.f;
            pts[p].pos.at<float>(0) = x; pts[p].pos.at<float>(1) = y; pts[p].pos.at<float>(2) = z;
            pts[p].normal.at<float>(0) = sgn * x / len; pts[p].normal.at<float>(1) = sgn * y / len; pts[p].normal.at<float>(2) = sgn * z / len;
            pts[p].desc = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + (size_t)p * 32)).clone();
            pts[p].plevel = predicted_level[p]; pts[p].mind = min_dist[p]; pts[p].maxd = max_dist[p];
            pts[p].bad = state[p] == 2;
            if (state[p] == 3) matched[found_at[p]] = &pts[p];
            vp[p] = &pts[p];
            const float invz = 1 / z;
            proj_uv[2 * p] = fx * (x * invz) + cx; proj_uv[2 * p + 1] = fy * (y * invz) + cy;          // :417-421
        }
        const std::vector<MapPoint*> before = matched;
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchByProjection(&K, cv::Mat::eye(4, 4, CV_32F), vp, matched, th);
        for (int j = 0; j < kfv->n; ++j) kf_match[j] = (matched[j] != before[j] && matched[j]) ? (int)(matched[j] - &pts[0]) : -1;
    }
    return nm;
}

// ORBmatcher::SearchBySim3   ORBmatcher.cc:1290.  Both KeyFrames at the world origin, s12 = 1, R12 = I, t12 = 0, so every map point's camera
// coordinates in either KeyFrame are its world coordinates (xyz1 / xyz2).  state: 0 no map point, 1 good, 2 bad; already12[i] >= 0: vpMatches12[i]
// holds, on entry, the map point of pKF2's feature already12[i] (=> vbAlreadyMatched1[i] and vbAlreadyMatched2[already12[i]]).
namespace {
void fill_keyframe(KeyFrame& K, Frame& G, const FrameView* v, float fx, float fy, float cx, float cy) {
    fill_frame(G, v);
    K.N = v->n; K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy;
    K.mvKeysUn = G.mvKeysUn; K.mDescriptors = G.mDescriptors; K.mvScaleFactors = G.mvScaleFactors;
    K.mnMinX = (int)Frame::mnMinX; K.mnMinY = (int)Frame::mnMinY; K.mnMaxX = (int)Frame::mnMaxX; K.mnMaxY = (int)Frame::mnMaxY;
    K.mfGridElementWidthInv = Frame::mfGridElementWidthInv; K.mfGridElementHeightInv = Frame::mfGridElementHeightInv;
    K.mGrid.resize(FRAME_GRID_COLS);
    for (int i = 0; i < FRAME_GRID_COLS; ++i) { K.mGrid[i].resize(FRAME_GRID_ROWS); for (int j = 0; j < FRAME_GRID_ROWS; ++j) K.mGrid[i][j] = G.mGrid[i][j]; }
}
void fill_points(std::vector<MapPoint>& pts, KeyFrame& K, int n, const float* xyz, const int* level, const unsigned char* desc, const unsigned char* state, const float* mind, const float* maxd) {
    K.mvpMapPoints.assign(n, (MapPoint*)NULL);
    for (int i = 0; i < n; ++i) {
        pts[i].pos = cv::Mat(3, 1, CV_32F); for (int k = 0; k < 3; ++k) pts[i].pos.at<float>(k) = xyz[3 * i + k];
        pts[i].desc = cv::Mat(1, 32, CV_8U, (void*)(desc + (size_t)i * 32)).clone();
        pts[i].plevel = level[i]; pts[i].mind = mind[i]; pts[i].maxd = maxd[i]; pts[i].bad = state[i] == 2;
        pts[i].in_kf = &K; pts[i].in_idx = i;
        if (state[i]) K.mvpMapPoints[i] = &pts[i];
    }
}
}
int ref_search_by_sim3(const FrameView* v1, const FrameView* v2, const float* xyz1, const int* level1, const unsigned char* desc1, const unsigned char* state1,
                       const float* mind1, const float* maxd1, const float* xyz2, const int* level2, const unsigned char* desc2, const unsigned char* state2,
                       const float* mind2, const float* maxd2, const int* already12, float th, float fx, float fy, float cx, float cy, int* match12) {
    ArenaScope scope;
    int nf;
    {
        Frame G1, G2; KeyFrame K1, K2;
        fill_keyframe(K1, G1, v1, fx, fy, cx, cy); fill_keyframe(K2, G2, v2, fx, fy, cx, cy);
        std::vector<MapPoint> p1(v1->n), p2(v2->n);
        fill_points(p1, K1, v1->n, xyz1, level1, desc1, state1, mind1, maxd1);
        fill_points(p2, K2, v2->n, xyz2, level2, desc2, state2, mind2, maxd2);
        std::vector<MapPoint*> m12(v1->n, (MapPoint*)NULL);
        for (int i = 0; i < v1->n; ++i) if (already12[i] >= 0) m12[i] = &p2[already12[i]];
        const std::vector<MapPoint*> before = m12;
        ORBmatcher matcher(0.6f, true);
        nf = matcher.SearchBySim3(&K1, &K2, m12, 1.0f, cv::Mat::eye(3, 3, CV_32F), cv::Mat::zeros(3, 1, CV_32F), th);
        for (int i = 0; i < v1->n; ++i) match12[i] = (m12[i] && m12[i] != before[i]) ? (int)(m12[i] - &p2[0]) : -1;
    }
    return nf;
}

// ORBmatcher::SearchForTriangulation   ORBmatcher.cc:810 (+ CheckDistEpipolarLine :188).  KeyFrame 1 at the world origin, KeyFrame 2 with
// rotation I and translation c2 (= camera 1's centre in camera 2, from which the body derives the epipole).  has_mp: the feature holds a
// map point.  Output match12[i] = second of the pair (i, .) in vMatchedPairs, -1 if absent; epi[2] = the epipole the body computed.
int ref_search_for_triangulation(float nnratio, int checkOri, int n1, const cv::KeyPoint* keys1, const unsigned char* desc1, const unsigned char* has_mp1, const float* ur1,
                                 int n_fv1, const int* fv1_nodes, const int* fv1_offsets, const int* fv1_idx,
                                 int n2, const cv::KeyPoint* keys2, const unsigned char* desc2, const unsigned char* has_mp2, const float* ur2,
                                 int n_fv2, const int* fv2_nodes, const int* fv2_offsets, const int* fv2_idx,
                                 const float* F12, const float* c2, float fx, float fy, float cx, float cy, int nlevels, const float* scale_factors, const float* level_sigma2,
                                 int only_stereo, int* match12, float* epi) {
    ArenaScope scope;
    int nm;
    {
        std::vector<MapPoint> p1(n1), p2(n2);
        KeyFrame K1, K2;
        K1.N = n1; K1.mvKeysUn.assign(keys1, keys1 + n1); K1.mDescriptors = cv::Mat(n1, 32, CV_8U, (void*)desc1).clone(); K1.mvuRight.assign(ur1, ur1 + n1);
        K2.N = n2; K2.mvKeysUn.assign(keys2, keys2 + n2); K2.mDescriptors = cv::Mat(n2, 32, CV_8U, (void*)desc2).clone(); K2.mvuRight.assign(ur2, ur2 + n2);
        K1.mvpMapPoints.assign(n1, (MapPoint*)NULL); K2.mvpMapPoints.assign(n2, (MapPoint*)NULL);
        for (int i = 0; i < n1; ++i) if (has_mp1[i]) K1.mvpMapPoints[i] = &p1[i];
        for (int j = 0; j < n2; ++j) if (has_mp2[j]) K2.mvpMapPoints[j] = &p2[j];
        fill_fv(K1.mFeatVec, n_fv1, fv1_nodes, fv1_offsets, fv1_idx); fill_fv(K2.mFeatVec, n_fv2, fv2_nodes, fv2_offsets, fv2_idx);
        K2.fx = fx; K2.fy = fy; K2.cx = cx; K2.cy = cy;
        K2.mvScaleFactors.assign(scale_factors, scale_factors + nlevels); K2.mvLevelSigma2.assign(level_sigma2, level_sigma2 + nlevels);
        K2.tcw = cv::Mat(3, 1, CV_32F); for (int k = 0; k < 3; ++k) K2.tcw.at<float>(k) = c2[k];
        cv::Mat F(3, 3, CV_32F); for (int k = 0; k < 9; ++k) F.at<float>(k / 3, k % 3) = F12[k];
        const float invz = 1.0f / c2[2];                                 // the body's own epipole (:822-825), restated for the caller
        epi[0] = fx * c2[0] * invz + cx; epi[1] = fy * c2[1] * invz + cy;
        for (int i = 0; i < n1; ++i) match12[i] = -1;
        std::vector<std::pair<size_t, size_t> > pairs;
        ORBmatcher matcher(nnratio, checkOri != 0);
        nm = matcher.SearchForTriangulation(&K1, &K2, F, pairs, only_stereo != 0);
        for (size_t q = 0; q < pairs.size(); ++q) match12[pairs[q].first] = (int)pairs[q].second;
    }
    return nm;
}

// ORBmatcher::Fuse, both forms (ORBmatcher.cc:1020, 1179), KeyFrame at the world origin / Scw = identity.  The KeyFrame's own map points
// (kf_has_mp) are good and never in vpMapPoints; state[p]: 1 good, 2 bad, 3 already observed in the KeyFrame; facing as above.
// Output best_idx[p] = the KeyFrame feature the body fused point p with (from the recorded Replace / AddObservation / vpReplacePoint), -1 = none.
int ref_fuse(int sim3_form, const FrameView* kfv, const float* kf_uright, const unsigned char* kf_has_mp, int n_points, const float* cam_xyz, const int* predicted_level,
             const unsigned char* mp_desc, const unsigned char* state, const unsigned char* facing, const float* min_dist, const float* max_dist,
             const float* inv_level_sigma2, float th, float bf, float fx, float fy, float cx, float cy, float* proj_uv, float* proj_ur, int* best_idx) {
    ArenaScope scope;
    int nf;
    {
        Frame G; KeyFrame K;
        fill_keyframe(K, G, kfv, fx, fy, cx, cy);
        K.mbf = bf; K.mvuRight.assign(kf_uright, kf_uright + kfv->n); K.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + kfv->nlevels);
        std::vector<MapPoint> own(kfv->n), pts(n_points);
        K.mvpMapPoints.assign(kfv->n, (MapPoint*)NULL);
        for (int j = 0; j < kfv->n; ++j) if (kf_has_mp[j]) { K.mvpMapPoints[j] = &own[j]; own[j].in_kf = &K; own[j].in_idx = j; own[j].nobs = 1 + j % 3; }   // with the points' 1 + p % 4 both Replace directions occur
        std::vector<MapPoint*> vp(n_points);
        for (int p = 0; p < n_points; ++p) {
            const float x = cam_xyz[3 * p], y = cam_xyz[3 * p + 1], z = cam_xyz[3 * p + 2];
            pts[p].pos = cv::Mat(3, 1, CV_32F); pts[p].normal = cv::Mat(3, 1, CV_32F);
            const float len = std::sqrt(x * x + y * y + z * z), sgn = facing[p] ? 1.f : -1.f;
            pts[p].pos.at<float>(0) = x; pts[p].pos.at<float>(1) = y; pts[p].pos.at<float>(2) = z;
            pts[p].normal.at<float>(0) = sgn * x / len; pts[p].normal.at<float>(1) = sgn * y / len; pts[p].normal.at<float>(2) = sgn * z / len;
            pts[p].desc = cv::Mat(1, 32, CV_8U, (void*)(mp_desc + (size_t)p * 32)).clone();
            pts[p].plevel = predicted_level[p]; pts[p].mind = min_dist[p]; pts[p].maxd = max_dist[p]; pts[p].nobs = 1 + p % 4;
            pts[p].bad = state[p] == 2 || (sim3_form && state[p] == 3);    // "already in the KeyFrame" is a caller-side skip in both forms
            if (state[p] == 3) pts[p].in_kf = &K;
            vp[p] = &pts[p];
            const float invz = 1 / z;
            proj_uv[2 * p] = fx * (x * invz) + cx; proj_uv[2 * p + 1] = fy * (y * invz) + cy; proj_ur[p] = proj_uv[2 * p] - bf * invz;     // :1054-1065
            best_idx[p] = -1;
        }
        ORBmatcher matcher(0.6f, true);
        g_replace_log.clear();
        if (!sim3_form) {
            nf = matcher.Fuse(&K, vp, th);
            // which KeyFrame feature did each point meet?  AddObservation names it directly; a Replace pairs the point with the map point that
            // held the feature (a KeyFrame-owned one, or an earlier point that was added there)
            for (size_t q = 0; q < g_replace_log.size(); ++q) {
                MapPoint *a = g_replace_log[q].first, *b = g_replace_log[q].second;
                for (int pass = 0; pass < 2; ++pass, std::swap(a, b)) {
                    const bool a_is_point = a >= &pts[0] && a < &pts[0] + n_points;
                    if (!a_is_point || a->fused_idx >= 0 || best_idx[a - &pts[0]] >= 0) continue;
                    const bool b_is_own = kfv->n && b >= &own[0] && b < &own[0] + kfv->n;
                    const int j = b_is_own ? (int)(b - &own[0]) : ((b >= &pts[0] && b < &pts[0] + n_points) ? b->fused_idx : -1);
                    if (j >= 0) { best_idx[a - &pts[0]] = j; break; }
                }
            }
            for (int p = 0; p < n_points; ++p) if (pts[p].fused_idx >= 0) best_idx[p] = pts[p].fused_idx;
        } else {
            std::vector<MapPoint*> repl(n_points, (MapPoint*)NULL);
            nf = matcher.Fuse(&K, cv::Mat::eye(4, 4, CV_32F), vp, th, repl);
            for (int p = 0; p < n_points; ++p) {
                if (pts[p].fused_idx >= 0) best_idx[p] = pts[p].fused_idx;                       // AddObservation(pKF, bestIdx)  (:1299)
                else if (repl[p]) best_idx[p] = (repl[p] >= &pts[0] && repl[p] < &pts[0] + n_points) ? repl[p]->fused_idx : repl[p]->in_idx;   // vpReplacePoint (:1294)
            }
        }
    }
    return nf;
}

// MapPoint::ComputeDistinctiveDescriptors   MapPoint.cc:359.  Point p is observed in KeyFrames kf_of[offsets[p] .. offsets[p + 1]) (ascending, which
// is also the address order of the harness's KeyFrame array and therefore the iteration order of mObservations) at feature 0 of a one-row
// descriptor matrix each; kf_bad marks bad KeyFrames.  Output: chosen[p][32] = mDescriptor after the call (zeros if it was never set).
int ref_distinctive_descriptors(int n_points, const int* offsets, const unsigned char* obs_desc, int n_kf, const int* kf_of, const unsigned char* kf_bad, unsigned char* chosen) {
    ArenaScope scope;
    {
        const int total = offsets[n_points];
        // one KeyFrame object per observation slot keeps "descriptor row of this observation" simple: slot s lives in KeyFrame s
        std::vector<KeyFrame> kfs(total);
        for (int s = 0; s < total; ++s) { kfs[s].mDescriptors = cv::Mat(1, 32, CV_8U, (void*)(obs_desc + (size_t)s * 32)).clone(); kfs[s].kf_bad = kf_bad[kf_of[s]] != 0; }
        std::vector<MapPoint> pts(n_points);
        for (int p = 0; p < n_points; ++p) {
            for (int s = offsets[p]; s < offsets[p + 1]; ++s) pts[p].mObservations[&kfs[s]] = 0;
            pts[p].ComputeDistinctiveDescriptors();
            if (!pts[p].mDescriptor.empty()) std::memcpy(chosen + (size_t)p * 32, pts[p].mDescriptor.ptr(), 32); else std::memset(chosen + (size_t)p * 32, 0, 32);
        }
    }
    return 0;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)  ORBmatcher.cc:230  and  (KeyFrame*, KeyFrame*, ...)  :656.
// Feature vectors arrive flattened (nodes ascending, offsets, feature indices) as ref_voc_transform returns them.
// valid1 / valid2: the feature holds a map point that is not bad.  Outputs: match21[j] = side-1 index assigned to side-2
// feature j (KF x Frame: vpMapPointMatches[j] = KF's map point), match12[i] = side-2 index (KF x KF: vpMatches12[i]).
int ref_search_by_bow(float nnratio, int checkOri, int kf_kf, int n1, const cv::KeyPoint* keys1, const unsigned char* desc1, const unsigned char* valid1,
                      int n_fv1, const int* fv1_nodes, const int* fv1_offsets, const int* fv1_idx,
                      int n2, const cv::KeyPoint* keys2, const unsigned char* desc2, const unsigned char* valid2,
                      int n_fv2, const int* fv2_nodes, const int* fv2_offsets, const int* fv2_idx, int* match12, int* match21) {
    ArenaScope scope;
    int nm;
    {
        std::vector<MapPoint> p1(n1), p2(n2);
        KeyFrame K1; K1.mvKeysUn.assign(keys1, keys1 + n1); K1.mDescriptors = cv::Mat(n1, 32, CV_8U, (void*)desc1).clone();
        K1.mvpMapPoints.assign(n1, (MapPoint*)NULL);
        for (int i = 0; i < n1; ++i) if (valid1[i]) K1.mvpMapPoints[i] = &p1[i];
        fill_fv(K1.mFeatVec, n_fv1, fv1_nodes, fv1_offsets, fv1_idx);
        for (int i = 0; i < n1; ++i) match12[i] = -1;
        for (int j = 0; j < n2; ++j) match21[j] = -1;
        ORBmatcher matcher(nnratio, checkOri != 0);
        if (kf_kf) {
            KeyFrame K2; K2.mvKeysUn.assign(keys2, keys2 + n2); K2.mDescriptors = cv::Mat(n2, 32, CV_8U, (void*)desc2).clone();
            K2.mvpMapPoints.assign(n2, (MapPoint*)NULL);
            for (int j = 0; j < n2; ++j) if (valid2[j]) K2.mvpMapPoints[j] = &p2[j];
            fill_fv(K2.mFeatVec, n_fv2, fv2_nodes, fv2_offsets, fv2_idx);
            std::vector<MapPoint*> m12;
            nm = matcher.SearchByBoW(&K1, &K2, m12);
            for (int i = 0; i < n1; ++i) if (m12[i]) { match12[i] = (int)(m12[i] - &p2[0]); match21[match12[i]] = i; }
        } else {
            Frame F; F.N = n2; F.mvKeys.assign(keys2, keys2 + n2); F.mDescriptors = cv::Mat(n2, 32, CV_8U, (void*)desc2).clone();
            fill_fv(F.mFeatVec, n_fv2, fv2_nodes, fv2_offsets, fv2_idx);
            std::vector<MapPoint*> mf;
            nm = matcher.SearchByBoW(&K1, F, mf);
            for (int j = 0; j < n2; ++j) if (mf[j]) { match21[j] = (int)(mf[j] - &p1[0]); match12[match21[j]] = j; }
        }
    }
    return nm;
}

// Frame::ComputeStereoMatches   Frame.cc:1179.  left / right: handles from ref_extractor_create whose last
// ref_extract call processed the left / right image (their mvImagePyramid is read).
int ref_compute_stereo_matches(void* left, void* right, const cv::KeyPoint* keys_left, const unsigned char* desc_left, int nl,
                               const cv::KeyPoint* keys_right, const unsigned char* desc_right, int nr, float mb, float mbf,
                               float* u_right, float* depth) {
    ArenaScope scope;
    {
        Frame F;
        F.mpORBextractorLeft = (ORBextractor*)left; F.mpORBextractorRight = (ORBextractor*)right;
        F.N = nl; F.mb = mb; F.mbf = mbf;
        F.mvKeys.assign(keys_left, keys_left + nl); F.mvKeysRight.assign(keys_right, keys_right + nr);
        F.mDescriptors = cv::Mat(nl, 32, CV_8U, (void*)desc_left).clone(); F.mDescriptorsRight = cv::Mat(nr, 32, CV_8U, (void*)desc_right).clone();
        F.mvScaleFactors = F.mpORBextractorLeft->GetScaleFactors(); F.mvInvScaleFactors = F.mpORBextractorLeft->GetInverseScaleFactors();
        F.ComputeStereoMatches();
        for (int i = 0; i < nl; ++i) { u_right[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
    }
    return 0;
}

}  // extern "C"
