// ORBmatcher_b200.cc -- B200-native bodies of the ORBmatcher / Frame methods that lie on the hot path
// (SURVEY.md 8a rows a14-a18).  They are member definitions of the reference's OWN classes, declared in
// the reference's unchanged include/ORBmatcher.h and include/Frame.h, so Tracking / Initializer call
// sites stay as they are; a maintainer compiles this file and removes (or #ifdef's out) the four
// bodies it replaces in src/ORBmatcher.cc / src/Frame.cc:
//
//   ORBmatcher::SearchForInitialization(Frame&, Frame&, vector<Point2f>&, vector<int>&, int)   src/ORBmatcher.cc:515-643
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, float)                    src/ORBmatcher.cc:70-175
//   ORBmatcher::SearchByProjection(Frame&, const Frame&, float, bool)                          src/ORBmatcher.cc:1569-1728
//   Frame::ComputeStereoMatches()                                                               src/Frame.cc:1179-1573
//   ORBmatcher::SearchByProjection(Frame&, KeyFrame*, const set<MapPoint*>&, float, int)       src/ORBmatcher.cc:1731-1863  (8f rank 3)
//   ORBmatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, vector<MapPoint*>&, int)   src/ORBmatcher.cc:388-512  (8f rank 3)
//   ORBmatcher::SearchBySim3(KeyFrame*, KeyFrame*, vector<MapPoint*>&, const float&, const cv::Mat&, const cv::Mat&, float)   src/ORBmatcher.cc:1314-1555  (8f rank 3)
//   ORBmatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, float)  and  Fuse(KeyFrame*, cv::Mat Scw, const vector<MapPoint*>&, float, vector<MapPoint*>&)   src/ORBmatcher.cc:1020-1310  (8f rank 3)
//
// Each body only flattens the object graph (Frame / MapPoint) into the plain arrays of the C ABI
// (include/orbx_b200.h), calls the CUDA implementation and writes the results back into the same
// members the reference writes.  Descriptor distances, grid lookups, ratio tests, rotation histograms
// and the SAD refinement all run on the GPU; nothing is matched on the CPU here.
#include "ORBmatcher.h"
#include "../../include/orbx_b200.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be the 28-byte POD the C ABI uses");

namespace ORB_SLAM2
{
namespace
{

void check(int rc, const char* what) {
    if (rc != ORBX_OK) throw std::runtime_error(std::string(what) + ": " + orbx_last_error());
}

// ORBmatcher objects are short-lived locals in the reference (Tracking.cc:1492, 1744, ...): keep one device
// matcher per (thread, nnratio, checkOri) instead of creating a stream and scratch buffers per call.
struct MatcherCache {
    std::map<std::pair<float, bool>, orbx_matcher*> m;
    ~MatcherCache() { for (auto& kv : m) orbx_matcher_destroy(kv.second); }
    orbx_matcher* get(float nnratio, bool checkOri) {
        auto key = std::make_pair(nnratio, checkOri);
        auto it = m.find(key);
        if (it != m.end()) return it->second;
        const char* e = std::getenv("ORBX_DEVICE");
        orbx_matcher* h = nullptr;
        check(orbx_matcher_create(nnratio, checkOri ? 1 : 0, e ? std::atoi(e) : 0, &h), "orbx_matcher_create");
        m[key] = h;
        return h;
    }
};
thread_local MatcherCache t_matchers;

// what the matchers read from a Frame (Frame.h: mvKeysUn, mDescriptors, mvuRight, image bounds, grid constants)
struct FrameFlat {
    orbx_frame_view v;
    std::vector<unsigned char> desc;     // only used when mDescriptors is not continuous
    explicit FrameFlat(const Frame& F) {
        v.n = F.N;
        v.keys_un = reinterpret_cast<const orbx_keypoint*>(F.mvKeysUn.data());
        if (F.mDescriptors.isContinuous()) v.descriptors = F.mDescriptors.ptr();
        else {
            desc.resize((size_t)F.N * 32);
            for (int i = 0; i < F.N; ++i) std::memcpy(&desc[(size_t)i * 32], F.mDescriptors.ptr(i), 32);
            v.descriptors = desc.data();
        }
        v.u_right = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
        v.min_x = Frame::mnMinX; v.min_y = Frame::mnMinY; v.max_x = Frame::mnMaxX; v.max_y = Frame::mnMaxY;
        v.grid_element_width_inv = Frame::mfGridElementWidthInv; v.grid_element_height_inv = Frame::mfGridElementHeightInv;
        v.nlevels = (int)F.mvScaleFactors.size(); v.scale_factors = F.mvScaleFactors.data();
    }
};

// the device-resident frame built by Frame_b200.cc, when the Frame carries one that matches its current keypoints
#ifdef ORBX_FRAME_HAS_DEVICE
const orbx_frame* device_of(const Frame& F) {
    const orbx_frame* f = F.mpDeviceFrame.get();
    return (f && orbx_frame_size(f) == F.N) ? f : nullptr;
}
#else
const orbx_frame* device_of(const Frame&) { return nullptr; }
#endif

// mvpMapPoints[j] already holds a map point with Observations() > 0   (ORBmatcher.cc:124-126, 1658-1660)
std::vector<unsigned char> occupied_flags(const Frame& F) {
    std::vector<unsigned char> occ(F.N, 0);
    for (int j = 0; j < F.N; ++j) { MapPoint* p = F.mvpMapPoints[j]; if (p && p->Observations() > 0) occ[j] = 1; }
    return occ;
}

}  // namespace

int ORBmatcher::SearchForInitialization(Frame &F1, Frame &F2, std::vector<cv::Point2f> &vbPrevMatched, std::vector<int> &vnMatches12, int windowSize)
{
    vnMatches12.assign(F1.mvKeysUn.size(), -1);
    int nmatches = 0;
    static_assert(sizeof(cv::Point2f) == 8, "cv::Point2f must be two floats");
    const orbx_frame *d1 = device_of(F1), *d2 = device_of(F2);
    if (d1 && d2) {
        check(orbx_search_for_initialization_frames(t_matchers.get(mfNNratio, mbCheckOrientation), d1, d2,
                                                    vbPrevMatched.empty() ? nullptr : reinterpret_cast<float*>(vbPrevMatched.data()),
                                                    vnMatches12.data(), windowSize, &nmatches), "orbx_search_for_initialization_frames");
        return nmatches;
    }
    FrameFlat a(F1), b(F2);
    check(orbx_search_for_initialization(t_matchers.get(mfNNratio, mbCheckOrientation), &a.v, &b.v,
                                         vbPrevMatched.empty() ? nullptr : reinterpret_cast<float*>(vbPrevMatched.data()),
                                         vnMatches12.data(), windowSize, &nmatches), "orbx_search_for_initialization");
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame &F, const std::vector<MapPoint*> &vpMapPoints, const float th)
{
    // per map point that passes the reference's two filters (:82-86): projection, predicted level, viewing cosine, descriptor
    std::vector<MapPoint*> pts; pts.reserve(vpMapPoints.size());
    for (size_t i = 0; i < vpMapPoints.size(); ++i) { MapPoint* p = vpMapPoints[i]; if (p->mbTrackInView && !p->isBad()) pts.push_back(p); }
    const int n = (int)pts.size();
    std::vector<float> uv((size_t)n * 2), ur(n), vc(n);
    std::vector<int> lvl(n);
    std::vector<unsigned char> desc((size_t)n * 32), obs(n);
    for (int i = 0; i < n; ++i) {
        MapPoint* p = pts[i];
        uv[2 * i] = p->mTrackProjX; uv[2 * i + 1] = p->mTrackProjY; ur[i] = p->mTrackProjXR;
        lvl[i] = p->mnTrackScaleLevel; vc[i] = p->mTrackViewCos;
        const cv::Mat d = p->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(), 32);
        obs[i] = p->Observations() > 0;
    }
    std::vector<unsigned char> occ = occupied_flags(F);
    std::vector<int> fmatch(F.N, -1);
    int nmatches = 0;
    if (const orbx_frame* df = device_of(F))
        check(orbx_search_by_projection_points_dev(t_matchers.get(mfNNratio, mbCheckOrientation), df, n, uv.data(), ur.data(), lvl.data(), vc.data(),
                                                   desc.data(), obs.data(), occ.data(), th, fmatch.data(), &nmatches), "orbx_search_by_projection_points_dev");
    else {
        FrameFlat f(F);
        check(orbx_search_by_projection_points(t_matchers.get(mfNNratio, mbCheckOrientation), &f.v, n, uv.data(), ur.data(), lvl.data(), vc.data(),
                                               desc.data(), obs.data(), occ.data(), th, fmatch.data(), &nmatches), "orbx_search_by_projection_points");
    }
    for (int j = 0; j < F.N; ++j) if (fmatch[j] >= 0) F.mvpMapPoints[j] = pts[fmatch[j]];       // :168
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
{
    // the pose of the two frames decides forward / backward once per call (:1580-1592); the per-point projection (:1597-1623) runs on the
    // device from the flattened world positions -- with cv::gemm's float arithmetic, see orbx_search_by_projection_frame_pose
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat twc = -Rcw.t() * tcw;
    const cv::Mat Rlw = LastFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tlw = LastFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat tlc = Rlw * twc + tlw;
    const bool bForward = tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    const bool bBackward = -tlc.at<float>(2) > CurrentFrame.mb && !bMono;
    float R9[9], t3[3];
    for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) R9[3 * r + c] = Rcw.at<float>(r, c); t3[r] = tcw.at<float>(r); }

    const int n = LastFrame.N;
    std::vector<float> world((size_t)n * 3, 0.f), angle(n, 0.f);
    std::vector<int> octave(n, 0);
    std::vector<unsigned char> desc((size_t)n * 32, 0), has(n, 0), obs(n, 0);
    for (int i = 0; i < n; ++i) {
        MapPoint* pMP = LastFrame.mvpMapPoints[i];
        if (!pMP || LastFrame.mvbOutlier[i]) continue;
        const cv::Mat x3Dw = pMP->GetWorldPos();
        world[3 * i] = x3Dw.at<float>(0); world[3 * i + 1] = x3Dw.at<float>(1); world[3 * i + 2] = x3Dw.at<float>(2);
        octave[i] = LastFrame.mvKeys[i].octave; angle[i] = LastFrame.mvKeysUn[i].angle;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(), 32);
        obs[i] = pMP->Observations() > 0;
        has[i] = 1;
    }
    std::vector<unsigned char> occ = occupied_flags(CurrentFrame);
    std::vector<int> cmatch(CurrentFrame.N, -1);
    int nmatches = 0;
    const orbx_frame* dc = device_of(CurrentFrame);
    FrameFlat c(CurrentFrame);                                                           // (cheap when the device frame is used: only consulted if dc == NULL)
    check(orbx_search_by_projection_frame_pose(t_matchers.get(mfNNratio, mbCheckOrientation), dc ? NULL : &c.v, dc, n, world.data(), has.data(), R9, t3,
                                               CurrentFrame.fx, CurrentFrame.fy, CurrentFrame.cx, CurrentFrame.cy, octave.data(), angle.data(), desc.data(), obs.data(), occ.data(),
                                               th, bForward ? 1 : 0, bBackward ? 1 : 0, CurrentFrame.mbf, cmatch.data(), &nmatches, NULL, NULL, NULL),
          "orbx_search_by_projection_frame_pose");
    for (int j = 0; j < CurrentFrame.N; ++j) {
        if (cmatch[j] >= 0) CurrentFrame.mvpMapPoints[j] = LastFrame.mvpMapPoints[cmatch[j]];   // :1685
        else if (cmatch[j] == -2) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint*>(NULL); // :1719
    }
    return nmatches;
}

int ORBmatcher::SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, const std::set<MapPoint*> &sAlreadyFound, const float th, const int ORBdist)
{
    // pose algebra, scale prediction and the set lookup stay on the host: they need the MapPoint objects and are O(N) (:1737-1789)
    const cv::Mat Rcw = CurrentFrame.mTcw.rowRange(0, 3).colRange(0, 3);
    const cv::Mat tcw = CurrentFrame.mTcw.rowRange(0, 3).col(3);
    const cv::Mat Ow = -Rcw.t() * tcw;
    const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
    const int n = (int)vpMPs.size();
    std::vector<float> uv((size_t)n * 2, 0.f), angle(n, 0.f);
    std::vector<int> level(n, 0);
    std::vector<unsigned char> desc((size_t)n * 32, 0), valid(n, 0);
    for (int i = 0; i < n; ++i) {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
        cv::Mat x3Dw = pMP->GetWorldPos();
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        const float u = CurrentFrame.fx * xc * invzc + CurrentFrame.cx;
        const float v = CurrentFrame.fy * yc * invzc + CurrentFrame.cy;
        if (u < CurrentFrame.mnMinX || u > CurrentFrame.mnMaxX || v < CurrentFrame.mnMinY || v > CurrentFrame.mnMaxY) continue;
        cv::Mat PO = x3Dw - Ow;
        float dist3D = cv::norm(PO);
        if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
        uv[2 * i] = u; uv[2 * i + 1] = v;
        level[i] = pMP->PredictScale(dist3D, &CurrentFrame);
        angle[i] = pKF->mvKeysUn[i].angle;
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(), 32);
        valid[i] = 1;
    }
    std::vector<unsigned char> occ(CurrentFrame.N, 0);
    for (int j = 0; j < CurrentFrame.N; ++j) if (CurrentFrame.mvpMapPoints[j]) occ[j] = 1;                                    // :1808: any map point blocks
    std::vector<int> cmatch(CurrentFrame.N ? CurrentFrame.N : 1, -1);
    int nmatches = 0;
    if (const orbx_frame* dc = device_of(CurrentFrame))
        check(orbx_search_by_projection_keyframe_dev(t_matchers.get(mfNNratio, mbCheckOrientation), dc, n, uv.data(), level.data(), angle.data(), desc.data(), valid.data(),
                                                     occ.data(), th, ORBdist, cmatch.data(), &nmatches), "orbx_search_by_projection_keyframe_dev");
    else {
        FrameFlat c(CurrentFrame);
        check(orbx_search_by_projection_keyframe(t_matchers.get(mfNNratio, mbCheckOrientation), &c.v, n, uv.data(), level.data(), angle.data(), desc.data(), valid.data(),
                                                 occ.data(), th, ORBdist, cmatch.data(), &nmatches), "orbx_search_by_projection_keyframe");
    }
    for (int j = 0; j < CurrentFrame.N; ++j) {
        if (cmatch[j] >= 0) CurrentFrame.mvpMapPoints[j] = vpMPs[cmatch[j]];                                                 // :1822
        else if (cmatch[j] == -2) CurrentFrame.mvpMapPoints[j] = static_cast<MapPoint*>(NULL);                               // :1853
    }
    return nmatches;
}

int ORBmatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*> &vpPoints, std::vector<MapPoint*> &vpMatched, int th)
{
    // Sim3 decomposition, projection, distance / viewing-angle gates and scale prediction stay on the host (:391-446)
    const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());
    spAlreadyFound.erase(static_cast<MapPoint*>(NULL));
    const int n = (int)vpPoints.size();
    std::vector<float> uv((size_t)n * 2, 0.f);
    std::vector<int> level(n, 0);
    std::vector<unsigned char> desc((size_t)n * 32, 0), valid(n, 0);
    for (int i = 0; i < n; ++i) {
        MapPoint* pMP = vpPoints[i];
        if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
        cv::Mat p3Dw = pMP->GetWorldPos();
        cv::Mat p3Dc = Rcw * p3Dw + tcw;
        if (p3Dc.at<float>(2) < 0.0) continue;
        const float invz = 1 / p3Dc.at<float>(2);
        const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
        const float u = fx * x + cx, v = fy * y + cy;
        if (!pKF->IsInImage(u, v)) continue;
        cv::Mat PO = p3Dw - Ow;
        const float dist = cv::norm(PO);
        if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
        cv::Mat Pn = pMP->GetNormal();
        if (PO.dot(Pn) < 0.5 * dist) continue;
        uv[2 * i] = u; uv[2 * i + 1] = v;
        level[i] = pMP->PredictScale(dist, pKF);
        const cv::Mat d = pMP->GetDescriptor();
        std::memcpy(&desc[(size_t)i * 32], d.ptr(), 32);
        valid[i] = 1;
    }
    const int nk = (int)pKF->mvKeysUn.size();
    std::vector<unsigned char> taken(nk, 0);
    for (int j = 0; j < nk; ++j) if (vpMatched[j]) taken[j] = 1;                                                            // :459
    // the KeyFrame as a frame view: its grid is the Frame's (src/KeyFrame.cc:58-66), its bounds are the integer members (include/KeyFrame.h:408-411)
    orbx_frame_view kv;
    std::vector<unsigned char> dtmp;
    kv.n = nk; kv.keys_un = reinterpret_cast<const orbx_keypoint*>(pKF->mvKeysUn.data());
    if (pKF->mDescriptors.isContinuous()) kv.descriptors = pKF->mDescriptors.ptr();
    else { dtmp.resize((size_t)nk * 32); for (int j = 0; j < nk; ++j) std::memcpy(&dtmp[(size_t)j * 32], pKF->mDescriptors.ptr(j), 32); kv.descriptors = dtmp.data(); }
    kv.u_right = NULL;
    kv.min_x = (float)pKF->mnMinX; kv.min_y = (float)pKF->mnMinY; kv.max_x = (float)pKF->mnMaxX; kv.max_y = (float)pKF->mnMaxY;
    kv.grid_element_width_inv = pKF->mfGridElementWidthInv; kv.grid_element_height_inv = pKF->mfGridElementHeightInv;
    kv.nlevels = (int)pKF->mvScaleFactors.size(); kv.scale_factors = pKF->mvScaleFactors.data();
    std::vector<int> kmatch(nk ? nk : 1, -1);
    int nmatches = 0;
    check(orbx_search_by_projection_keyframe_points(t_matchers.get(mfNNratio, mbCheckOrientation), &kv, n, uv.data(), level.data(), desc.data(), valid.data(), taken.data(),
                                                    (float)th, kmatch.data(), &nmatches), "orbx_search_by_projection_keyframe_points");
    for (int j = 0; j < nk; ++j) if (kmatch[j] >= 0) vpMatched[j] = vpPoints[kmatch[j]];                                    // :483
    return nmatches;
}

namespace
{
// a KeyFrame as the frame view the C ABI takes: its grid is the Frame's (src/KeyFrame.cc:58-66), its bounds are the integer members (include/KeyFrame.h:408-411)
struct KeyFrameFlat {
    orbx_frame_view v; std::vector<unsigned char> desc;
    explicit KeyFrameFlat(KeyFrame* pKF) {
        v.n = (int)pKF->mvKeysUn.size(); v.keys_un = reinterpret_cast<const orbx_keypoint*>(pKF->mvKeysUn.data());
        if (pKF->mDescriptors.isContinuous()) v.descriptors = pKF->mDescriptors.ptr();
        else { desc.resize((size_t)v.n * 32); for (int j = 0; j < v.n; ++j) std::memcpy(&desc[(size_t)j * 32], pKF->mDescriptors.ptr(j), 32); v.descriptors = desc.data(); }
        v.u_right = NULL;
        v.min_x = (float)pKF->mnMinX; v.min_y = (float)pKF->mnMinY; v.max_x = (float)pKF->mnMaxX; v.max_y = (float)pKF->mnMaxY;
        v.grid_element_width_inv = pKF->mfGridElementWidthInv; v.grid_element_height_inv = pKF->mfGridElementHeightInv;
        v.nlevels = (int)pKF->mvScaleFactors.size(); v.scale_factors = pKF->mvScaleFactors.data();
    }
};
// one direction of SearchBySim3 (:1367-1398 / :1443-1480): which map points of `from` are searched in `to`, and where
struct Sim3Queries {
    std::vector<float> uv; std::vector<int> level; std::vector<unsigned char> desc, valid;
    Sim3Queries(const std::vector<MapPoint*>& mps, const std::vector<bool>& already, const cv::Mat& Rfw, const cv::Mat& tfw, const cv::Mat& sR, const cv::Mat& t,
                KeyFrame* to, float fx, float fy, float cx, float cy) {
        const int n = (int)mps.size();
        uv.assign((size_t)n * 2, 0.f); level.assign(n, 0); desc.assign((size_t)n * 32, 0); valid.assign(n, 0);
        for (int i = 0; i < n; ++i) {
            MapPoint* pMP = mps[i];
            if (!pMP || already[i] || pMP->isBad()) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dcf = Rfw * p3Dw + tfw;
            cv::Mat p3Dct = sR * p3Dcf + t;
            if (p3Dct.at<float>(2) < 0.0) continue;
            const float invz = 1.0 / p3Dct.at<float>(2);
            const float x = p3Dct.at<float>(0) * invz, y = p3Dct.at<float>(1) * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            if (!to->IsInImage(u, v)) continue;
            const float dist3D = cv::norm(p3Dct);
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            uv[2 * i] = u; uv[2 * i + 1] = v; level[i] = pMP->PredictScale(dist3D, to);
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(&desc[(size_t)i * 32], d.ptr(), 32);
            valid[i] = 1;
        }
    }
};
}  // namespace

int ORBmatcher::SearchBySim3(KeyFrame *pKF1, KeyFrame *pKF2, std::vector<MapPoint*> &vpMatches12, const float &s12, const cv::Mat &R12, const cv::Mat &t12, const float th)
{
    const float &fx = pKF1->fx, &fy = pKF1->fy, &cx = pKF1->cx, &cy = pKF1->cy;
    cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation(), R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
    cv::Mat sR12 = s12 * R12;
    cv::Mat sR21 = (1.0 / s12) * R12.t();
    cv::Mat t21 = -sR21 * t12;
    const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
    const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
    std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
    for (int i = 0; i < N1; i++) {                                                                                          // :1344-1355
        MapPoint* pMP = vpMatches12[i];
        if (pMP) { vbAlreadyMatched1[i] = true; const int idx2 = pMP->GetIndexInKeyFrame(pKF2); if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true; }
    }
    Sim3Queries q1(vpMapPoints1, vbAlreadyMatched1, R1w, t1w, sR21, t21, pKF2, fx, fy, cx, cy), q2(vpMapPoints2, vbAlreadyMatched2, R2w, t2w, sR12, t12, pKF1, fx, fy, cx, cy);
    KeyFrameFlat k1(pKF1), k2(pKF2);
    std::vector<int> m12(N1 ? N1 : 1, -1);
    int nFound = 0;
    check(orbx_search_by_sim3(t_matchers.get(mfNNratio, mbCheckOrientation), &k1.v, &k2.v, q1.uv.data(), q1.level.data(), q1.desc.data(), q1.valid.data(),
                              q2.uv.data(), q2.level.data(), q2.desc.data(), q2.valid.data(), th, m12.data(), &nFound), "orbx_search_by_sim3");
    for (int i1 = 0; i1 < N1; ++i1) if (m12[i1] >= 0) vpMatches12[i1] = vpMapPoints2[m12[i1]];                               // :1544
    return nFound;
}

namespace
{
// caller-side part of both Fuse forms (:1036-1085 / :1200-1243): projection, gates, predicted level; returns the queries of orbx_fuse_search
struct FuseQueries {
    std::vector<float> uv, ur; std::vector<int> level; std::vector<unsigned char> desc, valid;
    FuseQueries(KeyFrame* pKF, const std::vector<MapPoint*>& pts, const cv::Mat& Rcw, const cv::Mat& tcw, const cv::Mat& Ow, const std::set<MapPoint*>* found, bool pose_form) {
        const float &fx = pKF->fx, &fy = pKF->fy, &cx = pKF->cx, &cy = pKF->cy;
        const int n = (int)pts.size();
        uv.assign((size_t)n * 2, 0.f); ur.assign(n, 0.f); level.assign(n, 0); desc.assign((size_t)n * 32, 0); valid.assign(n, 0);
        for (int i = 0; i < n; ++i) {
            MapPoint* pMP = pts[i];
            if (!pMP || pMP->isBad()) continue;
            if (pose_form ? pMP->IsInKeyFrame(pKF) : (found->count(pMP) != 0)) continue;
            cv::Mat p3Dw = pMP->GetWorldPos();
            cv::Mat p3Dc = Rcw * p3Dw + tcw;
            if (p3Dc.at<float>(2) < 0.0f) continue;
            const float invz = pose_form ? 1 / p3Dc.at<float>(2) : (float)(1.0 / p3Dc.at<float>(2));
            const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
            const float u = fx * x + cx, v = fy * y + cy;
            if (!pKF->IsInImage(u, v)) continue;
            cv::Mat PO = p3Dw - Ow;
            const float dist3D = cv::norm(PO);
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            cv::Mat Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist3D) continue;
            uv[2 * i] = u; uv[2 * i + 1] = v; if (pose_form) ur[i] = u - pKF->mbf * invz;
            level[i] = pMP->PredictScale(dist3D, pKF);
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(&desc[(size_t)i * 32], d.ptr(), 32);
            valid[i] = 1;
        }
    }
};
}  // namespace

int ORBmatcher::Fuse(KeyFrame *pKF, const std::vector<MapPoint *> &vpMapPoints, const float th)
{
    cv::Mat Rcw = pKF->GetRotation(), tcw = pKF->GetTranslation(), Ow = pKF->GetCameraCenter();
    const int nMPs = (int)vpMapPoints.size();
    FuseQueries q(pKF, vpMapPoints, Rcw, tcw, Ow, NULL, true);
    KeyFrameFlat k(pKF);
    k.v.u_right = pKF->mvuRight.empty() ? NULL : pKF->mvuRight.data();
    std::vector<int> best(nMPs ? nMPs : 1, -1);
    check(orbx_fuse_search(t_matchers.get(mfNNratio, mbCheckOrientation), &k.v, nMPs, q.uv.data(), q.ur.data(), q.level.data(), q.desc.data(), q.valid.data(),
                           pKF->mvInvLevelSigma2.data(), th, best.data()), "orbx_fuse_search");
    // the map surgery of :1143-1172, point by point in the reference's order; a point touched by an earlier Replace is re-checked as the reference would see it
    int nFused = 0;
    for (int i = 0; i < nMPs; ++i) {
        if (best[i] < 0) continue;
        MapPoint* pMP = vpMapPoints[i];
        if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;
        MapPoint* pMPinKF = pKF->GetMapPoint(best[i]);
        if (pMPinKF) {
            if (!pMPinKF->isBad()) {
                if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                else pMPinKF->Replace(pMP);
            }
        } else {
            pMP->AddObservation(pKF, best[i]);
            pKF->AddMapPoint(pMP, best[i]);
        }
        nFused++;
    }
    return nFused;
}

int ORBmatcher::Fuse(KeyFrame *pKF, cv::Mat Scw, const std::vector<MapPoint *> &vpPoints, float th, std::vector<MapPoint *> &vpReplacePoint)
{
    cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
    const float scw = sqrt(sRcw.row(0).dot(sRcw.row(0)));
    cv::Mat Rcw = sRcw / scw;
    cv::Mat tcw = Scw.rowRange(0, 3).col(3) / scw;
    cv::Mat Ow = -Rcw.t() * tcw;
    const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();
    const int nPoints = (int)vpPoints.size();
    FuseQueries q(pKF, vpPoints, Rcw, tcw, Ow, &spAlreadyFound, false);
    KeyFrameFlat k(pKF);
    std::vector<int> best(nPoints ? nPoints : 1, -1);
    check(orbx_fuse_search(t_matchers.get(mfNNratio, mbCheckOrientation), &k.v, nPoints, q.uv.data(), NULL, q.level.data(), q.desc.data(), q.valid.data(), NULL, th, best.data()),
          "orbx_fuse_search");
    int nFused = 0;
    for (int iMP = 0; iMP < nPoints; ++iMP) {                                                                              // :1287-1305
        if (best[iMP] < 0) continue;
        MapPoint* pMP = vpPoints[iMP];
        MapPoint* pMPinKF = pKF->GetMapPoint(best[iMP]);
        if (pMPinKF) { if (!pMPinKF->isBad()) vpReplacePoint[iMP] = pMPinKF; }
        else { pMP->AddObservation(pKF, best[iMP]); pKF->AddMapPoint(pMP, best[iMP]); }
        nFused++;
    }
    return nFused;
}

void Frame::ComputeStereoMatches()
{
    mvuRight = std::vector<float>(N, -1.0f);                                                   // src/Frame.cc:1187-1188
    mvDepth = std::vector<float>(N, -1.0f);
    if (N == 0) return;
    // both pyramids are still resident in the two extractor handles (ExtractORB ran just before, Frame.cc:165-176)
    const unsigned char* dl = mDescriptors.ptr();
    const unsigned char* dr = mDescriptorsRight.ptr();
    std::vector<unsigned char> tl, tr;
    if (!mDescriptors.isContinuous()) { tl.resize((size_t)N * 32); for (int i = 0; i < N; ++i) std::memcpy(&tl[(size_t)i * 32], mDescriptors.ptr(i), 32); dl = tl.data(); }
    const int nr = (int)mvKeysRight.size();
    if (!mDescriptorsRight.isContinuous()) { tr.resize((size_t)nr * 32); for (int i = 0; i < nr; ++i) std::memcpy(&tr[(size_t)i * 32], mDescriptorsRight.ptr(i), 32); dr = tr.data(); }
    check(orbx_compute_stereo_matches(t_matchers.get(0.6f, true), mpORBextractorLeft->handle(), mpORBextractorRight->handle(),
                                      reinterpret_cast<const orbx_keypoint*>(mvKeys.data()), dl, N,
                                      reinterpret_cast<const orbx_keypoint*>(mvKeysRight.data()), dr, nr,
                                      mb, mbf, mvuRight.data(), mvDepth.data()), "orbx_compute_stereo_matches");
}

} //namespace ORB_SLAM2
