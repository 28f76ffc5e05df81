// tests/host/host_dropin_main.cc -- drives the drop-in C++ classes (amos-slam_b200/host) the way the reference's
// Frame / Tracking code drives its own ORBextractor / ORBmatcher, and dumps every result to a binary file that
// tests/test_gpu_host_dropin.py compares with the C-ABI results (themselves bit-exact against the oracle).
// cv:: types come from the OpenCV-free shim (oracle/shim) -- test infrastructure, not product.
//
// usage: host_dropin <in.bin> <out.bin> [vocabulary.txt]
//   in.bin : int32 w,h ; u8 A[h*w] ; u8 B[h*w] ; int32 sw,sh ; u8 L[sh*sw] ; u8 R[sh*sw] ; f32 cam[10] (fx fy cx cy k1 k2 p1 p2 k3 bf) ; f32 bounds[6]
#include "ORBextractor.h"
#include "ORBmatcher.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include <chrono>
#include <functional>

using namespace ORB_SLAM2;
namespace ORB_SLAM2 { void RegisterDeviceVocabulary(const ORBVocabulary* voc, const std::string& file); }   // amos-slam_b200/host/BoW_b200.cc

float Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv, Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY;
const int ORBmatcher::TH_HIGH = 100, ORBmatcher::TH_LOW = 50, ORBmatcher::HISTO_LENGTH = 30;

static FILE* g_out;
static void put(const void* p, size_t n) { fwrite(p, 1, n, g_out); }
static void put_i(int v) { put(&v, 4); }
static void put_kps(const std::vector<cv::KeyPoint>& k, const cv::Mat& d) {
    put_i((int)k.size());
    if (!k.empty()) put(k.data(), k.size() * sizeof(cv::KeyPoint));
    for (int i = 0; i < (int)k.size(); ++i) put(d.ptr(i), 32);
}
static cv::Mat read_img(FILE* f, int w, int h) {
    cv::Mat m(h, w, CV_8UC1);
    if (fread(m.ptr(), 1, (size_t)w * h, f) != (size_t)w * h) { fprintf(stderr, "short read\n"); exit(2); }
    return m;
}
static void fill_frame(Frame& F, ORBextractor& e, const std::vector<cv::KeyPoint>& k, const cv::Mat& d, int w, int h) {
    F.N = (int)k.size(); F.mvKeys = k; F.mvKeysUn = k; F.mDescriptors = d;
    F.mvpMapPoints.assign(F.N, (MapPoint*)NULL); F.mvbOutlier.assign(F.N, false);
    F.mvuRight.assign(F.N, -1.f); F.mvDepth.assign(F.N, -1.f);
    F.mvScaleFactors = e.GetScaleFactors(); F.mnScaleLevels = e.GetLevels();
    Frame::mnMinX = 0.f; Frame::mnMaxX = (float)w; Frame::mnMinY = 0.f; Frame::mnMaxY = (float)h;       // Frame.cc:1148-1151 (no distortion)
    Frame::mfGridElementWidthInv = 64.f / (Frame::mnMaxX - Frame::mnMinX);                              // Frame.cc:219-220
    Frame::mfGridElementHeightInv = 48.f / (Frame::mnMaxY - Frame::mnMinY);
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb"); g_out = fopen(argv[2], "wb");
    if (!f || !g_out) return 2;
    int w, h; if (fread(&w, 4, 1, f) != 1 || fread(&h, 4, 1, f) != 1) return 2;
    cv::Mat A = read_img(f, w, h), B = read_img(f, w, h);
    int sw, sh; if (fread(&sw, 4, 1, f) != 1 || fread(&sh, 4, 1, f) != 1) return 2;
    cv::Mat L = read_img(f, sw, sh), R = read_img(f, sw, sh);

    // ---- operator()(image, mask, keypoints, descriptors) as Frame::ExtractORB calls it (Frame.cc:464-472)
    ORBextractor ext(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> ka, kb; cv::Mat da, db;
    ext(A, cv::Mat(), ka, da);
    put_kps(ka, da);
    put_i(ext.GetLevels());
    for (int l = 0; l < ext.GetLevels(); ++l) {                          // mvImagePyramid: ROI views inside 19-px padded buffers
        const cv::Mat& m = ext.mvImagePyramid[l];
        put_i(m.rows); put_i(m.cols);
        for (int y = -19; y < m.rows + 19; ++y) put(m.data + (ptrdiff_t)y * (ptrdiff_t)m.step - 19, (size_t)m.cols + 38);
    }
    // ---- Amos two-stage path (Frame.cc:474-498, 633-636): detect -> MovingKeyPoints -> ProcessDesp
    std::vector<std::vector<cv::KeyPoint> > per_level;
    ext(A, cv::Mat(), per_level);
    cv::Mat mask(h, w, CV_8UC1), label(h, w, CV_64F);
    for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) {
        mask.at<uchar>(y, x) = (x > w / 3 && x < w / 2 && y > h / 4 && y < h / 2) ? 255 : 0;
        label.at<double>(y, x) = 1 + (x / 80) + 8 * (y / 80);
    }
    std::vector<center> centers(64);
    for (int i = 0; i < 64; ++i) { centers[i].id = i; }
    std::vector<int> rm(64, 0); rm[5] = 1; rm[17] = 1;
    std::vector<cv::KeyPoint> culled = ext.MovingKeyPoints(A, mask, label, centers, rm, std::vector<bool>(), per_level);
    std::vector<cv::KeyPoint> kc; cv::Mat dc;
    ext.ProcessDesp(A, cv::Mat(), per_level, kc, dc);
    put_i((int)culled.size()); if (!culled.empty()) put(culled.data(), culled.size() * sizeof(cv::KeyPoint));
    put_kps(kc, dc);
    // ---- ORBmatcher::SearchForInitialization as Tracking::MonocularInitialization calls it (Tracking.cc:1492-1500)
    ext.SetExportPyramid(false);
    ext(B, cv::Mat(), kb, db);
    Frame FA, FB; fill_frame(FA, ext, ka, da, w, h); fill_frame(FB, ext, kb, db, w, h);
    std::vector<cv::Point2f> prev(ka.size());
    for (size_t i = 0; i < ka.size(); ++i) prev[i] = ka[i].pt;
    std::vector<int> m12;
    ORBmatcher matcher(0.9, true);
    int nm = matcher.SearchForInitialization(FA, FB, prev, m12, 100);
    put_i(nm); put_i((int)m12.size()); if (!m12.empty()) put(m12.data(), m12.size() * 4);
    // ---- ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)  (Tracking.cc:2378-2389)
    std::vector<MapPoint> store(ka.size());
    std::vector<MapPoint*> pts;
    for (size_t i = 0; i < ka.size(); ++i) {
        MapPoint& p = store[i];
        p.mTrackProjX = ka[i].pt.x + 7.f; p.mTrackProjY = ka[i].pt.y - 4.f; p.mTrackProjXR = -1.f;
        p.mbTrackInView = (i % 3) != 0; p.mbBad = (i % 11) == 0;
        p.mnTrackScaleLevel = ka[i].octave; p.mTrackViewCos = (i % 2) ? 0.9995f : 0.9f;
        p.mDescriptor = da.row((int)i); p.nObs = (i % 5) ? 1 : 0;
        pts.push_back(&p);
    }
    nm = ORBmatcher(0.8, true).SearchByProjection(FB, pts, 3.f);
    put_i(nm);
    for (int j = 0; j < FB.N; ++j) put_i(FB.mvpMapPoints[j] ? (int)(FB.mvpMapPoints[j] - store.data()) : -1);
    // ---- Frame::ComputeStereoMatches (Frame.cc:165-176 then :1179)
    ORBextractor eL(2000, 1.2f, 8, 20, 7), eR(2000, 1.2f, 8, 20, 7);
    eL.SetExportPyramid(false); eR.SetExportPyramid(false);
    Frame S; S.mpORBextractorLeft = &eL; S.mpORBextractorRight = &eR;
    eL(L, cv::Mat(), S.mvKeys, S.mDescriptors);
    eR(R, cv::Mat(), S.mvKeysRight, S.mDescriptorsRight);
    S.N = (int)S.mvKeys.size(); S.mb = 0.f; S.mbf = 386.1448f;
    S.ComputeStereoMatches();
    put_i(S.N); put(S.mvuRight.data(), (size_t)S.N * 4); put(S.mvDepth.data(), (size_t)S.N * 4);
    // ---- device-resident Frame: UndistortKeyPoints / ComputeStereoFromRGBD / AssignFeaturesToGrid as Frame::CalDyna calls them
    //      (Frame.cc:636-645), then the matchers on frames that carry a device frame
    float cam[10], bounds[6];
    if (fread(cam, 4, 10, f) != 10 || fread(bounds, 4, 6, f) != 6) return 2;
    {
        Frame D; D.mpORBextractorLeft = &ext;
        ext(A, cv::Mat(), D.mvKeys, D.mDescriptors);                      // ExtractORBDesp: the result stays on the device
        D.N = (int)D.mvKeys.size(); D.mbf = cam[9]; D.mvScaleFactors = ext.GetScaleFactors();
        D.mK = cv::Mat::eye(3, 3, CV_32F);
        D.mK.at<float>(0, 0) = cam[0]; D.mK.at<float>(1, 1) = cam[1]; D.mK.at<float>(0, 2) = cam[2]; D.mK.at<float>(1, 2) = cam[3];
        D.mDistCoef = cv::Mat(5, 1, CV_32F);
        for (int i = 0; i < 5; ++i) D.mDistCoef.at<float>(i) = cam[4 + i];
        Frame::mnMinX = bounds[0]; Frame::mnMaxX = bounds[1]; Frame::mnMinY = bounds[2]; Frame::mnMaxY = bounds[3];   // ComputeImageBounds stays the reference's
        Frame::mfGridElementWidthInv = bounds[4]; Frame::mfGridElementHeightInv = bounds[5];
        cv::Mat imD(h, w, CV_32F);
        for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) imD.at<float>(y, x) = (float)((x * 7 + y * 13) % 97) / 16.0f - 0.5f;
        D.UndistortKeyPoints();
        D.ComputeStereoFromRGBD(imD);
        D.AssignFeaturesToGrid();
        put_i(D.N);
        put(D.mvKeysUn.data(), (size_t)D.N * sizeof(cv::KeyPoint)); put(D.mvuRight.data(), (size_t)D.N * 4); put(D.mvDepth.data(), (size_t)D.N * 4);
        for (int x = 0; x < 64; ++x) for (int y = 0; y < 48; ++y) {
            put_i((int)D.mGrid[x][y].size());
            for (size_t q = 0; q < D.mGrid[x][y].size(); ++q) put_i((int)D.mGrid[x][y][q]);
        }
        // a frame whose keypoints did NOT just come from the extractor (N differs from its device result): host upload path
        Frame H = D; H.mpDeviceFrame.reset(); H.mpORBextractorLeft = &ext;
        H.N = D.N - 7; H.mvKeys.resize(H.N); H.mDescriptors = D.mDescriptors.rowRange(0, H.N);
        H.UndistortKeyPoints();
        put_i(H.N); put(H.mvKeysUn.data(), (size_t)H.N * sizeof(cv::KeyPoint));
    }
    {   // rectified camera (k1 == 0), mono constructor order (Frame.cc:367-428): the matchers must give what they gave on host frames
        Frame D1, D2;
        Frame* Fs[2] = {&D1, &D2}; cv::Mat* imgs[2] = {&A, &B};
        for (int q = 0; q < 2; ++q) {
            Frame& D = *Fs[q];
            D.mpORBextractorLeft = &ext;
            ext(*imgs[q], cv::Mat(), D.mvKeys, D.mDescriptors);
            fill_frame(D, ext, D.mvKeys, D.mDescriptors, w, h);
            D.mK = cv::Mat::eye(3, 3, CV_32F); D.mK.at<float>(0, 0) = cam[0]; D.mK.at<float>(1, 1) = cam[1]; D.mK.at<float>(0, 2) = cam[2]; D.mK.at<float>(1, 2) = cam[3];
            D.mDistCoef = cv::Mat::zeros(4, 1, CV_32F);
            D.UndistortKeyPoints();
            D.mvuRight = std::vector<float>(D.N, -1); D.mvDepth = std::vector<float>(D.N, -1);         // :381-382
            D.AssignFeaturesToGrid();
        }
        std::vector<cv::Point2f> prev2(D1.mvKeysUn.size());
        for (size_t i = 0; i < prev2.size(); ++i) prev2[i] = D1.mvKeysUn[i].pt;
        std::vector<int> m12b;
        int nm2 = ORBmatcher(0.9, true).SearchForInitialization(D1, D2, prev2, m12b, 100);
        put_i(nm2); put_i((int)m12b.size()); if (!m12b.empty()) put(m12b.data(), m12b.size() * 4);
        nm2 = ORBmatcher(0.8, true).SearchByProjection(D2, pts, 3.f);
        put_i(nm2);
        for (int j = 0; j < D2.N; ++j) put_i(D2.mvpMapPoints[j] ? (int)(D2.mvpMapPoints[j] - store.data()) : -1);
    }
    // ---- ORBmatcher::SearchByProjection(Frame&, KeyFrame*, set&, th, ORBdist)  (relocalisation, Tracking.cc:2663): identity pose,
    //      map points placed so that they project onto A's keypoints shifted by the known motion
    {
        Frame C; fill_frame(C, ext, kb, db, w, h);
        C.fx = 517.3f; C.fy = 516.5f; C.cx = 318.6f; C.cy = 255.3f;
        C.mTcw = cv::Mat::eye(4, 4, CV_32F);
        std::vector<MapPoint> mps(ka.size());
        KeyFrame K; K.mvKeysUn = ka; K.mvpMapPoints.assign(ka.size(), (MapPoint*)NULL);
        std::set<MapPoint*> found;
        for (size_t i = 0; i < ka.size(); ++i) {
            MapPoint& p = mps[i];
            const float z = 1.f + (float)(i % 7);
            p.mWorldPos = cv::Mat(3, 1, CV_32F);
            p.mWorldPos.at<float>(0) = (ka[i].pt.x + 7.f - C.cx) / C.fx * z; p.mWorldPos.at<float>(1) = (ka[i].pt.y - 4.f - C.cy) / C.fy * z; p.mWorldPos.at<float>(2) = z;
            p.mDescriptor = da.row((int)i); p.nPredictedLevel = ka[i].octave; p.mbBad = (i % 11) == 0;
            if (i % 13 == 0) p.mfMaxDistance = 0.5f;                                   // out of its scale-invariance range
            if (i % 4) K.mvpMapPoints[i] = &p;
            if (i % 9 == 0) found.insert(&p);
        }
        for (int j = 0; j < C.N; ++j) if (j % 10 == 0) C.mvpMapPoints[j] = &mps[0];      // pre-existing assignments block
        int nk = ORBmatcher(0.9, true).SearchByProjection(C, &K, found, 10.f, 100);
        put_i(nk);
        for (int j = 0; j < C.N; ++j) put_i((C.mvpMapPoints[j] && j % 10) ? (int)(C.mvpMapPoints[j] - mps.data()) : -1);
    }
    // ---- ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)  (LoopClosing): Scw = identity
    {
        const float fx = 517.3f, fy = 516.5f, cx = 318.6f, cy = 255.3f;
        KeyFrame K; K.mvKeysUn = kb; K.mDescriptors = db; K.fx = fx; K.fy = fy; K.cx = cx; K.cy = cy;
        K.mnMinX = 0; K.mnMinY = 0; K.mnMaxX = w; K.mnMaxY = h; K.mfGridElementWidthInv = 64.f / (float)w; K.mfGridElementHeightInv = 48.f / (float)h;
        K.mvScaleFactors = ext.GetScaleFactors();
        std::vector<MapPoint> mps(ka.size()); MapPoint dummy;
        std::vector<MapPoint*> pts(ka.size()), matched(kb.size(), (MapPoint*)NULL);
        for (size_t i = 0; i < ka.size(); ++i) {
            MapPoint& p = mps[i];
            const float z = 1.f + (float)(i % 7);
            const float X = (ka[i].pt.x + 7.f - cx) / fx * z, Y = (ka[i].pt.y - 4.f - cy) / fy * z;
            p.mWorldPos = cv::Mat(3, 1, CV_32F); p.mWorldPos.at<float>(0) = X; p.mWorldPos.at<float>(1) = Y; p.mWorldPos.at<float>(2) = z;
            const float len = std::sqrt(X * X + Y * Y + z * z), sgn = (i % 6) ? 1.f : -1.f;
            p.mNormalVector = cv::Mat(3, 1, CV_32F); p.mNormalVector.at<float>(0) = sgn * X / len; p.mNormalVector.at<float>(1) = sgn * Y / len; p.mNormalVector.at<float>(2) = sgn * z / len;
            p.mDescriptor = da.row((int)i); p.nPredictedLevel = ka[i].octave; p.mbBad = (i % 11) == 0;
            if (i % 13 == 0) p.mfMaxDistance = 0.5f;
            pts[i] = &p;
        }
        for (size_t j = 0; j < kb.size(); ++j) if (j % 10 == 0) matched[j] = &dummy;
        int nk = ORBmatcher(0.75, true).SearchByProjection(&K, cv::Mat::eye(4, 4, CV_32F), pts, matched, 10);
        put_i(nk);
        for (size_t j = 0; j < kb.size(); ++j) put_i((matched[j] && matched[j] != &dummy) ? (int)(matched[j] - mps.data()) : -1);
    }
    // ---- bag of words: Frame::ComputeBoW / KeyFrame::ComputeBoW and both SearchByBoW forms (Tracking.cc:1740-1752, LoopClosing)
    if (argc > 3) {
        ORBVocabulary voc;
        RegisterDeviceVocabulary(&voc, argv[3]);                                                   // System.cc:84, once
        Frame F2; fill_frame(F2, ext, kb, db, w, h); F2.mpORBvocabulary = &voc;
        F2.ComputeBoW();
        put_i((int)F2.mBowVec.size());
        for (DBoW2::BowVector::const_iterator it = F2.mBowVec.begin(); it != F2.mBowVec.end(); ++it) { put_i((int)it->first); put(&it->second, 8); }
        put_i((int)F2.mFeatVec.size());
        for (DBoW2::FeatureVector::const_iterator it = F2.mFeatVec.begin(); it != F2.mFeatVec.end(); ++it) {
            put_i((int)it->first); put_i((int)it->second.size());
            for (size_t q = 0; q < it->second.size(); ++q) put_i((int)it->second[q]);
        }
        std::vector<MapPoint> mp1(ka.size()), mp2(kb.size());
        KeyFrame K1; K1.mvKeysUn = ka; K1.mDescriptors = da; K1.mpORBvocabulary = &voc; K1.mvpMapPoints.assign(ka.size(), (MapPoint*)NULL);
        for (size_t i = 0; i < ka.size(); ++i) { if (i % 5) K1.mvpMapPoints[i] = &mp1[i]; mp1[i].mbBad = (i % 11) == 0; }
        K1.ComputeBoW();
        KeyFrame K2; K2.mvKeysUn = kb; K2.mDescriptors = db; K2.mpORBvocabulary = &voc; K2.mvpMapPoints.assign(kb.size(), (MapPoint*)NULL);
        for (size_t j = 0; j < kb.size(); ++j) { if (j % 7) K2.mvpMapPoints[j] = &mp2[j]; mp2[j].mbBad = (j % 13) == 0; }
        K2.ComputeBoW();
        std::vector<MapPoint*> mf, m12;
        int nb1 = ORBmatcher(0.7, true).SearchByBoW(&K1, F2, mf);
        put_i(nb1); put_i((int)mf.size());
        for (size_t j = 0; j < mf.size(); ++j) put_i(mf[j] ? (int)(mf[j] - mp1.data()) : -1);
        int nb2 = ORBmatcher(0.75, true).SearchByBoW(&K1, &K2, m12);
        put_i(nb2); put_i((int)m12.size());
        for (size_t i = 0; i < m12.size(); ++i) put_i(m12[i] ? (int)(m12[i] - mp2.data()) : -1);
    }
    fclose(g_out); fclose(f);
    // ---- per-call latency of the drop-in classes as C++ code sees them (everything the body does on the host included: flattening the
    //      arguments, the C-ABI call with its copies, scattering the results back into the reference's containers)
    {
        ext.SetExportPyramid(false);
        std::vector<cv::KeyPoint> kt; cv::Mat dt;
        auto timeit = [&](const char* name, int reps, const std::function<void()>& fn) {
            for (int i = 0; i < 5; ++i) fn();
            const auto t0 = std::chrono::steady_clock::now();
            for (int i = 0; i < reps; ++i) fn();
            const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
            printf("timing_us %s %.1f\n", name, us);
        };
        timeit("ORBextractor::operator()_640x480_1000", 200, [&] { ext(A, cv::Mat(), kt, dt); });
        timeit("ORBmatcher::SearchForInitialization_host_frames", 200, [&] {
            std::vector<cv::Point2f> pv(ka.size()); for (size_t i = 0; i < ka.size(); ++i) pv[i] = ka[i].pt;
            std::vector<int> mm; ORBmatcher(0.9, true).SearchForInitialization(FA, FB, pv, mm, 100); });
        timeit("ORBmatcher::SearchByProjection(Frame,MapPoints)_incl_flatten", 200, [&] {
            FB.mvpMapPoints.assign(FB.N, (MapPoint*)NULL); ORBmatcher(0.8, true).SearchByProjection(FB, pts, 3.f); });
        timeit("Frame::ComputeStereoMatches_2000x2000", 100, [&] { S.ComputeStereoMatches(); });
    }
    printf("host drop-in ok: %zu / %zu keypoints, %d stereo keypoints\n", ka.size(), kb.size(), S.N);
    return 0;
}
