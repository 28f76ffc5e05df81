// orb_port.cpp -- TEST INFRASTRUCTURE ONLY (oracle "port").  Not part of the shipped product path;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// A plain CPU restatement of the reference's ORB extractor (Amos-SLAM / ORB-SLAM2), function by
// function, each citing the reference file:line it follows.  It is pinned two ways:
//   * against oracle/_ref (the reference's own ORBextractor.cc compiled unmodified, monotonic allocator)
//   * against cv2 4.13 for the OpenCV primitives (tests/test_oracle_cvlite.py, tests/golden/)
//
// Differences from the reference, all deliberate and all stated in DESIGN.md:
//   1. DistributeOctTree tie-break is the canonical rule "equal size => newest node first" made
//      explicit with creation ids (the reference sorts by heap address, ORBextractor.cc:948).
//   2. cos/sin for the BRIEF rotation: det_sincos() restates glibc's sincosf (FMA variant), the function the
//      reference build calls, operation by operation; tests pin it against the live libm (0 differences).
//   3. Per-cell FAST is stated in its "score map" form: S(p) once per pixel, candidates are the
//      strict 3x3 local maxima of S inside the cell zone, threshold applied afterwards
//      (equivalent to two cv::FAST calls).
#include "../cvlite/cvlite.hpp"
#include <list>
#include <vector>
#include <cmath>
#include <cstring>

namespace port {

typedef cv::KeyPoint KeyPoint;

static const int PATCH_SIZE = 31;        // ORBextractor.cc:91
static const int HALF_PATCH_SIZE = 15;   // ORBextractor.cc:92
static const int EDGE_THRESHOLD = 19;    // ORBextractor.cc:93

#include "brief_pattern.inc"             // static const signed char bit_pattern_31[256*4]  (ORBextractor.cc:231-489)

struct Image { int w = 0, h = 0; std::vector<unsigned char> px; const unsigned char* row(int y) const { return px.data() + (size_t)y * w; } unsigned char* row(int y) { return px.data() + (size_t)y * w; } };

// ---------------------------------------------------------------------------------------------
// det_sincos: sin/cos of a float angle in radians exactly as the reference computes them.  The reference writes
// `cos(angle)`, `sin(angle)` on a float under `using namespace std` (ORBextractor.cc:178-181); GCC merges the pair into
// one sincosf call (oracle/_ref imports `sincosf` and nothing else from libm), which glibc >= 2.28 evaluates as a
// double-precision polynomial after a pi/2 reduction (sysdeps/ieee754/flt-32/s_sincosf.c, sincosf_poly.h; constants
// from __sincosf_table).  The contract is glibc's FMA ifunc variant (__sincosf_fma, selected on every FMA + AVX2
// CPU): in the disassembly of glibc 2.39's libm every `a + b * c` of sincosf_poly and the `x - n * hpi` of
// reduce_fast is one fused multiply-add and every bare product is rounded -- restated here with explicit
// std::fma (this file is compiled with -ffp-contract=off) and mirrored operation by operation by the CUDA
// kernel (amos-slam_b200/csrc/det_math.cuh).  Pinned against the live libm by tests/test_oracle_cvlite.py.
// ---------------------------------------------------------------------------------------------
static inline void det_sincos(float y, float* sinp, float* cosp) {
    // __sincosf_table[0] / [1] (the second one negates the cosine polynomial for quadrants 2, 3)
    static const double SIGN[4] = {1.0, -1.0, -1.0, 1.0};
    const double HPI_INV = 0x1.45F306DC9C883p+23;     // 2/pi * 2^24
    const double HPI = 0x1.921FB54442D18p0;
    const double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10, C4 = 0x1.99343027bf8c3p-16;
    const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
    static const uint32_t INV_PIO4[24] = {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1,
                                          0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
    uint32_t xi; std::memcpy(&xi, &y, 4);
    const uint32_t top = (xi >> 20) & 0x7ff;          // abstop12
    double x = (double)y, s = 1.0, cneg = 1.0;        // cneg = -1 selects __sincosf_table[1]
    int n = 0;
    if (top < 0x3f4) {                                // |y| < pi/4
        if (top < 0x398) { *sinp = y; *cosp = 1.0f; return; }   // |y| < 2^-12
    } else if (top < 0x42f) {                         // |y| < 120: reduce_fast
        const double r = x * HPI_INV;
        n = ((int32_t)r + 0x800000) >> 24;
        x = std::fma(-(double)n, HPI, x);
        s = SIGN[n & 3];
        if (n & 2) cneg = -1.0;
    } else if (top < 0x7f8) {                         // reduce_large
        const uint32_t* arr = &INV_PIO4[(xi >> 26) & 15];
        const int shift = (xi >> 23) & 7, sign = xi >> 31;
        uint32_t m = ((xi & 0xffffff) | 0x800000) << shift;
        uint64_t res0 = (uint32_t)(m * arr[0]), res1 = (uint64_t)m * arr[4], res2 = (uint64_t)m * arr[8];
        res0 = (res2 >> 32) | (res0 << 32);
        res0 += res1;
        const uint64_t nn = (res0 + (1ULL << 61)) >> 62;
        res0 -= nn << 62;
        x = (double)(int64_t)res0 * 0x1.921FB54442D18p-62;
        n = (int)nn;
        s = SIGN[(n + sign) & 3];
        if ((n + sign) & 2) cneg = -1.0;
    } else { *sinp = *cosp = y - y; return; }         // inf / nan
    // sincosf_poly(x * s, x * x, p, n, sinp, cosp)
    const double xs = x * s, x2 = x * x;
    const double c0 = cneg * C0, c1 = cneg * C1, c2 = cneg * C2, c3 = cneg * C3, c4 = cneg * C4;     // exact sign flips
    const double s1v = std::fma(x2, S3, S2), c2v = std::fma(x2, c4, c3);
    const double x3 = x2 * xs, x4 = x2 * x2;
    const double x5 = x2 * x3, x6 = x2 * x4;
    const double c1v = std::fma(x2, c1, c0);
    const double sv = std::fma(x3, S1, xs), cv = std::fma(x4, c2, c1v);
    const float fs = (float)std::fma(s1v, x5, sv), fc = (float)std::fma(c2v, x6, cv);
    if (n & 1) { *sinp = fc; *cosp = fs; } else { *sinp = fs; *cosp = fc; }
}

// ---------------------------------------------------------------------------------------------
struct Extractor {
    int nfeatures; double scaleFactor; int nlevels, iniThFAST, minThFAST;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mnFeaturesPerLevel, umax;
    std::vector<Image> pyramid;                    // mvImagePyramid (ROI only; the 19-px pad is never read)
    std::vector<std::vector<KeyPoint> > lastCandidates;   // vToDistributeKeys per level of the last detect (stage parity)

    // ORBextractor::ORBextractor   ORBextractor.cc:492-609
    Extractor(int _nfeatures, float _scaleFactor, int _nlevels, int _ini, int _min)
        : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_ini), minThFAST(_min) {
        mvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvScaleFactor[0] = 1.0f; mvLevelSigma2[0] = 1.0f;
        for (int i = 1; i < nlevels; i++) {
            mvScaleFactor[i] = (float)(mvScaleFactor[i - 1] * scaleFactor);      // double product stored to float (:512, A.6)
            mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
        }
        mvInvScaleFactor.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        for (int i = 0; i < nlevels; i++) { mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i]; mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i]; }
        pyramid.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels);
        float factor = (float)(1.0f / scaleFactor);                                  // :534 (float = 1.0f/double)
        float nDesired = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));   // :539
        int sum = 0;
        for (int level = 0; level < nlevels - 1; level++) {
            mnFeaturesPerLevel[level] = cv::cvRound(nDesired);
            sum += mnFeaturesPerLevel[level];
            nDesired *= factor;
        }
        mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sum, 0);
        // circular patch row extents  :579-608
        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = cv::cvFloor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = cv::cvCeil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = cv::cvRound(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0; ++v0;
        }
    }

    // ORBextractor::ComputePyramid   ORBextractor.cc:1826-1886 ; levels are CHAINED (l from l-1)
    void ComputePyramid(const unsigned char* img, int rows, int cols, int step) {
        for (int level = 0; level < nlevels; ++level) {
            float scale = mvInvScaleFactor[level];
            int w = cv::cvRound((float)cols * scale), h = cv::cvRound((float)rows * scale);    // :1834
            Image& im = pyramid[level];
            im.w = w; im.h = h; im.px.resize((size_t)w * h);
            if (level == 0) { for (int y = 0; y < rows; ++y) std::memcpy(im.row(y), img + (size_t)y * step, (size_t)cols); }
            else { const Image& p = pyramid[level - 1]; cv::cvl_resize_linear_u8(p.px.data(), p.w, p.h, (size_t)p.w, im.px.data(), w, h, (size_t)w); }
        }
    }

    // One level of ORBextractor::ComputeKeyPointsOctTree's cell loop   ORBextractor.cc:1064-1157
    // + cv::FAST semantics (A.3) in score-map form.
    void LevelCandidates(int level, std::vector<KeyPoint>& vToDistributeKeys) const {
        const Image& im = pyramid[level];
        vToDistributeKeys.clear();
        const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
        const int maxBorderX = im.w - EDGE_THRESHOLD + 3, maxBorderY = im.h - EDGE_THRESHOLD + 3;
        const float W = 30;
        const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
        const int nCols = (int)(width / W), nRows = (int)(height / W);
        if (nCols <= 0 || nRows <= 0) return;   // reference would divide by zero; images this small are rejected upstream
        const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
        std::vector<unsigned char> S;
        for (int i = 0; i < nRows; i++) {
            const float iniY = (float)(minBorderY + i * hCell);
            float maxY = iniY + hCell + 6;
            if (iniY >= maxBorderY - 3) continue;
            if (maxY > maxBorderY) maxY = (float)maxBorderY;
            for (int j = 0; j < nCols; j++) {
                const float iniX = (float)(minBorderX + j * wCell);
                float maxX = iniX + wCell + 6;
                if (iniX >= maxBorderX - 6) continue;
                if (maxX > maxBorderX) maxX = (float)maxBorderX;
                // cell ROI = rows [iniY,maxY) x cols [iniX,maxX); FAST zone = ROI shrunk by 3
                const int x0 = (int)iniX, y0 = (int)iniY, cw = (int)maxX - x0, ch = (int)maxY - y0;
                if (cw < 7 || ch < 7) continue;
                const int zw = cw - 6, zh = ch - 6;
                S.assign((size_t)(zw + 2) * (zh + 2), 0);   // 1-px zero ring = "outside the zone counts as 0"
                for (int y = 0; y < zh; ++y)
                    for (int x = 0; x < zw; ++x) {
                        int s = cv::cvl_fast_S(im.row(y0 + 3 + y) + x0 + 3 + x, (size_t)im.w);
                        S[(size_t)(y + 1) * (zw + 2) + x + 1] = (unsigned char)(s > minThFAST ? s : 0);
                    }
                // strict local maxima; threshold iniTh, retry with minTh iff the cell produced nothing
                size_t start = vToDistributeKeys.size();
                for (int pass = 0; pass < 2; ++pass) {
                    const int th = pass == 0 ? iniThFAST : minThFAST;
                    for (int y = 0; y < zh; ++y)
                        for (int x = 0; x < zw; ++x) {
                            const unsigned char* s = &S[(size_t)(y + 1) * (zw + 2) + x + 1];
                            const int st = zw + 2, c = s[0];
                            if (c <= th) continue;
                            if (c > s[-1] && c > s[1] && c > s[-st - 1] && c > s[-st] && c > s[-st + 1] && c > s[st - 1] && c > s[st] && c > s[st + 1]) {
                                // pt in cell-ROI coords (x+3,y+3), shifted by j*wCell, i*hCell  (:1150-1151)
                                vToDistributeKeys.push_back(KeyPoint((float)(x + 3 + j * wCell), (float)(y + 3 + i * hCell), 7.f, -1.f, (float)(c - 1)));
                            }
                        }
                    if (vToDistributeKeys.size() != start) break;
                    if (minThFAST >= iniThFAST) break;
                }
            }
        }
    }

    // ExtractorNode + DistributeOctTree   ORBextractor.cc:635-1049, canonical tie-break
    struct Node {
        int ULx, ULy, URx, BLy, BRx, BRy;   // UL=(ULx,ULy) UR=(URx,ULy) BL=(ULx,BLy) BR=(BRx,BRy)
        std::vector<KeyPoint> vKeys; bool bNoMore = false; long id = 0;
    };
    static void DivideNode(const Node& p, Node& n1, Node& n2, Node& n3, Node& n4) {   // :635-703
        const int halfX = (int)std::ceil(static_cast<float>(p.URx - p.ULx) / 2);
        const int halfY = (int)std::ceil(static_cast<float>(p.BRy - p.ULy) / 2);
        // n1.UL=UL n1.UR=(UL.x+halfX,UL.y) n1.BL=(UL.x,UL.y+halfY) n1.BR=(UL.x+halfX,UL.y+halfY)
        n1.ULx = p.ULx; n1.ULy = p.ULy; n1.URx = p.ULx + halfX; n1.BLy = p.ULy + halfY; n1.BRx = p.ULx + halfX; n1.BRy = p.ULy + halfY;
        // n2.UL=n1.UR n2.UR=UR n2.BL=n1.BR n2.BR=(UR.x,UL.y+halfY)
        n2.ULx = n1.URx; n2.ULy = p.ULy; n2.URx = p.URx; n2.BLy = n1.BRy; n2.BRx = p.URx; n2.BRy = p.ULy + halfY;
        // n3.UL=n1.BL n3.UR=n1.BR n3.BL=BL n3.BR=(n1.BR.x,BL.y)
        n3.ULx = p.ULx; n3.ULy = n1.BLy; n3.URx = n1.BRx; n3.BLy = p.BLy; n3.BRx = n1.BRx; n3.BRy = p.BLy;
        // n4.UL=n3.UR n4.UR=n2.BR n4.BL=n3.BR n4.BR=BR
        n4.ULx = n3.URx; n4.ULy = n3.ULy; n4.URx = n2.BRx; n4.BLy = n3.BRy; n4.BRx = p.BRx; n4.BRy = p.BRy;
        for (size_t i = 0; i < p.vKeys.size(); i++) {
            const KeyPoint& kp = p.vKeys[i];
            if (kp.pt.x < n1.URx) { if (kp.pt.y < n1.BRy) n1.vKeys.push_back(kp); else n3.vKeys.push_back(kp); }
            else if (kp.pt.y < n1.BRy) n2.vKeys.push_back(kp);
            else n4.vKeys.push_back(kp);
        }
        n1.bNoMore = n1.vKeys.size() == 1; n2.bNoMore = n2.vKeys.size() == 1;
        n3.bNoMore = n3.vKeys.size() == 1; n4.bNoMore = n4.vKeys.size() == 1;
    }
    typedef std::list<Node> NodeList;
    struct SizeNode { int size; long id; NodeList::iterator it; };

    std::vector<KeyPoint> DistributeOctTree(const std::vector<KeyPoint>& vToDistributeKeys, int minX, int maxX, int minY, int maxY, int N) const {
        const int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));   // :719
        const float hX = static_cast<float>(maxX - minX) / nIni;                            // :722
        NodeList lNodes; long nextId = 0;
        std::vector<NodeList::iterator> vpIniNodes(std::max(nIni, 0));
        for (int i = 0; i < nIni; i++) {
            Node ni;
            ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;                          // :741 float->int truncation
            ni.URx = (int)(hX * static_cast<float>(i + 1));
            ni.BLy = maxY - minY; ni.BRx = ni.URx; ni.BRy = maxY - minY;
            ni.id = nextId++;
            lNodes.push_back(ni);
            vpIniNodes[i] = --lNodes.end();
        }
        for (size_t i = 0; i < vToDistributeKeys.size(); i++) {
            const KeyPoint& kp = vToDistributeKeys[i];
            vpIniNodes[(size_t)(kp.pt.x / hX)]->vKeys.push_back(kp);                         // :766
        }
        for (NodeList::iterator lit = lNodes.begin(); lit != lNodes.end();) {             // :770-788
            if (lit->vKeys.size() == 1) { lit->bNoMore = true; ++lit; }
            else if (lit->vKeys.empty()) lit = lNodes.erase(lit);
            else ++lit;
        }
        bool bFinish = false;
        std::vector<SizeNode> vSizeAndNode;
        // helper: push the non-empty children to the FRONT of the list in order n1..n4  (:846-892)
        auto pushChildren = [&](Node* ch[4], int& nToExpand) {
            for (int c = 0; c < 4; ++c) {
                if (ch[c]->vKeys.size() > 0) {
                    ch[c]->id = nextId++;
                    lNodes.push_front(*ch[c]);
                    if (ch[c]->vKeys.size() > 1) { nToExpand++; SizeNode sn; sn.size = (int)ch[c]->vKeys.size(); sn.id = lNodes.front().id; sn.it = lNodes.begin(); vSizeAndNode.push_back(sn); }
                }
            }
        };
        while (!bFinish) {
            int prevSize = (int)lNodes.size();
            NodeList::iterator lit = lNodes.begin();
            int nToExpand = 0;
            vSizeAndNode.clear();
            while (lit != lNodes.end()) {                                                   // :824-899
                if (lit->bNoMore) { ++lit; continue; }
                Node n1, n2, n3, n4; Node* ch[4] = {&n1, &n2, &n3, &n4};
                DivideNode(*lit, n1, n2, n3, n4);
                pushChildren(ch, nToExpand);
                lit = lNodes.erase(lit);
            }
            if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;   // :907
            else if (((int)lNodes.size() + nToExpand * 3) > N) {                            // :929
                while (!bFinish) {
                    prevSize = (int)lNodes.size();
                    std::vector<SizeNode> vPrev = vSizeAndNode;
                    vSizeAndNode.clear();
                    // :948 sort ascending by (size, address); canonical: address order == creation order
                    std::sort(vPrev.begin(), vPrev.end(), [](const SizeNode& a, const SizeNode& b) { return a.size != b.size ? a.size < b.size : a.id < b.id; });
                    for (int j = (int)vPrev.size() - 1; j >= 0; j--) {                      // :950 largest first, newest first
                        Node n1, n2, n3, n4; Node* ch[4] = {&n1, &n2, &n3, &n4};
                        DivideNode(*vPrev[j].it, n1, n2, n3, n4);
                        int dummy = 0;
                        pushChildren(ch, dummy);
                        lNodes.erase(vPrev[j].it);
                        if ((int)lNodes.size() >= N) break;
                    }
                    if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
                }
            }
        }
        std::vector<KeyPoint> vResultKeys;                                                  // :1018-1048
        for (NodeList::iterator lit = lNodes.begin(); lit != lNodes.end(); ++lit) {
            const std::vector<KeyPoint>& k = lit->vKeys;
            const KeyPoint* pKP = &k[0]; float maxResponse = pKP->response;
            for (size_t i = 1; i < k.size(); i++) if (k[i].response > maxResponse) { pKP = &k[i]; maxResponse = k[i].response; }
            vResultKeys.push_back(*pKP);
        }
        return vResultKeys;
    }

    // IC_Angle   ORBextractor.cc:108-161
    float IC_Angle(const Image& im, float ptx, float pty) const {
        int m_01 = 0, m_10 = 0;
        const int cx = cv::cvRound(ptx), cy = cv::cvRound(pty);
        const unsigned char* center = im.row(cy) + cx;
        for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
        const int step = im.w;
        for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
            int v_sum = 0, d = umax[v];
            for (int u = -d; u <= d; ++u) {
                int val_plus = center[u + v * step], val_minus = center[u - v * step];
                v_sum += (val_plus - val_minus);
                m_10 += u * (val_plus + val_minus);
            }
            m_01 += v * v_sum;
        }
        return cv::fastAtan2((float)m_01, (float)m_10);
    }

    // ComputeKeyPointsOctTree   ORBextractor.cc:1052-1199
    void ComputeKeyPointsOctTree(std::vector<std::vector<KeyPoint> >& all) {
        all.assign(nlevels, std::vector<KeyPoint>());
        lastCandidates.assign(nlevels, std::vector<KeyPoint>());
        for (int level = 0; level < nlevels; ++level) {
            const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
            const int maxBorderX = pyramid[level].w - EDGE_THRESHOLD + 3, maxBorderY = pyramid[level].h - EDGE_THRESHOLD + 3;
            LevelCandidates(level, lastCandidates[level]);
            std::vector<KeyPoint>& keypoints = all[level];
            keypoints = DistributeOctTree(lastCandidates[level], minBorderX, maxBorderX, minBorderY, maxBorderY, mnFeaturesPerLevel[level]);
            const int scaledPatchSize = (int)(PATCH_SIZE * mvScaleFactor[level]);            // :1175 float->int
            for (size_t i = 0; i < keypoints.size(); i++) {
                keypoints[i].pt.x += minBorderX; keypoints[i].pt.y += minBorderY;
                keypoints[i].octave = level; keypoints[i].size = (float)scaledPatchSize;
            }
        }
        for (int level = 0; level < nlevels; ++level)
            for (size_t i = 0; i < all[level].size(); ++i)
                all[level][i].angle = IC_Angle(pyramid[level], all[level][i].pt.x, all[level][i].pt.y);
    }

    // computeOrbDescriptor   ORBextractor.cc:173-227  (on the blurred level)
    static CVL_NOFMA void computeOrbDescriptor(const KeyPoint& kpt, const Image& img, unsigned char* desc) {
        const float factorPI = (float)(CV_PI / 180.f);
        float angle = (float)kpt.angle * factorPI;
        float a, b;
        det_sincos(angle, &b, &a);                       // a = cos, b = sin
        const unsigned char* center = img.row(cv::cvRound(kpt.pt.y)) + cv::cvRound(kpt.pt.x);
        const int step = img.w;
        const signed char* pat = bit_pattern_31;
        for (int i = 0; i < 32; ++i) {
            int val = 0;
            for (int k = 0; k < 8; ++k, pat += 4) {
                volatile float xb0 = pat[0] * b, ya0 = pat[1] * a, xa0 = pat[0] * a, yb0 = pat[1] * b;
                volatile float xb1 = pat[2] * b, ya1 = pat[3] * a, xa1 = pat[2] * a, yb1 = pat[3] * b;
                int t0 = center[cv::cvRound(xb0 + ya0) * step + cv::cvRound(xa0 - yb0)];
                int t1 = center[cv::cvRound(xb1 + ya1) * step + cv::cvRound(xa1 - yb1)];
                val |= (t0 < t1) << k;
            }
            desc[i] = (unsigned char)val;
        }
    }

    // second half of operator() / ProcessDesp   ORBextractor.cc:1578-1667, 1747-1820
    int Describe(std::vector<std::vector<KeyPoint> >& all, std::vector<KeyPoint>& flat, std::vector<unsigned char>& desc) const {
        int nkeypoints = 0;
        for (int level = 0; level < nlevels; ++level) nkeypoints += (int)all[level].size();
        flat.clear(); desc.assign((size_t)nkeypoints * 32, 0);
        int offset = 0;
        const std::vector<int> q = cv::cvl_gauss_kernel_q8(7, 2.0);
        for (int level = 0; level < nlevels; ++level) {
            std::vector<KeyPoint>& keypoints = all[level];
            int n = (int)keypoints.size();
            if (n == 0) continue;
            const Image& im = pyramid[level];
            Image working; working.w = im.w; working.h = im.h; working.px.resize(im.px.size());
            cv::cvl_gauss_blur_u8(im.px.data(), im.w, im.h, (size_t)im.w, working.px.data(), (size_t)im.w, q, cv::BORDER_REFLECT_101);
            for (int i = 0; i < n; ++i) computeOrbDescriptor(keypoints[i], working, &desc[(size_t)(offset + i) * 32]);
            offset += n;
            if (level != 0) { float scale = mvScaleFactor[level]; for (int i = 0; i < n; ++i) { keypoints[i].pt.x *= scale; keypoints[i].pt.y *= scale; } }
            flat.insert(flat.end(), keypoints.begin(), keypoints.end());
        }
        return nkeypoints;
    }

    // MovingKeyPoints   ORBextractor.cc:1688-1745
    std::vector<KeyPoint> MovingKeyPoints(const unsigned char* mask, const double* label, int rows, int cols,
                                          const int* centers_id, const int* rm_vector, std::vector<std::vector<KeyPoint> >& keys) const {
        cv::Mat imS(rows, cols, CV_8UC1, (void*)mask), dil, closed;
        cv::Mat kernel = cv::getStructuringElement(cv::MORPH_ELLIPSE, cv::Size(31, 31), cv::Point(15, 15));
        cv::dilate(imS, dil, kernel);
        cv::erode(dil, closed, kernel);
        std::vector<KeyPoint> dyna;
        for (int level = 0; level < nlevels; ++level) {
            std::vector<KeyPoint>& k = keys[level];
            if (k.empty()) continue;
            float scale = level != 0 ? mvScaleFactor[level] : 1.f;
            std::vector<KeyPoint> keep;
            for (size_t i = 0; i < k.size(); ++i) {
                float sx = k[i].pt.x * scale, sy = k[i].pt.y * scale;
                double super_pixel = label[(size_t)(int)sy * cols + (int)sx];
                int dyna_flag = rm_vector[centers_id[(size_t)(super_pixel - 1)]] == 1;      // :1727 (index = double->size_t)
                int label_coord = closed.at<unsigned char>((int)sy, (int)sx);
                if (label_coord != 0 || dyna_flag == 1) dyna.push_back(k[i]); else keep.push_back(k[i]);
            }
            k.swap(keep);
        }
        return dyna;
    }
};

}  // namespace port

// ================================= C API (ctypes) ============================================
using port::Extractor; using port::KeyPoint;

extern "C" {

void* port_extractor_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) { return new Extractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST); }
void port_extractor_destroy(void* h) { delete (Extractor*)h; }
int port_extractor_info(void* h, int* nlevels, float* scale_factors, int* features_per_level, int* umax16) {
    Extractor* e = (Extractor*)h; *nlevels = e->nlevels;
    for (int i = 0; i < e->nlevels; ++i) { scale_factors[i] = e->mvScaleFactor[i]; features_per_level[i] = e->mnFeaturesPerLevel[i]; }
    for (int i = 0; i < 16; ++i) umax16[i] = e->umax[i];
    return 0;
}
int port_extract(void* h, const unsigned char* img, int rows, int cols, int step, KeyPoint* kp_out, unsigned char* desc_out, int cap) {
    Extractor* e = (Extractor*)h;
    if (!img || rows <= 0 || cols <= 0) return 0;
    e->ComputePyramid(img, rows, cols, step);
    std::vector<std::vector<KeyPoint> > all; e->ComputeKeyPointsOctTree(all);
    std::vector<KeyPoint> flat; std::vector<unsigned char> desc;
    int n = e->Describe(all, flat, desc);
    if (n > cap) return -n;
    if (n) { std::memcpy(kp_out, flat.data(), sizeof(KeyPoint) * n); std::memcpy(desc_out, desc.data(), (size_t)n * 32); }
    return n;
}
int port_detect(void* h, const unsigned char* img, int rows, int cols, int step, KeyPoint* kp_out, int* level_counts, int cap) {
    Extractor* e = (Extractor*)h;
    e->ComputePyramid(img, rows, cols, step);
    std::vector<std::vector<KeyPoint> > all; e->ComputeKeyPointsOctTree(all);
    int n = 0; for (int l = 0; l < e->nlevels; ++l) { level_counts[l] = (int)all[l].size(); n += level_counts[l]; }
    if (n > cap) return -n;
    int o = 0; for (int l = 0; l < e->nlevels; ++l) { if (!all[l].empty()) std::memcpy(kp_out + o, all[l].data(), sizeof(KeyPoint) * all[l].size()); o += (int)all[l].size(); }
    return n;
}
int port_pyramid_level(void* h, int level, unsigned char* out, int* rows, int* cols) {
    Extractor* e = (Extractor*)h;
    if (level < 0 || level >= e->nlevels || e->pyramid[level].px.empty()) return -1;
    *rows = e->pyramid[level].h; *cols = e->pyramid[level].w;
    if (out) std::memcpy(out, e->pyramid[level].px.data(), e->pyramid[level].px.size());
    return 0;
}
// vToDistributeKeys of the last detect/extract for one level (stage parity with the GPU FAST kernel)
int port_level_candidates(void* h, int level, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h;
    if (level < 0 || level >= (int)e->lastCandidates.size()) return -1;
    int n = (int)e->lastCandidates[level].size();
    if (n > cap) return -n;
    if (n) std::memcpy(out, e->lastCandidates[level].data(), sizeof(KeyPoint) * n);
    return n;
}
int port_distribute_octtree(void* h, const KeyPoint* cand, int ncand, int minX, int maxX, int minY, int maxY, int N, int level, KeyPoint* out, int cap) {
    Extractor* e = (Extractor*)h; (void)level;
    std::vector<KeyPoint> v(cand, cand + ncand);
    std::vector<KeyPoint> r = e->DistributeOctTree(v, minX, maxX, minY, maxY, N);
    int n = (int)r.size();
    if (n > cap) return -n;
    if (n) std::memcpy(out, r.data(), sizeof(KeyPoint) * n);
    return n;
}
int port_moving_keypoints(void* h, const unsigned char* mask, const double* label, int rows, int cols,
                          const int* centers_id, int ncenters, const int* rm_vector, int nrm,
                          KeyPoint* kp_inout, int* level_counts, KeyPoint* culled_out) {
    Extractor* e = (Extractor*)h; (void)ncenters; (void)nrm;
    std::vector<std::vector<KeyPoint> > keys(e->nlevels);
    int o = 0; for (int l = 0; l < e->nlevels; ++l) { keys[l].assign(kp_inout + o, kp_inout + o + level_counts[l]); o += level_counts[l]; }
    std::vector<KeyPoint> dyn = e->MovingKeyPoints(mask, label, rows, cols, centers_id, rm_vector, keys);
    o = 0; for (int l = 0; l < e->nlevels; ++l) { level_counts[l] = (int)keys[l].size(); if (!keys[l].empty()) std::memcpy(kp_inout + o, keys[l].data(), sizeof(KeyPoint) * keys[l].size()); o += level_counts[l]; }
    if (!dyn.empty() && culled_out) std::memcpy(culled_out, dyn.data(), sizeof(KeyPoint) * dyn.size());
    return (int)dyn.size();
}
int port_process_desp(void* h, const KeyPoint* kp_in, const int* level_counts, KeyPoint* kp_out, unsigned char* desc_out, int cap) {
    Extractor* e = (Extractor*)h;
    std::vector<std::vector<KeyPoint> > keys(e->nlevels);
    int o = 0; for (int l = 0; l < e->nlevels; ++l) { keys[l].assign(kp_in + o, kp_in + o + level_counts[l]); o += level_counts[l]; }
    std::vector<KeyPoint> flat; std::vector<unsigned char> desc;
    int n = e->Describe(keys, flat, desc);
    if (n > cap) return -n;
    if (n) { std::memcpy(kp_out, flat.data(), sizeof(KeyPoint) * n); std::memcpy(desc_out, desc.data(), (size_t)n * 32); }
    return n;
}

// ---- primitive entry points (pinned against cv2 by tests/test_oracle_cvlite.py) --------------
void cvl_c_resize(const unsigned char* src, int sw, int sh, unsigned char* dst, int dw, int dh) { cv::cvl_resize_linear_u8(src, sw, sh, (size_t)sw, dst, dw, dh, (size_t)dw); }
void cvl_c_blur7(const unsigned char* src, int w, int h, unsigned char* dst) { cv::cvl_gauss_blur_u8(src, w, h, (size_t)w, dst, (size_t)w, cv::cvl_gauss_kernel_q8(7, 2.0), cv::BORDER_REFLECT_101); }
int cvl_c_fast(const unsigned char* img, int w, int h, int threshold, int nms, KeyPoint* out, int cap) {
    cv::Mat m(h, w, CV_8UC1, (void*)img); std::vector<KeyPoint> k; cv::FAST(m, k, threshold, nms != 0);
    int n = (int)k.size(); if (n > cap) return -n; if (n) std::memcpy(out, k.data(), sizeof(KeyPoint) * n); return n;
}
void cvl_c_fast_smap(const unsigned char* img, int w, int h, unsigned char* S) {
    std::memset(S, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; ++y) for (int x = 3; x < w - 3; ++x) { int s = cv::cvl_fast_S(img + (size_t)y * w + x, (size_t)w); S[(size_t)y * w + x] = (unsigned char)std::max(s, 0); }
}
// D = A (ar x ac) * B (ac x bc) [+ C] through cvlite's Mat algebra (the expression the matcher bodies write: Rcw * x3Dw + tcw)
void cvl_c_gemm(const float* a, int ar, int ac, const float* b, int bc, const float* c, float* d) {
    cv::Mat A(ar, ac, cv::CV_32F, (void*)a), B(ac, bc, cv::CV_32F, (void*)b);
    cv::Mat D = A * B;
    if (c) { cv::Mat Cm(ar, bc, cv::CV_32F, (void*)c); D = D + Cm; }
    for (int i = 0; i < ar; ++i) for (int j = 0; j < bc; ++j) d[i * bc + j] = D.at<float>(i, j);
}
void cvl_c_atan2(const float* y, const float* x, float* out, int n) { for (int i = 0; i < n; ++i) out[i] = cv::fastAtan2(y[i], x[i]); }
void cvl_c_close31(const unsigned char* mask, int w, int h, unsigned char* out) {
    cv::Mat m(h, w, CV_8UC1, (void*)mask), d, c; cv::Mat k = cv::getStructuringElement(cv::MORPH_ELLIPSE, cv::Size(31, 31), cv::Point(15, 15));
    cv::dilate(m, d, k); cv::erode(d, c, k);
    for (int y = 0; y < h; ++y) std::memcpy(out + (size_t)y * w, c.ptr(y), (size_t)w);
}
void cvl_c_ellipse31(unsigned char* out) { cv::Mat k = cv::getStructuringElement(cv::MORPH_ELLIPSE, cv::Size(31, 31), cv::Point(15, 15)); for (int y = 0; y < 31; ++y) std::memcpy(out + y * 31, k.ptr(y), 31); }
void cvl_c_border101(const unsigned char* src, int w, int h, int b, unsigned char* dst) {
    cv::Mat s(h, w, CV_8UC1, (void*)src), d; cv::copyMakeBorder(s, d, b, b, b, b, cv::BORDER_REFLECT_101);
    for (int y = 0; y < d.rows; ++y) std::memcpy(dst + (size_t)y * d.cols, d.ptr(y), (size_t)d.cols);
}
void port_det_sincos(const float* x, float* s, float* c, int n) { for (int i = 0; i < n; ++i) port::det_sincos(x[i], &s[i], &c[i]); }
void port_libm_sincosf(const float* x, float* s, float* c, int n) { for (int i = 0; i < n; ++i) sincosf(x[i], &s[i], &c[i]); }   // the libm entry point the reference build imports
// exhaustive pin: every float whose bit pattern lies in [lo, hi) through det_sincos and the live libm sincosf; returns the number of
// inputs where either output differs in any bit (the first `cap` such bit patterns go to `bad`)
long long port_sincos_sweep(unsigned lo, unsigned hi, unsigned* bad, int cap) {
    long long nbad = 0;
    for (unsigned long long b = lo; b < hi; ++b) {
        const unsigned u = (unsigned)b; float x, s1, c1, s2, c2;
        std::memcpy(&x, &u, 4);
        port::det_sincos(x, &s1, &c1); sincosf(x, &s2, &c2);
        if (std::memcmp(&s1, &s2, 4) || std::memcmp(&c1, &c2, 4)) { if (nbad < cap) bad[nbad] = u; ++nbad; }
    }
    return nbad;
}

}  // extern "C"
