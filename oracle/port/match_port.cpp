// match_port.cpp -- TEST INFRASTRUCTURE ONLY (oracle "port"): plain restatement of the reference's matcher path.
//   ORBmatcher::DescriptorDistance / SearchForInitialization / SearchByProjection x2 / ComputeThreeMaxima
//     (/root/reference/src/ORBmatcher.cc:1913-1933, 515-643, 1569-1728, 70-175, 1866-1908)
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea / ComputeStereoMatches
//     (/root/reference/src/Frame.cc:431-461, 1007-1030, 894-1003, 1179-1573)
//   Frame::UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD   (/root/reference/src/Frame.cc:1052-1176, 1576-1614)
// Pinned against oracle/_ref (the reference's own bodies) by tests/test_oracle_matcher.py and by the golden
// vectors in tests/golden/ref_match.npz.  Same argument layout as the product's C ABI (include/orbx_b200.h).
#include "../cvlite/cvlite.hpp"
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

namespace port {

typedef cv::KeyPoint KeyPoint;
static const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 30;    // ORBmatcher.cc:49-51
static const int GRID_COLS = 64, GRID_ROWS = 48;                   // Frame.h:56-61

struct FrameView {
    int n; const KeyPoint* keys_un; const unsigned char* descriptors; const float* u_right;
    float min_x, min_y, max_x, max_y, gw_inv, gh_inv; int nlevels; const float* scale_factors;
};

// ORBmatcher::DescriptorDistance   ORBmatcher.cc:1913-1933
static inline int DescriptorDistance(const unsigned char* a, const unsigned char* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        unsigned int x, y; std::memcpy(&x, a + 4 * i, 4); std::memcpy(&y, b + 4 * i, 4);
        unsigned int v = x ^ y;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    const FrameView* v;
    // Frame::AssignFeaturesToGrid + PosInGrid   Frame.cc:431-461, 1007-1030
    explicit Grid(const FrameView* fv) : v(fv) {
        for (int i = 0; i < v->n; ++i) {
            const KeyPoint& kp = v->keys_un[i];
            int posX = (int)std::round((kp.pt.x - v->min_x) * v->gw_inv);      // round(): half away from zero, double
            int posY = (int)std::round((kp.pt.y - v->min_y) * v->gh_inv);
            if (posX < 0 || posX >= GRID_COLS || posY < 0 || posY >= GRID_ROWS) continue;
            cell[posX][posY].push_back(i);
        }
    }
    // Frame::GetFeaturesInArea   Frame.cc:894-1003
    void area(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
        out.clear();
        const int nMinCellX = std::max(0, (int)std::floor((x - v->min_x - r) * v->gw_inv));
        if (nMinCellX >= GRID_COLS) return;
        const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - v->min_x + r) * v->gw_inv));
        if (nMaxCellX < 0) return;
        const int nMinCellY = std::max(0, (int)std::floor((y - v->min_y - r) * v->gh_inv));
        if (nMinCellY >= GRID_ROWS) return;
        const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - v->min_y + r) * v->gh_inv));
        if (nMaxCellY < 0) return;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<int>& vCell = cell[ix][iy];
                for (size_t j = 0; j < vCell.size(); j++) {
                    const KeyPoint& kpUn = v->keys_un[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0 && kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.pt.x - x, disty = kpUn.pt.y - y;
                    if (std::fabs(distx) < r && std::fabs(disty) < r) out.push_back(vCell[j]);
                }
            }
    }
};

// ORBmatcher::ComputeThreeMaxima   ORBmatcher.cc:1866-1908
static void ComputeThreeMaxima(const std::vector<int>* histo, int L, int& ind1, int& ind2, int& ind3) {
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) ind3 = -1;
}
static inline int rot_bin(float a1, float a2) {      // :597-602
    const float factor = HISTO_LENGTH / 360.0f;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

// ORBmatcher::SearchForInitialization   ORBmatcher.cc:515-643
static int SearchForInitialization(float nnratio, bool checkOri, const FrameView* F1, const FrameView* F2, float* prev_xy, int* vnMatches12, int windowSize) {
    int nmatches = 0;
    for (int i = 0; i < F1->n; ++i) vnMatches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> vMatchedDistance(F2->n, INT_MAX), vnMatches21(F2->n, -1);
    Grid g2(F2);
    std::vector<int> vIndices2;
    for (int i1 = 0; i1 < F1->n; i1++) {
        const KeyPoint& kp1 = F1->keys_un[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        g2.area(prev_xy[2 * i1], prev_xy[2 * i1 + 1], (float)windowSize, level1, level1, vIndices2);
        if (vIndices2.empty()) continue;
        const unsigned char* d1 = F1->descriptors + (size_t)i1 * 32;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); ++k) {
            const int i2 = vIndices2[k];
            const int dist = DescriptorDistance(d1, F2->descriptors + (size_t)i2 * 32);
            if (vMatchedDistance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW) {
            if (bestDist < (float)bestDist2 * nnratio) {
                if (vnMatches21[bestIdx2] >= 0) { vnMatches12[vnMatches21[bestIdx2]] = -1; nmatches--; }
                vnMatches12[i1] = bestIdx2; vnMatches21[bestIdx2] = i1; vMatchedDistance[bestIdx2] = bestDist; nmatches++;
                if (checkOri) rotHist[rot_bin(F1->keys_un[i1].angle, F2->keys_un[bestIdx2].angle)].push_back(i1);
            }
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0; j < rotHist[i].size(); j++) { int idx1 = rotHist[i][j]; if (vnMatches12[idx1] >= 0) { vnMatches12[idx1] = -1; nmatches--; } }
        }
    }
    for (int i1 = 0; i1 < F1->n; i1++)
        if (vnMatches12[i1] >= 0) { prev_xy[2 * i1] = F2->keys_un[vnMatches12[i1]].pt.x; prev_xy[2 * i1 + 1] = F2->keys_un[vnMatches12[i1]].pt.y; }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)   ORBmatcher.cc:1569-1728
// (projection u, v, invz computed by the caller as at :1605-1623; `valid` = has a non-outlier map point,
//  invz >= 0 and (u,v) inside the image bounds)
static int SearchByProjectionFrame(bool checkOri, const FrameView* cur, int n_last, const float* proj_uv, const float* proj_invz, const int* last_octave,
                                   const float* last_angle, const unsigned char* mp_desc, const unsigned char* valid, const unsigned char* mp_observed,
                                   const unsigned char* cur_occupied, float th, bool bForward, bool bBackward, float mbf, int* cur_match) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    // occupied[j]: CurrentFrame.mvpMapPoints[j] is set AND that map point has Observations() > 0 (:1658-1660).
    // Pre-existing holders come from cur_occupied; a point assigned in this loop blocks later ones iff mp_observed[i].
    std::vector<unsigned char> occupied(cur->n, 0);
    for (int j = 0; j < cur->n; ++j) { cur_match[j] = -1; occupied[j] = cur_occupied ? (cur_occupied[j] != 0) : 0; }
    Grid g(cur);
    std::vector<int> vIndices2;
    for (int i = 0; i < n_last; i++) {
        if (!valid[i]) continue;
        const float u = proj_uv[2 * i], v = proj_uv[2 * i + 1], invzc = proj_invz[i];
        if (invzc < 0) continue;
        if (u < cur->min_x || u > cur->max_x) continue;
        if (v < cur->min_y || v > cur->max_y) continue;
        const int nLastOctave = last_octave[i];
        const float radius = th * cur->scale_factors[nLastOctave];
        if (bForward) g.area(u, v, radius, nLastOctave, -1, vIndices2);
        else if (bBackward) g.area(u, v, radius, 0, nLastOctave, vIndices2);
        else g.area(u, v, radius, nLastOctave - 1, nLastOctave + 1, vIndices2);
        if (vIndices2.empty()) continue;
        const unsigned char* dMP = mp_desc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < vIndices2.size(); ++k) {
            const int i2 = vIndices2[k];
            if (occupied[i2]) continue;
            if (cur->u_right && cur->u_right[i2] > 0) {
                const float ur = u - mbf * invzc;
                const float er = std::fabs(ur - cur->u_right[i2]);
                if (er > radius) continue;
            }
            const int dist = DescriptorDistance(dMP, cur->descriptors + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            cur_match[bestIdx2] = i;                                  // may overwrite an unobserved earlier holder
            occupied[bestIdx2] = mp_observed ? (mp_observed[i] != 0) : 0;
            nmatches++;
            if (checkOri) rotHist[rot_bin(last_angle[i], cur->keys_un[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) { cur_match[rotHist[i][j]] = -1; nmatches--; }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(Frame& Cur, KeyFrame*, const set<MapPoint*>&, th, ORBdist)   ORBmatcher.cc:1731-1863
// (per KeyFrame map point the caller supplies the projection, the predicted level and valid = not NULL / bad / already found,
//  inside the image, within its distance range)
static int SearchByProjectionKeyFrame(bool checkOri, const FrameView* cur, int n_kf, const float* proj_uv, const int* predicted_level, const float* kf_angle,
                                      const unsigned char* mp_desc, const unsigned char* valid, const unsigned char* cur_occupied, float th, int ORBdist, int* cur_match) {
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<unsigned char> taken(cur->n, 0);                       // CurrentFrame.mvpMapPoints[i2] != NULL, before or during the call (:1808)
    for (int j = 0; j < cur->n; ++j) { cur_match[j] = -1; taken[j] = cur_occupied ? (cur_occupied[j] != 0) : 0; }
    Grid g(cur);
    std::vector<int> cands;
    for (int i = 0; i < n_kf; i++) {
        if (!valid[i]) continue;
        const float u = proj_uv[2 * i], v = proj_uv[2 * i + 1];
        if (u < cur->min_x || u > cur->max_x || v < cur->min_y || v > cur->max_y) continue;
        const int lvl = predicted_level[i];
        g.area(u, v, th * cur->scale_factors[lvl], lvl - 1, lvl + 1, cands);
        int bestDist = 256, bestIdx2 = -1;
        for (size_t k = 0; k < cands.size(); ++k) {
            if (taken[cands[k]]) continue;
            const int dist = DescriptorDistance(mp_desc + (size_t)i * 32, cur->descriptors + (size_t)cands[k] * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = cands[k]; }
        }
        if (bestDist <= ORBdist) {
            cur_match[bestIdx2] = i; taken[bestIdx2] = 1; nmatches++;
            if (checkOri) rotHist[rot_bin(kf_angle[i], cur->keys_un[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (checkOri) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++)
            if (i != ind1 && i != ind2 && i != ind3)
                for (size_t j = 0; j < rotHist[i].size(); j++) { cur_match[rotHist[i][j]] = -1; nmatches--; }
    }
    return nmatches;
}

// ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)   ORBmatcher.cc:388-512  (KeyFrame::GetFeaturesInArea,
// src/KeyFrame.cc:752-797, walks the same grid in the same order as Frame's, without a level filter; the level test is at :462-463)
static int SearchByProjectionKeyFramePoints(const FrameView* kf, int n_points, const float* proj_uv, const int* predicted_level, const unsigned char* mp_desc,
                                            const unsigned char* valid, const unsigned char* kf_matched, float th, int* kf_match) {
    int nmatches = 0;
    std::vector<unsigned char> taken(kf->n, 0);
    for (int j = 0; j < kf->n; ++j) { kf_match[j] = -1; taken[j] = kf_matched ? (kf_matched[j] != 0) : 0; }
    Grid g(kf);
    std::vector<int> cands;
    for (int p = 0; p < n_points; ++p) {
        if (!valid[p]) continue;
        const int lvl = predicted_level[p];
        g.area(proj_uv[2 * p], proj_uv[2 * p + 1], th * kf->scale_factors[lvl], -1, -1, cands);
        int bestDist = 256, bestIdx = -1;
        for (size_t k = 0; k < cands.size(); ++k) {
            const int idx = cands[k];
            if (taken[idx]) continue;
            const int kpLevel = kf->keys_un[idx].octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            const int dist = DescriptorDistance(mp_desc + (size_t)p * 32, kf->descriptors + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) { kf_match[bestIdx] = p; taken[bestIdx] = 1; nmatches++; }
    }
    return nmatches;
}

// ORBmatcher::SearchBySim3   ORBmatcher.cc:1290-1555: one pass per direction (independent best feature on level predicted - 1 or predicted,
// distance <= TH_HIGH), then the pairs that agree both ways
static void Sim3Pass(const FrameView* target, int nq, const float* proj_uv, const int* predicted_level, const unsigned char* mp_desc, const unsigned char* valid, float th, std::vector<int>& best) {
    best.assign(nq, -1);
    Grid g(target);
    std::vector<int> cands;
    for (int q = 0; q < nq; ++q) {
        if (!valid[q]) continue;
        const int lvl = predicted_level[q];
        g.area(proj_uv[2 * q], proj_uv[2 * q + 1], th * target->scale_factors[lvl], -1, -1, cands);
        int bestDist = INT_MAX, bestIdx = -1;
        for (size_t k = 0; k < cands.size(); ++k) {
            const int oct = target->keys_un[cands[k]].octave;
            if (oct < lvl - 1 || oct > lvl) continue;
            const int dist = DescriptorDistance(mp_desc + (size_t)q * 32, target->descriptors + (size_t)cands[k] * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = cands[k]; }
        }
        if (bestDist <= TH_HIGH) best[q] = bestIdx;
    }
}

// The search inside ORBmatcher::Fuse (pose form :1020-1175 with its chi-square gate when proj_ur is given, Sim3 form :1179-1310 without)
static void FuseSearch(const FrameView* kf, int n_points, const float* proj_uv, const float* proj_ur, const int* predicted_level, const unsigned char* mp_desc,
                       const unsigned char* valid, const float* inv_sigma2, float th, int* best_idx) {
    Grid g(kf);
    std::vector<int> cands;
    for (int p = 0; p < n_points; ++p) {
        best_idx[p] = -1;
        if (!valid[p]) continue;
        const float u = proj_uv[2 * p], v = proj_uv[2 * p + 1];
        const int lvl = predicted_level[p];
        g.area(u, v, th * kf->scale_factors[lvl], -1, -1, cands);
        int bestDist = 256, bestIdx = -1;
        for (size_t k = 0; k < cands.size(); ++k) {
            const int idx = cands[k];
            const KeyPoint& kp = kf->keys_un[idx];
            if (kp.octave < lvl - 1 || kp.octave > lvl) continue;
            if (proj_ur) {
                volatile float ex = u - kp.pt.x, ey = v - kp.pt.y;
                if (kf->u_right && kf->u_right[idx] >= 0) {
                    volatile float er = proj_ur[p] - kf->u_right[idx];
                    volatile float e2 = ex * ex + ey * ey + er * er, s = e2 * inv_sigma2[kp.octave];
                    if ((double)s > 7.8) continue;
                } else {
                    volatile float e2 = ex * ex + ey * ey, s = e2 * inv_sigma2[kp.octave];
                    if ((double)s > 5.99) continue;
                }
            }
            const int dist = DescriptorDistance(mp_desc + (size_t)p * 32, kf->descriptors + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW) best_idx[p] = bestIdx;
    }
}

// ORBmatcher::SearchByProjection(Frame& F, const vector<MapPoint*>&, th)   ORBmatcher.cc:70-175
static int SearchByProjectionPoints(float nnratio, const FrameView* F, int n_points, const float* track_uv, const float* track_ur, const int* track_level,
                                    const float* track_view_cos, const unsigned char* mp_desc, const unsigned char* mp_observed, const unsigned char* f_occupied,
                                    float th, int* f_match) {
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<unsigned char> occupied(F->n, 0);
    for (int j = 0; j < F->n; ++j) { f_match[j] = -1; occupied[j] = f_occupied ? (f_occupied[j] != 0) : 0; }
    Grid g(F);
    std::vector<int> vIndices;
    for (int iMP = 0; iMP < n_points; iMP++) {
        const int nPredictedLevel = track_level[iMP];
        float r = track_view_cos[iMP] > 0.998 ? 2.5f : 4.0f;          // RadiusByViewingCos :178-185
        if (bFactor) r *= th;
        g.area(track_uv[2 * iMP], track_uv[2 * iMP + 1], r * F->scale_factors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel, vIndices);
        if (vIndices.empty()) continue;
        const unsigned char* d = mp_desc + (size_t)iMP * 32;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (size_t k = 0; k < vIndices.size(); ++k) {
            const int idx = vIndices[k];
            if (occupied[idx]) continue;                              // mvpMapPoints[idx] with Observations() > 0 (:124-126)
            if (F->u_right && F->u_right[idx] > 0) {
                const float er = std::fabs(track_ur[iMP] - F->u_right[idx]);
                if (er > r * F->scale_factors[nPredictedLevel]) continue;
            }
            const int dist = DescriptorDistance(d, F->descriptors + (size_t)idx * 32);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F->keys_un[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = F->keys_un[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            f_match[bestIdx] = iMP;                                   // may overwrite an unobserved earlier holder
            occupied[bestIdx] = mp_observed ? (mp_observed[iMP] != 0) : 0;
            nmatches++;
        }
    }
    return nmatches;
}

struct Pyr { int nlevels; std::vector<int> w, h; std::vector<const unsigned char*> px; };

// Frame::ComputeStereoMatches   Frame.cc:1179-1573
static void ComputeStereoMatches(const Pyr& PL, const Pyr& PR, const float* scale, const float* inv_scale,
                                 const KeyPoint* keysL, const unsigned char* descL, int N, const KeyPoint* keysR, const unsigned char* descR, int Nr,
                                 float mb, float mbf, float* mvuRight, float* mvDepth) {
    for (int i = 0; i < N; ++i) { mvuRight[i] = -1.0f; mvDepth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = PL.h[0];
    std::vector<std::vector<int> > vRowIndices(nRows);
    for (int iR = 0; iR < Nr; iR++) {
        const KeyPoint& kp = keysR[iR];
        const float kpY = kp.pt.y;
        const float r = 2.0f * scale[kp.octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++) if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);   // (reference indexes unchecked)
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;      // mb == 0 at call time => maxD = +inf
    std::vector<std::pair<int, int> > vDistIdx;
    for (int iL = 0; iL < N; iL++) {
        const KeyPoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.pt.y, uL = kpL.pt.x;
        const int row = (int)vL;
        if (row < 0 || row >= nRows) continue;
        const std::vector<int>& vCandidates = vRowIndices[row];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH; size_t bestIdxR = 0;
        const unsigned char* dL = descL + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const int iR = vCandidates[iC];
            const KeyPoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.pt.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = DescriptorDistance(dL, descR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = keysR[bestIdxR].pt.x;
            const float scaleFactor = inv_scale[kpL.octave];
            const float scaleduL = std::round(kpL.pt.x * scaleFactor), scaledvL = std::round(kpL.pt.y * scaleFactor), scaleduR0 = std::round(uR0 * scaleFactor);
            const int w = 5, L = 5;
            const int lv = kpL.octave;
            const unsigned char* IL = PL.px[lv]; const unsigned char* IR = PR.px[lv];
            const int wl = PL.w[lv], wr = PR.w[lv];
            const int cy = (int)scaledvL, cxl = (int)scaleduL, cxr = (int)scaleduR0;
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= wr) continue;
            int bestDistS = INT_MAX, bestincR = 0;
            float vDists[2 * 5 + 1];
            const int cl = IL[(size_t)cy * wl + cxl];
            for (int incR = -L; incR <= +L; incR++) {
                const int cr = IR[(size_t)cy * wr + cxr + incR];
                // cv::norm(IL - IL(w,w), IR - IR(w,w), NORM_L1) over the 11x11 windows: integers, exact
                float dist = 0;
                { long long s = 0;
                  for (int dy = -w; dy <= w; ++dy) for (int dx = -w; dx <= w; ++dx) {
                      const int a = (int)IL[(size_t)(cy + dy) * wl + cxl + dx] - cl, b = (int)IR[(size_t)(cy + dy) * wr + cxr + incR + dx] - cr;
                      s += std::abs(a - b);
                  }
                  dist = (float)s; }
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; bestuR = uL - 0.01f; }
                mvDepth[iL] = mbf / disparity; mvuRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return;      // (reference reads vDistIdx[0] of an empty vector: undefined)
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        mvuRight[vDistIdx[i].second] = -1; mvDepth[vDistIdx[i].second] = -1;
    }
}

// ---- the frame steps between extractor and matchers ----
struct Camera { float fx, fy, cx, cy, k1, k2, p1, p2, k3, bf; };

// cv::undistortPoints(pts, pts, mK, mDistCoef, Mat(), mK) for one point, written out for the 5-coefficient model
// (cvlite.hpp "Primitive 7" is the general restatement; both are pinned against cv2 goldens)
static CVL_NOFMA void undistort_one(const Camera& c, float uf, float vf, float& ox, float& oy) {
    const double fx = c.fx, fy = c.fy, cx = c.cx, cy = c.cy, k1 = c.k1, k2 = c.k2, p1 = c.p1, p2 = c.p2, k3 = c.k3;
    const double u = uf, v = vf;
    volatile double x = (u - cx) * (1. / fx), y = (v - cy) * (1. / fy);
    const double x0 = x, y0 = y;
    for (int it = 0; it < 5; ++it) {
        volatile double r2 = x * x + y * y;
        volatile double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);      // numerator 1 + ((0 r2 + 0) r2 + 0) r2 == 1 exactly
        if (icdist < 0) { x = x0; y = y0; break; }
        volatile double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);              // the zero thin-prism terms add +0.0
        volatile double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
        x = (x0 - dX) * icdist; y = (y0 - dY) * icdist;
    }
    volatile double xx = fx * x + cx, yy = fy * y + cy;                            // + 0 * y, * (1 / 1): exact no-ops
    ox = (float)xx; oy = (float)yy;
}

// Frame.cc:1120-1176 + :301-302
static void ImageBounds(const Camera& c, int rows, int cols, float b[6]) {
    if (c.k1 != 0.0f) {
        float x[4], y[4];
        undistort_one(c, 0.f, 0.f, x[0], y[0]); undistort_one(c, (float)cols, 0.f, x[1], y[1]);
        undistort_one(c, 0.f, (float)rows, x[2], y[2]); undistort_one(c, (float)cols, (float)rows, x[3], y[3]);
        b[0] = std::min(x[0], x[2]); b[1] = std::max(x[1], x[3]); b[2] = std::min(y[0], y[1]); b[3] = std::max(y[2], y[3]);
    } else { b[0] = 0.f; b[1] = (float)cols; b[2] = 0.f; b[3] = (float)rows; }
    b[4] = (float)GRID_COLS / (b[1] - b[0]); b[5] = (float)GRID_ROWS / (b[3] - b[2]);
}

// UndistortKeyPoints (:1052-1117), ComputeStereoFromRGBD (:1576-1614), AssignFeaturesToGrid (:431-461)
static void FrameBuild(const KeyPoint* keys, int n, const Camera& c, int rows, int cols, const float* depth_img,
                       KeyPoint* keys_un, float* u_right, float* depth_out, float* bounds, int* cell_start, int* entries) {
    ImageBounds(c, rows, cols, bounds);
    for (int i = 0; i < n; ++i) {
        keys_un[i] = keys[i];
        if (c.k1 != 0.0f) undistort_one(c, keys[i].pt.x, keys[i].pt.y, keys_un[i].pt.x, keys_un[i].pt.y);
        u_right[i] = -1.f; depth_out[i] = -1.f;
        if (depth_img) {
            const float d = depth_img[(size_t)(int)keys[i].pt.y * cols + (int)keys[i].pt.x];      // at<float>(v, u): truncation
            if (d > 0) { depth_out[i] = d; volatile float q = c.bf / d; u_right[i] = keys_un[i].pt.x - q; }
        }
    }
    FrameView v; v.n = n; v.keys_un = keys_un; v.descriptors = nullptr; v.u_right = u_right;
    v.min_x = bounds[0]; v.max_x = bounds[1]; v.min_y = bounds[2]; v.max_y = bounds[3]; v.gw_inv = bounds[4]; v.gh_inv = bounds[5];
    v.nlevels = 0; v.scale_factors = nullptr;
    Grid g(&v);
    int o = 0;
    for (int x = 0; x < GRID_COLS; ++x) for (int y = 0; y < GRID_ROWS; ++y) {
        cell_start[x * GRID_ROWS + y] = o;
        for (size_t q = 0; q < g.cell[x][y].size(); ++q) entries[o++] = g.cell[x][y][q];
    }
    cell_start[GRID_COLS * GRID_ROWS] = o;
}

}  // namespace port

extern "C" {
int port_frame_build(const cv::KeyPoint* keys, int n, const float* cam9, int ndist, float bf, int rows, int cols, const float* depth_img,
                     cv::KeyPoint* keys_un, float* u_right, float* depth_out, float* bounds, int* cell_start, int* entries) {
    port::Camera c = {cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], ndist > 4 ? cam9[8] : 0.f, bf};
    port::FrameBuild(keys, n, c, rows, cols, depth_img, keys_un, u_right, depth_out, bounds, cell_start, entries);
    return 0;
}
void port_undistort_points(const float* pts, int n, const float* cam9, int ndist, float* out) {
    port::Camera c = {cam9[0], cam9[1], cam9[2], cam9[3], cam9[4], cam9[5], cam9[6], cam9[7], ndist > 4 ? cam9[8] : 0.f, 0.f};
    for (int i = 0; i < n; ++i) port::undistort_one(c, pts[2 * i], pts[2 * i + 1], out[2 * i], out[2 * i + 1]);
}
using port::FrameView; using port::KeyPoint;

void port_descriptor_distance(const unsigned char* a, const unsigned char* b, int n, int* out) { for (int i = 0; i < n; ++i) out[i] = port::DescriptorDistance(a + (size_t)i * 32, b + (size_t)i * 32); }
int port_get_features_in_area(const FrameView* v, float x, float y, float r, int minLevel, int maxLevel, int* out, int cap) {
    port::Grid g(v); std::vector<int> idx; g.area(x, y, r, minLevel, maxLevel, idx);
    for (int i = 0; i < (int)idx.size() && i < cap; ++i) out[i] = idx[i];
    return (int)idx.size();
}
int port_search_for_initialization(float nnratio, int checkOri, const FrameView* v1, const FrameView* v2, float* prev_matched_xy, int* matches12, int windowSize) {
    return port::SearchForInitialization(nnratio, checkOri != 0, v1, v2, prev_matched_xy, matches12, windowSize);
}
int port_search_by_projection_frame(float nnratio, int checkOri, const FrameView* cur, int n_last, const float* proj_uv, const float* proj_invz, const int* last_octave,
                                    const float* last_angle, const unsigned char* mp_desc, const unsigned char* valid, const unsigned char* mp_observed,
                                    const unsigned char* cur_occupied, float th, int forward, int backward, float mbf, int* cur_match) {
    (void)nnratio;
    return port::SearchByProjectionFrame(checkOri != 0, cur, n_last, proj_uv, proj_invz, last_octave, last_angle, mp_desc, valid, mp_observed, cur_occupied, th, forward != 0, backward != 0, mbf, cur_match);
}
int port_search_by_projection_keyframe(float nnratio, int checkOri, const FrameView* cur, int n_kf, const float* proj_uv, const int* predicted_level, const float* kf_angle,
                                       const unsigned char* mp_desc, const unsigned char* valid, const unsigned char* cur_occupied, float th, int orb_dist, int* cur_match) {
    (void)nnratio;
    return port::SearchByProjectionKeyFrame(checkOri != 0, cur, n_kf, proj_uv, predicted_level, kf_angle, mp_desc, valid, cur_occupied, th, orb_dist, cur_match);
}
int port_search_by_projection_keyframe_points(const FrameView* kf, int n_points, const float* proj_uv, const int* predicted_level, const unsigned char* mp_desc,
                                              const unsigned char* valid, const unsigned char* kf_matched, float th, int* kf_match) {
    return port::SearchByProjectionKeyFramePoints(kf, n_points, proj_uv, predicted_level, mp_desc, valid, kf_matched, th, kf_match);
}
int port_search_by_sim3(const FrameView* kf1, const FrameView* kf2, const float* proj_uv1, const int* level1, const unsigned char* desc1, const unsigned char* valid1,
                        const float* proj_uv2, const int* level2, const unsigned char* desc2, const unsigned char* valid2, float th, int* match12) {
    std::vector<int> m1, m2;
    port::Sim3Pass(kf2, kf1->n, proj_uv1, level1, desc1, valid1, th, m1);
    port::Sim3Pass(kf1, kf2->n, proj_uv2, level2, desc2, valid2, th, m2);
    int nFound = 0;
    for (int i1 = 0; i1 < kf1->n; ++i1) {
        match12[i1] = -1;
        if (m1[i1] >= 0 && m2[m1[i1]] == i1) { match12[i1] = m1[i1]; ++nFound; }
    }
    return nFound;
}
void port_fuse_search(const FrameView* kf, int n_points, const float* proj_uv, const float* proj_ur, const int* predicted_level, const unsigned char* mp_desc,
                      const unsigned char* valid, const float* inv_sigma2, float th, int* best_idx) {
    port::FuseSearch(kf, n_points, proj_uv, proj_ur, predicted_level, mp_desc, valid, inv_sigma2, th, best_idx);
}
int port_search_by_projection_points(float nnratio, int checkOri, const FrameView* F, int n_points, const float* track_uv, const float* track_ur, const int* track_level,
                                     const float* track_view_cos, const unsigned char* mp_desc, const unsigned char* mp_observed, const unsigned char* f_occupied,
                                     float th, int* f_match) {
    (void)checkOri;
    return port::SearchByProjectionPoints(nnratio, F, n_points, track_uv, track_ur, track_level, track_view_cos, mp_desc, mp_observed, f_occupied, th, f_match);
}
// pyramids: levels packed tightly, level l at px_l with (w_l, h_l)
int port_compute_stereo_matches(int nlevels, const int* wl, const int* hl, const unsigned char* const* pxl, const int* wr, const int* hr, const unsigned char* const* pxr,
                                const float* scale, const float* inv_scale, const KeyPoint* keys_left, const unsigned char* desc_left, int nl,
                                const KeyPoint* keys_right, const unsigned char* desc_right, int nr, float mb, float mbf, float* u_right, float* depth) {
    port::Pyr L, R; L.nlevels = R.nlevels = nlevels;
    for (int i = 0; i < nlevels; ++i) { L.w.push_back(wl[i]); L.h.push_back(hl[i]); L.px.push_back(pxl[i]); R.w.push_back(wr[i]); R.h.push_back(hr[i]); R.px.push_back(pxr[i]); }
    port::ComputeStereoMatches(L, R, scale, inv_scale, keys_left, desc_left, nl, keys_right, desc_right, nr, mb, mbf, u_right, depth);
    return 0;
}
}  // extern "C"
