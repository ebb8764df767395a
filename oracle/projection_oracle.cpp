// projection_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
//
// Line-by-line restatement of the reference's projection-guided searches over the 64x48 frame grid:
//   Frame::AssignFeaturesToGrid     src/Frame.cc:387-418      Frame::PosInGrid          src/Frame.cc:755-766
//   Frame::GetFeaturesInArea        src/Frame.cc:687-753      FRAME_GRID_ROWS/COLS      include/Frame.h:44-45
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th, bFarPoints, thFarPoints)     src/ORBmatcher1.cc:45-215
//   ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last, th, bMono)                src/ORBmatcher3.cc:256-467
//   ORBmatcher::SearchByProjection(Frame& Current, KeyFrame*, sAlreadyFound, th, ORBdist)       src/ORBmatcher3.cc:469-578
// Only the Nleft == -1 paths (one camera, rectified stereo, RGB-D) are restated; the two-fisheye rig (Nleft != -1) is out of
// scope.  Eigen/Sophus arithmetic (the projection of the map point into the frame) stays with the caller: the
// functions take the projected coordinates.  Parity pinning: PINNED — tests/test_ref_pin.py checks every function of this
// file against the reference's own AssignFeaturesToGrid / GetFeaturesInArea / PosInGrid / SearchByProjection code, compiled
// from /root/reference by oracle/build_ref.sh (oracle/_ref/libref.so), and tests/golden/ref_golden.json freezes its answers.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

constexpr int FRAME_GRID_ROWS = 48;   // include/Frame.h:44
constexpr int FRAME_GRID_COLS = 64;   // include/Frame.h:45
constexpr int TH_HIGH = 100;          // src/ORBmatcher1.cc:37
constexpr int HISTO_LENGTH = 30;      // src/ORBmatcher1.cc:39

struct KeyPoint {                     // cv::KeyPoint — include/OpenCL/Kernel/key_point.hpp:22-29
    float x, y, size, angle, response;
    int octave, class_id;
};

// ORBmatcher::DescriptorDistance — src/ORBmatcher3.cc:637-653
int DescriptorDistance(const uint8_t* a, const uint8_t* b)
{
    const int32_t* pa = (const int32_t*)a;
    const int32_t* pb = (const int32_t*)b;
    int dist = 0;
    for (int i = 0; i < 8; i++, pa++, pb++) {
        unsigned int v = *pa ^ *pb;
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// ORBmatcher::ComputeThreeMaxima — src/ORBmatcher3.cc:592-633
void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3)
{
    int max1 = 0, max2 = 0, max3 = 0;
    for (int i = 0; i < L; i++) {
        const int s = (int)histo[i].size();
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
}

// The slice of ORB_SLAM3::Frame these functions read (Nleft == -1).
struct Frame {
    int N = 0;
    const KeyPoint* mvKeysUn = nullptr;
    const uint8_t* mDescriptors = nullptr;
    const float* mvuRight = nullptr;                 // may be NULL = all -1
    const float* mvScaleFactors = nullptr;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    std::vector<size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS];

    // src/Frame.cc:755-766
    bool PosInGrid(const KeyPoint& kp, int& posX, int& posY) const
    {
        posX = round((kp.x - mnMinX) * mfGridElementWidthInv);
        posY = round((kp.y - mnMinY) * mfGridElementHeightInv);
        if (posX < 0 || posX >= FRAME_GRID_COLS || posY < 0 || posY >= FRAME_GRID_ROWS) return false;
        return true;
    }
    // src/Frame.cc:387-418
    void AssignFeaturesToGrid()
    {
        for (int i = 0; i < N; i++) {
            const KeyPoint& kp = mvKeysUn[i];
            int nGridPosX, nGridPosY;
            if (PosInGrid(kp, nGridPosX, nGridPosY)) mGrid[nGridPosX][nGridPosY].push_back(i);
        }
    }
    // src/Frame.cc:687-753
    std::vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1) const
    {
        std::vector<size_t> vIndices;
        float factorX = r;
        float factorY = r;
        const int nMinCellX = std::max(0, (int)floor((x - mnMinX - factorX) * mfGridElementWidthInv));
        if (nMinCellX >= FRAME_GRID_COLS) return vIndices;
        const int nMaxCellX = std::min((int)FRAME_GRID_COLS - 1, (int)ceil((x - mnMinX + factorX) * mfGridElementWidthInv));
        if (nMaxCellX < 0) return vIndices;
        const int nMinCellY = std::max(0, (int)floor((y - mnMinY - factorY) * mfGridElementHeightInv));
        if (nMinCellY >= FRAME_GRID_ROWS) return vIndices;
        const int nMaxCellY = std::min((int)FRAME_GRID_ROWS - 1, (int)ceil((y - mnMinY + factorY) * mfGridElementHeightInv));
        if (nMaxCellY < 0) return vIndices;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = nMinCellX; ix <= nMaxCellX; ix++) {
            for (int iy = nMinCellY; iy <= nMaxCellY; iy++) {
                const std::vector<size_t>& vCell = mGrid[ix][iy];
                if (vCell.empty()) continue;
                for (size_t j = 0, jend = vCell.size(); j < jend; j++) {
                    const KeyPoint& kpUn = mvKeysUn[vCell[j]];
                    if (bCheckLevels) {
                        if (kpUn.octave < minLevel) continue;
                        if (maxLevel >= 0)
                            if (kpUn.octave > maxLevel) continue;
                    }
                    const float distx = kpUn.x - x;
                    const float disty = kpUn.y - y;
                    if (fabs(distx) < factorX && fabs(disty) < factorY) vIndices.push_back(vCell[j]);
                }
            }
        }
        return vIndices;
    }
};

void fill_frame(Frame& F, const KeyPoint* kp, const uint8_t* desc, const float* uright, int n, const float* bounds_grid, const float* scale)
{
    F.N = n; F.mvKeysUn = kp; F.mDescriptors = desc; F.mvuRight = uright; F.mvScaleFactors = scale;
    F.mnMinX = bounds_grid[0]; F.mnMinY = bounds_grid[1]; F.mnMaxX = bounds_grid[2]; F.mnMaxY = bounds_grid[3];
    F.mfGridElementWidthInv = bounds_grid[4]; F.mfGridElementHeightInv = bounds_grid[5];
    F.AssignFeaturesToGrid();
}

// src/ORBmatcher1.cc:217-223
float RadiusByViewingCos(const float& viewCos)
{
    if (viewCos > 0.998) return 2.5;
    else return 4.0;
}

// The rotation-consistency tail shared by the two 1-NN variants (src/ORBmatcher3.cc:442-464, 555-575).
int apply_rotation_filter(std::vector<int>* rotHist, std::vector<int>& mvpMapPoints, int nmatches)
{
    int ind1 = -1, ind2 = -1, ind3 = -1;
    ComputeThreeMaxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
    for (int i = 0; i < HISTO_LENGTH; i++) {
        if (i != ind1 && i != ind2 && i != ind3) {
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { mvpMapPoints[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace

extern "C" {

// Frame::AssignFeaturesToGrid as CSR: cell id = ix * 48 + iy, cell_start[3073], items[n] (in-cell order = push_back order).
void orbo_assign_features_to_grid(const KeyPoint* kp, int n, const float* bounds_grid, int32_t* cell_start, int32_t* items)
{
    Frame F;
    fill_frame(F, kp, nullptr, nullptr, n, bounds_grid, nullptr);
    int t = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
        for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
            cell_start[ix * FRAME_GRID_ROWS + iy] = t;
            for (size_t v : F.mGrid[ix][iy]) items[t++] = (int32_t)v;
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = t;
}

// Frame::GetFeaturesInArea; returns the count, indices in the reference's order.
int orbo_get_features_in_area(const KeyPoint* kp, int n, const float* bounds_grid, float x, float y, float r, int minLevel, int maxLevel,
                              int32_t* out)
{
    Frame F;
    fill_frame(F, kp, nullptr, nullptr, n, bounds_grid, nullptr);
    const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size(); ++i) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, th, bFarPoints, thFarPoints) — src/ORBmatcher1.cc:45-215,
// Nleft == -1.  Map point i is described by mbTrackInView (in_view), isBad (bad), mTrackProjX/Y/XR, mTrackViewCos, mTrackDepth,
// mnTrackScaleLevel, Observations() and GetDescriptor().  mvpMapPoints[idx] = map-point index or -1; frame_obs[idx] != 0 <=>
// the frame's initial map point at idx has Observations() > 0.  Output mvpMapPoints after the call.
int orbo_search_by_projection_map(const KeyPoint* kp, const uint8_t* desc, const float* uright, const uint8_t* frame_obs, int n,
                                  const float* bounds_grid, const float* scale, const uint8_t* in_view, const uint8_t* bad,
                                  const float* projx, const float* projy, const float* projxr, const float* viewcos,
                                  const float* trackdepth, const int32_t* level, const int32_t* nobs, const uint8_t* mpdesc, int nmp,
                                  float th, int bFarPoints, float thFarPoints, float mfNNratio, int32_t* mvpMapPoints_out)
{
    Frame F;
    fill_frame(F, kp, desc, uright, n, bounds_grid, scale);
    // mvpMapPoints[idx]: -1 = NULL, -2 = the frame's own initial map point, >= 0 = map point iMP
    std::vector<int> mvpMapPoints(n, -1);
    auto observed = [&](int idx) {
        const int m = mvpMapPoints[idx];
        if (m == -1) return false;
        if (m == -2) return frame_obs[idx] != 0;
        return nobs[m] > 0;
    };
    for (int i = 0; i < n; ++i) if (frame_obs && frame_obs[i]) mvpMapPoints[i] = -2;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    for (int iMP = 0; iMP < nmp; iMP++) {
        if (!in_view[iMP]) continue;
        if (bFarPoints && trackdepth[iMP] > thFarPoints) continue;
        if (bad[iMP]) continue;
        const int nPredictedLevel = level[iMP];
        float r = RadiusByViewingCos(viewcos[iMP]);
        if (bFactor) r *= th;
        const std::vector<size_t> vIndices =
            F.GetFeaturesInArea(projx[iMP], projy[iMP], r * F.mvScaleFactors[nPredictedLevel], nPredictedLevel - 1, nPredictedLevel);
        if (!vIndices.empty()) {
            const uint8_t* MPdescriptor = mpdesc + (size_t)iMP * 32;
            int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
            for (auto vit = vIndices.begin(), vend = vIndices.end(); vit != vend; vit++) {
                const size_t idx = *vit;
                if (observed((int)idx)) continue;
                if (F.mvuRight && F.mvuRight[idx] > 0) {
                    const float er = fabs(projxr[iMP] - F.mvuRight[idx]);
                    if (er > r * F.mvScaleFactors[nPredictedLevel]) continue;
                }
                const int dist = DescriptorDistance(MPdescriptor, F.mDescriptors + idx * 32);
                if (dist < bestDist) {
                    bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.mvKeysUn[idx].octave; bestIdx = (int)idx;
                } else if (dist < bestDist2) {
                    bestLevel2 = F.mvKeysUn[idx].octave; bestDist2 = dist;
                }
            }
            if (bestDist <= TH_HIGH) {
                if (bestLevel == bestLevel2 && bestDist > mfNNratio * bestDist2) continue;
                if (bestLevel != bestLevel2 || bestDist <= mfNNratio * bestDist2) { mvpMapPoints[bestIdx] = iMP; nmatches++; }
            }
        }
    }
    for (int i = 0; i < n; ++i) mvpMapPoints_out[i] = mvpMapPoints[i] == -2 ? -1 : mvpMapPoints[i];
    return nmatches;
}

// SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) — src/ORBmatcher3.cc:256-467, Nleft == -1.
// Last-frame feature i: valid (pMP && !mvbOutlier[i]), (u, v) = mpCamera->project(Tcw * x3Dw), invzc = 1/x3Dc(2) (by the
// caller), octave and angle of the last frame's key point, Observations() of its map point, the map point's descriptor.
// cur_obs[idx] != 0 <=> CurrentFrame.mvpMapPoints[idx] is set and has Observations() > 0 on entry.
int orbo_search_by_projection_last(const KeyPoint* kp, const uint8_t* desc, const float* uright, const uint8_t* cur_obs, int n,
                                   const float* bounds_grid, const float* scale, float mbf, const uint8_t* valid, const float* u,
                                   const float* v, const float* invz, const int32_t* octave, const float* angle, const int32_t* nobs,
                                   const uint8_t* mpdesc, int nlast, float th, int bForward, int bBackward, int mbCheckOrientation,
                                   int32_t* mvpMapPoints_out)
{
    Frame CurrentFrame;
    fill_frame(CurrentFrame, kp, desc, uright, n, bounds_grid, scale);
    std::vector<int> mvpMapPoints(n, -1);
    for (int i = 0; i < n; ++i) if (cur_obs && cur_obs[i]) mvpMapPoints[i] = -2;
    auto observed = [&](int idx) {
        const int m = mvpMapPoints[idx];
        if (m == -1) return false;
        if (m == -2) return true;
        return nobs[m] > 0;
    };
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < nlast; i++) {
        if (!valid[i]) continue;
        const float invzc = invz[i];
        if (invzc < 0) continue;
        const float uv0 = u[i], uv1 = v[i];
        if (uv0 < CurrentFrame.mnMinX || uv0 > CurrentFrame.mnMaxX) continue;
        if (uv1 < CurrentFrame.mnMinY || uv1 > CurrentFrame.mnMaxY) continue;
        const int nLastOctave = octave[i];
        float radius = th * CurrentFrame.mvScaleFactors[nLastOctave];
        std::vector<size_t> vIndices2;
        if (bForward) vIndices2 = CurrentFrame.GetFeaturesInArea(uv0, uv1, radius, nLastOctave);
        else if (bBackward) vIndices2 = CurrentFrame.GetFeaturesInArea(uv0, uv1, radius, 0, nLastOctave);
        else vIndices2 = CurrentFrame.GetFeaturesInArea(uv0, uv1, radius, nLastOctave - 1, nLastOctave + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = mpdesc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (auto vit = vIndices2.begin(), vend = vIndices2.end(); vit != vend; vit++) {
            const size_t i2 = *vit;
            if (observed((int)i2)) continue;
            if (CurrentFrame.mvuRight && CurrentFrame.mvuRight[i2] > 0) {
                const float ur = uv0 - mbf * invzc;
                const float er = fabs(ur - CurrentFrame.mvuRight[i2]);
                if (er > radius) continue;
            }
            const int dist = DescriptorDistance(dMP, CurrentFrame.mDescriptors + i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
        }
        if (bestDist <= TH_HIGH) {
            mvpMapPoints[bestIdx2] = i;
            nmatches++;
            if (mbCheckOrientation) {
                float rot = angle[i] - CurrentFrame.mvKeysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (mbCheckOrientation) nmatches = apply_rotation_filter(rotHist, mvpMapPoints, nmatches);
    for (int i = 0; i < n; ++i) mvpMapPoints_out[i] = mvpMapPoints[i] == -2 ? -1 : mvpMapPoints[i];
    return nmatches;
}

// SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, sAlreadyFound, th, ORBdist) — src/ORBmatcher3.cc:469-578.
// Key-frame map point i: valid (pMP && !isBad() && !sAlreadyFound.count(pMP)), (u, v) projected by the caller, dist3D = |x3Dw - Ow|,
// min/max distance invariance, nPredictedLevel = pMP->PredictScale(dist3D, &CurrentFrame) (by the caller: it needs log()),
// the key frame's key-point angle.  cur_has[idx] != 0 <=> CurrentFrame.mvpMapPoints[idx] != NULL on entry (:533).
int orbo_search_by_projection_kf(const KeyPoint* kp, const uint8_t* desc, const uint8_t* cur_has, int n, const float* bounds_grid,
                                 const float* scale, const uint8_t* valid, const float* u, const float* v, const float* dist3d,
                                 const float* mindist, const float* maxdist, const int32_t* level, const float* angle,
                                 const uint8_t* mpdesc, int nkf, float th, int ORBdist, int mbCheckOrientation, int32_t* mvpMapPoints_out)
{
    Frame CurrentFrame;
    fill_frame(CurrentFrame, kp, desc, nullptr, n, bounds_grid, scale);
    std::vector<int> mvpMapPoints(n, -1);
    for (int i = 0; i < n; ++i) if (cur_has && cur_has[i]) mvpMapPoints[i] = -2;
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < nkf; i++) {
        if (!valid[i]) continue;
        const float uv0 = u[i], uv1 = v[i];
        if (uv0 < CurrentFrame.mnMinX || uv0 > CurrentFrame.mnMaxX) continue;
        if (uv1 < CurrentFrame.mnMinY || uv1 > CurrentFrame.mnMaxY) continue;
        const float d3 = dist3d[i];
        if (d3 < mindist[i] || d3 > maxdist[i]) continue;
        const int nPredictedLevel = level[i];
        const float radius = th * CurrentFrame.mvScaleFactors[nPredictedLevel];
        const std::vector<size_t> vIndices2 = CurrentFrame.GetFeaturesInArea(uv0, uv1, radius, nPredictedLevel - 1, nPredictedLevel + 1);
        if (vIndices2.empty()) continue;
        const uint8_t* dMP = mpdesc + (size_t)i * 32;
        int bestDist = 256, bestIdx2 = -1;
        for (auto vit = vIndices2.begin(); vit != vIndices2.end(); vit++) {
            const size_t i2 = *vit;
            if (mvpMapPoints[i2] != -1) continue;
            const int dist = DescriptorDistance(dMP, CurrentFrame.mDescriptors + i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = (int)i2; }
        }
        if (bestDist <= ORBdist) {
            mvpMapPoints[bestIdx2] = i;
            nmatches++;
            if (mbCheckOrientation) {
                float rot = angle[i] - CurrentFrame.mvKeysUn[bestIdx2].angle;
                if (rot < 0.0) rot += 360.0f;
                int bin = round(rot * factor);
                if (bin == HISTO_LENGTH) bin = 0;
                rotHist[bin].push_back(bestIdx2);
            }
        }
    }
    if (mbCheckOrientation) nmatches = apply_rotation_filter(rotHist, mvpMapPoints, nmatches);
    for (int i = 0; i < n; ++i) mvpMapPoints_out[i] = mvpMapPoints[i] == -2 ? -1 : mvpMapPoints[i];
    return nmatches;
}

}  // extern "C"
