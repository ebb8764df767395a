// ref_wrapper.cpp — C entry points over the REFERENCE's own functions (sliced into oracle/_ref/gen/*.inc by oracle/build_ref.sh).
// TEST INFRASTRUCTURE ONLY.  Nothing in this file restates reference logic: it declares the few classes the slices are members
// of (data members only, names as in the reference headers), includes the slices, and converts plain arrays to and from them.
#include <climits>
#include <map>
#include <set>
#include <stdexcept>
#include <tuple>

#include "cv_shim.h"

// ---- stand-ins for what the slices mention but this path never executes ------------------------------------------------
namespace opencl { struct Manager { static Manager& the() { static Manager m; return m; } }; }
namespace Eigen {
struct Vector3f {
    float v[3] = {0, 0, 0};
    Vector3f() {}
    Vector3f(float a, float b, float c) { v[0] = a; v[1] = b; v[2] = c; }
    float operator()(int i) const { return v[i]; }
    Vector3f operator-(const Vector3f& o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    float norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Matrix3f {
    float m[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    float operator()(int i, int j) const { return m[i][j]; }
};
struct Vector2f { float v[2] = {0, 0}; Vector2f() {} Vector2f(float a, float b) { v[0] = a; v[1] = b; } float operator()(int i) const { return v[i]; } };
}  // namespace Eigen
namespace Sophus {
// translation-only rigid motion: enough to drive bForward / bBackward; the test harness uses identity rotation so that the
// projected coordinates handed to the oracle are exactly the ones the reference code computes
struct SE3f {
    Eigen::Vector3f t;
    SE3f inverse() const { SE3f r; r.t = Eigen::Vector3f(-t(0), -t(1), -t(2)); return r; }
    Eigen::Vector3f translation() const { return t; }
    Eigen::Matrix3f rotationMatrix() const { return Eigen::Matrix3f(); }
    SE3f operator*(const SE3f& o) const { SE3f r; r.t = Eigen::Vector3f(t(0) + o.t(0), t(1) + o.t(1), t(2) + o.t(2)); return r; }
    Eigen::Vector3f operator*(const Eigen::Vector3f& p) const { return Eigen::Vector3f(p(0) + t(0), p(1) + t(1), p(2) + t(2)); }
};
}  // namespace Sophus

namespace DBoW2 { class FeatureVector : public std::map<unsigned int, std::vector<unsigned int>> {}; }   // Thirdparty/DBoW2/DBoW2/FeatureVector.h:23-25

#define FRAME_GRID_ROWS 48   /* include/Frame.h:44 (value checked by tests against the slice's behaviour) */
#define FRAME_GRID_COLS 64   /* include/Frame.h:45 */
#define TO_SIZE_T(x) (x > 0 ? x : 0)   /* src/ORBmatcher1.cc:30 */

using namespace std;
using cv::KeyPoint; using cv::Mat; using cv::Point; using cv::Point2f; using cv::InputArray;

namespace ORB_SLAM3 {

// ===== extractor =====================================================================================================
#include "ORBextractor_classes.inc"
#include "orb_constants_descriptor.inc"
#include "orb_bit_pattern.inc"
#include "orb_ctor.inc"
#include "orb_divide_compare.inc"
#include "orb_distribute_octree.inc"
#include "orb_tile_calc_keypoints.inc"

struct RefExtractor : ORBextractor {                    // opens the protected members
    using ORBextractor::ORBextractor;
    using ORBextractor::DistributeOctTree;
    using ORBextractor::mnFeaturesPerLevel; using ORBextractor::umax; using ORBextractor::pattern;
    using ORBextractor::mvScaleFactor; using ORBextractor::mvInvScaleFactor; using ORBextractor::mvLevelSigma2; using ORBextractor::mvInvLevelSigma2;
};

// ===== matcher + frame grid ==========================================================================================
class Frame;
class MapPoint {                                        // the members SearchByProjection reads (include/MapPoint.h)
public:
    bool mbTrackInView = false, mbTrackInViewR = false, mbBad = false;
    float mTrackDepth = 0, mTrackViewCos = 0, mTrackViewCosR = 0, mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = -1, nObs = 0, predictedLevel = 0;
    float mfMinDistance = 0, mfMaxDistance = 0;
    Eigen::Vector3f mWorldPos;
    cv::Mat mDescriptor;
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return mDescriptor; }
    Eigen::Vector3f GetWorldPos() { return mWorldPos; }
    float GetMinDistanceInvariance() { return mfMinDistance; }
    float GetMaxDistanceInvariance() { return mfMaxDistance; }
    int PredictScale(const float&, Frame*) { return predictedLevel; }
};
struct GeometricCamera {
    Eigen::Matrix3f mF12;                                // K1^-T [t12]x R12 K2^-1, evaluated by the caller (Eigen) in the reference
    Eigen::Vector2f project(const Eigen::Vector3f& p) { return Eigen::Vector2f(p(0), p(1)); }
    // Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:107-129): the slice starts after the Eigen evaluation of F12
    bool epipolarConstrain(GeometricCamera* /*pCamera2*/, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f& /*R12*/,
                           const Eigen::Vector3f& /*t12*/, const float sigmaLevel, const float unc)
    {
        const Eigen::Matrix3f& F12 = mF12;
#include "pinhole_epipolar_tail.inc"
    }
};

class Frame {                                           // the members the slices read (include/Frame.h)
public:
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight, mvKeysUn;
    std::vector<float> mvuRight, mvDepth, mvScaleFactors, mvInvScaleFactors;
    cv::Mat mDescriptorsRight;
    ORBextractor *mpORBextractorLeft = nullptr, *mpORBextractorRight = nullptr;
    DBoW2::FeatureVector mFeatVec;
    GeometricCamera* mpCamera2 = nullptr;
    void ComputeStereoMatches();
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    cv::Mat mDescriptors;
    float mb = 0, mbf = 0;
    Sophus::SE3f mTcw, mTrl;
    GeometricCamera cam, *mpCamera = &cam;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    Sophus::SE3f GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() const { return mTrl; }
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    vector<size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1,
                                     const bool bRight = false) const;
};
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
inline int extractorParenthesis_unused();               // (the slice of AssignFeaturesToGrid ends before the next function)

class KeyFrame {
public:
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<cv::KeyPoint> mvKeysUn, mvKeys, mvKeysRight;
    DBoW2::FeatureVector mFeatVec;
    cv::Mat mDescriptors;
    int NLeft = -1, N = 0;
    GeometricCamera *mpCamera = nullptr, *mpCamera2 = nullptr;
    std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2;
    Sophus::SE3f mTcw, mTwc, mTrw, mTwr;
    Eigen::Vector3f mOw;
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTwc; }
    Sophus::SE3f GetRightPose() { return mTrw; }
    Sophus::SE3f GetRightPoseInverse() { return mTwr; }
    Eigen::Vector3f GetCameraCenter() { return mOw; }
    MapPoint* GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
};

class ORBmatcher {                                      // include/ORBmatcher.h:40-102
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b);
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches);
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12);
    int SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo,
                               const bool bCoarse = false);
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f);
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono);
    int SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th, const int ORBdist);
    static const int TH_LOW, TH_HIGH, HISTO_LENGTH;
    float RadiusByViewingCos(const float& viewCos);
    void ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3);
    float mfNNratio;
    bool mbCheckOrientation;
};
const int ORBmatcher::TH_HIGH = 100;                    // src/ORBmatcher1.cc:37-39
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;

#include "frame_assign_grid.inc"
#include "frame_area_posingrid.inc"
#include "matcher_maxima_distance.inc"
#include "matcher_projection_map.inc"
#include "matcher_projection_frames.inc"
#include "matcher_bow_kf_frame.inc"
#include "matcher_bow_kf_kf.inc"
#include "frame_stereo_matches.inc"
#include "matcher_triangulation.inc"

}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

namespace {
struct KP28 { float x, y, size, angle, response; int octave, class_id; };
cv::KeyPoint to_cv(const KP28& k) { return cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id); }
KP28 from_cv(const cv::KeyPoint& k) { return KP28{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id}; }

void fill_frame(Frame& F, const KP28* kp, const uint8_t* desc, const float* uright, int n, const float* bounds_grid, const float* scale, int nlev)
{
    F.N = n; F.Nleft = -1;
    F.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) F.mvKeysUn[i] = to_cv(kp[i]);
    F.mvKeys = F.mvKeysUn;
    F.mvuRight.assign(n, -1.0f);
    if (uright) for (int i = 0; i < n; ++i) F.mvuRight[i] = uright[i];
    if (scale) F.mvScaleFactors.assign(scale, scale + nlev);
    F.mvpMapPoints.assign(n, nullptr);
    F.mvbOutlier.assign(n, false);
    if (desc) F.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)desc);
    Frame::mnMinX = bounds_grid[0]; Frame::mnMinY = bounds_grid[1]; Frame::mnMaxX = bounds_grid[2]; Frame::mnMaxY = bounds_grid[3];
    Frame::mfGridElementWidthInv = bounds_grid[4]; Frame::mfGridElementHeightInv = bounds_grid[5];
    F.AssignFeaturesToGrid();
}
}  // namespace

extern "C" {

int refc_frame_grid_dims(int* cols, int* rows) { *cols = FRAME_GRID_COLS; *rows = FRAME_GRID_ROWS; return 0; }

// ---- extractor ---------------------------------------------------------------------------------------------------------
// ORBextractor::ORBextractor: the scale / sigma tables, features per level, umax and the 512-point pattern.
void refc_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv, float* sigma2, float* invsigma2, int* nfeat,
                 int* umax16, int* pattern1024)
{
    RefExtractor e(nfeatures, scaleFactor, nlevels, 20, 7);
    for (int i = 0; i < nlevels; ++i) {
        scale[i] = e.mvScaleFactor[i]; inv[i] = e.mvInvScaleFactor[i]; sigma2[i] = e.mvLevelSigma2[i]; invsigma2[i] = e.mvInvLevelSigma2[i];
        nfeat[i] = e.mnFeaturesPerLevel[i];
    }
    for (int i = 0; i < 16; ++i) umax16[i] = e.umax[i];
    for (int i = 0; i < 512; ++i) { pattern1024[2 * i] = e.pattern[i].x; pattern1024[2 * i + 1] = e.pattern[i].y; }
}

// ORBextractor::DistributeOctTree on candidates (x, y, response) given relative to (minX, minY) as the caller does.
int refc_distribute_octree(const int* xs, const int* ys, const int* scores, int n, int minX, int maxX, int minY, int maxY, int N,
                           int level, int* out_x, int* out_y, int* out_score, int cap)
{
    RefExtractor e(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> v(n);
    for (int i = 0; i < n; ++i) v[i] = cv::KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)scores[i]);
    const std::vector<cv::KeyPoint> r = e.DistributeOctTree(v, minX, maxX, minY, maxY, N, level);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) { out_x[i] = (int)r[i].pt.x; out_y[i] = (int)r[i].pt.y; out_score[i] = (int)r[i].response; }
    return (int)r.size();
}

// tileCalcKeypoints(img, level, border = EDGE_THRESHOLD - 3, nfeatures, iniThFAST, minThFAST) on a pyramid level (interior image).
int refc_tile_calc_keypoints(const uint8_t* img, int w, int h, size_t step, int nfeatures, int iniTh, int minTh, int* xs, int* ys,
                             int* scores, int cap)
{
    cv::Mat m(h, w, CV_8U, (void*)img, step);
    const std::vector<cv::KeyPoint> r = tileCalcKeypoints(m, 0, EDGE_THRESHOLD - 3, nfeatures, iniTh, minTh);
    for (int i = 0; i < (int)r.size() && i < cap; ++i) { xs[i] = (int)r[i].pt.x; ys[i] = (int)r[i].pt.y; scores[i] = (int)r[i].response; }
    return (int)r.size();
}

// computeOrbDescriptor(kpt, img, pattern, desc) with the reference's own bit_pattern_31_.
void refc_descriptor(const uint8_t* blurred, int w, int h, size_t step, float x, float y, float angle_deg, uint8_t* out32)
{
    cv::Mat m(h, w, CV_8U, (void*)blurred, step);
    cv::KeyPoint k(x, y, 31.f, angle_deg);
    computeOrbDescriptor(k, m, (const cv::Point*)bit_pattern_31_, out32);
}

// The placement loop of ORBextractor::operator() (scale by mvScaleFactor[level], lapping-area keypoints from the back).
int refc_pack_level(KP28* level_kps, const uint8_t* level_desc, int nlevel, int level, float scale_of_level, int lap0, int lap1,
                    KP28* out_kps, uint8_t* out_desc, int nkeypoints, int* monoIndex_io, int* stereoIndex_io)
{
    std::vector<cv::KeyPoint> keypoints(nlevel);
    for (int i = 0; i < nlevel; ++i) keypoints[i] = to_cv(level_kps[i]);
    std::vector<cv::KeyPoint> _keypoints(nkeypoints);
    for (int i = 0; i < nkeypoints; ++i) _keypoints[i] = to_cv(out_kps[i]);
    cv::Mat desc(nlevel, 32, CV_8U, (void*)level_desc), descriptors(nkeypoints, 32, CV_8U, (void*)out_desc);
    std::vector<float> mvScaleFactor(level + 1, 1.0f);
    mvScaleFactor[level] = scale_of_level;
    std::vector<int> vLappingArea = {lap0, lap1};
    int monoIndex = *monoIndex_io, stereoIndex = *stereoIndex_io;
    {
#include "orb_pack_loop.inc"
    }
    for (int i = 0; i < nkeypoints; ++i) out_kps[i] = from_cv(_keypoints[i]);
    *monoIndex_io = monoIndex; *stereoIndex_io = stereoIndex;
    return 0;
}

// ---- matcher -----------------------------------------------------------------------------------------------------------
int refc_descriptor_distance(const uint8_t* a, const uint8_t* b)
{
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void*)a), cv::Mat(1, 32, CV_8U, (void*)b));
}

void refc_three_maxima(const int* counts, int L, int* ind)
{
    std::vector<std::vector<int>> h(L);
    for (int i = 0; i < L; ++i) h[i].assign(counts[i], 0);
    ORBmatcher m;
    ind[0] = ind[1] = ind[2] = -1;
    m.ComputeThreeMaxima(h.data(), L, ind[0], ind[1], ind[2]);
}

void refc_assign_features_to_grid(const KP28* kp, int n, const float* bounds_grid, int32_t* cell_start, int32_t* items)
{
    Frame F;
    fill_frame(F, kp, nullptr, nullptr, n, bounds_grid, nullptr, 0);
    int t = 0;
    for (int ix = 0; ix < FRAME_GRID_COLS; ++ix)
        for (int iy = 0; iy < FRAME_GRID_ROWS; ++iy) {
            cell_start[ix * FRAME_GRID_ROWS + iy] = t;
            for (size_t v : F.mGrid[ix][iy]) items[t++] = (int32_t)v;
        }
    cell_start[FRAME_GRID_COLS * FRAME_GRID_ROWS] = t;
}

int refc_get_features_in_area(const KP28* kp, int n, const float* bounds_grid, float x, float y, float r, int minLevel, int maxLevel, int32_t* out)
{
    Frame F;
    fill_frame(F, kp, nullptr, nullptr, n, bounds_grid, nullptr, 0);
    const std::vector<size_t> v = F.GetFeaturesInArea(x, y, r, minLevel, maxLevel);
    for (size_t i = 0; i < v.size(); ++i) out[i] = (int32_t)v[i];
    return (int)v.size();
}

// Initial occupancy: `occupied[idx]` puts a map point with Observations() == occ_obs on the feature (1 = observed; for the
// key-frame variant any non-NULL pointer blocks).  Outputs: the index of the map point of THIS call a feature ends up with.
int refc_search_by_projection_map(const KP28* kp, const uint8_t* desc, const float* uright, const uint8_t* occupied, int n,
                                  const float* bounds_grid, const float* scale, int nlev, const uint8_t* in_view, const uint8_t* bad,
                                  const float* projx, const float* projy, const float* projxr, const float* viewcos, const float* depth,
                                  const int32_t* level, const int32_t* nobs, const uint8_t* mpdesc, int nmp, float th, int bFarPoints,
                                  float thFarPoints, float nnratio, int32_t* out)
{
    Frame F;
    fill_frame(F, kp, desc, uright, n, bounds_grid, scale, nlev);
    MapPoint own; own.nObs = 1;
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) F.mvpMapPoints[i] = &own;
    std::vector<MapPoint> mps(nmp);
    std::vector<MapPoint*> vp(nmp);
    for (int i = 0; i < nmp; ++i) {
        MapPoint& m = mps[i];
        m.mbTrackInView = in_view[i]; m.mbBad = bad[i]; m.mTrackProjX = projx[i]; m.mTrackProjY = projy[i]; m.mTrackProjXR = projxr[i];
        m.mTrackViewCos = viewcos[i]; m.mTrackDepth = depth[i]; m.mnTrackScaleLevel = level[i]; m.nObs = nobs[i];
        m.mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mpdesc + (size_t)i * 32));
        vp[i] = &m;
    }
    ORBmatcher matcher(nnratio, true);
    const int nm = matcher.SearchByProjection(F, vp, th, bFarPoints != 0, thFarPoints);
    for (int i = 0; i < n; ++i) out[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &own) ? (int32_t)(F.mvpMapPoints[i] - mps.data()) : -1;
    return nm;
}

// SearchByProjection(CurrentFrame, LastFrame, th, bMono): world point i = (u, v, z) under identity poses, so the slice computes
// uv = (u, v) and invzc = 1.0 / z itself; last_tz moves the last frame along z to switch bForward / bBackward on (mb = 1).
int refc_search_by_projection_last(const KP28* kp, const uint8_t* desc, const float* uright, const uint8_t* occupied, int n,
                                   const float* bounds_grid, const float* scale, int nlev, float mbf, const uint8_t* valid,
                                   const float* u, const float* v, const float* z, const int32_t* octave, const float* angle,
                                   const int32_t* nobs, const uint8_t* mpdesc, int nlast, float th, float last_tz, int bMono,
                                   int checkOri, float* invz_out, int32_t* out)
{
    Frame Cur, Last;
    fill_frame(Cur, kp, desc, uright, n, bounds_grid, scale, nlev);
    Cur.mb = 1.0f; Cur.mbf = mbf;
    MapPoint own; own.nObs = 1;
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) Cur.mvpMapPoints[i] = &own;
    Last.N = nlast; Last.Nleft = -1;
    Last.mvKeys.resize(nlast); Last.mvKeysUn.resize(nlast);
    Last.mvpMapPoints.assign(nlast, nullptr); Last.mvbOutlier.assign(nlast, false);
    Last.mTcw.t = Eigen::Vector3f(0, 0, last_tz);
    std::vector<MapPoint> mps(nlast);
    for (int i = 0; i < nlast; ++i) {
        MapPoint& m = mps[i];
        m.mWorldPos = Eigen::Vector3f(u[i], v[i], z[i]); m.nObs = nobs[i];
        m.mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mpdesc + (size_t)i * 32));
        Last.mvKeys[i].octave = octave[i]; Last.mvKeys[i].angle = angle[i];
        Last.mvKeysUn[i] = Last.mvKeys[i];
        if (valid[i]) Last.mvpMapPoints[i] = &m;
        if (invz_out) invz_out[i] = 1.0 / z[i];          // the expression of src/ORBmatcher3.cc:291, for the oracle's input
    }
    ORBmatcher matcher(0.9f, checkOri != 0);
    const int nm = matcher.SearchByProjection(Cur, Last, th, bMono != 0);
    for (int i = 0; i < n; ++i) out[i] = (Cur.mvpMapPoints[i] && Cur.mvpMapPoints[i] != &own) ? (int32_t)(Cur.mvpMapPoints[i] - mps.data()) : -1;
    return nm;
}

// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): world point i = (u, v, z), identity pose (Ow = 0).
int refc_search_by_projection_kf(const KP28* kp, const uint8_t* desc, const uint8_t* occupied, int n, const float* bounds_grid,
                                 const float* scale, int nlev, const uint8_t* valid, const uint8_t* already_found, const float* u,
                                 const float* v, const float* z, const float* mind, const float* maxd, const int32_t* level,
                                 const float* angle, const uint8_t* mpdesc, int nkf, float th, int ORBdist, int checkOri,
                                 float* dist3d_out, int32_t* out)
{
    Frame Cur;
    fill_frame(Cur, kp, desc, nullptr, n, bounds_grid, scale, nlev);
    MapPoint own; own.nObs = 0;                          // any non-NULL pointer blocks in this variant
    for (int i = 0; i < n; ++i) if (occupied && occupied[i]) Cur.mvpMapPoints[i] = &own;
    KeyFrame KF;
    KF.mvpMapPoints.assign(nkf, nullptr); KF.mvKeysUn.resize(nkf);
    std::vector<MapPoint> mps(nkf);
    std::set<MapPoint*> found;
    for (int i = 0; i < nkf; ++i) {
        MapPoint& m = mps[i];
        m.mWorldPos = Eigen::Vector3f(u[i], v[i], z[i]); m.mfMinDistance = mind[i]; m.mfMaxDistance = maxd[i]; m.predictedLevel = level[i];
        m.mDescriptor = cv::Mat(1, 32, CV_8U, (void*)(mpdesc + (size_t)i * 32));
        KF.mvKeysUn[i].angle = angle[i];
        if (valid[i]) KF.mvpMapPoints[i] = &m;
        if (already_found && already_found[i]) found.insert(&m);
        if (dist3d_out) dist3d_out[i] = m.mWorldPos.norm();
    }
    ORBmatcher matcher(0.9f, checkOri != 0);
    const int nm = matcher.SearchByProjection(Cur, &KF, found, th, ORBdist);
    for (int i = 0; i < n; ++i) out[i] = (Cur.mvpMapPoints[i] && Cur.mvpMapPoints[i] != &own) ? (int32_t)(Cur.mvpMapPoints[i] - mps.data()) : -1;
    return nm;
}

// Frame::ComputeStereoMatches on two extracted images: pyramids as bordered buffers (interior at +border), mb / mbf as the
// reference reads them.  Outputs mvuRight / mvDepth.
void refc_compute_stereo_matches(const KP28* kpL, const uint8_t* descL, int nL, const KP28* kpR, const uint8_t* descR, int nR,
                                 const uint8_t* const* pyrL, const uint8_t* const* pyrR, const size_t* steps, const int* widths,
                                 const int* heights, int nlevels, int border, const float* scale, const float* inv, float mb, float mbf,
                                 float* uRight, float* depth)
{
    RefExtractor exL(1000, 1.2f, nlevels, 20, 7), exR(1000, 1.2f, nlevels, 20, 7);
    for (int l = 0; l < nlevels; ++l) {
        exL.mvImagePyramid[l] = cv::Mat(heights[l], widths[l], CV_8U, (void*)(pyrL[l] + (size_t)border * steps[l] + border), steps[l]);
        exR.mvImagePyramid[l] = cv::Mat(heights[l], widths[l], CV_8U, (void*)(pyrR[l] + (size_t)border * steps[l] + border), steps[l]);
    }
    Frame F;
    F.N = nL;
    F.mvKeys.resize(nL); F.mvKeysRight.resize(nR);
    for (int i = 0; i < nL; ++i) F.mvKeys[i] = to_cv(kpL[i]);
    for (int i = 0; i < nR; ++i) F.mvKeysRight[i] = to_cv(kpR[i]);
    F.mDescriptors = cv::Mat(nL, 32, CV_8U, (void*)descL);
    F.mDescriptorsRight = cv::Mat(nR, 32, CV_8U, (void*)descR);
    F.mvScaleFactors.assign(scale, scale + nlevels); F.mvInvScaleFactors.assign(inv, inv + nlevels);
    F.mb = mb; F.mbf = mbf;
    F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
    F.ComputeStereoMatches();
    for (int i = 0; i < nL; ++i) { uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i]; }
}

namespace {
void fill_featvec(DBoW2::FeatureVector& fv, int n_nodes, const uint32_t* nodes, const int32_t* offs, const uint32_t* idx)
{
    for (int i = 0; i < n_nodes; ++i) fv[nodes[i]] = std::vector<unsigned int>(idx + offs[i], idx + offs[i + 1]);
}
}  // namespace

// SearchByBoW(KeyFrame* pKF, Frame& F, vpMapPointMatches): out[j] = index of the key-frame feature whose map point F's feature j got.
int refc_search_by_bow_kf_frame(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* valid_kf, int n_kf, int fvk_n,
                                const uint32_t* fvk_nodes, const int32_t* fvk_off, const uint32_t* fvk_idx, const uint8_t* desc_f,
                                const float* angle_f, int n_f, int fvf_n, const uint32_t* fvf_nodes, const int32_t* fvf_off,
                                const uint32_t* fvf_idx, int nleft, float nnratio, int checkOri, int32_t* out)
{
    KeyFrame KF; Frame F;
    GeometricCamera cam2;
    std::vector<MapPoint> mps(n_kf);
    KF.mvpMapPoints.assign(n_kf, nullptr);
    for (int i = 0; i < n_kf; ++i) if (valid_kf[i]) KF.mvpMapPoints[i] = &mps[i];
    KF.mvKeysUn.resize(n_kf);
    for (int i = 0; i < n_kf; ++i) KF.mvKeysUn[i].angle = angle_kf[i];
    KF.mvKeys = KF.mvKeysUn;
    KF.mDescriptors = cv::Mat(n_kf, 32, CV_8U, (void*)desc_kf);
    fill_featvec(KF.mFeatVec, fvk_n, fvk_nodes, fvk_off, fvk_idx);
    F.N = n_f; F.Nleft = nleft;
    F.mvKeys.resize(nleft >= 0 ? nleft : n_f);
    for (size_t j = 0; j < F.mvKeys.size(); ++j) F.mvKeys[j].angle = angle_f[j];
    if (nleft >= 0) {                                    // two-camera frame: features >= Nleft live in mvKeysRight
        F.mvKeysRight.resize(n_f - nleft);
        for (int j = nleft; j < n_f; ++j) F.mvKeysRight[j - nleft].angle = angle_f[j];
        F.mpCamera2 = &cam2;
    }
    F.mDescriptors = cv::Mat(n_f, 32, CV_8U, (void*)desc_f);
    fill_featvec(F.mFeatVec, fvf_n, fvf_nodes, fvf_off, fvf_idx);
    ORBmatcher matcher(nnratio, checkOri != 0);
    std::vector<MapPoint*> matches;
    const int nm = matcher.SearchByBoW(&KF, F, matches);
    for (int j = 0; j < n_f; ++j) out[j] = matches[j] ? (int32_t)(matches[j] - mps.data()) : -1;
    return nm;
}

// SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vpMatches12): out[i] = index of KF2's feature whose map point KF1's feature i got.
int refc_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, int fv1_n, const uint32_t* fv1_nodes,
                             const int32_t* fv1_off, const uint32_t* fv1_idx, const uint8_t* desc2, const float* angle2,
                             const uint8_t* valid2, int n2, int fv2_n, const uint32_t* fv2_nodes, const int32_t* fv2_off,
                             const uint32_t* fv2_idx, float nnratio, int checkOri, int32_t* out)
{
    KeyFrame K1, K2;
    std::vector<MapPoint> m1(n1), m2(n2);
    K1.mvpMapPoints.assign(n1, nullptr); K2.mvpMapPoints.assign(n2, nullptr);
    for (int i = 0; i < n1; ++i) if (valid1[i]) K1.mvpMapPoints[i] = &m1[i];
    for (int i = 0; i < n2; ++i) if (valid2[i]) K2.mvpMapPoints[i] = &m2[i];
    K1.mvKeysUn.resize(n1); K2.mvKeysUn.resize(n2);
    for (int i = 0; i < n1; ++i) K1.mvKeysUn[i].angle = angle1[i];
    for (int i = 0; i < n2; ++i) K2.mvKeysUn[i].angle = angle2[i];
    K1.mDescriptors = cv::Mat(n1, 32, CV_8U, (void*)desc1); K2.mDescriptors = cv::Mat(n2, 32, CV_8U, (void*)desc2);
    fill_featvec(K1.mFeatVec, fv1_n, fv1_nodes, fv1_off, fv1_idx);
    fill_featvec(K2.mFeatVec, fv2_n, fv2_nodes, fv2_off, fv2_idx);
    ORBmatcher matcher(nnratio, checkOri != 0);
    std::vector<MapPoint*> matches;
    const int nm = matcher.SearchByBoW(&K1, &K2, matches);
    for (int i = 0; i < n1; ++i) out[i] = matches[i] ? (int32_t)(matches[i] - m2.data()) : -1;
    return nm;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:329-401) from the point where the observed descriptors are gathered.
int refc_distinctive_descriptor(const uint8_t* desc, int n)
{
    std::vector<cv::Mat> vDescriptors;
    for (int i = 0; i < n; ++i) vDescriptors.push_back(cv::Mat(1, 32, CV_8U, (void*)(desc + (size_t)i * 32)));
#include "mappoint_distinctive_core.inc"
    return BestIdx;
}

// SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse), single pinhole camera on both key frames.  F12 and the
// epipole are supplied (the reference evaluates them with Eigen/Sophus); out[i] = matched feature of KF2 or -1.
int refc_search_for_triangulation(const KP28* kp1, const uint8_t* desc1, const uint8_t* free1, const uint8_t* stereo1, int n1, int fv1_n,
                                  const uint32_t* fv1_nodes, const int32_t* fv1_off, const uint32_t* fv1_idx, const KP28* kp2,
                                  const uint8_t* desc2, const uint8_t* free2, const uint8_t* stereo2, int n2, int fv2_n,
                                  const uint32_t* fv2_nodes, const int32_t* fv2_off, const uint32_t* fv2_idx, const float* F12,
                                  const float* ep, const float* scale2, const float* sigma2_2, int nlev, int bOnlyStereo, int bCoarse,
                                  int checkOri, int32_t* out)
{
    KeyFrame K1, K2;
    GeometricCamera cam1, cam2;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cam1.mF12.m[r][c] = F12[r * 3 + c];
    K1.mpCamera = &cam1; K2.mpCamera = &cam2;
    MapPoint some;
    auto fill = [&](KeyFrame& K, const KP28* kp, const uint8_t* desc, const uint8_t* fr, const uint8_t* st, int n) {
        K.N = n;
        K.mvKeysUn.resize(n);
        for (int i = 0; i < n; ++i) K.mvKeysUn[i] = to_cv(kp[i]);
        K.mvKeys = K.mvKeysUn;
        K.mvpMapPoints.assign(n, nullptr);
        K.mvuRight.assign(n, -1.0f);
        for (int i = 0; i < n; ++i) { if (!fr[i]) K.mvpMapPoints[i] = &some; if (st[i]) K.mvuRight[i] = 1.0f; }
        K.mDescriptors = cv::Mat(n, 32, CV_8U, (void*)desc);
    };
    fill(K1, kp1, desc1, free1, stereo1, n1);
    fill(K2, kp2, desc2, free2, stereo2, n2);
    K2.mvScaleFactors.assign(scale2, scale2 + nlev); K2.mvLevelSigma2.assign(sigma2_2, sigma2_2 + nlev);
    K1.mvScaleFactors = K2.mvScaleFactors; K1.mvLevelSigma2 = K2.mvLevelSigma2;
    K1.mOw = Eigen::Vector3f(ep[0], ep[1], 1.0f);          // identity poses: C2 = Cw, project() keeps (x, y) => ep as given
    fill_featvec(K1.mFeatVec, fv1_n, fv1_nodes, fv1_off, fv1_idx);
    fill_featvec(K2.mFeatVec, fv2_n, fv2_nodes, fv2_off, fv2_idx);
    ORBmatcher matcher(0.6f, checkOri != 0);
    std::vector<std::pair<size_t, size_t>> pairs;
    const int nm = matcher.SearchForTriangulation(&K1, &K2, pairs, bOnlyStereo != 0, bCoarse != 0);
    for (int i = 0; i < n1; ++i) out[i] = -1;
    for (auto& pr : pairs) out[pr.first] = (int32_t)pr.second;
    return nm;
}

}  // extern "C"
