// Compiles csrc/adapter/ORBmatcher.h — `class ORB_SLAM3::ORBmatcher` with the reference's declarations — against stand-ins of
// ORB_SLAM3::Frame / KeyFrame / MapPoint (only the members the reference functions read; names as in include/Frame.h,
// KeyFrame.h, MapPoint.h; a three-line SE3 / pinhole stand-in for Sophus and GeometricCamera) and drives it the way
// Tracking::Relocalization / LoopClosing (SearchByBoW) and Tracking::TrackWithMotionModel (SearchByProjection) do.
// Scene in, matches out through files written / read by tests/test_adapter_cpp.py.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

#include <opencv2/core.hpp>

namespace ORB_SLAM3 {
struct Vec2 { float d[2]; float operator()(int i) const { return d[i]; } };
struct Vec3 { float d[3]; float operator()(int i) const { return d[i]; } };
struct SE3f {                                     // rotation (row-major) + translation; enough of Sophus::SE3f for the call sites
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t[3] = {0, 0, 0};
    Vec3 operator*(const Vec3& p) const
    {
        Vec3 o;
        for (int r = 0; r < 3; ++r) o.d[r] = R[3 * r] * p.d[0] + R[3 * r + 1] * p.d[1] + R[3 * r + 2] * p.d[2] + t[r];
        return o;
    }
    SE3f inverse() const
    {
        SE3f o;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) o.R[3 * r + c] = R[3 * c + r];
        for (int r = 0; r < 3; ++r) o.t[r] = -(o.R[3 * r] * t[0] + o.R[3 * r + 1] * t[1] + o.R[3 * r + 2] * t[2]);
        return o;
    }
    Vec3 translation() const { return Vec3{{t[0], t[1], t[2]}}; }
};
struct Pinhole {
    float fx, fy, cx, cy;
    Vec2 project(const Vec3& p) const { return Vec2{{fx * p(0) / p(2) + cx, fy * p(1) / p(2) + cy}}; }
};
class MapPoint {
public:
    bool mbTrackInView = false, mbBad = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0, mTrackDepth = 0;
    int mnTrackScaleLevel = 0, nObs = 1;
    unsigned char descriptor[32];
    Vec3 pos{{0, 0, 1}};
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return cv::Mat(1, 32, CV_8U, descriptor); }
    Vec3 GetWorldPos() { return pos; }
};
typedef std::map<unsigned int, std::vector<unsigned int>> FeatureVector;      // DBoW2::FeatureVector
class Frame {
public:
    int N = 0, Nleft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mDescriptors;
    FeatureVector mFeatVec;
    float mbf = 0, mb = 0;
    SE3f pose;
    Pinhole* mpCamera = nullptr;
    SE3f GetPose() const { return pose; }
    static float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
};
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
class KeyFrame {
public:
    std::vector<cv::KeyPoint> mvKeysUn;
    cv::Mat mDescriptors;
    FeatureVector mFeatVec;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<MapPoint*> GetMapPointMatches() { return mvpMapPoints; }
};
}  // namespace ORB_SLAM3

#include "ORBmatcher.h"

using namespace ORB_SLAM3;

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { std::perror("read"); exit(2); } return v; }

static FeatureVector read_fv(FILE* f)
{
    const int nn = rd<int32_t>(f, 1)[0];
    const std::vector<uint32_t> nodes = rd<uint32_t>(f, nn);
    const std::vector<int32_t> off = rd<int32_t>(f, nn + 1);
    const std::vector<uint32_t> idx = rd<uint32_t>(f, off[nn]);
    FeatureVector fv;
    for (int k = 0; k < nn; ++k) fv[nodes[k]] = std::vector<unsigned int>(idx.begin() + off[k], idx.begin() + off[k + 1]);
    return fv;
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s scene.bin result.bin\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    const std::vector<int32_t> hdr = rd<int32_t>(f, 4);                    // mode (0 KF-Frame, 1 KF-KF), nA, nB, check orientation
    const int mode = hdr[0], nA = hdr[1], nB = hdr[2];
    const float ratio = rd<float>(f, 1)[0];
    std::vector<unsigned char> dA = rd<unsigned char>(f, (size_t)nA * 32), dB = rd<unsigned char>(f, (size_t)nB * 32);
    const std::vector<float> aA = rd<float>(f, nA), aB = rd<float>(f, nB);
    const std::vector<unsigned char> vA = rd<unsigned char>(f, nA), vB = rd<unsigned char>(f, nB);
    const FeatureVector fvA = read_fv(f), fvB = read_fv(f);
    std::fclose(f);

    std::vector<MapPoint> mpA(nA), mpB(nB);
    KeyFrame kf1;
    kf1.mvKeysUn.resize(nA); kf1.mDescriptors = cv::Mat(nA, 32, CV_8U, dA.data()); kf1.mFeatVec = fvA; kf1.mvpMapPoints.assign(nA, nullptr);
    for (int i = 0; i < nA; ++i) { kf1.mvKeysUn[i].angle = aA[i]; if (vA[i] & 1) kf1.mvpMapPoints[i] = &mpA[i]; mpA[i].mbBad = (vA[i] & 2) != 0; }
    ORBmatcher matcher(ratio, hdr[3] != 0);
    std::vector<MapPoint*> out;
    int nm = 0;
    std::vector<int32_t> res;
    if (mode == 0) {
        Frame F;
        F.N = nB; F.mvKeys.resize(nB); F.mvKeysUn.resize(nB); F.mDescriptors = cv::Mat(nB, 32, CV_8U, dB.data()); F.mFeatVec = fvB;
        for (int j = 0; j < nB; ++j) { F.mvKeys[j].angle = aB[j]; F.mvKeysUn[j].angle = aB[j]; }
        nm = matcher.SearchByBoW(&kf1, F, out);                           // vpMapPointMatches[j] = KF map point
        res.assign(nB, -1);
        for (int j = 0; j < nB; ++j) if (out[j]) res[j] = (int32_t)(out[j] - mpA.data());
    } else {
        KeyFrame kf2;
        kf2.mvKeysUn.resize(nB); kf2.mDescriptors = cv::Mat(nB, 32, CV_8U, dB.data()); kf2.mFeatVec = fvB; kf2.mvpMapPoints.assign(nB, nullptr);
        for (int j = 0; j < nB; ++j) { kf2.mvKeysUn[j].angle = aB[j]; if (vB[j] & 1) kf2.mvpMapPoints[j] = &mpB[j]; mpB[j].mbBad = (vB[j] & 2) != 0; }
        nm = matcher.SearchByBoW(&kf1, &kf2, out);                        // vpMatches12[i] = KF2 map point
        res.assign(nA, -1);
        for (int i = 0; i < nA; ++i) if (out[i]) res[i] = (int32_t)(out[i] - mpB.data());
    }
    const int d01 = nA >= 2 ? ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, dA.data()), cv::Mat(1, 32, CV_8U, dA.data() + 32)) : -1;
    FILE* g = std::fopen(argv[2], "wb");
    std::fwrite(&nm, 4, 1, g);
    std::fwrite(&d01, 4, 1, g);
    std::fwrite(res.data(), 4, res.size(), g);
    std::fclose(g);
    std::printf("nmatches=%d TH_LOW=%d TH_HIGH=%d HISTO_LENGTH=%d\n", nm, ORBmatcher::TH_LOW, ORBmatcher::TH_HIGH, ORBmatcher::HISTO_LENGTH);
    // the projection overloads are instantiated too (compile check of the Sophus-side lines against the stand-ins)
    if (argc > 3) {
        Frame::mnMinX = 0; Frame::mnMinY = 0; Frame::mnMaxX = 752; Frame::mnMaxY = 480;
        Frame::mfGridElementWidthInv = 64.f / 752.f; Frame::mfGridElementHeightInv = 48.f / 480.f;
        Frame C, L; Pinhole cam{435.f, 435.f, 367.f, 252.f}; C.mpCamera = &cam;
        C.mvScaleFactors.assign(8, 1.f); L.mvScaleFactors.assign(8, 1.f);
        std::vector<MapPoint*> none;
        try {
            const int a = matcher.SearchByProjection(C, none, 3.f, false, 50.f);
            const int b = matcher.SearchByProjection(C, L, 15.f, false);
            std::printf("empty projection searches: %d %d\n", a, b);
        } catch (const std::exception& e) { std::printf("empty projection searches rejected: %s\n", e.what()); }
    }
    return 0;
}
