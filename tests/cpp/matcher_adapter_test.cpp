// Compiles csrc/adapter/ORBmatcherProjection.h against stand-ins of ORB_SLAM3::Frame / MapPoint (only the members the reference
// functions read; names as in include/Frame.h, include/MapPoint.h) and runs SearchByProjection(Frame, MapPoints) and the
// last-frame overload the way Tracking does.  The scene comes from a file written by tests/test_adapter_cpp.py, the resulting
// Frame::mvpMapPoints (as map-point indices) go back through another file and are compared with the ctypes / oracle results.
#include <cstdint>
#include <cstdio>
#include <vector>

#include <opencv2/core.hpp>

#include "ORBmatcherProjection.h"

namespace ORB_SLAM3 {
class MapPoint {
public:
    bool mbTrackInView = false, mbBad = false;
    float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackViewCos = 0, mTrackDepth = 0;
    int mnTrackScaleLevel = 0, nObs = 0;
    unsigned char descriptor[32];
    bool isBad() { return mbBad; }
    int Observations() { return nObs; }
    cv::Mat GetDescriptor() { return cv::Mat(1, 32, CV_8U, descriptor); }
};
class Frame {
public:
    int N = 0;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn;
    std::vector<float> mvuRight, mvScaleFactors;
    std::vector<MapPoint*> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    cv::Mat mDescriptors;
    float mbf = 0;
    static float mnMinX, mnMaxX, mnMinY, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
};
float Frame::mnMinX, Frame::mnMaxX, Frame::mnMinY, Frame::mnMaxY, Frame::mfGridElementWidthInv, Frame::mfGridElementHeightInv;
}  // namespace ORB_SLAM3

using namespace ORB_SLAM3;

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { std::perror("read"); exit(2); } return v; }

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s scene.bin result.bin\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { std::perror(argv[1]); return 2; }
    const std::vector<int32_t> hdr = rd<int32_t>(f, 4);                    // n features, n map points, n levels, mode (0 map, 1 last)
    const int n = hdr[0], m = hdr[1], nlev = hdr[2], mode = hdr[3];
    const std::vector<float> fl = rd<float>(f, 10);                        // bounds(4), grid inv(2), th, ratio / mbf, far, th_far
    Frame F;
    F.N = n;
    F.mvKeysUn = rd<cv::KeyPoint>(f, n); F.mvKeys = F.mvKeysUn;
    std::vector<unsigned char> desc = rd<unsigned char>(f, (size_t)n * 32);
    F.mDescriptors = cv::Mat(n, 32, CV_8U, desc.data());
    F.mvuRight = rd<float>(f, n);
    const std::vector<unsigned char> occ = rd<unsigned char>(f, n);
    F.mvScaleFactors = rd<float>(f, nlev);
    Frame::mnMinX = fl[0]; Frame::mnMinY = fl[1]; Frame::mnMaxX = fl[2]; Frame::mnMaxY = fl[3];
    Frame::mfGridElementWidthInv = fl[4]; Frame::mfGridElementHeightInv = fl[5];
    MapPoint own; own.nObs = 1;
    F.mvpMapPoints.assign(n, nullptr);
    for (int i = 0; i < n; ++i) if (occ[i]) F.mvpMapPoints[i] = &own;
    // map points / last-frame features: flags(2 bytes each), 6 floats each, 2 ints each, descriptor
    const std::vector<unsigned char> fA = rd<unsigned char>(f, m), fB = rd<unsigned char>(f, m);
    const std::vector<float> x = rd<float>(f, m), y = rd<float>(f, m), xr = rd<float>(f, m), vc = rd<float>(f, m), dep = rd<float>(f, m), ang = rd<float>(f, m);
    const std::vector<int32_t> lvl = rd<int32_t>(f, m), nobs = rd<int32_t>(f, m);
    const std::vector<unsigned char> md = rd<unsigned char>(f, (size_t)m * 32);
    std::fclose(f);
    std::vector<MapPoint> mps(m);
    for (int i = 0; i < m; ++i) {
        MapPoint& p = mps[i];
        p.mbTrackInView = fA[i]; p.mbBad = fB[i]; p.mTrackProjX = x[i]; p.mTrackProjY = y[i]; p.mTrackProjXR = xr[i]; p.mTrackViewCos = vc[i];
        p.mTrackDepth = dep[i]; p.mnTrackScaleLevel = lvl[i]; p.nObs = nobs[i];
        for (int b = 0; b < 32; ++b) p.descriptor[b] = md[(size_t)i * 32 + b];
    }
    int nm = 0;
    if (mode == 0) {
        std::vector<MapPoint*> vp(m);
        for (int i = 0; i < m; ++i) vp[i] = &mps[i];
        nm = orbx_adapter::SearchByProjection(F, vp, fl[6], fl[8] != 0, fl[9], fl[7]);
    } else {
        Frame Last;
        Last.N = m;
        Last.mvKeys.resize(m); Last.mvKeysUn.resize(m);
        Last.mvpMapPoints.assign(m, nullptr); Last.mvbOutlier.assign(m, false);
        for (int i = 0; i < m; ++i) {
            Last.mvKeys[i].octave = lvl[i]; Last.mvKeysUn[i].angle = ang[i];
            if (fA[i]) Last.mvpMapPoints[i] = &mps[i];       // fA = valid, x / y = uv, dep = invz in this mode
        }
        F.mbf = fl[7];
        nm = orbx_adapter::SearchByProjectionLast(F, Last, x, y, dep, fl[6], false, false, true);
    }
    std::vector<int32_t> out(n);
    for (int i = 0; i < n; ++i) out[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i] != &own) ? (int32_t)(F.mvpMapPoints[i] - mps.data()) : -1;
    FILE* g = std::fopen(argv[2], "wb");
    std::fwrite(&nm, 4, 1, g);
    std::fwrite(out.data(), 4, n, g);
    std::fclose(g);
    std::printf("nmatches=%d\n", nm);
    return 0;
}
