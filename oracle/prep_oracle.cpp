// prep_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
//
// Image / key-point preparation either side of the extractor:
//   Frame::UndistortKeyPoints           src/Frame.cc:777-810  -> cv::undistortPoints(mat, mat, K, mDistCoef, cv::Mat(), mK)
//   System::TrackStereo rectification   src/System.cc:253-260 -> cv::remap(im, out, M1, M2, cv::INTER_LINEAR) with the CV_32FC1
//                                       maps of cv::initUndistortRectifyMap (src/Settings.cc:488-491)
// Both are OpenCV-owned arithmetic (system OpenCV >= 4.4, un-vendored): restated here from the published algorithm and PINNED
// bit-exact against cv2 4.13 in tests/test_oracle_cv2.py (undistortPoints on random points / distortion sets, remap on random
// maps including out-of-image samples).
#include <cmath>
#include <cstdint>
#include <cstring>

namespace {
struct KeyPoint { float x, y, size, angle, response; int octave, class_id; };
inline int cvRoundF(float v) { return (int)lrintf(v); }
}  // namespace

extern "C" {

// cv::undistortPoints for the pinhole + radial-tangential model, R = identity, P = new camera matrix (fx, fy, cx, cy):
// normalise, 5 fixed-point iterations of the inverse distortion (TermCriteria(MAX_ITER, 5, 0.01)), re-project.  All in double,
// results stored as float.  k = (k1, k2, p1, p2[, k3[, k4, k5, k6]]), missing coefficients are 0.
void orbo_undistort_points(const float* src_xy, int n, const double* K /*fx fy cx cy*/, const double* dist, int n_dist,
                           const double* P /*fx fy cx cy*/, float* dst_xy)
{
    double k[14] = {0};
    for (int i = 0; i < n_dist && i < 14; ++i) k[i] = dist[i];
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        double x = src_xy[2 * i], y = src_xy[2 * i + 1];
        const double u = x, v = y;
        x = (x - cx) * ifx;
        y = (y - cy) * ify;
        if (n_dist > 0) {
            const double x0 = x, y0 = y;
            for (int j = 0; j < 5; ++j) {
                const double r2 = x * x + y * y;
                double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
                const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
                const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
                x = (x0 - deltaX) * icdist;
                y = (y0 - deltaY) * icdist;
            }
        }
        // RR = P * I: xx = P00 x + P01 y + P02, ww = 1 / (P20 x + P21 y + P22) = 1
        const double xx = P[0] * x + 0.0 * y + P[2];
        const double yy = 0.0 * x + P[1] * y + P[3];
        const double ww = 1. / (0.0 * x + 0.0 * y + 1.0);
        dst_xy[2 * i] = (float)(xx * ww);
        dst_xy[2 * i + 1] = (float)(yy * ww);
    }
}

// Frame::UndistortKeyPoints — src/Frame.cc:777-810: copies the key points when mDistCoef[0] == 0, else undistorts pt.
void orbo_undistort_keypoints(const KeyPoint* kps, int n, const double* K, const float* dist, int n_dist, const double* newK, KeyPoint* out)
{
    if (n_dist == 0 || dist[0] == 0.0) { if (n) memcpy(out, kps, (size_t)n * sizeof(KeyPoint)); return; }
    double d[14] = {0};
    for (int i = 0; i < n_dist && i < 14; ++i) d[i] = dist[i];          // mDistCoef is CV_32F; OpenCV converts to double
    for (int i = 0; i < n; ++i) {
        float s[2] = {kps[i].x, kps[i].y}, t[2];
        orbo_undistort_points(s, 1, K, d, n_dist, newK, t);
        out[i] = kps[i];
        out[i].x = t[0]; out[i].y = t[1];
    }
}

// cv::remap(src, dst, mapx, mapy, INTER_LINEAR, BORDER_CONSTANT, 0) for CV_8UC1 images and CV_32FC1 maps.
// Map coordinates are quantised to 1/32 pixel (cvRound(m * 32)); the four weights are the exact products
// (32 - fy)(32 - fx) * 32 ... in 15-bit fixed point (the (0,0) entry saturates to 32767 and the missing 1 lands on the
// opposite corner, which cannot change an 8-bit result); out-of-image taps read the border value 0.
void orbo_remap_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, const float* mapx, const float* mapy, size_t map_step_floats,
                          uint8_t* dst, int dw, int dh, size_t dstep)
{
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            const int sxq = cvRoundF(mapx[(size_t)y * map_step_floats + x] * 32.0f);
            const int syq = cvRoundF(mapy[(size_t)y * map_step_floats + x] * 32.0f);
            const int fx = sxq & 31, fy = syq & 31;
            int sx = sxq >> 5, sy = syq >> 5;
            sx = sx > 32767 ? 32767 : sx < -32768 ? -32768 : sx;             // saturate_cast<short>
            sy = sy > 32767 ? 32767 : sy < -32768 ? -32768 : sy;
            int w[4] = {(32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32};
            if (w[0] == 32768) { w[0] = 32767; w[3] = 1; }
            auto px = [&](int xx, int yy) -> int {
                return (xx >= 0 && xx < sw && yy >= 0 && yy < sh) ? src[(size_t)yy * sstep + xx] : 0;
            };
            const int v = px(sx, sy) * w[0] + px(sx + 1, sy) * w[1] + px(sx, sy + 1) * w[2] + px(sx + 1, sy + 1) * w[3];
            dst[(size_t)y * dstep + x] = (uint8_t)((v + (1 << 14)) >> 15);
        }
}

}  // extern "C"
