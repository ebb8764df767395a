// orb_oracle.cpp — CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
//
// A self-contained C++17 restatement of the *CPU semantics* of the reference ORB front end
// (kpmrozowski/wut-cuda-orb-slam3).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this library, and only as the checker / baseline.
//
// Parity pinning status: PINNED on both sides.
//  * OpenCV-owned arithmetic (resize, copyMakeBorder, FAST, GaussianBlur, fastAtan2, BFMatcher.knnMatch, cvtColor) is checked
//    bit-exact against cv2 4.13 — the same third-party code the reference calls — in tests/test_oracle_cv2.py.
//  * Reference-owned logic (constructor tables, DistributeOctTree incl. its std::sort tie behaviour, the cell loop
//    tileCalcKeypoints, computeOrbDescriptor + bit_pattern_31_, operator()'s placement loop, DescriptorDistance,
//    ComputeThreeMaxima, ComputeStereoMatches, both SearchByBoW overloads, SearchForTriangulation, ComputeDistinctiveDescriptors,
//    the DBoW2 transform / L1 score / text loader) is checked against the REFERENCE'S OWN FUNCTIONS, compiled from the sources
//    where they lie by oracle/build_ref.sh into oracle/_ref/libref.so (tests/test_ref_pin.py), and against
//    tests/golden/ref_golden.json, frozen from that library (tests/golden/make_ref_golden.py).
//  * Not pinned against compiled reference code (restatement only): IC_Angle (the fork deleted the CPU function; the oracle
//    follows upstream's summation pattern and OpenCV's fastAtan2, which is cv2-pinned) and ComputePyramid's glue.
//
// Every function cites the reference file:line (paths relative to /root/reference) it follows.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <chrono>
#include <vector>

namespace {

constexpr int PATCH_SIZE = 31;       // src/ORBextractor.cc:99
constexpr int HALF_PATCH_SIZE = 15;  // src/ORBextractor.cc:100
constexpr int EDGE_THRESHOLD = 19;   // src/ORBextractor.cc:101

static const int8_t kPattern[1024] = {
#include "brief_pattern.inc"
};

// cv::KeyPoint / key_point_t layout — include/OpenCL/Kernel/key_point.hpp:22-29.
struct KeyPoint {
    float x, y, size, angle, response;
    int octave, class_id;
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint is a 28-byte POD");

inline int cvRoundF(float v) { return (int)lrintf(v); }   // cvRound: round-half-to-even (SSE cvtss2si)
inline int cvRoundD(double v) { return (int)lrint(v); }
inline int cvFloorD(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeilD(double v) { int i = (int)v; return i + (i < v); }

// ---------------------------------------------------------------------------------------------
// OpenCV primitives (third-party, system OpenCV >= 4.4; restated, verified vs cv2 4.13 in tests)
// ---------------------------------------------------------------------------------------------

// cv::resize(..., INTER_LINEAR) for CV_8UC1 — call site src/ORBextractor.cc:1320.
// Fixed-point bilinear: 11-bit coefficients, horizontal pass into int32, vertical pass
// (((b0*(r0>>4))>>16) + ((b1*(r1>>4))>>16) + 2) >> 2.  Exact 2x decimation takes OpenCV's
// INTER_AREA fast path ((a+b+c+d+2)>>2).
void resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep)
{
    // cv::resize: inv_scale = (double)dsize/ssize; cv::hal::resize: scale = 1./inv_scale
    double scale_x = 1. / ((double)dw / sw), scale_y = 1. / ((double)dh / sh);
    {
        int iscale_x = cvRoundD(scale_x), iscale_y = cvRoundD(scale_y);   // saturate_cast<int>(double)
        bool is_area_fast = std::abs(scale_x - iscale_x) < DBL_EPSILON && std::abs(scale_y - iscale_y) < DBL_EPSILON;
        if (is_area_fast && iscale_x == 2 && iscale_y == 2) {
            for (int y = 0; y < dh; ++y)
                for (int x = 0; x < dw; ++x) {
                    const uint8_t* s0 = src + (size_t)(2 * y) * sstep + 2 * x;
                    const uint8_t* s1 = s0 + sstep;
                    dst[y * dstep + x] = (uint8_t)((s0[0] + s0[1] + s1[0] + s1[1] + 2) >> 2);
                }
            return;
        }
    }
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> a0(dw), a1(dw), b0(dh), b1(dh);
    auto coeffs = [](int d, double scale, int n, int& s, short& c0, short& c1) {
        float f = (float)((d + 0.5) * scale - 0.5);
        s = cvFloorD(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= n - 1) { s = n - 1; f = 0.f; }
        c0 = (short)cvRoundF((1.f - f) * 2048.f);
        c1 = (short)cvRoundF(f * 2048.f);
    };
    for (int x = 0; x < dw; ++x) coeffs(x, scale_x, sw, xofs[x], a0[x], a1[x]);
    // The vertical direction is NOT clamped like the horizontal one: cv::resize keeps sy and fy as computed and the row
    // loop clips the two row indices into the image (resizeGeneric_Invoker: clip(sy + k, 0, ssize.height)).  Identical for
    // every down-scale (the pyramid); on up-scales the first/last rows blend a row with itself with split truncation.
    for (int y = 0; y < dh; ++y) {
        float f = (float)((y + 0.5) * scale_y - 0.5);
        const int s = cvFloorD(f);
        f -= s;
        yofs[y] = s;
        b0[y] = (short)cvRoundF((1.f - f) * 2048.f);
        b1[y] = (short)cvRoundF(f * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    for (int y = 0; y < dh; ++y) {
        int sy0 = std::min(std::max(yofs[y], 0), sh - 1), sy1 = std::min(std::max(yofs[y] + 1, 0), sh - 1);
        const uint8_t* S0 = src + (size_t)sy0 * sstep;
        const uint8_t* S1 = src + (size_t)sy1 * sstep;
        for (int x = 0; x < dw; ++x) {
            int sx0 = xofs[x], sx1 = std::min(sx0 + 1, sw - 1);
            r0[x] = S0[sx0] * a0[x] + S0[sx1] * a1[x];
            r1[x] = S1[sx0] * a0[x] + S1[sx1] * a1[x];
        }
        for (int x = 0; x < dw; ++x)
            dst[y * dstep + x] = (uint8_t)((((b0[y] * (r0[x] >> 4)) >> 16) + ((b1[y] * (r1[x] >> 4)) >> 16) + 2) >> 2);
    }
}

inline int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

// cv::copyMakeBorder(..., BORDER_REFLECT_101) — call sites src/ORBextractor.cc:1322-1326.
// `interior` points at pixel (0,0) of a w x h image that lives inside a buffer with >= border pixels
// of slack on every side; the slack is filled from the interior.
void fill_border_reflect101(uint8_t* interior, int w, int h, size_t step, int border)
{
    for (int y = -border; y < h + border; ++y) {
        int sy = reflect101(y, h);
        uint8_t* drow = interior + (ptrdiff_t)y * (ptrdiff_t)step;
        const uint8_t* srow = interior + (ptrdiff_t)sy * (ptrdiff_t)step;
        for (int x = -border; x < w + border; ++x) {
            if (y >= 0 && y < h && x >= 0 && x < w) continue;
            drow[x] = srow[reflect101(x, w)];
        }
    }
}

// cv::FAST(img, kps, threshold, nonmaxSuppression=true) TYPE_9_16 — call sites src/ORBextractor.cc:908,925.
// score(p) = max over the 16 arcs of 9 contiguous circle pixels of min(v - c_k) [centre brighter] or
// min(c_k - v) [centre darker], minus 1; corner iff score >= threshold; rows/cols 3..n-4 only;
// 3x3 strict '>' NMS with non-corners counting as 0; output in (y, x) order.
static const int kCircle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

inline int fast_score(const uint8_t* p, size_t step)
{
    int v = p[0];
    int d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - p[(ptrdiff_t)kCircle[k][1] * (ptrdiff_t)step + kCircle[k][0]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = INT_MIN;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

struct Cand { int x, y, score; };

void fast9(const uint8_t* img, int w, int h, size_t step, int threshold, bool nms, std::vector<Cand>& out)
{
    out.clear();
    if (w < 7 || h < 7) return;
    static thread_local std::vector<uint8_t> sc;
    sc.assign((size_t)w * h, 0);
    const ptrdiff_t st = (ptrdiff_t)step;
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            // Cheap exact rejection (speed only; OpenCV does the same kind of early-out): a 9-arc always contains two
            // adjacent compass points of the circle, so a corner at `threshold` needs two adjacent ones beyond it.
            const uint8_t* p = img + (size_t)y * step + x;
            const int v = p[0], t = threshold;
            const int d0 = v - p[3 * st], d4 = v - p[3], d8 = v - p[-3 * st], d12 = v - p[-3];
            const bool b = ((d0 > t) | (d8 > t)) & ((d4 > t) | (d12 > t));
            const bool k = ((d0 < -t) | (d8 < -t)) & ((d4 < -t) | (d12 < -t));
            if (!(b | k)) continue;
            int s = fast_score(p, step);
            if (s >= threshold) sc[(size_t)y * w + x] = (uint8_t)s;   // OpenCV stores (uchar)score
        }
    // A pixel with score >= threshold but threshold == 0 and score == 0 would be stored as 0 and lost
    // by OpenCV's NMS as well (it compares against a zeroed buffer); thresholds here are >= 1.
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (s < threshold || s == 0) continue;
            if (nms) {
                const uint8_t* c = &sc[(size_t)y * w + x];
                if (!(s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] && s > c[w - 1] &&
                      s > c[w] && s > c[w + 1]))
                    continue;
            }
            out.push_back({x, y, s});
        }
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for CV_8UC1 — call site
// src/ORBextractor.cc:1270-1273.  Bit-exact fixed-point kernel [18,34,48,56,48,34,18]/256 per axis.
void gaussian_blur7(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
    static const int k[7] = {18, 34, 48, 56, 48, 34, 18};
    // reflect-101 padded copy (3 px) so that the two passes are plain loops
    const int pw = w + 6, ph = h + 6;
    static thread_local std::vector<uint8_t> pad;
    static thread_local std::vector<uint16_t> hbuf;
    pad.resize((size_t)pw * ph);
    hbuf.resize((size_t)w * ph);
    for (int y = 0; y < ph; ++y) {
        const uint8_t* srow = src + (size_t)reflect101(y - 3, h) * sstep;
        uint8_t* prow = &pad[(size_t)y * pw];
        for (int x = 0; x < 3; ++x) prow[x] = srow[reflect101(x - 3, w)];
        memcpy(prow + 3, srow, w);
        for (int x = 0; x < 3; ++x) prow[w + 3 + x] = srow[reflect101(w + x, w)];
    }
    for (int y = 0; y < ph; ++y) {
        const uint8_t* p = &pad[(size_t)y * pw];
        uint16_t* hrow = &hbuf[(size_t)y * w];
        for (int x = 0; x < w; ++x)
            hrow[x] = (uint16_t)(k[0] * (p[x] + p[x + 6]) + k[1] * (p[x + 1] + p[x + 5]) + k[2] * (p[x + 2] + p[x + 4]) + k[3] * p[x + 3]);
    }
    for (int y = 0; y < h; ++y) {
        const uint16_t* r0 = &hbuf[(size_t)y * w];
        uint8_t* drow = dst + (size_t)y * dstep;
        for (int x = 0; x < w; ++x) {
            uint32_t s = 0;
            for (int t = 0; t < 7; ++t) s += (uint32_t)k[t] * r0[(size_t)t * w + x];
            drow[x] = (uint8_t)((s + 32768u) >> 16);
        }
    }
}

// cv::fastAtan2(y, x) — imported at src/ORBextractor.cc:85 (the CPU IC_Angle that called it was deleted
// by the fork; upstream ORB-SLAM3 IC_Angle ends with `return fastAtan2((float)m_01, (float)m_10);`).
float fast_atan2(float y, float x)
{
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// ---------------------------------------------------------------------------------------------
// Reference-owned logic
// ---------------------------------------------------------------------------------------------

struct Tables {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    std::vector<float> scale, inv, sigma2, invsigma2;
    std::vector<int> nfeat;
    std::vector<int> umax;
};

// ORBextractor::ORBextractor — src/ORBextractor.cc:410-468.
void make_tables(Tables& t, int nfeatures, float scaleFactorF, int nlevels, int iniTh, int minTh)
{
    t.nfeatures = nfeatures; t.nlevels = nlevels; t.iniTh = iniTh; t.minTh = minTh;
    t.scaleFactor = scaleFactorF;                       // member is double (include/ORBextractor.h:105)
    t.scale.assign(nlevels, 1.f); t.sigma2.assign(nlevels, 1.f);
    for (int i = 1; i < nlevels; ++i) {
        t.scale[i] = (float)(t.scale[i - 1] * t.scaleFactor);   // float * double -> double -> float
        t.sigma2[i] = t.scale[i] * t.scale[i];
    }
    t.inv.resize(nlevels); t.invsigma2.resize(nlevels);
    for (int i = 0; i < nlevels; ++i) { t.inv[i] = 1.0f / t.scale[i]; t.invsigma2[i] = 1.0f / t.sigma2[i]; }
    t.nfeat.assign(nlevels, 0);
    float factor = (float)(1.0f / t.scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; ++level) {
        t.nfeat[level] = cvRoundF(nDesired);
        sum += t.nfeat[level];
        nDesired *= factor;
    }
    t.nfeat[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax — src/ORBextractor.cc:453-467
    t.umax.assign(HALF_PATCH_SIZE + 1, 0);
    int v, v0, vmax = cvFloorD(HALF_PATCH_SIZE * sqrt(2.f) / 2 + 1);
    int vmin = cvCeilD(HALF_PATCH_SIZE * sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) t.umax[v] = cvRoundD(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (t.umax[v0] == t.umax[v0 + 1]) ++v0;
        t.umax[v] = v0;
        ++v0;
    }
}

struct Level {
    int w = 0, h = 0;
    size_t step = 0;                 // bordered row pitch
    std::vector<uint8_t> buf;        // (h+38) x (w+38)
    std::vector<uint8_t> blurred;    // h x w, pitch w
    uint8_t* interior() { return buf.data() + (size_t)EDGE_THRESHOLD * step + EDGE_THRESHOLD; }
    const uint8_t* interior() const { return buf.data() + (size_t)EDGE_THRESHOLD * step + EDGE_THRESHOLD; }
};

// ORBextractor::ComputePyramid — src/ORBextractor.cc:1309-1329.
void compute_pyramid(const Tables& t, const uint8_t* img, int rows, int cols, size_t step, std::vector<Level>& pyr)
{
    pyr.resize(t.nlevels);
    for (int level = 0; level < t.nlevels; ++level) {
        float scale = t.inv[level];
        Level& L = pyr[level];
        L.w = cvRoundF((float)cols * scale);
        L.h = cvRoundF((float)rows * scale);
        L.step = (size_t)L.w + 2 * EDGE_THRESHOLD;
        L.buf.assign(L.step * (size_t)(L.h + 2 * EDGE_THRESHOLD), 0);
        if (level != 0) {
            const Level& P = pyr[level - 1];
            resize_linear_u8(P.interior(), P.w, P.h, P.step, L.interior(), L.w, L.h, L.step);
        } else {
            for (int y = 0; y < rows; ++y) memcpy(L.interior() + (size_t)y * L.step, img + (size_t)y * step, cols);
        }
        fill_border_reflect101(L.interior(), L.w, L.h, L.step, EDGE_THRESHOLD);
    }
}

// tileCalcKeypoints — src/ORBextractor.cc:867-950 (CPU cell loop; the semantics the fork's broken
// OpenCL kernel was meant to reproduce), window set up at src/ORBextractor.cc:958-966.
// Output coordinates are relative to (minBorderX, minBorderY) = (16,16), order = cell row-major then (y,x).
void cell_fast(const uint8_t* interior, int w, int h, size_t step, int iniTh, int minTh, std::vector<Cand>& out)
{
    out.clear();
    const int border = EDGE_THRESHOLD - 3;
    const int minBorderX = border, minBorderY = border;
    const float W = 35;
    const int maxBorderX = w - EDGE_THRESHOLD + 3;
    const int maxBorderY = h - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX);
    const float height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W);
    const int nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;   // reference would divide by zero; treated as "no keypoints"
    const int wCell = (int)ceil(width / nCols);
    const int hCell = (int)ceil(height / nRows);
    std::vector<Cand> cell;
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
            const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const uint8_t* roi = interior + (size_t)y0 * step + x0;
            fast9(roi, x1 - x0, y1 - y0, step, iniTh, true, cell);
            if (cell.empty()) fast9(roi, x1 - x0, y1 - y0, step, minTh, true, cell);
            for (const Cand& c : cell) out.push_back({c.x + j * wCell, c.y + i * hCell, c.score});
        }
    }
}

// ---- octree: ExtractorNode (include/ORBextractor.h:39-50), DivideNode (src/ORBextractor.cc:515-567),
// compareNodes (569-582), DistributeOctTree (584-774).  Keys carry their input index.
struct Key { float x, y, response; int idx; };
struct Pt { int x, y; };
struct Node {
    std::vector<Key> vKeys;
    Pt UL{0, 0}, UR{0, 0}, BL{0, 0}, BR{0, 0};
    std::list<Node>::iterator lit;
    bool bNoMore = false;
    void Divide(Node& n1, Node& n2, Node& n3, Node& n4) const;
};

void Node::Divide(Node& n1, Node& n2, Node& n3, Node& n4) const
{
    const int halfX = (int)ceil(static_cast<float>(UR.x - UL.x) / 2);
    const int halfY = (int)ceil(static_cast<float>(BR.y - UL.y) / 2);
    n1.UL = UL; n1.UR = {UL.x + halfX, UL.y}; n1.BL = {UL.x, UL.y + halfY}; n1.BR = {UL.x + halfX, UL.y + halfY};
    n2.UL = n1.UR; n2.UR = UR; n2.BL = n1.BR; n2.BR = {UR.x, UL.y + halfY};
    n3.UL = n1.BL; n3.UR = n1.BR; n3.BL = BL; n3.BR = {n1.BR.x, BL.y};
    n4.UL = n3.UR; n4.UR = n2.BR; n4.BL = n3.BR; n4.BR = BR;
    for (const Key& kp : vKeys) {
        if (kp.x < n1.UR.x) { if (kp.y < n1.BR.y) n1.vKeys.push_back(kp); else n3.vKeys.push_back(kp); }
        else if (kp.y < n1.BR.y) n2.vKeys.push_back(kp);
        else n4.vKeys.push_back(kp);
    }
    if (n1.vKeys.size() == 1) n1.bNoMore = true;
    if (n2.vKeys.size() == 1) n2.bNoMore = true;
    if (n3.vKeys.size() == 1) n3.bNoMore = true;
    if (n4.vKeys.size() == 1) n4.bNoMore = true;
}

static bool compareNodes(std::pair<int, Node*>& e1, std::pair<int, Node*>& e2)
{
    if (e1.first < e2.first) return true;
    else if (e1.first > e2.first) return false;
    else return e1.second->UL.x < e2.second->UL.x;
}

void distribute_octree(const std::vector<Cand>& cands, int minX, int maxX, int minY, int maxY, int N, std::vector<int>& outIdx)
{
    outIdx.clear();
    const int nIni = (int)round(static_cast<float>(maxX - minX) / (maxY - minY));
    if (nIni <= 0) return;   // reference quirk (division by zero for portrait windows); "no keypoints"
    const float hX = static_cast<float>(maxX - minX) / nIni;
    std::list<Node> lNodes;
    std::vector<Node*> vpIniNodes(nIni);
    for (int i = 0; i < nIni; i++) {
        Node ni;
        ni.UL = {(int)(hX * static_cast<float>(i)), 0};
        ni.UR = {(int)(hX * static_cast<float>(i + 1)), 0};
        ni.BL = {ni.UL.x, maxY - minY};
        ni.BR = {ni.UR.x, maxY - minY};
        lNodes.push_back(ni);
        vpIniNodes[i] = &lNodes.back();
    }
    for (size_t i = 0; i < cands.size(); i++) {
        Key k{(float)cands[i].x, (float)cands[i].y, (float)cands[i].score, (int)i};
        vpIniNodes[(size_t)(k.x / hX)]->vKeys.push_back(k);
    }
    auto lit = lNodes.begin();
    while (lit != lNodes.end()) {
        if (lit->vKeys.size() == 1) { lit->bNoMore = true; lit++; }
        else if (lit->vKeys.empty()) lit = lNodes.erase(lit);
        else lit++;
    }
    bool bFinish = false;
    std::vector<std::pair<int, Node*>> vSizeAndPointerToNode;
    auto push_children = [&](Node* ch[4], int* nToExpand) {
        for (int c = 0; c < 4; ++c) {
            if (ch[c]->vKeys.size() > 0) {
                lNodes.push_front(*ch[c]);
                if (ch[c]->vKeys.size() > 1) {
                    if (nToExpand) (*nToExpand)++;
                    vSizeAndPointerToNode.push_back(std::make_pair((int)ch[c]->vKeys.size(), &lNodes.front()));
                    lNodes.front().lit = lNodes.begin();
                }
            }
        }
    };
    while (!bFinish) {
        int prevSize = (int)lNodes.size();
        lit = lNodes.begin();
        int nToExpand = 0;
        vSizeAndPointerToNode.clear();
        while (lit != lNodes.end()) {
            if (lit->bNoMore) { lit++; continue; }
            Node n1, n2, n3, n4;
            lit->Divide(n1, n2, n3, n4);
            Node* ch[4] = {&n1, &n2, &n3, &n4};
            push_children(ch, &nToExpand);
            lit = lNodes.erase(lit);
        }
        if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) {
            bFinish = true;
        } else if (((int)lNodes.size() + nToExpand * 3) > N) {
            while (!bFinish) {
                prevSize = (int)lNodes.size();
                std::vector<std::pair<int, Node*>> vPrev = vSizeAndPointerToNode;
                vSizeAndPointerToNode.clear();
                std::sort(vPrev.begin(), vPrev.end(), compareNodes);   // libstdc++ introsort; ties matter
                for (int j = (int)vPrev.size() - 1; j >= 0; j--) {
                    Node n1, n2, n3, n4;
                    vPrev[j].second->Divide(n1, n2, n3, n4);
                    Node* ch[4] = {&n1, &n2, &n3, &n4};
                    push_children(ch, nullptr);
                    lNodes.erase(vPrev[j].second->lit);
                    if ((int)lNodes.size() >= N) break;
                }
                if ((int)lNodes.size() >= N || (int)lNodes.size() == prevSize) bFinish = true;
            }
        }
    }
    for (auto it = lNodes.begin(); it != lNodes.end(); it++) {
        const std::vector<Key>& v = it->vKeys;
        const Key* p = &v[0];
        float maxResponse = p->response;
        for (size_t k = 1; k < v.size(); k++)
            if (v[k].response > maxResponse) { p = &v[k]; maxResponse = v[k].response; }
        outIdx.push_back(p->idx);
    }
}

// IC_Angle — summation pattern src/OpenCL/Kernel/Angle.cl:24-53 with umax from src/ORBextractor.cc:455-467,
// closed with cv::fastAtan2 as in the CPU path the fork deleted (src/ORBextractor.cc:85).
float ic_angle(const uint8_t* img, size_t step, int x, int y, const std::vector<int>& umax)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = img + (size_t)y * step + x;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + (ptrdiff_t)v * (ptrdiff_t)step], val_minus = center[u - (ptrdiff_t)v * (ptrdiff_t)step];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
}

// computeOrbDescriptor — src/ORBextractor.cc:105-149.
void orb_descriptor(const uint8_t* img, size_t step_, int x, int y, float angle_deg, uint8_t* desc)
{
    const float factorPI = (float)(M_PI / 180.f);
    float angle = angle_deg * factorPI;
    float a = (float)cos(angle), b = (float)sin(angle);
    const uint8_t* center = img + (size_t)y * step_ + x;
    const int step = (int)step_;
    const int8_t* pat = kPattern;
    auto get = [&](int idx) -> int {
        float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        // volatile-free but un-contracted: this TU is built with -ffp-contract=off
        int iy = cvRoundF(px * b + py * a);
        int ix = cvRoundF(px * a - py * b);
        return center[iy * step + ix];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            int t0 = get(2 * k), t1 = get(2 * k + 1);
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// ORBmatcher::DescriptorDistance — src/ORBmatcher3.cc:637-653 (SWAR bit-hack, 8 x int32).
int descriptor_distance(const uint8_t* a, const uint8_t* b)
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

inline int hamming_popcnt(const uint8_t* a, const uint8_t* b)
{
    const uint64_t* pa = (const uint64_t*)a;
    const uint64_t* pb = (const uint64_t*)b;
    return __builtin_popcountll(pa[0] ^ pb[0]) + __builtin_popcountll(pa[1] ^ pb[1]) +
           __builtin_popcountll(pa[2] ^ pb[2]) + __builtin_popcountll(pa[3] ^ pb[3]);
}

struct Extractor {
    Tables t;
    std::vector<Level> pyr;
    std::vector<std::vector<Cand>> cands;        // per level, window-relative
    std::vector<std::vector<KeyPoint>> levelKps; // per level, level coordinates (border added, angle set)
    std::vector<std::vector<uint8_t>> levelDesc;
    // wall time per stage of the calls so far (bench.py's cpu_baseline breaks the frame time down with it): pyramid, cell FAST,
    // octree + border + orientation, blur, descriptors + packing
    double stage_ms[5] = {0, 0, 0, 0, 0};
};

static inline double now_ms()
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// =============================================================================================
// C ABI (ctypes-friendly).  Mirrors include/orbx.h so tests can drive both back ends alike.
// =============================================================================================
extern "C" {

typedef KeyPoint orbo_keypoint;

int orbo_tables(int nfeatures, float scaleFactor, int nlevels, float* scale, float* inv, float* sigma2,
                float* invsigma2, int* nfeat_per_level, int* umax16)
{
    Tables t;
    make_tables(t, nfeatures, scaleFactor, nlevels, 20, 7);
    for (int i = 0; i < nlevels; ++i) {
        if (scale) scale[i] = t.scale[i];
        if (inv) inv[i] = t.inv[i];
        if (sigma2) sigma2[i] = t.sigma2[i];
        if (invsigma2) invsigma2[i] = t.invsigma2[i];
        if (nfeat_per_level) nfeat_per_level[i] = t.nfeat[i];
    }
    if (umax16) for (int i = 0; i < 16; ++i) umax16[i] = t.umax[i];
    return 0;
}

void orbo_resize_linear_u8(const uint8_t* src, int sw, int sh, size_t sstep, uint8_t* dst, int dw, int dh, size_t dstep)
{
    resize_linear_u8(src, sw, sh, sstep, dst, dw, dh, dstep);
}

void orbo_fill_border_reflect101(uint8_t* interior, int w, int h, size_t step, int border)
{
    fill_border_reflect101(interior, w, h, step, border);
}

int orbo_fast9(const uint8_t* img, int w, int h, size_t step, int threshold, int nms, int* xs, int* ys, int* scores, int cap)
{
    std::vector<Cand> out;
    fast9(img, w, h, step, threshold, nms != 0, out);
    int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) { xs[i] = out[i].x; ys[i] = out[i].y; scores[i] = out[i].score; }
    return n;
}

int orbo_cell_fast(const uint8_t* interior, int w, int h, size_t step, int iniTh, int minTh, int* xs, int* ys, int* scores, int cap)
{
    std::vector<Cand> out;
    cell_fast(interior, w, h, step, iniTh, minTh, out);
    int n = (int)out.size();
    for (int i = 0; i < n && i < cap; ++i) { xs[i] = out[i].x; ys[i] = out[i].y; scores[i] = out[i].score; }
    return n;
}

int orbo_octree(const int* xs, const int* ys, const int* scores, int n, int minX, int maxX, int minY, int maxY, int N,
                int* out_idx, int cap)
{
    std::vector<Cand> c(n);
    for (int i = 0; i < n; ++i) c[i] = {xs[i], ys[i], scores[i]};
    std::vector<int> out;
    distribute_octree(c, minX, maxX, minY, maxY, N, out);
    int m = (int)out.size();
    for (int i = 0; i < m && i < cap; ++i) out_idx[i] = out[i];
    return m;
}

float orbo_fast_atan2(float y, float x) { return fast_atan2(y, x); }

float orbo_ic_angle(const uint8_t* img, size_t step, int x, int y)
{
    Tables t;
    make_tables(t, 1000, 1.2f, 8, 20, 7);
    return ic_angle(img, step, x, y, t.umax);
}

void orbo_gaussian_blur7(const uint8_t* src, int w, int h, size_t sstep, uint8_t* dst, size_t dstep)
{
    gaussian_blur7(src, w, h, sstep, dst, dstep);
}

void orbo_descriptor(const uint8_t* blurred, size_t step, int x, int y, float angle_deg, uint8_t* out32)
{
    orb_descriptor(blurred, step, x, y, angle_deg, out32);
}

void* orbo_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    Extractor* e = new Extractor();
    make_tables(e->t, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
    return e;
}

void orbo_destroy(void* h) { delete (Extractor*)h; }

// ORBextractor::operator() — src/ORBextractor.cc:1227-1307, with ComputeKeyPointsOctTree (958-1027)
// taken on its CPU path (tileCalcKeypoints) and orientation as upstream IC_Angle.
// Returns 0, or -1 for an empty image (reference returns -1 from operator()).
int orbo_extract(void* h, const uint8_t* img, int rows, int cols, size_t step, int lap0, int lap1,
                 orbo_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* n_mono)
{
    Extractor& e = *(Extractor*)h;
    const Tables& t = e.t;
    if (!img || rows <= 0 || cols <= 0) return -1;
    double t0 = now_ms(), t1;
    compute_pyramid(t, img, rows, cols, step, e.pyr);
    t1 = now_ms(); e.stage_ms[0] += t1 - t0; t0 = t1;
    e.cands.assign(t.nlevels, {});
    e.levelKps.assign(t.nlevels, {});
    e.levelDesc.assign(t.nlevels, {});
    const int border = EDGE_THRESHOLD - 3;
    for (int level = 0; level < t.nlevels; ++level) {
        Level& L = e.pyr[level];
        const int maxBorderX = L.w - EDGE_THRESHOLD + 3;
        const int maxBorderY = L.h - EDGE_THRESHOLD + 3;
        cell_fast(L.interior(), L.w, L.h, L.step, t.iniTh, t.minTh, e.cands[level]);
        t1 = now_ms(); e.stage_ms[1] += t1 - t0; t0 = t1;
        std::vector<int> keep;
        distribute_octree(e.cands[level], border, maxBorderX, border, maxBorderY, t.nfeat[level], keep);
        const int scaledPatchSize = (int)(PATCH_SIZE * t.scale[level]);
        for (int idx : keep) {
            const Cand& c = e.cands[level][idx];
            KeyPoint kp;
            kp.x = (float)c.x + border;          // AddBorder.cl:3-20 / src/ORBextractor.cc:1004-1012
            kp.y = (float)c.y + border;
            kp.size = (float)scaledPatchSize;
            kp.response = (float)c.score;
            kp.octave = level;
            kp.class_id = -1;
            kp.angle = ic_angle(L.interior(), L.step, (int)kp.x, (int)kp.y, t.umax);
            e.levelKps[level].push_back(kp);
        }
        t1 = now_ms(); e.stage_ms[2] += t1 - t0; t0 = t1;
    }
    int nkeypoints = 0;
    for (int level = 0; level < t.nlevels; ++level) nkeypoints += (int)e.levelKps[level].size();
    *n_out = nkeypoints;
    if (nkeypoints > cap) return -2;
    int monoIndex = 0, stereoIndex = nkeypoints - 1;
    for (int level = 0; level < t.nlevels; ++level) {
        std::vector<KeyPoint>& v = e.levelKps[level];
        if (v.empty()) continue;
        Level& L = e.pyr[level];
        L.blurred.resize((size_t)L.w * L.h);
        t0 = now_ms();
        gaussian_blur7(L.interior(), L.w, L.h, L.step, L.blurred.data(), L.w);
        t1 = now_ms(); e.stage_ms[3] += t1 - t0; t0 = t1;
        e.levelDesc[level].resize(v.size() * 32);
        for (size_t i = 0; i < v.size(); ++i)
            orb_descriptor(L.blurred.data(), L.w, cvRoundF(v[i].x), cvRoundF(v[i].y), v[i].angle, &e.levelDesc[level][i * 32]);
        float scale = t.scale[level];
        for (size_t i = 0; i < v.size(); ++i) {
            KeyPoint kp = v[i];
            if (level != 0) { kp.x *= scale; kp.y *= scale; }
            int dst;
            if (kp.x >= lap0 && kp.x <= lap1) dst = stereoIndex--;
            else dst = monoIndex++;
            kps[dst] = kp;
            memcpy(desc + (size_t)dst * 32, &e.levelDesc[level][i * 32], 32);
        }
        t1 = now_ms(); e.stage_ms[4] += t1 - t0;
    }
    *n_mono = monoIndex;
    return 0;
}

// accumulated wall time per stage (ms) since creation / the last reset: pyramid, FAST, octree + orientation, blur, descriptors
void orbo_stage_times(void* h, double* ms5, int reset)
{
    Extractor& e = *(Extractor*)h;
    for (int i = 0; i < 5; ++i) { ms5[i] = e.stage_ms[i]; if (reset) e.stage_ms[i] = 0; }
}

int orbo_level_size(void* h, int level, int* w, int* hh)
{
    Extractor& e = *(Extractor*)h;
    if (level < 0 || level >= (int)e.pyr.size()) return -1;
    *w = e.pyr[level].w; *hh = e.pyr[level].h;
    return 0;
}

// with_border: copy the full (h+38)x(w+38) bordered buffer, else the interior.
int orbo_get_pyramid_level(void* h, int level, uint8_t* dst, size_t dst_step, int with_border)
{
    Extractor& e = *(Extractor*)h;
    if (level < 0 || level >= (int)e.pyr.size()) return -1;
    Level& L = e.pyr[level];
    if (with_border) {
        for (int y = 0; y < L.h + 2 * EDGE_THRESHOLD; ++y) memcpy(dst + (size_t)y * dst_step, L.buf.data() + (size_t)y * L.step, L.w + 2 * EDGE_THRESHOLD);
    } else {
        for (int y = 0; y < L.h; ++y) memcpy(dst + (size_t)y * dst_step, L.interior() + (size_t)y * L.step, L.w);
    }
    return 0;
}

int orbo_get_blurred_level(void* h, int level, uint8_t* dst, size_t dst_step)
{
    Extractor& e = *(Extractor*)h;
    if (level < 0 || level >= (int)e.pyr.size()) return -1;
    Level& L = e.pyr[level];
    if (L.blurred.empty()) {   // levels without keypoints are never blurred by the reference; do it on demand
        L.blurred.resize((size_t)L.w * L.h);
        gaussian_blur7(L.interior(), L.w, L.h, L.step, L.blurred.data(), L.w);
    }
    for (int y = 0; y < L.h; ++y) memcpy(dst + (size_t)y * dst_step, L.blurred.data() + (size_t)y * L.w, L.w);
    return 0;
}

int orbo_get_candidates(void* h, int level, int* xs, int* ys, int* scores, int cap)
{
    Extractor& e = *(Extractor*)h;
    if (level < 0 || level >= (int)e.cands.size()) return -1;
    int n = (int)e.cands[level].size();
    for (int i = 0; i < n && i < cap; ++i) { xs[i] = e.cands[level][i].x; ys[i] = e.cands[level][i].y; scores[i] = e.cands[level][i].score; }
    return n;
}

int orbo_get_level_keypoints(void* h, int level, orbo_keypoint* kps, uint8_t* desc, int cap)
{
    Extractor& e = *(Extractor*)h;
    if (level < 0 || level >= (int)e.levelKps.size()) return -1;
    int n = (int)e.levelKps[level].size();
    for (int i = 0; i < n && i < cap; ++i) {
        kps[i] = e.levelKps[level][i];
        if (desc && !e.levelDesc[level].empty()) memcpy(desc + (size_t)i * 32, &e.levelDesc[level][(size_t)i * 32], 32);
    }
    return n;
}

int orbo_hamming_swar(const uint8_t* a, const uint8_t* b) { return descriptor_distance(a, b); }
int orbo_hamming(const uint8_t* a, const uint8_t* b) { return hamming_popcnt(a, b); }

// Brute-force 2-NN — cv::BFMatcher(NORM_HAMMING).knnMatch(k=2) as used at src/Frame.cc:1174, identical in
// (d1,i1,d2) to the strict-'<' scan of src/ORBmatcher1.cc:283-300 over the whole set: ascending
// distance, ties -> lower train index.  idx/dist are [nq][2]; missing neighbours are (-1, 256... INT_MAX).
// `use_swar` selects the reference's bit-hack (1) or hardware popcnt (0) — same results, different speed.
void orbo_knn2(const uint8_t* q, int nq, const uint8_t* db, int ndb, int32_t* idx, int32_t* dist, int use_swar)
{
    for (int i = 0; i < nq; ++i) {
        int d1 = INT_MAX, d2 = INT_MAX, i1 = -1, i2 = -1;
        const uint8_t* a = q + (size_t)i * 32;
        for (int j = 0; j < ndb; ++j) {
            int d = use_swar ? descriptor_distance(a, db + (size_t)j * 32) : hamming_popcnt(a, db + (size_t)j * 32);
            if (d < d1) { d2 = d1; i2 = i1; d1 = d; i1 = j; }
            else if (d < d2) { d2 = d; i2 = j; }
        }
        idx[2 * i] = i1; idx[2 * i + 1] = i2; dist[2 * i] = d1; dist[2 * i + 1] = d2;
    }
}

// Ratio-test acceptance — src/ORBmatcher1.cc:329-333 (TH_LOW gate + mfNNratio) when th_low >= 0, and
// src/Frame.cc:1181 (`d1 < d2 * 0.7`, double arithmetic on float distances) when th_low < 0.
int orbo_ratio_accept(int d1, int d2, float ratio, int th_low)
{
    if (th_low >= 0) return d1 <= th_low && (float)d1 < ratio * (float)d2;
    return (float)d1 < (float)d2 * (double)ratio;
}

// Frame::ComputeStereoMatches — src/Frame.cc:841-1011.  Pyramids are passed as arrays of interior pointers
// (un-blurred levels; both images' levels have the same sizes).  mb/mbf quirk: maxD is explicit.
// Returns number of matches surviving the median filter.
int orbo_stereo_match(const orbo_keypoint* kpL, const uint8_t* descL, int nL, const orbo_keypoint* kpR,
                      const uint8_t* descR, int nR, const uint8_t* const* pyrL, const uint8_t* const* pyrR,
                      const size_t* steps, const int* widths, int nRows, const float* scaleFactors,
                      const float* invScaleFactors, float mbf, float maxD_, float* uRight, float* depth)
{
    const int TH_HIGH = 100, TH_LOW = 50;           // src/ORBmatcher1.cc:37-38
    for (int i = 0; i < nL; ++i) { uRight[i] = -1.0f; depth[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    std::vector<std::vector<size_t>> vRowIndices(nRows);
    for (int iR = 0; iR < nR; iR++) {
        const float kpY = kpR[iR].y;
        const float r = 2.0f * scaleFactors[kpR[iR].octave];
        const int maxr = (int)ceil(kpY + r);
        const int minr = (int)floor(kpY - r);
        for (int yi = minr; yi <= maxr; yi++)
            if (yi >= 0 && yi < nRows) vRowIndices[yi].push_back(iR);   // reference indexes unchecked
    }
    const float minD = 0;
    const float maxD = maxD_;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < nL; iL++) {
        const orbo_keypoint& kL = kpL[iL];
        const int levelL = kL.octave;
        const float vL = kL.y, uL = kL.x;
        const std::vector<size_t>& vCandidates = vRowIndices[(size_t)vL];
        if (vCandidates.empty()) continue;
        const float minU = uL - maxD;
        const float maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH;
        size_t bestIdxR = 0;
        const uint8_t* dL = descL + (size_t)iL * 32;
        for (size_t iC = 0; iC < vCandidates.size(); iC++) {
            const size_t iR = vCandidates[iC];
            const orbo_keypoint& kR = kpR[iR];
            if (kR.octave < levelL - 1 || kR.octave > levelL + 1) continue;
            const float uR = kR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = descriptor_distance(dL, descR + iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = kpR[bestIdxR].x;
            const float scaleFactor = invScaleFactors[kL.octave];
            const float scaleduL = roundf(kL.x * scaleFactor);
            const float scaledvL = roundf(kL.y * scaleFactor);
            const float scaleduR0 = roundf(uR0 * scaleFactor);
            const int w = 5;
            const uint8_t* IL = pyrL[kL.octave];
            const uint8_t* IRm = pyrR[kL.octave];
            const size_t st = steps[kL.octave];
            int bestDistS = INT_MAX;
            int bestincR = 0;
            const int L = 5;
            float vDists[2 * 5 + 1];
            const float iniu = scaleduR0 + L - w;
            const float endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= widths[kL.octave]) continue;
            for (int incR = -L; incR <= +L; incR++) {
                double acc = 0;    // cv::norm(IL, IR, NORM_L1) over 11x11 u8
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        int a = IL[(ptrdiff_t)((int)scaledvL + dy) * (ptrdiff_t)st + ((int)scaleduL + dx)];
                        int b = IRm[(ptrdiff_t)((int)scaledvL + dy) * (ptrdiff_t)st + ((int)scaleduR0 + incR + dx)];
                        acc += std::abs(a - b);
                    }
                float dist = (float)acc;
                if (dist < bestDistS) { bestDistS = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1];
            const float dist2 = vDists[L + bestincR];
            const float dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scaleFactors[kL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01f; bestuR = uL - 0.01f; }
                depth[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestDistS, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;   // reference would read vDistIdx[0] of an empty vector (UB)
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; i--) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1;
        depth[vDistIdx[i].second] = -1;
        kept--;
    }
    return kept;
}

// Rotation histogram + ComputeThreeMaxima — src/ORBmatcher1.cc:236-238, 344-356, 408-427; src/ORBmatcher3.cc:592-633.
void orbo_rotation_consistency(const float* angle_a, const float* angle_b, int n, uint8_t* keep)
{
    const int HISTO_LENGTH = 30;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    for (int i = 0; i < n; ++i) {
        float rot = angle_a[i] - angle_b[i];
        if (rot < 0.0) rot += 360.0f;
        int bin = (int)round(rot * factor);
        if (bin == HISTO_LENGTH) bin = 0;
        rotHist[bin].push_back(i);
    }
    int ind1 = -1, ind2 = -1, ind3 = -1;
    {
        int max1 = 0, max2 = 0, max3 = 0;
        for (int i = 0; i < HISTO_LENGTH; i++) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    }
    for (int i = 0; i < n; ++i) keep[i] = 1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
        if (i == ind1 || i == ind2 || i == ind3) continue;
        for (size_t j = 0; j < rotHist[i].size(); j++) keep[rotHist[i][j]] = 0;
    }
}

// MapPoint::ComputeDistinctiveDescriptors — src/MapPoint.cc:368-395.
int orbo_distinctive_descriptor(const uint8_t* desc, int N)
{
    std::vector<std::vector<float>> Distances(N, std::vector<float>(N, 0.f));
    for (int i = 0; i < N; i++) {
        Distances[i][i] = 0;
        for (int j = i + 1; j < N; j++) {
            int distij = descriptor_distance(desc + (size_t)i * 32, desc + (size_t)j * 32);
            Distances[i][j] = (float)distij;
            Distances[j][i] = (float)distij;
        }
    }
    int BestMedian = INT_MAX, BestIdx = 0;
    for (int i = 0; i < N; i++) {
        std::vector<int> vDists(Distances[i].begin(), Distances[i].end());
        std::sort(vDists.begin(), vDists.end());
        int median = vDists[(size_t)(0.5 * (N - 1))];
        if (median < BestMedian) { BestMedian = median; BestIdx = i; }
    }
    return BestIdx;
}

// std::sort with a comparator that looks only at the bits above bit 24 (the GPU octree's replay of libstdc++'s
// introsort is checked against this, tests/test_introsort.py).
void orbo_std_sort_hi40(unsigned long long* items, int n)
{
    std::sort(items, items + n, [](unsigned long long& a, unsigned long long& b) { return (a >> 24) < (b >> 24); });
}

}  // extern "C"

// =====================================================================================================================
// Bag of words (DBoW2) and the vocabulary-guided searches.  DBoW2 is vendored in the reference tree
// (Thirdparty/DBoW2) but needs OpenCV + boost::serialization headers to compile, which this image lacks, so it is
// restated here like the rest; the std::map containers are the same ones DBoW2 derives from (BowVector.h:56-57,
// FeatureVector.h:23-25), so iteration and floating-point summation order are identical by construction.
// =====================================================================================================================
#include <map>

namespace {

// TemplatedVocabulary<FORB::TDescriptor, FORB>::Node — Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:182-218
struct VocNode {
    int id = 0, parent = 0;
    double weight = 0;
    std::vector<int> children;
    uint8_t descriptor[32] = {0};
    int word_id = -1;
    bool isLeaf() const { return children.empty(); }
};
struct Vocabulary {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<VocNode> nodes;
    int n_words = 0;
};

// FORB::distance — Thirdparty/DBoW2/DBoW2/FORB.cpp:81-101
int forb_distance(const uint8_t* a, const uint8_t* b)
{
    int32_t pa[8], pb[8];
    memcpy(pa, a, 32); memcpy(pb, b, 32);
    int dist = 0;
    for (int i = 0; i < 8; i++) {
        unsigned int v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

// transform(feature, word_id, weight, nid, levelsup) — TemplatedVocabulary.h:1217-1259
void voc_transform_one(const Vocabulary& V, const uint8_t* feature, unsigned& word_id, double& weight, unsigned* nid, int levelsup)
{
    std::vector<int> nodes;
    const int nid_level = V.L - levelsup;
    if (nid_level <= 0 && nid != nullptr) *nid = 0;   // root
    int final_id = 0;
    int current_level = 0;
    do {
        ++current_level;
        nodes = V.nodes[final_id].children;
        final_id = nodes[0];
        double best_d = forb_distance(feature, V.nodes[final_id].descriptor);
        for (size_t n = 1; n < nodes.size(); ++n) {
            const int id = nodes[n];
            const double d = forb_distance(feature, V.nodes[id].descriptor);
            if (d < best_d) { best_d = d; final_id = id; }
        }
        if (nid != nullptr && current_level == nid_level) *nid = final_id;
    } while (!V.nodes[final_id].isLeaf());
    word_id = V.nodes[final_id].word_id;
    weight = V.nodes[final_id].weight;
}

// DBoW2::FeatureVector in CSR form (node ids ascending)
struct FeatVec {
    std::map<unsigned, std::vector<unsigned>> m;
    FeatVec(int n_nodes, const uint32_t* node_ids, const int32_t* offsets, const uint32_t* indices)
    {
        for (int i = 0; i < n_nodes; ++i) m[node_ids[i]] = std::vector<unsigned>(indices + offsets[i], indices + offsets[i + 1]);
    }
};

const int TH_LOW = 50;          // src/ORBmatcher1.cc:37
const int HISTO_LENGTH = 30;    // src/ORBmatcher1.cc:39

// ORBmatcher::ComputeThreeMaxima — src/ORBmatcher3.cc:592-633
void compute_three_maxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3)
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

}  // namespace

extern "C" {

void* orbo_vocab_create(int n_nodes, const int32_t* parent, const uint8_t* descriptors, const double* weights, int k, int L,
                        int scoring, int weighting)
{
    // node numbering, children order and word ids as built by loadFromTextFile — TemplatedVocabulary.h:1378-1418
    Vocabulary* V = new Vocabulary;
    V->k = k; V->L = L; V->scoring = scoring; V->weighting = weighting;
    V->nodes.resize(n_nodes);
    for (int nid = 1; nid < n_nodes; ++nid) {
        V->nodes[nid].id = nid;
        V->nodes[nid].parent = parent[nid];
        V->nodes[parent[nid]].children.push_back(nid);
        memcpy(V->nodes[nid].descriptor, descriptors + (size_t)nid * 32, 32);
        V->nodes[nid].weight = weights[nid];
    }
    for (int nid = 1; nid < n_nodes; ++nid)
        if (V->nodes[nid].isLeaf()) V->nodes[nid].word_id = V->n_words++;
    return V;
}
void orbo_vocab_destroy(void* v) { delete (Vocabulary*)v; }

void orbo_bow_transform(void* v, const uint8_t* desc, int n, int levelsup, uint32_t* word_id, double* weight, uint32_t* node_id)
{
    const Vocabulary& V = *(Vocabulary*)v;
    for (int i = 0; i < n; ++i) {
        unsigned id = 0, nid = 0; double w = 0;
        voc_transform_one(V, desc + (size_t)i * 32, id, w, &nid, levelsup);
        word_id[i] = id; weight[i] = w; node_id[i] = nid;
    }
}

// transform(features, BowVector, FeatureVector, levelsup) — TemplatedVocabulary.h:1126-1204; BowVector.cpp:34-86;
// FeatureVector.cpp:32-46.  Called by Frame::ComputeBoW with levelsup = 4 (src/Frame.cc:768-775).
void orbo_compute_bow(void* v, const uint8_t* desc, int n, int levelsup, uint32_t* bow_ids, double* bow_vals, int* n_bow,
                      uint32_t* fv_nodes, int32_t* fv_offsets, uint32_t* fv_indices, int* n_fv)
{
    const Vocabulary& V = *(Vocabulary*)v;
    std::map<unsigned, double> bow;
    std::map<unsigned, std::vector<unsigned>> fv;
    const bool must = V.scoring != 5;                 // every scoring class but DotProduct normalises (ScoringObject.h:73-89)
    const bool l2 = V.scoring == 1;
    if (V.weighting == 1 || V.weighting == 0) {       // TF || TF_IDF
        unsigned i_feature = 0;
        for (int f = 0; f < n; ++f, ++i_feature) {
            unsigned id, nid = 0; double w;
            voc_transform_one(V, desc + (size_t)f * 32, id, w, &nid, levelsup);
            if (w > 0) {
                auto vit = bow.lower_bound(id);       // BowVector::addWeight
                if (vit != bow.end() && !(bow.key_comp()(id, vit->first))) vit->second += w;
                else bow.insert(vit, std::map<unsigned, double>::value_type(id, w));
                auto fit = fv.lower_bound(nid);       // FeatureVector::addFeature
                if (fit != fv.end() && fit->first == nid) fit->second.push_back(i_feature);
                else { fit = fv.insert(fit, std::map<unsigned, std::vector<unsigned>>::value_type(nid, std::vector<unsigned>())); fit->second.push_back(i_feature); }
            }
        }
        if (!bow.empty() && !must) {
            const double nd = bow.size();
            for (auto vit = bow.begin(); vit != bow.end(); vit++) vit->second /= nd;
        }
    } else {                                          // IDF || BINARY
        unsigned i_feature = 0;
        for (int f = 0; f < n; ++f, ++i_feature) {
            unsigned id, nid = 0; double w;
            voc_transform_one(V, desc + (size_t)f * 32, id, w, &nid, levelsup);
            if (w > 0) {
                auto vit = bow.lower_bound(id);       // BowVector::addIfNotExist
                if (vit == bow.end() || (bow.key_comp()(id, vit->first))) bow.insert(vit, std::map<unsigned, double>::value_type(id, w));
                auto fit = fv.lower_bound(nid);
                if (fit != fv.end() && fit->first == nid) fit->second.push_back(i_feature);
                else { fit = fv.insert(fit, std::map<unsigned, std::vector<unsigned>>::value_type(nid, std::vector<unsigned>())); fit->second.push_back(i_feature); }
            }
        }
    }
    if (must) {                                       // BowVector::normalize
        double norm = 0.0;
        if (!l2) { for (auto it = bow.begin(); it != bow.end(); ++it) norm += fabs(it->second); }
        else { for (auto it = bow.begin(); it != bow.end(); ++it) norm += it->second * it->second; norm = sqrt(norm); }
        if (norm > 0.0) for (auto it = bow.begin(); it != bow.end(); ++it) it->second /= norm;
    }
    int nb = 0;
    for (auto& kv : bow) { bow_ids[nb] = kv.first; bow_vals[nb] = kv.second; ++nb; }
    *n_bow = nb;
    int nf = 0, off = 0;
    fv_offsets[0] = 0;
    for (auto& kv : fv) { fv_nodes[nf] = kv.first; for (unsigned i : kv.second) fv_indices[off++] = i; fv_offsets[++nf] = off; }
    *n_fv = nf;
}

// L1Scoring::score — Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:24-65
double orbo_bow_score_l1(const uint32_t* ids_a, const double* vals_a, int na, const uint32_t* ids_b, const double* vals_b, int nb)
{
    std::map<unsigned, double> v1, v2;
    for (int i = 0; i < na; ++i) v1[ids_a[i]] = vals_a[i];
    for (int i = 0; i < nb; ++i) v2[ids_b[i]] = vals_b[i];
    auto v1_it = v1.begin(), v2_it = v2.begin();
    double score = 0;
    while (v1_it != v1.end() && v2_it != v2.end()) {
        const double& vi = v1_it->second;
        const double& wi = v2_it->second;
        if (v1_it->first == v2_it->first) { score += fabs(vi - wi) - fabs(vi) - fabs(wi); ++v1_it; ++v2_it; }
        else if (v1_it->first < v2_it->first) v1_it = v1.lower_bound(v2_it->first);
        else v2_it = v2.lower_bound(v1_it->first);
    }
    score = -score / 2.0;
    return score;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF, Frame& F, vector<MapPoint*>& vpMapPointMatches) — src/ORBmatcher1.cc:225-427.
// valid_kf[i] = pMP && !pMP->isBad(); match_f[j] = index of the KF feature whose MapPoint ends up in vpMapPointMatches[j].
int orbo_search_by_bow_kf_frame(const uint8_t* desc_kf, const float* angle_kf, const uint8_t* valid_kf, int n_kf, int fvk_n,
                                const uint32_t* fvk_nodes, const int32_t* fvk_off, const uint32_t* fvk_idx, const uint8_t* desc_f,
                                const float* angle_f, int n_f, int fvf_n, const uint32_t* fvf_nodes, const int32_t* fvf_off,
                                const uint32_t* fvf_idx, int Nleft, float mfNNratio, int mbCheckOrientation, int32_t* match_f)
{
    (void)n_kf;
    FeatVec vFeatVecKF(fvk_n, fvk_nodes, fvk_off, fvk_idx), FFeatVec(fvf_n, fvf_nodes, fvf_off, fvf_idx);
    std::vector<int> vpMapPointMatches(n_f, -1);
    int nmatches = 0;
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    auto KFit = vFeatVecKF.m.begin(), Fit = FFeatVec.m.begin();
    const auto KFend = vFeatVecKF.m.end(), Fend = FFeatVec.m.end();
    while (KFit != KFend && Fit != Fend) {
        if (KFit->first == Fit->first) {
            const std::vector<unsigned> vIndicesKF = KFit->second, vIndicesF = Fit->second;
            for (size_t iKF = 0; iKF < vIndicesKF.size(); iKF++) {
                const unsigned realIdxKF = vIndicesKF[iKF];
                if (!valid_kf[realIdxKF]) continue;
                const uint8_t* dKF = desc_kf + (size_t)realIdxKF * 32;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                int bestDist1R = 256, bestIdxFR = -1, bestDist2R = 256;
                for (size_t iF = 0; iF < vIndicesF.size(); iF++) {
                    const unsigned realIdxF = vIndicesF[iF];
                    if (vpMapPointMatches[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(dKF, desc_f + (size_t)realIdxF * 32);
                    if (Nleft == -1) {
                        if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                        else if (dist < bestDist2) bestDist2 = dist;
                    } else {
                        if ((int)realIdxF < Nleft && dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                        else if ((int)realIdxF < Nleft && dist < bestDist2) bestDist2 = dist;
                        if ((int)realIdxF >= Nleft && dist < bestDist1R) { bestDist2R = bestDist1R; bestDist1R = dist; bestIdxFR = realIdxF; }
                        else if ((int)realIdxF >= Nleft && dist < bestDist2R) bestDist2R = dist;
                    }
                }
                if (bestDist1 <= TH_LOW) {
                    if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
                        vpMapPointMatches[bestIdxF] = realIdxKF;
                        if (mbCheckOrientation) {
                            float rot = angle_kf[realIdxKF] - angle_f[bestIdxF];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back(bestIdxF);
                        }
                        nmatches++;
                    }
                    if (bestDist1R <= TH_LOW) {
                        if (static_cast<float>(bestDist1R) < mfNNratio * static_cast<float>(bestDist2R) || true) {
                            vpMapPointMatches[bestIdxFR] = realIdxKF;
                            if (mbCheckOrientation) {
                                float rot = angle_kf[realIdxKF] - angle_f[bestIdxFR];
                                if (rot < 0.0) rot += 360.0f;
                                int bin = round(rot * factor);
                                if (bin == HISTO_LENGTH) bin = 0;
                                rotHist[bin].push_back(bestIdxFR);
                            }
                            nmatches++;
                        }
                    }
                }
            }
            KFit++; Fit++;
        } else if (KFit->first < Fit->first) KFit = vFeatVecKF.m.lower_bound(Fit->first);
        else Fit = FFeatVec.m.lower_bound(KFit->first);
    }
    if (mbCheckOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { vpMapPointMatches[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    for (int j = 0; j < n_f; ++j) match_f[j] = vpMapPointMatches[j];
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, vector<MapPoint*>& vpMatches12) — src/ORBmatcher2.cc:36-171.
// valid1[i] / valid2[j] = map point present and not bad (and index < mvKeysUn.size() when NLeft != -1);
// match12[i] = index of the KF2 feature whose MapPoint ends up in vpMatches12[i].
int orbo_search_by_bow_kf_kf(const uint8_t* desc1, const float* angle1, const uint8_t* valid1, int n1, int fv1_n,
                             const uint32_t* fv1_nodes, const int32_t* fv1_off, const uint32_t* fv1_idx, const uint8_t* desc2,
                             const float* angle2, const uint8_t* valid2, int n2, int fv2_n, const uint32_t* fv2_nodes,
                             const int32_t* fv2_off, const uint32_t* fv2_idx, float mfNNratio, int mbCheckOrientation, int32_t* match12)
{
    FeatVec vFeatVec1(fv1_n, fv1_nodes, fv1_off, fv1_idx), vFeatVec2(fv2_n, fv2_nodes, fv2_off, fv2_idx);
    std::vector<int> vpMatches12(n1, -1);
    std::vector<bool> vbMatched2(n2, false);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0;
    auto f1it = vFeatVec1.m.begin(), f2it = vFeatVec2.m.begin();
    const auto f1end = vFeatVec1.m.end(), f2end = vFeatVec2.m.end();
    while (f1it != f1end && f2it != f2end) {
        if (f1it->first == f2it->first) {
            for (size_t i1 = 0, iend1 = f1it->second.size(); i1 < iend1; i1++) {
                const size_t idx1 = f1it->second[i1];
                if (!valid1[idx1]) continue;
                const uint8_t* d1 = desc1 + idx1 * 32;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (size_t i2 = 0, iend2 = f2it->second.size(); i2 < iend2; i2++) {
                    const size_t idx2 = f2it->second[i2];
                    if (vbMatched2[idx2] || !valid2[idx2]) continue;
                    const int dist = descriptor_distance(d1, desc2 + idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = (int)idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if (static_cast<float>(bestDist1) < mfNNratio * static_cast<float>(bestDist2)) {
                        vpMatches12[idx1] = bestIdx2;
                        vbMatched2[bestIdx2] = true;
                        if (mbCheckOrientation) {
                            float rot = angle1[idx1] - angle2[bestIdx2];
                            if (rot < 0.0) rot += 360.0f;
                            int bin = round(rot * factor);
                            if (bin == HISTO_LENGTH) bin = 0;
                            rotHist[bin].push_back((int)idx1);
                        }
                        nmatches++;
                    }
                }
            }
            f1it++; f2it++;
        } else if (f1it->first < f2it->first) f1it = vFeatVec1.m.lower_bound(f2it->first);
        else f2it = vFeatVec2.m.lower_bound(f1it->first);
    }
    if (mbCheckOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { vpMatches12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    for (int i = 0; i < n1; ++i) match12[i] = vpMatches12[i];
    return nmatches;
}

// ORBmatcher::SearchForTriangulation — src/ORBmatcher2.cc:173-471, single pinhole camera (!mpCamera2 on both key frames);
// Pinhole::epipolarConstrain — src/CameraModels/Pinhole.cpp:107-129 with F12 = K1^-T [t12]x R12 K2^-1 supplied by the caller.
int orbo_search_for_triangulation(const KeyPoint* kp1s, const uint8_t* desc1, const uint8_t* free1, const uint8_t* stereo1, int n1,
                                  int fv1_n, const uint32_t* fv1_nodes, const int32_t* fv1_off, const uint32_t* fv1_idx,
                                  const KeyPoint* kp2s, const uint8_t* desc2, const uint8_t* free2, const uint8_t* stereo2, int n2,
                                  int fv2_n, const uint32_t* fv2_nodes, const int32_t* fv2_off, const uint32_t* fv2_idx,
                                  const float* F12, const float* ep, const float* mvScaleFactors2, const float* mvLevelSigma2_2,
                                  int bOnlyStereo, int bCoarse, int mbCheckOrientation, int32_t* match12)
{
    FeatVec vFeatVec1(fv1_n, fv1_nodes, fv1_off, fv1_idx), vFeatVec2(fv2_n, fv2_nodes, fv2_off, fv2_idx);
    int nmatches = 0;
    std::vector<bool> vbMatched2(n2, false);
    std::vector<int> vMatches12(n1, -1);
    std::vector<int> rotHist[HISTO_LENGTH];
    const float factor = 1.0f / HISTO_LENGTH;
    auto Fm = [&](int r, int c) { return F12[r * 3 + c]; };
    auto f1it = vFeatVec1.m.begin(), f2it = vFeatVec2.m.begin();
    const auto f1end = vFeatVec1.m.end(), f2end = vFeatVec2.m.end();
    while (f1it != f1end && f2it != f2end) {
        if (f1it->first == f2it->first) {
            for (size_t i1 = 0, iend1 = f1it->second.size(); i1 < iend1; i1++) {
                const size_t idx1 = f1it->second[i1];
                if (!free1[idx1]) continue;                       // already a MapPoint
                const bool bStereo1 = stereo1[idx1] != 0;
                if (bOnlyStereo) if (!bStereo1) continue;
                const KeyPoint& kp1 = kp1s[idx1];
                const uint8_t* d1 = desc1 + idx1 * 32;
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (size_t i2 = 0, iend2 = f2it->second.size(); i2 < iend2; i2++) {
                    const size_t idx2 = f2it->second[i2];
                    if (vbMatched2[idx2] || !free2[idx2]) continue;
                    const bool bStereo2 = stereo2[idx2] != 0;
                    if (bOnlyStereo) if (!bStereo2) continue;
                    const int dist = descriptor_distance(d1, desc2 + idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const KeyPoint& kp2 = kp2s[idx2];
                    if (!bStereo1 && !bStereo2) {
                        const float distex = ep[0] - kp2.x;
                        const float distey = ep[1] - kp2.y;
                        if (distex * distex + distey * distey < 100 * mvScaleFactors2[kp2.octave]) continue;
                    }
                    bool ok = bCoarse != 0;
                    if (!ok) {
                        const float unc = mvLevelSigma2_2[kp2.octave];
                        const float a = kp1.x * Fm(0, 0) + kp1.y * Fm(1, 0) + Fm(2, 0);
                        const float b = kp1.x * Fm(0, 1) + kp1.y * Fm(1, 1) + Fm(2, 1);
                        const float c = kp1.x * Fm(0, 2) + kp1.y * Fm(1, 2) + Fm(2, 2);
                        const float num = a * kp2.x + b * kp2.y + c;
                        const float den = a * a + b * b;
                        if (den == 0) ok = false;
                        else { const float dsqr = num * num / den; ok = dsqr < 3.84 * unc; }
                    }
                    if (ok) { bestIdx2 = (int)idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    const KeyPoint& kp2 = kp2s[bestIdx2];
                    vMatches12[idx1] = bestIdx2;
                    nmatches++;
                    if (mbCheckOrientation) {
                        float rot = kp1.angle - kp2.angle;
                        if (rot < 0.0) rot += 360.0f;
                        int bin = round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back((int)idx1);
                    }
                }
            }
            f1it++; f2it++;
        } else if (f1it->first < f2it->first) f1it = vFeatVec1.m.lower_bound(f2it->first);
        else f2it = vFeatVec2.m.lower_bound(f1it->first);
    }
    if (mbCheckOrientation) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        compute_three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; i++) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (size_t j = 0, jend = rotHist[i].size(); j < jend; j++) { vMatches12[rotHist[i][j]] = -1; nmatches--; }
        }
    }
    for (int i = 0; i < n1; ++i) match12[i] = vMatches12[i];
    return nmatches;
}

}  // extern "C"

// cv::cvtColor(src, dst, COLOR_{RGB,BGR,RGBA,BGRA}2GRAY) as called by Tracking::GrabImage* (src/Tracking2.cc:289-316, 347-361,
// 392-406).  OpenCV 8U arithmetic (un-vendored dependency; pinned bit-exact against cv2 4.13 in tests/test_oracle_cv2.py):
// gray = (R*9798 + G*19235 + B*3735 + (1 << 14)) >> 15.
extern "C" void orbo_cvt_gray(const uint8_t* src, int rows, int cols, size_t step, int channels, int rgb, uint8_t* dst, size_t dstep)
{
    const int ri = rgb ? 0 : 2, bi = rgb ? 2 : 0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const uint8_t* p = src + (size_t)y * step + (size_t)x * channels;
            dst[(size_t)y * dstep + x] = (uint8_t)((p[ri] * 9798 + p[1] * 19235 + p[bi] * 3735 + (1 << 14)) >> 15);
        }
}
