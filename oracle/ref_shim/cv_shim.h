// cv_shim.h — the few OpenCV TYPES the sliced reference functions touch, so that they compile in this image (no OpenCV C++
// headers here).  TEST INFRASTRUCTURE ONLY (oracle/_ref).  Containers and scalar helpers only: every OpenCV ALGORITHM a slice
// calls (cv::FAST) is forwarded to the oracle's restatement, which tests/test_oracle_cv2.py pins bit-exact against cv2.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <memory>
#include <vector>

typedef unsigned char uchar;
#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0

inline int cvRound(double v) { return (int)lrint(v); }          // SSE2 cvtsd2si: round half to even
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

extern "C" int orbo_fast9(const uint8_t* img, int w, int h, size_t step, int threshold, int nms, int* xs, int* ys, int* scores, int cap);

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    template <typename A, typename B> Point_(A x_, B y_) : x((T)x_), y((T)y_) {}       // Point2i(float, int): C++ conversion, as cv::Point_
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
template <> inline Point_<int>& Point_<int>::operator*=(float s) { x = cvRound(x * s); y = cvRound(y * s); return *this; }
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;

struct KeyPoint {                       // same field order and size as cv::KeyPoint (28 bytes)
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
        : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// 8-bit single-channel matrix header over shared storage (enough for at<uchar>, ptr, row, rowRange/colRange, copyTo).
struct Mat {
    std::shared_ptr<std::vector<uchar>> store;
    uchar* data = nullptr;
    int rows = 0, cols = 0;
    size_t step = 0;
    Mat() {}
    Mat(int r, int c, int /*type*/) : store(std::make_shared<std::vector<uchar>>((size_t)r * c)), data(store->data()), rows(r), cols(c), step((size_t)c) {}
    Mat(int r, int c, int /*type*/, void* d, size_t s = 0) : data((uchar*)d), rows(r), cols(c), step(s ? s : (size_t)c) {}
    void create(int r, int c, int /*type*/) { store = std::make_shared<std::vector<uchar>>((size_t)r * c); data = store->data(); rows = r; cols = c; step = (size_t)c; }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + x * sizeof(T)); }
    template <typename T = uchar> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
    template <typename T = uchar> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
    Mat row(int y) const { Mat m = *this; m.data = data + (size_t)y * step; m.rows = 1; return m; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.data = data + (size_t)a * step; m.rows = b - a; return m; }
    Mat colRange(int a, int b) const { Mat m = *this; m.data = data + a; m.cols = b - a; return m; }
    void copyTo(Mat dst) const { for (int y = 0; y < rows; ++y) memcpy(dst.data + (size_t)y * dst.step, data + (size_t)y * step, (size_t)cols); }
};
enum { NORM_L1 = 2 };
// cv::norm(a, b, NORM_L1) for 8-bit matrices: sum of absolute differences (an integer, returned as double like OpenCV)
inline double norm(const Mat& a, const Mat& b, int /*NORM_L1*/)
{
    long long s = 0;
    for (int y = 0; y < a.rows; ++y)
        for (int x = 0; x < a.cols; ++x) s += std::abs((int)a.data[(size_t)y * a.step + x] - (int)b.data[(size_t)y * b.step + x]);
    return (double)s;
}
typedef const Mat& InputArray;
typedef Mat& OutputArray;

// cv::FAST(image, keypoints, threshold, nonmaxSuppression) -> the oracle's cv2-pinned restatement; output exactly as OpenCV
// fills it: KeyPoint(x, y, 7.f, -1, score), rows top to bottom, columns left to right.
inline void FAST(InputArray img, std::vector<KeyPoint>& keypoints, int threshold, bool nms)
{
    keypoints.clear();
    if (img.rows < 7 || img.cols < 7) return;
    const int cap = img.rows * img.cols;
    std::vector<int> xs(cap), ys(cap), sc(cap);
    const int n = orbo_fast9(img.data, img.cols, img.rows, img.step, threshold, nms ? 1 : 0, xs.data(), ys.data(), sc.data(), cap);
    for (int i = 0; i < n; ++i) keypoints.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)sc[i]));
}

}  // namespace cv
