// Minimal stand-in for <opencv2/core.hpp> (OpenCV C++ headers are not installed in this image): just enough of
// cv::Mat / cv::KeyPoint / InputArray / OutputArray for the adapter header to compile and run in tests/cpp.
#pragma once
#include <cstddef>
#include <cstdlib>
#include <memory>
#include <vector>
#define CV_8U 0
#define CV_8UC1 0
namespace cv {
struct Point2f { float x, y; };
struct KeyPoint { Point2f pt; float size, angle, response; int octave, class_id; };
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int, void* d, size_t s = 0) : rows(r), cols(c), step(s ? s : (size_t)c), data((unsigned char*)d) {}
    void create(int r, int c, int) { owner.reset(new std::vector<unsigned char>((size_t)r * c)); rows = r; cols = c; step = (size_t)c; data = owner->data(); }
    void release() { owner.reset(); rows = cols = 0; step = 0; data = nullptr; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    unsigned char* ptr(int r = 0) { return data + (size_t)r * step; }
    const unsigned char* ptr(int r = 0) const { return data + (size_t)r * step; }
    Mat& getMat() { return *this; }
    const Mat& getMat() const { return *this; }
private:
    std::shared_ptr<std::vector<unsigned char>> owner;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv
