// Compiles the drop-in adapter (csrc/adapter/ORBextractor.h) against the cv stub and runs it exactly the way
// Frame::ExtractORB does (reference src/Frame.cc:420-455): (*extractor)(im, cv::Mat(), keys, descriptors, vLapping).
// Prints a checksum line that tests/test_adapter_cpp.py compares with the C-ABI result obtained through ctypes.
#include <cstdio>
#include <cstdint>
#include <vector>

#include "ORBextractor.h"

int main(int argc, char** argv)
{
    const int cols = 752, rows = 480;
    std::vector<unsigned char> buf((size_t)cols * rows);
    orbx_synth_image_host(argc > 1 ? (uint32_t)atoi(argv[1]) : 1u, 0, cols, rows, 48, buf.data(), cols);
    cv::Mat im(rows, cols, CV_8UC1, buf.data());
    ORB_SLAM3::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> keys;
    cv::Mat desc;
    std::vector<int> lap = {0, 1000};
    const int mono = ex(im, cv::Mat(), keys, desc, lap);
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; } };
    mix(keys.data(), keys.size() * sizeof(cv::KeyPoint));
    for (int i = 0; i < desc.rows; ++i) mix(desc.ptr(i), 32);
    uint64_t hp = 1469598103934665603ull;
    for (int l = 0; l < ex.GetLevels(); ++l)
        for (int y = 0; y < ex.mvImagePyramid[l].rows; ++y) {
            const unsigned char* r = ex.mvImagePyramid[l].ptr(y);
            for (int x = 0; x < ex.mvImagePyramid[l].cols; ++x) { hp ^= r[x]; hp *= 1099511628211ull; }
        }
    const int n = (int)keys.size();
    const int dist = n >= 2 ? ORB_SLAM3::ORBmatcherHamming::DescriptorDistance(cv::Mat(1, 32, CV_8U, desc.ptr(0)), cv::Mat(1, 32, CV_8U, desc.ptr(1))) : -1;
    cv::Mat empty;
    std::vector<cv::KeyPoint> k2;
    cv::Mat d2;
    const int rc_empty = ex(empty, cv::Mat(), k2, d2, lap);
    std::printf("mono=%d n=%d levels=%d scale1=%.9g kphash=%016llx pyrhash=%016llx empty=%d dist01=%d\n", mono, n, ex.GetLevels(),
                (double)ex.GetScaleFactors()[1], (unsigned long long)h, (unsigned long long)hp, rc_empty, dist);
    return 0;
}
