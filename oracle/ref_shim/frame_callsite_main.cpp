// frame_callsite_main.cpp — TEST INFRASTRUCTURE.  The reference's OWN call site of the extractor, compiled unchanged
// against the drop-in adapter (csrc/adapter/ORBextractor.h) instead of the reference's include/ORBextractor.h:
//   * src/Frame.cc:111-127  the stereo Frame constructor's accessor block (GetLevels .. GetInverseScaleSigmaSquares) and
//                           its two extraction threads                       -> gen/frame_ctor_scale_extract.inc
//   * src/Frame.cc:420-455  extractorParenthesis + Frame::ExtractORB         -> gen/frame_extract_orb.inc
// Both are sliced from /root/reference at build time by oracle/build_ref.sh (never committed).  The class below declares
// only the members those lines touch, with the reference's names and types (include/Frame.h).  The program runs a
// synthetic stereo pair through it on the GPU and prints checksums that tests/test_gpu_ref.py compares with the C ABI.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "ORBextractor.h"

// the fork wraps the call in its benchmark macro (include/orb/Benchmark.h:14); the timing layer is out of scope
#define MEASURE_RET_CALL(func, ...) func(__VA_ARGS__)

using namespace std;

namespace ORB_SLAM3 {

class Frame {
public:
    Frame(const cv::Mat& imLeft, const cv::Mat& imRight, ORBextractor* extractorLeft, ORBextractor* extractorRight);
    void ExtractORB(int flag, const cv::Mat& im, const int x0, const int x1);

    ORBextractor *mpORBextractorLeft, *mpORBextractorRight;
    int mnScaleLevels;
    float mfScaleFactor;
    float mfLogScaleFactor;
    vector<float> mvScaleFactors;
    vector<float> mvInvScaleFactors;
    vector<float> mvLevelSigma2;
    vector<float> mvInvLevelSigma2;
    int N;
    std::vector<cv::KeyPoint> mvKeys, mvKeysRight;
    cv::Mat mDescriptors, mDescriptorsRight;
    int monoLeft, monoRight;
};

Frame::Frame(const cv::Mat& imLeft, const cv::Mat& imRight, ORBextractor* extractorLeft, ORBextractor* extractorRight)
    : mpORBextractorLeft(extractorLeft), mpORBextractorRight(extractorRight)
{
#include "frame_ctor_scale_extract.inc"
    N = mvKeys.size();
}

#include "frame_extract_orb.inc"

}  // namespace ORB_SLAM3

static uint64_t fnv(uint64_t h, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv)
{
    const uint32_t seed = argc > 1 ? (uint32_t)atoi(argv[1]) : 31u;
    const int cols = 752, rows = 480, nfeatures = argc > 2 ? atoi(argv[2]) : 1200;
    std::vector<unsigned char> bl((size_t)cols * rows), br((size_t)cols * rows);
    orbx_synth_image_host(seed, 0, cols, rows, 40, bl.data(), cols);
    orbx_synth_image_host(seed, 1, cols, rows, 40, br.data(), cols);
    cv::Mat imLeft(rows, cols, CV_8UC1, bl.data()), imRight(rows, cols, CV_8UC1, br.data());
    ORB_SLAM3::ORBextractor exL(nfeatures, 1.2f, 8, 20, 7), exR(nfeatures, 1.2f, 8, 20, 7);
    for (int rep = 0; rep < 2; ++rep) {          // twice: the second frame replays the captured graph
        ORB_SLAM3::Frame F(imLeft, imRight, &exL, &exR);
        uint64_t hl = 1469598103934665603ull, hr = hl, ht = hl;
        hl = fnv(hl, F.mvKeys.data(), F.mvKeys.size() * sizeof(cv::KeyPoint));
        for (int i = 0; i < F.mDescriptors.rows; ++i) hl = fnv(hl, F.mDescriptors.ptr(i), 32);
        hr = fnv(hr, F.mvKeysRight.data(), F.mvKeysRight.size() * sizeof(cv::KeyPoint));
        for (int i = 0; i < F.mDescriptorsRight.rows; ++i) hr = fnv(hr, F.mDescriptorsRight.ptr(i), 32);
        ht = fnv(ht, F.mvScaleFactors.data(), 4 * F.mvScaleFactors.size());
        ht = fnv(ht, F.mvInvScaleFactors.data(), 4 * F.mvInvScaleFactors.size());
        ht = fnv(ht, F.mvLevelSigma2.data(), 4 * F.mvLevelSigma2.size());
        ht = fnv(ht, F.mvInvLevelSigma2.data(), 4 * F.mvInvLevelSigma2.size());
        std::printf("rep=%d N=%d Nright=%d monoLeft=%d monoRight=%d levels=%d scale=%.9g logscale=%.9g left=%016llx right=%016llx tables=%016llx\n",
                    rep, F.N, (int)F.mvKeysRight.size(), F.monoLeft, F.monoRight, F.mnScaleLevels, (double)F.mfScaleFactor,
                    (double)F.mfLogScaleFactor, (unsigned long long)hl, (unsigned long long)hr, (unsigned long long)ht);
    }
    return 0;
}
