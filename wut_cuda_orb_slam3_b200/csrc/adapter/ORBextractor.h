// ORBextractor.h / ORBmatcher (Hamming part) — drop-in C++ adapter over the C ABI (include/orbx.h).
//
// Same class name, namespace, constructor, operator() signature, accessors and public mvImagePyramid member as the
// reference's include/ORBextractor.h:52-120, so Frame / Tracking compile unchanged (Frame.cc:112-127, 420-455;
// Tracking1.cc:595-601).  Header-only; needs <opencv2/core.hpp> (the real one inside ORB-SLAM3, a 60-line stub in
// tests/cpp/cv_stub for this repo's own compile+run test) and links against liborbx.so.
//
// Error mapping (reference behaviour in brackets): empty image -> returns -1 [src/ORBextractor.cc:1231-1232];
// non-8UC1 image -> assert [1235]; any other library failure -> std::runtime_error [the fork throws on kernel launch
// failure, 498/861/973/1212].  The library itself never aborts or exits.
#ifndef ORBEXTRACTOR_H
#define ORBEXTRACTOR_H

#include <cassert>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include <opencv2/core.hpp>

#include "orbx.h"

namespace ORB_SLAM3 {

class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int cudaDevice = 0)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), iniThFAST(iniThFAST), minThFAST(minThFAST)
    {
        static_assert(sizeof(cv::KeyPoint) == sizeof(orbx_keypoint), "cv::KeyPoint must be the 28-byte POD");
        check(orbx_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, cudaDevice, 0, 0, 1, &handle), "orbx_create");
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
        check(orbx_get_tables(handle, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                              mnFeaturesPerLevel.data()), "orbx_get_tables");
        mvImagePyramid.resize(nlevels);
        check(orbx_set_pyramid_mirror(handle, 1), "orbx_set_pyramid_mirror");
    }
    ~ORBextractor() { orbx_destroy(handle); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image.  Mask is ignored (as in the reference).
    int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
                   cv::OutputArray _descriptors, std::vector<int>& vLappingArea)
    {
        if (_image.empty()) return -1;
        cv::Mat image = _image.getMat();
        assert(image.type() == CV_8UC1);
        if (bDownloadPyramid != mirrorOn) {      // the host copy of the pyramid rides along with the extraction (no blocking copies)
            mirrorOn = bDownloadPyramid;
            check(orbx_set_pyramid_mirror(handle, mirrorOn ? 1 : 0), "orbx_set_pyramid_mirror");
        }
        const int cap = orbx_max_keypoints_for(handle, image.rows, image.cols);
        // results land directly in the caller's vector (resized to the bound, trimmed to n) and in a reusable descriptor slab
        _keypoints.resize(cap);
        if ((int)descScratch.size() < cap * 32) descScratch.resize((size_t)cap * 32);
        int n = 0, nMono = 0;
        const int lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0, lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
        const int rc = orbx_extract(handle, image.data, image.rows, image.cols, (size_t)image.step, lap0, lap1,
                                    reinterpret_cast<orbx_keypoint*>(_keypoints.data()), descScratch.data(), cap, &n, &nMono);
        if (rc != ORBX_OK) { _keypoints.clear(); check(rc, "orbx_extract"); }
        _keypoints.resize(n);
        if (n == 0) {
            _descriptors.release();
        } else {
            _descriptors.create(n, 32, CV_8U);
            cv::Mat d = _descriptors.getMat();
            for (int i = 0; i < n; ++i) std::copy(descScratch.begin() + (size_t)i * 32, descScratch.begin() + (size_t)(i + 1) * 32, d.ptr(i));
        }
        if (mirrorOn) {
            for (int l = 0; l < nlevels; ++l) {
                const uint8_t* base = nullptr; size_t step = 0; int w = 0, h = 0;
                check(orbx_get_pyramid_mirror(handle, l, &base, &step, &w, &h), "orbx_get_pyramid_mirror");
                mvImagePyramid[l] = cv::Mat(h, w, CV_8UC1, const_cast<uint8_t*>(base) + (size_t)ORBX_EDGE_THRESHOLD * step + ORBX_EDGE_THRESHOLD, step);
            }
        }
        return nMono;
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // Interior views into the bordered host mirror the library keeps in pinned memory (filled level by level while the later
    // stages run), valid until the next call (reference: include/ORBextractor.h:92).  Frame::ComputeStereoMatches reads them on
    // the host (src/Frame.cc:848, 938-953); callers that use orbx_stereo_match / orbx_extract_stereo instead can set
    // bDownloadPyramid = false and skip the copy altogether.
    std::vector<cv::Mat> mvImagePyramid;
    bool bDownloadPyramid = true;

    orbx_extractor* nativeHandle() { return handle; }

protected:
    static void check(int rc, const char* what)
    {
        if (rc != ORBX_OK) throw std::runtime_error(std::string(what) + ": " + orbx_last_error());
    }

    int nfeatures;
    double scaleFactor;
    int nlevels;
    int iniThFAST;
    int minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<unsigned char> descScratch;
    bool mirrorOn = true;
    orbx_extractor* handle = nullptr;
};

// The Hamming part of ORBmatcher (reference include/ORBmatcher.h:40-43, src/ORBmatcher3.cc:637-653).  In ORB-SLAM3 this
// static member is the only thing to replace: `int ORBmatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b)
// { return orbx_descriptor_distance(a.ptr<uchar>(), b.ptr<uchar>()); }`.
struct ORBmatcherHamming {
    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 30;   // src/ORBmatcher1.cc:37-39
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orbx_descriptor_distance(a.ptr(0), b.ptr(0)); }
};

}  // namespace ORB_SLAM3

#endif  // ORBEXTRACTOR_H
