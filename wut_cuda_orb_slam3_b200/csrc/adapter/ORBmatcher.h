// ORBmatcher.h — drop-in C++ adapter: `class ORB_SLAM3::ORBmatcher` with the reference's declarations
// (include/ORBmatcher.h:36-60) for the searches this library implements, over the C ABI (include/orbx.h).
//
// Include it where the reference's include/ORBmatcher.h was included, AFTER the headers that define the complete types
// ORB_SLAM3::Frame, KeyFrame and MapPoint (include/Frame.h, KeyFrame.h, MapPoint.h in ORB-SLAM3; the stand-ins of
// tests/cpp/matcher_adapter_test.cpp here): the bodies read exactly the members the reference functions read.
// Single-camera rigs (Nleft == -1, no mpCamera2): a two-fisheye rig keeps the reference code for these calls.
//
//   static int DescriptorDistance(a, b)                                              src/ORBmatcher3.cc:637-653
//   int SearchByProjection(Frame&, const vector<MapPoint*>&, th, bFarPoints, thFar)  src/ORBmatcher1.cc:45-215
//   int SearchByProjection(Frame& Current, const Frame& Last, th, bMono)             src/ORBmatcher3.cc:256-467
//   int SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>& vpMapPointMatches)         src/ORBmatcher1.cc:225-427
//   int SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>& vpMatches12)            src/ORBmatcher2.cc:36-171
// Not declared (not on the tracking path / out of scope): the Sim3 SearchByProjection overloads,
// SearchForInitialization, SearchBySim3, Fuse; SearchForTriangulation is exported by the C ABI (orbx_search_for_triangulation)
// and takes F12 from the caller.
#ifndef ORBMATCHER_H
#define ORBMATCHER_H

#include <cstring>
#include <vector>

#include "ORBmatcherProjection.h"

namespace ORB_SLAM3 {

class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6, bool checkOri = true, int cudaDevice = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri), device(cudaDevice) {}

    // Computes the Hamming distance between two ORB descriptors
    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return orbx_descriptor_distance(a.ptr(0), b.ptr(0)); }

    // Search matches between Frame keypoints and projected MapPoints. Returns number of matches
    // Used to track the local map (Tracking)
    int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.0f)
    {
        return orbx_adapter::SearchByProjection(F, vpMapPoints, th, bFarPoints, thFarPoints, mfNNratio, device);
    }

    // Project MapPoints tracked in last frame into the current frame and search matches.
    // Used to track from previous frame (Tracking)
    int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
    {
        // the Sophus / Eigen lines of the reference stay as they are (src/ORBmatcher3.cc:266-296): Eigen's own arithmetic
        const auto Tcw = CurrentFrame.GetPose();
        const auto twc = Tcw.inverse().translation();
        const auto Tlw = LastFrame.GetPose();
        const auto tlc = Tlw * twc;
        const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
        const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
        const int M = LastFrame.N;
        std::vector<float> u(M, 0.f), v(M, 0.f), invz(M, -1.f);            // invz < 0: the reference skips the point (:290-291)
        for (int i = 0; i < M; i++) {
            MapPoint* pMP = LastFrame.mvpMapPoints[i];
            if (pMP && !LastFrame.mvbOutlier[i]) {
                const auto x3Dw = pMP->GetWorldPos();
                const auto x3Dc = Tcw * x3Dw;
                invz[i] = 1.0 / x3Dc(2);
                if (invz[i] < 0) continue;
                const auto uv = CurrentFrame.mpCamera->project(x3Dc);
                u[i] = uv(0); v[i] = uv(1);
            }
        }
        return orbx_adapter::SearchByProjectionLast(CurrentFrame, LastFrame, u, v, invz, th, bForward, bBackward, mbCheckOrientation, device);
    }

    // Search matches between MapPoints in a KeyFrame and ORB in a Frame.
    // Brute force constrained to ORB that belong to the same vocabulary node (at a certain level)
    // Used in Relocalisation and Loop Detection
    int SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches)
    {
        const std::vector<MapPoint*> vpMapPointsKF = pKF->GetMapPointMatches();
        vpMapPointMatches = std::vector<MapPoint*>(F.N, static_cast<MapPoint*>(NULL));
        const int nA = (int)vpMapPointsKF.size();
        std::vector<uint8_t> validA(nA);
        std::vector<float> angA(nA), angB(F.N);
        for (int i = 0; i < nA; ++i) { validA[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad(); angA[i] = pKF->mvKeysUn[i].angle; }   // :252-258, :335
        for (int j = 0; j < F.N; ++j) angB[j] = F.mvKeys[j].angle;                                                                 // :343
        Csr fa(pKF->mFeatVec), fb(F.mFeatVec);
        const orbx_feature_vector va = fa.view(), vb = fb.view();
        std::vector<int32_t> matchB(F.N, -1);
        int n = 0;
        orbx_adapter::check(orbx_search_by_bow(device, 0, pKF->mDescriptors.ptr(0), angA.data(), validA.data(), nA, &va, F.mDescriptors.ptr(0),
                                               angB.data(), nullptr, F.N, &vb, -1, mfNNratio, mbCheckOrientation, nullptr, matchB.data(), &n),
                            "SearchByBoW");
        for (int j = 0; j < F.N; ++j)
            if (matchB[j] >= 0) vpMapPointMatches[j] = vpMapPointsKF[matchB[j]];
        return n;
    }

    // Matching for the loop detection / place recognition between two key frames
    int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12)
    {
        const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches();
        const std::vector<MapPoint*> vpMapPoints2 = pKF2->GetMapPointMatches();
        const int nA = (int)vpMapPoints1.size(), nB = (int)vpMapPoints2.size();
        vpMatches12 = std::vector<MapPoint*>(nA, static_cast<MapPoint*>(NULL));
        std::vector<uint8_t> validA(nA), validB(nB);
        std::vector<float> angA(nA), angB(nB);
        for (int i = 0; i < nA; ++i) { validA[i] = vpMapPoints1[i] && !vpMapPoints1[i]->isBad(); angA[i] = pKF1->mvKeysUn[i].angle; }   // :73-82, :129
        for (int j = 0; j < nB; ++j) { validB[j] = vpMapPoints2[j] && !vpMapPoints2[j]->isBad(); angB[j] = pKF2->mvKeysUn[j].angle; }   // :95-103
        Csr fa(pKF1->mFeatVec), fb(pKF2->mFeatVec);
        const orbx_feature_vector va = fa.view(), vb = fb.view();
        std::vector<int32_t> matchA(nA, -1);
        int n = 0;
        orbx_adapter::check(orbx_search_by_bow(device, 1, pKF1->mDescriptors.ptr(0), angA.data(), validA.data(), nA, &va, pKF2->mDescriptors.ptr(0),
                                               angB.data(), validB.data(), nB, &vb, -1, mfNNratio, mbCheckOrientation, matchA.data(), nullptr, &n),
                            "SearchByBoW");
        for (int i = 0; i < nA; ++i)
            if (matchA[i] >= 0) vpMatches12[i] = vpMapPoints2[matchA[i]];
        return n;
    }

public:
    static const int TH_LOW = 50;          // src/ORBmatcher1.cc:37-39
    static const int TH_HIGH = 100;
    static const int HISTO_LENGTH = 30;

protected:
    // DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned int>>, Thirdparty/DBoW2/DBoW2/FeatureVector.h:23-25) -> CSR
    struct Csr {
        std::vector<uint32_t> nodes, indices;
        std::vector<int32_t> offsets;
        template <class FeatVecT> explicit Csr(const FeatVecT& fv)
        {
            offsets.push_back(0);
            for (auto it = fv.begin(); it != fv.end(); ++it) {
                nodes.push_back((uint32_t)it->first);
                for (unsigned int idx : it->second) indices.push_back((uint32_t)idx);
                offsets.push_back((int32_t)indices.size());
            }
        }
        orbx_feature_vector view() const { return orbx_feature_vector{(int)nodes.size(), nodes.data(), offsets.data(), indices.data()}; }
    };

    float mfNNratio;
    bool mbCheckOrientation;
    int device;
};

}  // namespace ORB_SLAM3

#endif  // ORBMATCHER_H
