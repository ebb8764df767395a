"""Python mirror of the Hamming-matching half of ORB_SLAM3::ORBmatcher / Frame (reference include/ORBmatcher.h:40-43,
src/ORBmatcher3.cc:637-653, src/Frame.cc:841-1011, 1156-1196) over the C ABI."""
import ctypes as C

import numpy as np

from .capi import check, lib, ptr


class ORBmatcher:
    TH_LOW = 50      # src/ORBmatcher1.cc:37-39
    TH_HIGH = 100
    HISTO_LENGTH = 30

    def __init__(self, nnratio=0.6, checkOri=True, device=0):
        self.mfNNratio = float(nnratio)
        self.mbCheckOrientation = bool(checkOri)
        self.device = device

    @staticmethod
    def DescriptorDistance(a, b):
        """static int ORBmatcher::DescriptorDistance(const cv::Mat&, const cv::Mat&) — 256-bit Hamming distance."""
        a = np.ascontiguousarray(a, np.uint8).reshape(-1); b = np.ascontiguousarray(b, np.uint8).reshape(-1)
        assert a.size == 32 and b.size == 32
        return lib().orbx_descriptor_distance(ptr(a), ptr(b))

    def knn2(self, queries, database):
        """Brute-force 2-NN (BFMatcher.knnMatch k=2 / the best-second scan of SearchByBoW) on the GPU; host arrays."""
        q = np.ascontiguousarray(queries, np.uint8).reshape(-1, 32); db = np.ascontiguousarray(database, np.uint8).reshape(-1, 32)
        idx = np.zeros((len(q), 2), np.int32); dist = np.zeros((len(q), 2), np.int32)
        check(lib().orbx_knn2(self.device, ptr(q), len(q), ptr(db), len(db), ptr(idx), ptr(dist)))
        return idx, dist

    def ratio_test(self, dist, mode=0, th_low=None, ratio=None):
        dist = np.ascontiguousarray(dist, np.int32)
        acc = np.zeros(len(dist), np.uint8)
        check(lib().orbx_ratio_test(ptr(dist), len(dist), self.mfNNratio if ratio is None else float(ratio),
                                    self.TH_LOW if th_low is None else int(th_low), int(mode), ptr(acc)))
        return acc.astype(bool)


def rotation_consistency(angle_a, angle_b):
    """The 30-bin rotation histogram + ComputeThreeMaxima filter of SearchByBoW (src/ORBmatcher1.cc:344-427): bool keep[n]."""
    a = np.ascontiguousarray(angle_a, np.float32); b = np.ascontiguousarray(angle_b, np.float32)
    keep = np.zeros(len(a), np.uint8)
    check(lib().orbx_rotation_consistency(ptr(a), ptr(b), len(a), ptr(keep)))
    return keep.astype(bool)


def distinctive_descriptor(descriptors):
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:329-401): index of the least-median-distance descriptor."""
    d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
    best = C.c_int(0)
    check(lib().orbx_distinctive_descriptor(ptr(d), len(d), C.byref(best)))
    return best.value


def knn2_device(d_q, nq, d_db, ndb, d_idx, d_dist, index_base=0, device=0, stream=None):
    check(lib().orbx_knn2_device(device, ptr(d_q), nq, ptr(d_db), ndb, index_base, ptr(d_idx), ptr(d_dist), ptr(stream) if stream else None))


def knn2_merge_device(d_idx_shards, d_dist_shards, n_shards, nq, d_idx, d_dist, device=0, stream=None):
    check(lib().orbx_knn2_merge_device(device, ptr(d_idx_shards), ptr(d_dist_shards), n_shards, nq, ptr(d_idx), ptr(d_dist),
                                       ptr(stream) if stream else None))


def compute_stereo_matches(exL, exR, kpL, descL, kpR, descR, bf, maxD, frameL=0, frameR=0):
    """Frame::ComputeStereoMatches (src/Frame.cc:841-1011): returns (mvuRight, mvDepth)."""
    kpL = np.ascontiguousarray(kpL); kpR = np.ascontiguousarray(kpR)
    descL = np.ascontiguousarray(descL, np.uint8); descR = np.ascontiguousarray(descR, np.uint8)
    u = np.full(len(kpL), -1.0, np.float32); d = np.full(len(kpL), -1.0, np.float32)
    check(lib().orbx_stereo_match(exL._h, frameL, exR._h, frameR, ptr(kpL), ptr(descL), len(kpL), ptr(kpR), ptr(descR), len(kpR),
                                  float(bf), float(maxD), ptr(u), ptr(d)))
    return u, d


def measure_popc_peak(device=0):
    v = C.c_double(0)
    check(lib().orbx_measure_popc_peak(device, C.byref(v)))
    return v.value
