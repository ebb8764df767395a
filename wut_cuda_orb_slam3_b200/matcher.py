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


def distinctive_descriptors(descriptors, offsets, device=0):
    """MapPoint::ComputeDistinctiveDescriptors for many map points in one GPU call: observations of point p are rows
    offsets[p]:offsets[p+1]; returns the index inside each point's own list (-1 for a point without observations)."""
    d = np.ascontiguousarray(descriptors, np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, np.int32)
    best = np.zeros(len(off) - 1, np.int32)
    check(lib().orbx_distinctive_descriptors(device, ptr(d), ptr(off), len(off) - 1, ptr(best)))
    return best


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


def extract_stereo(exL, exR, imgL, imgR, bf, maxD):
    """The stereo Frame constructor in one call (src/Frame.cc:124-143): returns (kpL, descL, kpR, descR, mvuRight, mvDepth)."""
    from .capi import KP_DTYPE
    imgL = np.ascontiguousarray(imgL, np.uint8); imgR = np.ascontiguousarray(imgR, np.uint8)
    rows, cols = imgL.shape
    assert imgR.shape == imgL.shape
    cap = max(exL.max_keypoints(rows, cols), exR.max_keypoints(rows, cols))
    kL = np.zeros(cap, KP_DTYPE); kR = np.zeros(cap, KP_DTYPE)
    dL = np.zeros((cap, 32), np.uint8); dR = np.zeros((cap, 32), np.uint8)
    u = np.full(cap, -1.0, np.float32); d = np.full(cap, -1.0, np.float32)
    nL = C.c_int(0); nR = C.c_int(0)
    check(lib().orbx_extract_stereo(exL._h, exR._h, ptr(imgL), ptr(imgR), rows, cols, imgL.strides[0], ptr(kL), ptr(dL), C.byref(nL),
                                    ptr(kR), ptr(dR), C.byref(nR), cap, float(bf), float(maxD), ptr(u), ptr(d)))
    return kL[:nL.value], dL[:nL.value], kR[:nR.value], dR[:nR.value], u[:nL.value], d[:nL.value]


class ShardedMatcher:
    """Database-sharded brute-force 2-NN over the GPUs of one node (BASELINE config 5): one process per GPU, NCCL under the
    C ABI (orbx_comm_* / orbx_knn2_sharded).  `broadcast_bytes(buf_or_None) -> bytes` hands rank 0's 128-byte id to every rank
    (torch.distributed, MPI, ...)."""

    def __init__(self, rank, world, device, broadcast_bytes):
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            check(lib().orbx_comm_unique_id(ptr(ident)))
        ident = np.frombuffer(broadcast_bytes(ident.tobytes() if rank == 0 else None), np.uint8).copy()
        self._h = C.c_void_p()
        check(lib().orbx_comm_create(device, rank, world, ptr(ident), C.byref(self._h)))
        self.rank, self.world, self.device = rank, world, device

    def nccl_version(self):
        v = C.c_int(0)
        check(lib().orbx_comm_info(self._h, None, None, C.byref(v)))
        return v.value

    @staticmethod
    def shard_rows(n_rows, world, rank):
        f = C.c_int64(0); c = C.c_int64(0)
        lib().orbx_shard_rows(n_rows, world, rank, C.byref(f), C.byref(c))
        return f.value, c.value

    def knn2(self, d_q, nq, d_db_shard, n_shard_rows, first_row, d_idx, d_dist, stream=None):
        check(lib().orbx_knn2_sharded(self._h, ptr(d_q), nq, ptr(d_db_shard), n_shard_rows, first_row, ptr(d_idx), ptr(d_dist),
                                      ptr(stream) if stream else None))

    def close(self):
        if self._h:
            lib().orbx_comm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def measure_popc_peak(device=0):
    v = C.c_double(0)
    check(lib().orbx_measure_popc_peak(device, C.byref(v)))
    return v.value
