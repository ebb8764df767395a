// api_shard.cu — the database-sharded brute-force 2-NN behind the C ABI (BASELINE config 5).
//
// Semantics: cv::BFMatcher(NORM_HAMMING).knnMatch(queries, database, k = 2) of the reference (src/Frame.cc:45, 1174) on the
// UNION of all ranks' database shards, ties -> lower global row index, so that the ratio test of src/Frame.cc:1181 sees the
// same (d1, d2) whatever the number of GPUs.  Queries are replicated, database rows are split in contiguous shards, every rank
// scans its shard (knn2_kernel), ONE ncclAllGather moves the packed per-query candidates (16 B: i1, i2, d1, d2) and
// knn2_merge_kernel folds the shards in lexicographic (distance, index) order.  Every rank ends with the full answer.
//
// NCCL is bound at run time (dlopen): a process that already carries a libnccl (PyTorch's, the caller's) is joined on that
// copy, so an ncclComm_t created by the caller can be wrapped with orbx_comm_from_nccl; otherwise libnccl.so.2 is loaded from
// the loader path (or $ORBX_NCCL_LIB).  liborbx.so has no link-time NCCL dependency and single-GPU users never load it.
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "orbx_internal.cuh"

namespace {

// the slice of nccl.h this file needs (ABI-stable since NCCL 2.0)
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;
typedef int NcclResult;                       // ncclSuccess == 0
constexpr int kNcclInt32 = 2;                 // ncclInt32 / ncclInt

struct NcclApi {
    NcclResult (*GetUniqueId)(NcclUniqueId*) = nullptr;
    NcclResult (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    NcclResult (*CommDestroy)(NcclComm) = nullptr;
    NcclResult (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(NcclResult) = nullptr;
    NcclResult (*GetVersion)(int*) = nullptr;
    bool ok = false;
    char where[256] = {0};
};

const NcclApi& nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        void* h = nullptr;
        const char* env = getenv("ORBX_NCCL_LIB");
        if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the process already uses, if any
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!h) { snprintf(a.where, sizeof a.where, "%s", dlerror() ? dlerror() : "libnccl.so.2 not found"); return a; }
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
        a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
        a.GetVersion = (decltype(a.GetVersion))dlsym(h, "ncclGetVersion");
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.GetErrorString;
        if (!a.ok) snprintf(a.where, sizeof a.where, "libnccl lacks a required symbol");
        return a;
    }();
    return api;
}

}  // namespace

using namespace orbx;

struct orbx_comm {
    NcclComm comm = nullptr;
    bool owned = false;
    int device = 0, rank = 0, world = 1;
    std::mutex mu;                 // one collective at a time per communicator (NCCL's own rule)
    int32_t* d_buf = nullptr;      // [4*nq] send | [world][4*nq] receive | knn2 chunk partials
    size_t buf_bytes = 0;
    cudaEvent_t done = nullptr;    // the last call's work on its stream: the buffer is reused only after it
};

#define NCCLCHK(call)                                                                                             \
    do {                                                                                                          \
        NcclResult r__ = (call);                                                                                  \
        if (r__ != 0) return fail(ORBX_ERR_CUDA, "%s failed: %s", #call, nccl().GetErrorString(r__));            \
    } while (0)
#define CUCHK(call)                                                                                               \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess)                                                                                   \
            return fail(e__ == cudaErrorMemoryAllocation ? ORBX_ERR_OOM : ORBX_ERR_CUDA, "%s failed: %s", #call, \
                        cudaGetErrorString(e__));                                                                 \
    } while (0)

extern "C" {

int orbx_comm_unique_id(uint8_t id[ORBX_COMM_ID_BYTES])
{
    if (!id) return fail(ORBX_ERR_INVALID_ARG, "id is NULL");
    if (!nccl().ok) return fail(ORBX_ERR_UNSUPPORTED, "NCCL unavailable: %s", nccl().where);
    NcclUniqueId u;
    NCCLCHK(nccl().GetUniqueId(&u));
    static_assert(sizeof u == ORBX_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, sizeof u);
    return ORBX_OK;
}

int orbx_comm_create(int device, int rank, int world, const uint8_t id[ORBX_COMM_ID_BYTES], orbx_comm** out)
{
    if (!out) return fail(ORBX_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world || !id) return fail(ORBX_ERR_INVALID_ARG, "bad rank %d / world %d", rank, world);
    int rc = set_device(device);
    if (rc) return rc;
    if (!nccl().ok) return fail(ORBX_ERR_UNSUPPORTED, "NCCL unavailable: %s", nccl().where);
    NcclUniqueId u;
    memcpy(&u, id, sizeof u);
    NcclComm c = nullptr;
    NCCLCHK(nccl().CommInitRank(&c, world, u, rank));
    orbx_comm* h = new orbx_comm();
    h->comm = c; h->owned = true; h->device = device; h->rank = rank; h->world = world;
    *out = h;
    return ORBX_OK;
}

int orbx_comm_from_nccl(int device, void* nccl_comm, int rank, int world, orbx_comm** out)
{
    if (!out) return fail(ORBX_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    if (!nccl_comm || world < 1 || rank < 0 || rank >= world) return fail(ORBX_ERR_INVALID_ARG, "bad communicator / rank / world");
    int rc = set_device(device);
    if (rc) return rc;
    if (!nccl().ok) return fail(ORBX_ERR_UNSUPPORTED, "NCCL unavailable: %s", nccl().where);
    orbx_comm* h = new orbx_comm();
    h->comm = (NcclComm)nccl_comm; h->owned = false; h->device = device; h->rank = rank; h->world = world;
    *out = h;
    return ORBX_OK;
}

void orbx_comm_destroy(orbx_comm* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->done) { cudaEventSynchronize(c->done); cudaEventDestroy(c->done); }
    if (c->d_buf) cudaFree(c->d_buf);
    if (c->owned && c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

int orbx_comm_info(const orbx_comm* c, int* rank, int* world, int* nccl_version)
{
    if (!c) return fail(ORBX_ERR_INVALID_ARG, "communicator is NULL");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) { *nccl_version = 0; if (nccl().GetVersion) nccl().GetVersion(nccl_version); }
    return ORBX_OK;
}

void orbx_shard_rows(int64_t n_rows, int world, int rank, int64_t* first, int64_t* count)
{
    const int64_t per = world > 0 ? (n_rows + world - 1) / world : n_rows;
    const int64_t f = std::min<int64_t>((int64_t)rank * per, n_rows);
    if (first) *first = f;
    if (count) *count = std::max<int64_t>(0, std::min<int64_t>(per, n_rows - f));
}

int orbx_knn2_sharded(orbx_comm* c, const uint8_t* d_queries, int nq, const uint8_t* d_db_shard, int64_t n_shard_rows,
                      int64_t first_row, int32_t* d_idx, int32_t* d_dist, void* stream)
{
    if (!c) return fail(ORBX_ERR_INVALID_ARG, "communicator is NULL");
    if (nq < 0 || n_shard_rows < 0 || first_row < 0 || (nq > 0 && (!d_queries || !d_idx || !d_dist)) || (n_shard_rows > 0 && !d_db_shard))
        return fail(ORBX_ERR_INVALID_ARG, "bad arguments");
    if (first_row + n_shard_rows > 0x7fffffffLL) return fail(ORBX_ERR_UNSUPPORTED, "global row indices must fit int32");
    if (nq == 0) return ORBX_OK;
    int rc = set_device(c->device);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t per_rank = (size_t)nq * 4;                                    // int32: idx[2nq] | dist[2nq]
    const size_t ws = knn2_workspace_bytes(nq, n_shard_rows);
    const size_t need = (per_rank * (1 + (size_t)c->world)) * sizeof(int32_t) + ws + 512;
    if (!c->done) CUCHK(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    if (c->buf_bytes < need) {
        CUCHK(cudaEventSynchronize(c->done));
        if (c->d_buf) cudaFree(c->d_buf);
        c->d_buf = nullptr; c->buf_bytes = 0;
        CUCHK(cudaMalloc(&c->d_buf, need));
        c->buf_bytes = need;
    } else {
        CUCHK(cudaStreamWaitEvent(st, c->done, 0));                            // a previous call on another stream still owns the buffer
    }
    int32_t* send = c->d_buf;
    int32_t* recv = send + per_rank;
    void* part = reinterpret_cast<uint8_t*>(recv + per_rank * c->world) + 256 - ((uintptr_t)(recv + per_rank * c->world) & 255);
    CUCHK(launch_knn2(d_queries, nq, d_db_shard, n_shard_rows, (int)first_row, send, send + 2 * (size_t)nq, st, ws ? part : nullptr));
    if (c->world == 1) {
        CUCHK(cudaMemcpyAsync(d_idx, send, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToDevice, st));
        CUCHK(cudaMemcpyAsync(d_dist, send + 2 * (size_t)nq, sizeof(int32_t) * 2 * nq, cudaMemcpyDeviceToDevice, st));
    } else {
        NCCLCHK(nccl().AllGather(send, recv, per_rank, kNcclInt32, c->comm, st));
        CUCHK(launch_knn2_merge(recv, recv + 2 * (size_t)nq, c->world, nq, d_idx, d_dist, st, per_rank));
    }
    CUCHK(cudaEventRecord(c->done, st));
    return ORBX_OK;
}

}  // extern "C"
