// api_proj.cu — host side of the frame grid and the projection-guided searches (C ABI of include/orbx.h).
//
// Reference interfaces replaced (paths relative to the reference root):
//   Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea          src/Frame.cc:387-418, 755-766, 687-753
//   ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, ...)      src/ORBmatcher1.cc:45-215   (Tracking::SearchLocalPoints)
//   ORBmatcher::SearchByProjection(Frame& Current, const Frame& Last..)  src/ORBmatcher3.cc:256-467  (Tracking::TrackWithMotionModel)
//   ORBmatcher::SearchByProjection(Frame& Current, KeyFrame*, ...)       src/ORBmatcher3.cc:469-578  (Tracking::Relocalization)
// Each entry point turns the reference's per-map-point pre-checks into window queries (float arithmetic in the reference's
// order, on the host: a handful of operations per point), ships everything in ONE pinned staging copy, runs the grid build and
// the window search (kernels_proj.cu) and reads the per-query assignment back in one copy.  The 30-bin rotation histogram runs
// on the host afterwards (it depends on the accepted matches only).  These calls sit on the tracking thread's per-frame path,
// so the device scratch, the pinned staging buffer and the stream live in a per-device grow-only arena (no cudaMalloc per call).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int GRID_CELLS = ORBX_FRAME_GRID_COLS * ORBX_FRAME_GRID_ROWS;
thread_local int t_rounds = 0;

struct ProjArena {
    std::mutex mu;
    unsigned char* d = nullptr; size_t d_bytes = 0;
    unsigned char* h = nullptr; size_t h_bytes = 0;
    cudaStream_t st = nullptr;
};
ProjArena g_arena[64];

// One call = one lease: a bump allocator that hands out matching (pinned host, device) ranges so that all inputs travel in
// one cudaMemcpyAsync.  Inputs are allocated first (the upload range), device-only scratch and outputs after.
struct Lease {
    ProjArena& A;
    std::unique_lock<std::mutex> lock;
    size_t off = 0, upload_end = 0;
    explicit Lease(int device) : A(g_arena[device & 63]), lock(A.mu) {}
    int reserve(size_t bytes)
    {
        bytes += 8192;
        if (!A.st) {
            cudaError_t e = cudaStreamCreateWithFlags(&A.st, cudaStreamNonBlocking);
            if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        if (A.d_bytes < bytes) {
            if (A.d) cudaFree(A.d);
            if (A.h) cudaFreeHost(A.h);
            A.d = A.h = nullptr; A.d_bytes = A.h_bytes = 0;
            const size_t want = bytes + bytes / 2;
            cudaError_t e = cudaMalloc((void**)&A.d, want);
            if (e == cudaSuccess) e = cudaMallocHost((void**)&A.h, want);
            if (e != cudaSuccess) return fail(ORBX_ERR_OOM, "projection scratch (%zu bytes): %s", want, cudaGetErrorString(e));
            A.d_bytes = A.h_bytes = want;
        }
        return ORBX_OK;
    }
    template <typename T> size_t take(size_t count)
    {
        off = (off + 255) & ~(size_t)255;
        const size_t o = off;
        off += std::max<size_t>(count, 1) * sizeof(T);
        return o;
    }
    template <typename T> T* dev(size_t o) const { return reinterpret_cast<T*>(A.d + o); }
    template <typename T> T* host(size_t o) const { return reinterpret_cast<T*>(A.h + o); }
};

size_t pad(size_t bytes) { return ((bytes + 255) & ~(size_t)255) + 256; }

int check_frame(const orbx_frame_view* f, bool need_desc)
{
    if (!f || f->n < 0) return fail(ORBX_ERR_INVALID_ARG, "bad frame view");
    if (f->n > 0 && (!f->keys_un || (need_desc && !f->descriptors))) return fail(ORBX_ERR_INVALID_ARG, "frame view: keys_un / descriptors missing");
    if (f->n >= (1 << 24)) return fail(ORBX_ERR_UNSUPPORTED, "frame view: too many features");
    if (!(f->grid_w_inv > 0.0f) || !(f->grid_h_inv > 0.0f)) return fail(ORBX_ERR_INVALID_ARG, "frame view: grid element sizes must be positive");
    if (need_desc && (!f->scale_factors || f->n_levels < 1 || f->n_levels > kMaxLevels)) return fail(ORBX_ERR_INVALID_ARG, "frame view: scale factors missing");
    return ORBX_OK;
}

// A frame + a set of window queries on the device.
struct Session {
    Lease L;
    ProjArgs A{};
    size_t o_kp = 0, o_desc = 0, o_ur = 0, o_taken = 0, o_q = 0, o_qd = 0, o_cs = 0, o_items = 0, o_cellof = 0, o_claim = 0, o_keys = 0,
           o_qm = 0, o_la = 0, o_lb = 0, o_rounds = 0, o_area = 0, o_area_n = 0;
    int n = 0, nq = 0, area_cap = 0;
    explicit Session(int device) : L(device) {}

    // Lay out and reserve; afterwards queries()/qdesc() point into the pinned staging buffer for the caller to fill.
    int begin(int device, const orbx_frame_view* f, int n_queries, bool with_desc, int area_capacity = 0)
    {
        int rc;
        if ((rc = set_device(device))) return rc;
        n = f->n; nq = n_queries; area_cap = area_capacity;
        const size_t total = pad((size_t)n * 28) + pad((size_t)n * 32) + pad((size_t)n * 4) + pad((size_t)n * 4) + pad((size_t)nq * sizeof(ProjQuery)) +
                             pad((size_t)nq * 32) + pad((GRID_CELLS + 1) * 4) + pad((size_t)n * 4) + pad((size_t)n * 2) + pad((size_t)n * 4) +
                             pad((size_t)nq * 16) + 3 * pad((size_t)nq * 4) + pad(4) + pad((size_t)area_cap * 8) + pad(4);
        if ((rc = L.reserve(total))) return rc;
        // upload range
        o_kp = L.take<orbx_keypoint>(n);
        o_desc = L.take<uint8_t>((size_t)n * 32);
        o_ur = L.take<float>(n);
        o_taken = L.take<int>(n);
        o_q = L.take<ProjQuery>(nq);
        o_qd = L.take<uint8_t>((size_t)nq * 32);
        o_qm = L.take<int>(nq);
        o_area_n = L.take<int>(1);
        L.upload_end = L.off;
        // device-only scratch / outputs
        o_cs = L.take<int>(GRID_CELLS + 1);
        o_items = L.take<int>(n);
        o_cellof = L.take<unsigned short>(n);
        o_claim = L.take<int>(n);
        o_keys = L.take<unsigned long long>((size_t)nq * 2);
        o_la = L.take<int>(nq);
        o_lb = L.take<int>(nq);
        o_rounds = L.take<int>(1);
        o_area = L.take<unsigned long long>(area_cap);
        if (n) memcpy(L.host<orbx_keypoint>(o_kp), f->keys_un, (size_t)n * sizeof(orbx_keypoint));
        if (n && with_desc) memcpy(L.host<uint8_t>(o_desc), f->descriptors, (size_t)n * 32);
        if (n && f->u_right) memcpy(L.host<float>(o_ur), f->u_right, (size_t)n * 4);
        for (int i = 0; i < n; ++i) L.host<int>(o_taken)[i] = f->occupied && f->occupied[i] ? -1 : INT_MAX;
        for (int i = 0; i < nq; ++i) L.host<int>(o_qm)[i] = -1;
        *L.host<int>(o_area_n) = 0;
        A.n = n; A.kp = L.dev<orbx_keypoint>(o_kp); A.desc = L.dev<uint8_t>(o_desc); A.u_right = f->u_right ? L.dev<float>(o_ur) : nullptr;
        A.min_x = f->min_x; A.min_y = f->min_y; A.grid_w_inv = f->grid_w_inv; A.grid_h_inv = f->grid_h_inv;
        A.cell_start = L.dev<int>(o_cs); A.items = L.dev<int>(o_items); A.cell_of = L.dev<unsigned short>(o_cellof);
        A.taken_by = L.dev<int>(o_taken); A.claim = L.dev<int>(o_claim);
        A.nq = nq; A.q = L.dev<ProjQuery>(o_q); A.qdesc = L.dev<uint8_t>(o_qd);
        A.keys = L.dev<unsigned long long>(o_keys); A.qmatch = L.dev<int>(o_qm); A.list_a = L.dev<int>(o_la); A.list_b = L.dev<int>(o_lb);
        A.rounds_out = L.dev<int>(o_rounds);
        return ORBX_OK;
    }
    ProjQuery* queries() { return L.host<ProjQuery>(o_q); }
    uint8_t* qdesc() { return L.host<uint8_t>(o_qd); }
    cudaStream_t stream() const { return L.A.st; }

    int upload_and_grid()
    {
        cudaError_t e = cudaMemcpyAsync(L.A.d, L.A.h, L.upload_end, cudaMemcpyHostToDevice, stream());
        if (e == cudaSuccess) e = launch_frame_grid(A, stream());
        if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, "frame grid: %s", cudaGetErrorString(e));
        return ORBX_OK;
    }
    // Run the search; on return qmatch()[i] = feature assigned to query i or -1.
    int search(int mode, int th_dist, float nn_ratio)
    {
        int rc;
        if ((rc = upload_and_grid())) return rc;
        A.mode = mode; A.th_dist = th_dist; A.nn_ratio = nn_ratio;
        cudaError_t e = launch_proj_search(A, stream());
        // qmatch and the round counter come back through the staging buffer (device-only offsets map 1:1 into it)
        if (e == cudaSuccess && nq) e = cudaMemcpyAsync(L.host<int>(o_qm), L.dev<int>(o_qm), (size_t)nq * 4, cudaMemcpyDeviceToHost, stream());
        if (e == cudaSuccess && nq) e = cudaMemcpyAsync(L.host<int>(o_rounds), L.dev<int>(o_rounds), 4, cudaMemcpyDeviceToHost, stream());
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream());
        if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, "projection search: %s", cudaGetErrorString(e));
        t_rounds = nq ? *L.host<int>(o_rounds) : 0;
        return ORBX_OK;
    }
    const int* qmatch() const { return L.host<int>(o_qm); }
};

// src/ORBmatcher1.cc:217-223
float radius_by_viewing_cos(const float& viewCos)
{
    if (viewCos > 0.998) return 2.5;
    else return 4.0;
}

// Shared tail of the two 1-NN variants: assignments in query order, then the rotation filter (src/ORBmatcher3.cc:442-464, 555-575).
int finish_1nn(const Session& S, const orbx_frame_view* cur, const float* angle_q, int check_orientation, int32_t* match_f, int* n_matches)
{
    const int H = 30;
    int nm = 0;
    for (int i = 0; i < cur->n; ++i) match_f[i] = -1;
    std::vector<int> rot[H];
    for (int i = 0; i < S.nq; ++i) {
        const int f = S.qmatch()[i];
        if (f < 0) continue;
        match_f[f] = i;
        ++nm;
        if (check_orientation) {
            const int b = rotation_bin(angle_q[i], cur->keys_un[f].angle);
            if (b < 0) return fail(ORBX_ERR_INVALID_ARG, "keypoint angles out of range (the reference asserts)");
            rot[b].push_back(f);
        }
    }
    if (check_orientation) {
        int count[H], i1, i2, i3;
        for (int b = 0; b < H; ++b) count[b] = (int)rot[b].size();
        three_maxima(count, H, i1, i2, i3);
        for (int b = 0; b < H; ++b) {
            if (b == i1 || b == i2 || b == i3) continue;
            for (int f : rot[b]) { match_f[f] = -1; --nm; }
        }
    }
    *n_matches = nm;
    return ORBX_OK;
}

}  // namespace

}  // namespace orbx

using namespace orbx;

extern "C" {

int orbx_projection_rounds(void) { return t_rounds; }

int orbx_assign_features_to_grid(int device, const orbx_frame_view* frame, int32_t* cell_start, int32_t* items)
{
    int rc;
    if ((rc = check_frame(frame, false))) return rc;
    if (!cell_start || (frame->n > 0 && !items)) return fail(ORBX_ERR_INVALID_ARG, "null output");
    Session S(device);
    if ((rc = S.begin(device, frame, 0, false))) return rc;
    if ((rc = S.upload_and_grid())) return rc;
    cudaError_t e = cudaMemcpyAsync(cell_start, S.A.cell_start, (GRID_CELLS + 1) * 4, cudaMemcpyDeviceToHost, S.stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(S.stream());
    if (e == cudaSuccess && cell_start[GRID_CELLS] > 0)
        e = cudaMemcpy(items, S.A.items, (size_t)cell_start[GRID_CELLS] * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, "assign_features_to_grid: %s", cudaGetErrorString(e));
    return ORBX_OK;
}

int orbx_get_features_in_area(int device, const orbx_frame_view* frame, float x, float y, float r, int min_level, int max_level,
                              int32_t* out, int capacity, int* n_out)
{
    int rc;
    if ((rc = check_frame(frame, false))) return rc;
    if (!n_out || capacity < 0 || (capacity > 0 && !out)) return fail(ORBX_ERR_INVALID_ARG, "bad output arguments");
    Session S(device);
    if ((rc = S.begin(device, frame, 0, false, frame->n))) return rc;
    if ((rc = S.upload_and_grid())) return rc;
    int* d_n = S.L.dev<int>(S.o_area_n);
    unsigned long long* d_out = S.L.dev<unsigned long long>(S.o_area);
    cudaError_t e = launch_features_in_area(S.A, x, y, r, min_level, max_level, d_out, frame->n, d_n, S.stream());
    int cnt = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&cnt, d_n, 4, cudaMemcpyDeviceToHost, S.stream());
    if (e == cudaSuccess) e = cudaStreamSynchronize(S.stream());
    std::vector<unsigned long long> keys(cnt);
    if (e == cudaSuccess && cnt) e = cudaMemcpy(keys.data(), d_out, (size_t)cnt * 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) return fail(ORBX_ERR_CUDA, "get_features_in_area: %s", cudaGetErrorString(e));
    *n_out = cnt;
    if (cnt > capacity) return fail(ORBX_ERR_CAPACITY, "get_features_in_area: %d candidates, capacity %d", cnt, capacity);
    std::sort(keys.begin(), keys.end());                      // (cell id, index) = the reference's visiting order
    for (int i = 0; i < cnt; ++i) out[i] = (int32_t)(uint32_t)keys[i];
    return ORBX_OK;
}

int orbx_search_by_projection_map(int device, const orbx_frame_view* frame, const orbx_track_points* p, float th, int far_points,
                                  float th_far_points, float nn_ratio, int32_t* match_f, int* n_matches)
{
    int rc;
    if ((rc = check_frame(frame, true))) return rc;
    if (!p || p->n < 0 || !n_matches || (frame->n > 0 && !match_f)) return fail(ORBX_ERR_INVALID_ARG, "bad arguments");
    if (p->n > 0 && (!p->in_view || !p->bad || !p->proj_x || !p->proj_y || !p->proj_xr || !p->view_cos || !p->scale_level || !p->n_obs ||
                     !p->descriptors || (far_points && !p->track_depth)))
        return fail(ORBX_ERR_INVALID_ARG, "track points: missing arrays");
    *n_matches = 0;
    for (int i = 0; i < frame->n; ++i) match_f[i] = -1;
    if (frame->n == 0 || p->n == 0) { t_rounds = 0; return ORBX_OK; }
    Session S(device);
    if ((rc = S.begin(device, frame, p->n, true))) return rc;
    ProjQuery* q = S.queries();
    const bool bFactor = th != 1.0;
    for (int i = 0; i < p->n; ++i) {
        ProjQuery w{};
        bool ok = p->in_view[i] != 0;                                        // src/ORBmatcher1.cc:54-55 (mbTrackInViewR: two-fisheye only)
        if (ok && far_points && p->track_depth[i] > th_far_points) ok = false;   // :57-58
        if (ok && p->bad[i]) ok = false;                                     // :60-61
        if (ok) {
            const int lvl = p->scale_level[i];
            if (lvl < 0 || lvl >= frame->n_levels) return fail(ORBX_ERR_INVALID_ARG, "track point %d: scale level %d out of range", i, lvl);
            float r = radius_by_viewing_cos(p->view_cos[i]);                 // :68
            if (bFactor) r *= th;                                            // :70-71
            w.x = p->proj_x[i]; w.y = p->proj_y[i];
            w.r = r * frame->scale_factors[lvl];                             // :74
            w.ur = p->proj_xr[i];
            w.min_level = lvl - 1; w.max_level = lvl;
            w.flags = PROJ_Q_VALID | PROJ_Q_CHECK_RIGHT | (p->n_obs[i] > 0 ? PROJ_Q_TAKES : 0);
        }
        q[i] = w;
    }
    memcpy(S.qdesc(), p->descriptors, (size_t)p->n * 32);
    if ((rc = S.search(0, 100 /* TH_HIGH, src/ORBmatcher1.cc:37 */, nn_ratio))) return rc;
    int nm = 0;
    for (int i = 0; i < p->n; ++i)
        if (S.qmatch()[i] >= 0) { match_f[S.qmatch()[i]] = i; ++nm; }       // later map points overwrite earlier un-observed ones
    *n_matches = nm;
    return ORBX_OK;
}

int orbx_search_by_projection_last(int device, const orbx_frame_view* cur, float mbf, int n_last, const uint8_t* valid, const float* u,
                                   const float* v, const float* invz, const int32_t* octave, const float* angle, const int32_t* n_obs,
                                   const uint8_t* desc, float th, int forward, int backward, int check_orientation, int32_t* match_f,
                                   int* n_matches)
{
    int rc;
    if ((rc = check_frame(cur, true))) return rc;
    if (n_last < 0 || !n_matches || (cur->n > 0 && !match_f) ||
        (n_last > 0 && (!valid || !u || !v || !invz || !octave || !n_obs || !desc || (check_orientation && !angle))))
        return fail(ORBX_ERR_INVALID_ARG, "bad arguments");
    *n_matches = 0;
    for (int i = 0; i < cur->n; ++i) match_f[i] = -1;
    if (cur->n == 0 || n_last == 0) { t_rounds = 0; return ORBX_OK; }
    Session S(device);
    if ((rc = S.begin(device, cur, n_last, true))) return rc;
    ProjQuery* q = S.queries();
    for (int i = 0; i < n_last; ++i) {
        ProjQuery w{};
        bool ok = valid[i] != 0;
        const float invzc = ok ? invz[i] : 0.0f;
        if (ok && invzc < 0) ok = false;                                                         // src/ORBmatcher3.cc:293-294
        if (ok && (u[i] < cur->min_x || u[i] > cur->max_x)) ok = false;                          // :298-299
        if (ok && (v[i] < cur->min_y || v[i] > cur->max_y)) ok = false;                          // :300-301
        if (ok) {
            const int oct = octave[i];
            if (oct < 0 || oct >= cur->n_levels) return fail(ORBX_ERR_INVALID_ARG, "last-frame feature %d: octave %d out of range", i, oct);
            w.x = u[i]; w.y = v[i];
            w.r = th * cur->scale_factors[oct];                                                  // :307
            if (forward) { w.min_level = oct; w.max_level = -1; }                                // :311-316
            else if (backward) { w.min_level = 0; w.max_level = oct; }
            else { w.min_level = oct - 1; w.max_level = oct + 1; }
            w.ur = u[i] - mbf * invzc;                                                           // :332
            w.flags = PROJ_Q_VALID | PROJ_Q_CHECK_RIGHT | (n_obs[i] > 0 ? PROJ_Q_TAKES : 0);
        }
        q[i] = w;
    }
    memcpy(S.qdesc(), desc, (size_t)n_last * 32);
    if ((rc = S.search(1, 100 /* TH_HIGH */, 0.0f))) return rc;
    return finish_1nn(S, cur, angle, check_orientation, match_f, n_matches);
}

int orbx_search_by_projection_kf(int device, const orbx_frame_view* cur, int n_kf, const uint8_t* valid, const float* u, const float* v,
                                 const float* dist3d, const float* min_dist, const float* max_dist, const int32_t* level,
                                 const float* angle, const uint8_t* desc, float th, int orb_dist, int check_orientation,
                                 int32_t* match_f, int* n_matches)
{
    int rc;
    if ((rc = check_frame(cur, true))) return rc;
    if (n_kf < 0 || !n_matches || (cur->n > 0 && !match_f) ||
        (n_kf > 0 && (!valid || !u || !v || !dist3d || !min_dist || !max_dist || !level || !desc || (check_orientation && !angle))))
        return fail(ORBX_ERR_INVALID_ARG, "bad arguments");
    *n_matches = 0;
    for (int i = 0; i < cur->n; ++i) match_f[i] = -1;
    if (cur->n == 0 || n_kf == 0) { t_rounds = 0; return ORBX_OK; }
    Session S(device);
    if ((rc = S.begin(device, cur, n_kf, true))) return rc;
    ProjQuery* q = S.queries();
    for (int i = 0; i < n_kf; ++i) {
        ProjQuery w{};
        bool ok = valid[i] != 0;
        if (ok && (u[i] < cur->min_x || u[i] > cur->max_x)) ok = false;                          // src/ORBmatcher3.cc:498-499
        if (ok && (v[i] < cur->min_y || v[i] > cur->max_y)) ok = false;                          // :500-501
        if (ok && (dist3d[i] < min_dist[i] || dist3d[i] > max_dist[i])) ok = false;              // :511-512
        if (ok) {
            const int lvl = level[i];
            if (lvl < 0 || lvl >= cur->n_levels) return fail(ORBX_ERR_INVALID_ARG, "key-frame point %d: level %d out of range", i, lvl);
            w.x = u[i]; w.y = v[i];
            w.r = th * cur->scale_factors[lvl];                                                  // :517
            w.min_level = lvl - 1; w.max_level = lvl + 1;                                        // :519
            w.flags = PROJ_Q_VALID | PROJ_Q_TAKES;                                               // any map point occupies (:533)
        }
        q[i] = w;
    }
    memcpy(S.qdesc(), desc, (size_t)n_kf * 32);
    if ((rc = S.search(1, orb_dist, 0.0f))) return rc;
    return finish_1nn(S, cur, angle, check_orientation, match_f, n_matches);
}

}  // extern "C"
