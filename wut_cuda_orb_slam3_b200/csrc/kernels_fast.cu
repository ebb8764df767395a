// kernels_fast.cu — grid FAST-9/16 with the iniThFAST/minThFAST two-threshold retry per 35-px cell.
//
// Semantics reproduced: the CPU cell loop `tileCalcKeypoints` (reference src/ORBextractor.cc:867-950, window at
// 958-966): every cell is an independent cv::FAST(roi, iniThFAST, nms=true) call on the (wCell+6)x(hCell+6) ROI and,
// only if that returns nothing, cv::FAST(roi, minThFAST, true).  cv::FAST evaluates ROI pixels >= 3 px from the ROI
// edge, its 3x3 strict-'>' NMS sees un-evaluated neighbours as 0, and keypoints come out in (y,x) order.
//
// One score pass is enough: the FAST score (max threshold for which the pixel is still a corner) does not depend on
// the detection threshold, so FAST(roi, ini) == {k in FAST(roi, min): score >= ini}.
//
// Mapping: one 128-thread CTA per cell (exactly the reference's unit of independence, so NMS is naturally
// cell-masked).  ROI -> shared memory, compass pre-test + shared-memory compaction so that the expensive arc test
// runs on dense warps, packed s16x2 min/max (VIMNMX.S16x2) computes the bright and dark arc scores at once, then
// ballot-based ordered compaction writes the survivors in (y,x) order into the cell's staging slot.
#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int kRoiPitch = 80;                   // >= 70 + 6
constexpr int kRoiRows = 76;
constexpr int kScorePitch = 76;                 // >= 70 + 2, multiple of 4
constexpr int kScoreRows = 72;
constexpr int kMaxEval = kMaxCellDim * kMaxCellDim;
constexpr int kMaxChunks = (kMaxEval + 31) / 32;

__device__ __forceinline__ uint32_t pk(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }

// FAST score of the pixel at `c` (shared memory, row pitch kRoiPitch): max over the 16 arcs of 9 contiguous circle
// pixels of min(v - ring) / min(ring - v), minus 1.  Low s16 lane carries v-ring ("centre brighter"), high lane ring-v.
__device__ __forceinline__ int fast_score(const uint8_t* c)
{
    constexpr int rp = kRoiPitch;
    const int v = c[0];
    int r[16];
    r[0] = c[3 * rp];       r[1] = c[3 * rp + 1];   r[2] = c[2 * rp + 2];   r[3] = c[rp + 3];
    r[4] = c[3];            r[5] = c[-rp + 3];      r[6] = c[-2 * rp + 2];  r[7] = c[-3 * rp + 1];
    r[8] = c[-3 * rp];      r[9] = c[-3 * rp - 1];  r[10] = c[-2 * rp - 2]; r[11] = c[-rp - 3];
    r[12] = c[-3];          r[13] = c[rp - 3];      r[14] = c[2 * rp - 2];  r[15] = c[3 * rp - 1];
    uint32_t P[16], m2[16], m4[16], m9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) P[k] = pk(v - r[k], r[k] - v);
#pragma unroll
    for (int k = 0; k < 16; ++k) m2[k] = __vmins2(P[k], P[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m4[k] = __vmins2(m2[k], m2[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) m9[k] = __vmins2(__vmins2(m4[k], m4[(k + 4) & 15]), P[(k + 8) & 15]);
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1)
#pragma unroll
        for (int k = 0; k < s; ++k) m9[k] = __vmaxs2(m9[k], m9[k + s]);
    const int a = (int)(short)(m9[0] & 0xffff), b = (int)(short)(m9[0] >> 16);
    return max(a, b) - 1;
}

}  // namespace

__global__ void __launch_bounds__(128) fast_cells_kernel(const __grid_constant__ FrameGeom fg, Workspace ws)
{
    __shared__ __align__(16) uint8_t roi[kRoiRows * kRoiPitch];
    __shared__ __align__(16) uint8_t sc[kScoreRows * kScorePitch];
    __shared__ uint16_t list[kMaxEval];
    __shared__ uint32_t selA[kMaxChunks], selH[kMaxChunks];
    __shared__ int off[kMaxChunks];
    __shared__ int n_list, total;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    int cell = blockIdx.x;
    int level = 0;
#pragma unroll 1
    for (int l = 1; l < fg.nlevels; ++l)
        if (cell >= fg.L[l].cell_base) level = l;
    const LevelGeom& g = fg.L[level];
    cell -= g.cell_base;
    const int ci = cell / g.nCols, cj = cell - ci * g.nCols;   // cell row, cell column
    int* count_out = ws.cell_count + (size_t)frame * fg.total_cells + g.cell_base + cell;

    const int maxBX = g.w - kWinBorder, maxBY = g.h - kWinBorder;
    const int iniX = kWinBorder + cj * g.wCell, iniY = kWinBorder + ci * g.hCell;
    const int maxX = min(iniX + g.wCell + 6, maxBX), maxY = min(iniY + g.hCell + 6, maxBY);
    const int rw = maxX - iniX, rh = maxY - iniY;
    const int ew = rw - 6, eh = rh - 6;
    if (iniY >= maxBY - 3 || iniX >= maxBX - 6 || ew <= 0 || eh <= 0) {   // src/ORBextractor.cc:892,899
        if (tid == 0) *count_out = 0;
        return;
    }
    const int minTh = max(fg.minTh, 1), iniTh = fg.iniTh;

    // 1. ROI -> shared memory; zero the score plane (with its 1-px ring)
    const uint8_t* src = level_interior((const uint8_t*)ws.pyr, g, frame) + (size_t)iniY * g.pitch + iniX;
    for (int i = tid; i < rh * rw; i += 128) {
        const int y = i / rw, x = i - y * rw;
        roi[y * kRoiPitch + x] = __ldg(src + (size_t)y * g.pitch + x);
    }
    for (int i = tid; i < (eh + 2) * (kScorePitch / 4); i += 128) reinterpret_cast<uint32_t*>(sc)[i] = 0;
    if (tid == 0) n_list = 0;
    __syncthreads();

    // 2. compass pre-test: any 9-arc contains two adjacent compass points (0,4,8,12) -> compact survivors
    const int npix = ew * eh;
    const int npad = (npix + 31) & ~31;
    for (int e = tid; e < npad; e += 128) {
        bool pass = false;
        if (e < npix) {
            const int ey = e / ew, ex = e - ey * ew;
            const uint8_t* c = roi + (ey + 3) * kRoiPitch + ex + 3;
            const int v = c[0];
            const int d0 = v - c[3 * kRoiPitch], d4 = v - c[3], d8 = v - c[-3 * kRoiPitch], d12 = v - c[-3];
            const bool b = (d0 > minTh) | (d8 > minTh), b2 = (d4 > minTh) | (d12 > minTh);
            const bool k = (d0 < -minTh) | (d8 < -minTh), k2 = (d4 < -minTh) | (d12 < -minTh);
            pass = (b & b2) | (k & k2);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, pass);
        int base = 0;
        if (lane == 0 && m) base = atomicAdd(&n_list, __popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pass) list[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)e;
    }
    __syncthreads();

    // 3. full arc score on the compacted list
    const int nl = n_list;
    for (int i = tid; i < nl; i += 128) {
        const int e = list[i];
        const int ey = e / ew, ex = e - ey * ew;
        const int s = fast_score(roi + (ey + 3) * kRoiPitch + ex + 3);
        if (s >= minTh) sc[(ey + 1) * kScorePitch + ex + 1] = (uint8_t)s;
    }
    __syncthreads();

    // 4. 3x3 strict NMS inside the cell, chunks of 32 consecutive pixels in (y,x) order
    const int nchunks = npad >> 5;
    for (int ch = warp; ch < nchunks; ch += 4) {
        const int e = ch * 32 + lane;
        bool keep = false, high = false;
        if (e < npix) {
            const int ey = e / ew, ex = e - ey * ew;
            const uint8_t* p = sc + (ey + 1) * kScorePitch + ex + 1;
            const int s = p[0];
            if (s > 0) {
                keep = s > p[-1] && s > p[1] && s > p[-kScorePitch - 1] && s > p[-kScorePitch] && s > p[-kScorePitch + 1] &&
                       s > p[kScorePitch - 1] && s > p[kScorePitch] && s > p[kScorePitch + 1];
                high = keep && s >= iniTh;
            }
        }
        const uint32_t mA = __ballot_sync(0xffffffffu, keep), mH = __ballot_sync(0xffffffffu, high);
        if (lane == 0) { selA[ch] = mA; selH[ch] = mH; }
    }
    __syncthreads();

    // 5. per-cell threshold selection (ini if it yields anything, else min) + exclusive offsets (warp 0)
    if (warp == 0) {
        uint32_t anyH = 0;
        for (int ch = lane; ch < nchunks; ch += 32) anyH |= selH[ch];
        anyH = __ballot_sync(0xffffffffu, anyH != 0);
        int running = 0;
        for (int base = 0; base < nchunks; base += 32) {
            const int ch = base + lane;
            uint32_t m = 0;
            if (ch < nchunks) { m = anyH ? selH[ch] : selA[ch]; selA[ch] = m; }
            int cnt = __popc(m), incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (ch < nchunks) off[ch] = running + incl - cnt;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { total = running; *count_out = running; }
    }
    __syncthreads();

    // 6. ordered scatter into the cell's staging slot
    if (total == 0) return;
    uint32_t* out = ws.cand + (size_t)frame * fg.cand_frame_stride + g.cand_off + (size_t)cell * g.cell_cap;
    const int xbase = cj * g.wCell + 3, ybase = ci * g.hCell + 3;
    for (int ch = warp; ch < nchunks; ch += 4) {
        const uint32_t m = selA[ch];
        if (!((m >> lane) & 1u)) continue;
        const int e = ch * 32 + lane;
        const int ey = e / ew, ex = e - ey * ew;
        const uint32_t s = sc[(ey + 1) * kScorePitch + ex + 1];
        out[off[ch] + __popc(m & ((1u << lane) - 1))] = (uint32_t)(xbase + ex) | ((uint32_t)(ybase + ey) << 12) | (s << 24);
    }
}

cudaError_t launch_fast(const FrameGeom& fg, const Workspace& ws, int n_frames, cudaStream_t st)
{
    if (fg.total_cells == 0) return cudaSuccess;
    dim3 grid(fg.total_cells, n_frames);
    fast_cells_kernel<<<grid, 128, 0, st>>>(fg, ws);
    count_launch();
    return cudaGetLastError();
}

}  // namespace orbx
