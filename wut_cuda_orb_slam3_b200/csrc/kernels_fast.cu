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
// cell-masked).  ROI -> shared memory with aligned 32-bit loads, 8-point pre-test + shared-memory compaction so that
// the expensive arc test runs on dense warps, packed s16x2 min/max (VIMNMX.S16x2) computes the bright and dark arc
// scores at once, NMS only visits pre-test survivors, and position-indexed bitmaps + a warp scan write the survivors
// in (y,x) order into the cell's staging slot (no atomics on the output order).
#include "orbx_internal.cuh"

namespace orbx {

namespace {

constexpr int kRoiPitch = 80;                   // >= 3 (misalignment) + 70 + 6, multiple of 16
constexpr int kRoiRows = 76;                    // >= 70 + 6
constexpr int kPlane = kRoiRows * kRoiPitch;    // ROI plane and score plane share one geometry
constexpr int kMaxEval = kMaxCellDim * kMaxCellDim;
constexpr int kBitWords = (kPlane + 31) / 32;   // survivor bitmaps are indexed by plane position

// FAST score of the pixel at `c` (shared memory, row pitch kRoiPitch): max over the 16 arcs of 9 contiguous circle
// pixels of min(v - ring) / min(ring - v), minus 1.  Both signs are carried in one register as s16x2
// (low = v - ring "centre brighter", high = ring - v) so one VIMNMX.S16x2 serves both.
//   packing:  A = (v+1, 1-v),  ~(r, -r) = r*0xFFFF - 1  ->  P = A + ~(r,-r) per half = (v - r, r - v)
//   windows:  prefix/suffix minima inside the two half-circles, window k = min(suffix(k), prefix(k+8))
__device__ __forceinline__ int fast_score(const uint8_t* c)
{
    constexpr int rp = kRoiPitch;
    const uint32_t v = c[0];
    const uint32_t A = v * 0xFFFF0001u + 0x00010001u;
    uint32_t r[16];
    r[0] = c[3 * rp];       r[1] = c[3 * rp + 1];   r[2] = c[2 * rp + 2];   r[3] = c[rp + 3];
    r[4] = c[3];            r[5] = c[-rp + 3];      r[6] = c[-2 * rp + 2];  r[7] = c[-3 * rp + 1];
    r[8] = c[-3 * rp];      r[9] = c[-3 * rp - 1];  r[10] = c[-2 * rp - 2]; r[11] = c[-rp - 3];
    r[12] = c[-3];          r[13] = c[rp - 3];      r[14] = c[2 * rp - 2];  r[15] = c[3 * rp - 1];
    uint32_t P[16], pf[16], sf[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) P[k] = __vadd2(A, r[k] * 0xFFFFu + 0xFFFFFFFFu);
    pf[0] = P[0]; pf[8] = P[8]; sf[7] = P[7]; sf[15] = P[15];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        pf[k] = __vmins2(pf[k - 1], P[k]);
        pf[8 + k] = __vmins2(pf[8 + k - 1], P[8 + k]);
        sf[7 - k] = __vmins2(sf[8 - k], P[7 - k]);
        sf[15 - k] = __vmins2(sf[16 - k], P[15 - k]);
    }
    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        w[k] = __vmins2(sf[k], pf[k + 8]);        // window k..k+8
        w[k + 8] = __vmins2(sf[k + 8], pf[k]);    // window k+8..k+16
    }
#pragma unroll
    for (int s = 8; s >= 1; s >>= 1)
#pragma unroll
        for (int k = 0; k < s; ++k) w[k] = __vmaxs2(w[k], w[k + s]);
    const int a = (int)(short)(w[0] & 0xffff), b = (int)(short)(w[0] >> 16);
    return max(a, b) - 1;
}

}  // namespace

__global__ void __launch_bounds__(128) fast_cells_kernel(const __grid_constant__ FrameGeom fg, Workspace ws)
{
    __shared__ __align__(16) uint8_t roi[kPlane];
    __shared__ __align__(16) uint8_t sc[kPlane];
    __shared__ uint16_t list[kMaxEval];
    __shared__ uint32_t selA[kBitWords], selH[kBitWords];
    __shared__ int off[kBitWords];
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

    // 1. ROI -> shared memory with aligned 32-bit loads (the row misalignment m is the same for every row because the
    //    pitch is a multiple of 16); zero the score plane and the survivor bitmaps
    const uint8_t* src = level_interior((const uint8_t*)ws.pyr, g, frame) + (size_t)iniY * g.pitch + iniX;
    const int m = (int)((uintptr_t)src & 3);
    const int nwords = (m + rw + 3) >> 2;                       // <= 20
    {
        const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src - m);
        const int wpitch = g.pitch >> 2;
        for (int y = warp; y < rh; y += 4)
            if (lane < nwords) reinterpret_cast<uint32_t*>(roi)[y * (kRoiPitch / 4) + lane] = __ldg(src4 + (size_t)y * wpitch + lane);
        const int nz = (rh * kRoiPitch + 15) >> 4;
        for (int i = tid; i < nz; i += 128) reinterpret_cast<uint4*>(sc)[i] = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < kBitWords; i += 128) { selA[i] = 0; selH[i] = 0; }
        if (tid == 0) n_list = 0;
    }
    __syncthreads();

    // 2. pre-test on the 8 even circle points: every 9-arc contains one point of each opposite pair, so a corner needs
    //    max_p min(r_a, r_b) < v - t  (bright) or  min_p max(r_a, r_b) > v + t  (dark).  Survivors are compacted so the
    //    expensive arc test runs on dense warps.
    const int npix = ew * eh;
    const int npad = (npix + 31) & ~31;
    {
        const int sdy = 128 / ew, sdx = 128 - sdy * ew;
        int ey = tid / ew, ex = tid - ey * ew;
        for (int e = tid; e < npad; e += 128) {
            bool pass = false;
            const int pos = (ey + 3) * kRoiPitch + m + ex + 3;
            if (e < npix) {
                const uint8_t* c = roi + pos;
                constexpr int rp = kRoiPitch;
                const int v = c[0];
                const int r0 = c[3 * rp], r8 = c[-3 * rp], r4 = c[3], r12 = c[-3];
                const int r2 = c[2 * rp + 2], r10 = c[-2 * rp - 2], r6 = c[-2 * rp + 2], r14 = c[2 * rp - 2];
                const int M1 = max(max(min(r0, r8), min(r4, r12)), max(min(r2, r10), min(r6, r14)));
                const int M2 = min(min(max(r0, r8), max(r4, r12)), min(max(r2, r10), max(r6, r14)));
                pass = (M1 < v - minTh) | (M2 > v + minTh);
            }
            const uint32_t mk = __ballot_sync(0xffffffffu, pass);
            if (mk) {
                int base = 0;
                if (lane == 0) base = atomicAdd(&n_list, __popc(mk));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pass) list[base + __popc(mk & ((1u << lane) - 1))] = (uint16_t)pos;
            }
            ex += sdx; ey += sdy;
            if (ex >= ew) { ex -= ew; ++ey; }
        }
    }
    __syncthreads();

    // 3. full arc score on the compacted list (order inside the list is irrelevant)
    const int nl = n_list;
    for (int i = tid; i < nl; i += 128) {
        const int pos = list[i];
        const int s = fast_score(roi + pos);
        if (s >= minTh) sc[pos] = (uint8_t)s;
    }
    __syncthreads();

    // 4. 3x3 strict NMS inside the cell (pixels outside the evaluated region hold score 0 = cv::FAST's zeroed buffer);
    //    survivors set their bit in position-indexed bitmaps
    for (int i = tid; i < nl; i += 128) {
        const int pos = list[i];
        const uint8_t* p = sc + pos;
        const int s = p[0];
        if (s > 0) {
            constexpr int sp = kRoiPitch;
            const bool keep = s > p[-1] && s > p[1] && s > p[-sp - 1] && s > p[-sp] && s > p[-sp + 1] && s > p[sp - 1] &&
                              s > p[sp] && s > p[sp + 1];
            if (keep) {
                atomicOr(&selA[pos >> 5], 1u << (pos & 31));
                if (s >= iniTh) atomicOr(&selH[pos >> 5], 1u << (pos & 31));
            }
        }
    }
    __syncthreads();

    // 5. per-cell threshold selection (ini if it yields anything, else min) + exclusive offsets (warp 0)
    const int nwordsB = ((rh * kRoiPitch) + 31) >> 5;
    if (warp == 0) {
        uint32_t anyH = 0;
        for (int w = lane; w < nwordsB; w += 32) anyH |= selH[w];
        anyH = __ballot_sync(0xffffffffu, anyH != 0);
        int running = 0;
        for (int base = 0; base < nwordsB; base += 32) {
            const int w = base + lane;
            uint32_t mk = 0;
            if (w < nwordsB) { mk = anyH ? selH[w] : selA[w]; selA[w] = mk; }
            const int cnt = __popc(mk);
            int incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (w < nwordsB) off[w] = running + incl - cnt;
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) { total = running; *count_out = running; }
    }
    __syncthreads();

    // 6. ordered scatter: plane position order == (y, x) order
    if (total == 0) return;
    uint32_t* out = ws.cand + (size_t)frame * fg.cand_frame_stride + g.cand_off + (size_t)cell * g.cell_cap;
    const int xbase = cj * g.wCell - m, ybase = ci * g.hCell;          // x_rel = cj*wCell + 3 + ex, ex = col - m - 3
    for (int w = tid; w < nwordsB; w += 128) {
        uint32_t mk = selA[w];
        int o = off[w];
        while (mk) {
            const int b = __ffs(mk) - 1;
            mk &= mk - 1;
            const int pos = (w << 5) + b;
            const int y = pos / kRoiPitch, x = pos - y * kRoiPitch;
            out[o++] = (uint32_t)(xbase + x) | ((uint32_t)(ybase + y) << 12) | ((uint32_t)sc[pos] << 24);
        }
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
