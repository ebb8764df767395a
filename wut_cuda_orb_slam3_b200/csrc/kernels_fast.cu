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
constexpr int kListCap = kMaxEval + 4 * 32;     // four warp-private regions, each rounded up to 32
constexpr int kBitWords = (kPlane + 31) / 32;   // survivor bitmaps are indexed by plane position

// Shared-memory accessors on explicit 32-bit shared addresses.  (nvcc re-derives the shared-window base — S2UR
// SR_CgaCtaId + ULEA — at every use inside divergent regions when it goes through C++ pointers; the first version of this
// kernel spent ~15 % of its issue slots on that.)  Offsets are compile-time immediates folded into the instruction.
#define ORBX_LDS_U8(dst, addr, off) asm volatile("ld.shared.u8 %0, [%1+" #off "];" : "=r"(dst) : "r"(addr))
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_zero16(uint32_t a) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(0u) : "memory"); }
__device__ __forceinline__ void atom_or(uint32_t a, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// FAST score of the pixel whose plane address is `c`: max over the 16 arcs of 9 contiguous circle pixels of
// min(v - ring) / min(ring - v), minus 1.  Both signs are carried in one register as s16x2 (low = v - ring "centre
// brighter", high = ring - v) so one VIMNMX.S16x2 serves both.
//   packing:  A = (v+1, 1-v),  ~(r, -r) = r*0xFFFF - 1  ->  P = A + ~(r,-r) per half = (v - r, r - v)
//   windows:  prefix/suffix minima inside the two half-circles, window k = min(suffix(k), prefix(k+8))
__device__ __forceinline__ int fast_score(uint32_t c)
{
    // base moved to the top-left of the 7x7 neighbourhood so that every offset is a non-negative immediate
    const uint32_t b = c - 3 * kRoiPitch - 3;
    uint32_t v, r[16];
    ORBX_LDS_U8(v, b, 243);       // (3,3)
    ORBX_LDS_U8(r[0], b, 483);    // ( 0, 3): row 6, col 3
    ORBX_LDS_U8(r[1], b, 484);    // ( 1, 3)
    ORBX_LDS_U8(r[2], b, 405);    // ( 2, 2): row 5, col 5
    ORBX_LDS_U8(r[3], b, 326);    // ( 3, 1): row 4, col 6
    ORBX_LDS_U8(r[4], b, 246);    // ( 3, 0)
    ORBX_LDS_U8(r[5], b, 166);    // ( 3,-1): row 2, col 6
    ORBX_LDS_U8(r[6], b, 85);     // ( 2,-2): row 1, col 5
    ORBX_LDS_U8(r[7], b, 4);      // ( 1,-3): row 0, col 4
    ORBX_LDS_U8(r[8], b, 3);      // ( 0,-3)
    ORBX_LDS_U8(r[9], b, 2);      // (-1,-3)
    ORBX_LDS_U8(r[10], b, 81);    // (-2,-2): row 1, col 1
    ORBX_LDS_U8(r[11], b, 160);   // (-3,-1): row 2, col 0
    ORBX_LDS_U8(r[12], b, 240);   // (-3, 0)
    ORBX_LDS_U8(r[13], b, 320);   // (-3, 1)
    ORBX_LDS_U8(r[14], b, 401);   // (-2, 2): row 5, col 1
    ORBX_LDS_U8(r[15], b, 482);   // (-1, 3): row 6, col 2
    const uint32_t A = v * 0xFFFF0001u + 0x00010001u;
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
    const int a = (int)(short)(w[0] & 0xffff), bb = (int)(short)(w[0] >> 16);
    return max(a, bb) - 1;
}
static_assert(kRoiPitch == 80, "fast_score() hard-codes the 7x7 offsets for an 80-byte pitch");

}  // namespace

__global__ void __launch_bounds__(128) fast_cells_kernel(const __grid_constant__ FrameGeom fg, Workspace ws)
{
    __shared__ __align__(16) uint8_t roi[kPlane];
    __shared__ __align__(16) uint8_t sc[kPlane];
    __shared__ __align__(4) uint16_t list[kListCap];
    __shared__ uint32_t selA[kBitWords], selH[kBitWords];
    __shared__ int off[kBitWords];
    __shared__ int wcnt[4], total;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int frame = blockIdx.y;
    // cell -> (level, cell row, cell column) from a small table (built by the host with the geometry)
    const uint32_t ct = __ldg(fg.cell_tab + blockIdx.x);
    const int level = ct & 15, ci = (ct >> 4) & 0xfff, cj = ct >> 16;
    const LevelGeom& g = fg.L[level];
    const int cell = (int)blockIdx.x - g.cell_base;
    int* count_out = ws.cell_count + (size_t)frame * fg.total_cells + blockIdx.x;

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
    const uint32_t roi_s = (uint32_t)__cvta_generic_to_shared(roi), sc_s = (uint32_t)__cvta_generic_to_shared(sc);
    const uint32_t list_s = (uint32_t)__cvta_generic_to_shared(list);
    const uint32_t selA_s = (uint32_t)__cvta_generic_to_shared(selA), selH_s = (uint32_t)__cvta_generic_to_shared(selH);

    // 1. ROI -> shared memory with aligned 32-bit loads (the row misalignment m is the same for every row because the
    //    pitch is a multiple of 16); zero the score plane and the survivor bitmaps
    const uint8_t* src = level_interior((const uint8_t*)ws.pyr, g, frame) + (size_t)iniY * g.pitch + iniX;
    const int m = (int)((uintptr_t)src & 3);
    {
        const int nwords = (m + rw + 3) >> 2;                   // <= 20
        const uint32_t* src4 = reinterpret_cast<const uint32_t*>(src - m) + lane;
        const int wpitch = g.pitch >> 2;
        // all global loads of this lane are issued before the first shared store (the stores are volatile asm, which would
        // otherwise serialise ~10 dependent L2 round trips at the start of every CTA)
        constexpr int kRowsPerWarp = (kRoiRows + 3) / 4;
        uint32_t tmp[kRowsPerWarp];
        const bool ld = lane < nwords;
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int y = warp + 4 * k;
            tmp[k] = (ld && y < rh) ? __ldg(src4 + (size_t)y * wpitch) : 0u;
        }
        const int nz = (rh * kRoiPitch + 15) >> 4;
        for (int i = tid; i < nz; i += 128) sts_zero16(sc_s + i * 16);
        for (int i = tid; i < kBitWords; i += 128) { sts_u32(selA_s + i * 4, 0); sts_u32(selH_s + i * 4, 0); }
#pragma unroll
        for (int k = 0; k < kRowsPerWarp; ++k) {
            const int y = warp + 4 * k;
            if (ld && y < rh) sts_u32(roi_s + y * kRoiPitch + lane * 4, tmp[k]);
        }
    }
    __syncthreads();

    // 2. pre-test on the 8 even circle points: every 9-arc contains one point of each opposite pair, so a corner needs
    //    max_p min(r_a, r_b) < v - t  (bright) or  min_p max(r_a, r_b) > v + t  (dark).  Each warp owns a contiguous
    //    quarter of the pixels and compacts its survivors (plane positions) into its own region of `list`: no atomics.
    const int npix = ew * eh;
    const int Q = (((npix + 3) >> 2) + 31) & ~31;               // pixels per warp, multiple of 32
    {
        const int e0 = warp * Q + lane;
        const int eend = min(warp * Q + Q, npix);               // warp-uniform end of this warp's range
        const float inv_ew = 1.0f / (float)ew;
        int ey = (int)(((float)e0 + 0.5f) * inv_ew);
        int ex = e0 - ey * ew;
        int pos = (ey + 3) * kRoiPitch + m + ex + 3;
        const int wrapfix = kRoiPitch - ew;
        const int safe = 3 * kRoiPitch + m + 3;                  // first evaluated pixel: a valid address for idle lanes
        uint32_t wl = list_s + (uint32_t)(warp * Q) * 2;        // write cursor of this warp (bytes)
        const int vlo = minTh, vhi = minTh;
        for (int base = warp * Q; base < eend; base += 32) {
            const bool valid = base + lane < eend;
            const uint32_t b = roi_s + (valid ? pos : safe) - 3 * kRoiPitch - 3;
            uint32_t v, r0, r2, r4, r6, r8, r10, r12, r14;
            ORBX_LDS_U8(v, b, 243);
            ORBX_LDS_U8(r0, b, 483);  ORBX_LDS_U8(r8, b, 3);
            ORBX_LDS_U8(r4, b, 246);  ORBX_LDS_U8(r12, b, 240);
            ORBX_LDS_U8(r2, b, 405);  ORBX_LDS_U8(r10, b, 81);
            ORBX_LDS_U8(r6, b, 85);   ORBX_LDS_U8(r14, b, 401);
            const int M1 = max(max(min(r0, r8), min(r4, r12)), max(min(r2, r10), min(r6, r14)));
            const int M2 = min(min(max(r0, r8), max(r4, r12)), min(max(r2, r10), max(r6, r14)));
            const bool pass = valid & ((M1 < (int)v - vlo) | (M2 > (int)v + vhi));
            const uint32_t mk = __ballot_sync(0xffffffffu, pass);
            if (pass) sts_u16(wl + 2 * __popc(mk & ((1u << lane) - 1)), (uint32_t)pos);
            wl += 2 * __popc(mk);
            ex += 32; pos += 32;
            while (ex >= ew) { ex -= ew; pos += wrapfix; }
        }
        if (lane == 0) wcnt[warp] = (int)((wl - list_s) >> 1) - warp * Q;
    }
    __syncthreads();

    // flat index over the four warp regions -> list slot
    const int c0 = wcnt[0], c1 = c0 + wcnt[1], c2 = c1 + wcnt[2], nl = c2 + wcnt[3];
    auto slot = [&](int i) { return i < c0 ? i : (i < c1 ? Q + i - c0 : (i < c2 ? 2 * Q + i - c1 : 3 * Q + i - c2)); };

    // 3. full arc score on the compacted list (order inside the list is irrelevant)
    for (int i = tid; i < nl; i += 128) {
        const uint32_t pos = lds_u16(list_s + 2 * slot(i));
        const int s = fast_score(roi_s + pos);
        if (s >= minTh) sts_u8(sc_s + pos, (uint32_t)s);
    }
    __syncthreads();

    // 4. 3x3 strict NMS inside the cell (pixels outside the evaluated region hold score 0 = cv::FAST's zeroed buffer);
    //    survivors set their bit in position-indexed bitmaps
    for (int i = tid; i < nl; i += 128) {
        const uint32_t pos = lds_u16(list_s + 2 * slot(i));
        const uint32_t b = sc_s + pos - kRoiPitch - 1;
        uint32_t s;
        ORBX_LDS_U8(s, b, 81);
        if (s > 0) {
            uint32_t n0, n1, n2, n3, n4, n5, n6, n7;
            ORBX_LDS_U8(n0, b, 0);   ORBX_LDS_U8(n1, b, 1);   ORBX_LDS_U8(n2, b, 2);
            ORBX_LDS_U8(n3, b, 80);  ORBX_LDS_U8(n4, b, 82);
            ORBX_LDS_U8(n5, b, 160); ORBX_LDS_U8(n6, b, 161); ORBX_LDS_U8(n7, b, 162);
            const uint32_t mx = max(max(max(n0, n1), max(n2, n3)), max(max(n4, n5), max(n6, n7)));
            if (s > mx) {
                atom_or(selA_s + (pos >> 5) * 4, 1u << (pos & 31));
                if ((int)s >= iniTh) atom_or(selH_s + (pos >> 5) * 4, 1u << (pos & 31));
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
