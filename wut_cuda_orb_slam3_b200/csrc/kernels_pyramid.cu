// kernels_pyramid.cu — ORBextractor::ComputePyramid (reference src/ORBextractor.cc:1309-1329) on sm_100a.
//
// Level 0 = copyMakeBorder(image, 19, BORDER_REFLECT_101); level l>0 = cv::resize(level l-1, INTER_LINEAR) followed
// by the same reflect-101 border.  The OpenCV 8U bilinear arithmetic is integer fixed point (11-bit coefficients,
// ((b0*(r0>>4))>>16)+((b1*(r1>>4))>>16)+2)>>2); the per-column / per-row source offsets and coefficients are computed
// once on the host with OpenCV's float/double formula (api.cu: build_resize_table) and read here from small tables,
// indexed by *bordered* coordinates so that the border needs no extra pass: a border pixel is simply the resized value
// at its reflected interior coordinate.
//
// Data layout: every level is a bordered buffer [rows+38][pitch] per frame, interior pixel (0,0) at byte
// 19*pitch + 32 (16-byte aligned), frames strided by pyr_frame_stride, levels by pyr_off.  Each thread produces one
// aligned 32-bit word (4 pixels) of a bordered row, so stores are fully coalesced 128-byte lines per warp.
#include "orbx_internal.cuh"

namespace orbx {

__device__ __forceinline__ int reflect101_dev(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// Level 0: copy + border.  grid = (ceil(words/128), rows_alloc, frames)
__global__ void __launch_bounds__(128) pyr_level0_kernel(const __grid_constant__ FrameGeom fg, Workspace ws,
                                                         const uint8_t* __restrict__ images, size_t frame_stride,
                                                         size_t in_pitch)
{
    const LevelGeom& g = fg.L[0];
    const int word = blockIdx.x * blockDim.x + threadIdx.x;
    const int brow = blockIdx.y;              // bordered row 0..h+37
    const int frame = blockIdx.z;
    if (word * 4 >= g.pitch) return;
    const int sy = reflect101_dev(brow - kEdge, g.h);
    const uint8_t* src = images + (size_t)frame * frame_stride + (size_t)sy * in_pitch;
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = word * 4 + j;           // buffer column
        const int x = c - kXPad;              // interior x
        uint32_t v = 0;
        if (x >= -kEdge && x < g.w + kEdge) v = __ldg(src + reflect101_dev(x, g.w));
        out |= v << (8 * j);
    }
    uint8_t* dst = ws.pyr + g.pyr_off + (size_t)frame * g.pyr_frame_stride + (size_t)brow * g.pitch;
    reinterpret_cast<uint32_t*>(dst)[word] = out;
}

// Level l >= 1 from level l-1.  grid = (ceil(words/128), rows_alloc, frames)
__global__ void __launch_bounds__(128) pyr_resize_kernel(const __grid_constant__ FrameGeom fg, Workspace ws, int level)
{
    const LevelGeom& g = fg.L[level];
    const LevelGeom& p = fg.L[level - 1];
    const int word = blockIdx.x * blockDim.x + threadIdx.x;
    const int brow = blockIdx.y;
    const int frame = blockIdx.z;
    if (word * 4 >= g.pitch) return;
    const uint2 yt = __ldg(g.ytab + brow);
    const int sy0 = yt.x & 0xffff, sy1 = yt.x >> 16;
    const int b0 = (int)(yt.y & 0xffff), b1 = (int)(yt.y >> 16);
    const uint8_t* S = level_interior((const uint8_t*)ws.pyr, p, frame);
    const uint8_t* S0 = S + (size_t)sy0 * p.pitch;
    const uint8_t* S1 = S + (size_t)sy1 * p.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = word * 4 + j;
        const int bc = c - (kXPad - kEdge);   // bordered column 0..w+37
        uint32_t v = 0;
        if (bc >= 0 && bc < g.w + 2 * kEdge) {
            const uint2 xt = __ldg(g.xtab + bc);
            const int sx0 = xt.x & 0xffff, sx1 = xt.x >> 16;
            const int a0 = (int)(xt.y & 0xffff), a1 = (int)(xt.y >> 16);
            const int p00 = S0[sx0], p01 = S0[sx1], p10 = S1[sx0], p11 = S1[sx1];
            if (g.area2x) {
                v = (uint32_t)((p00 + p01 + p10 + p11 + 2) >> 2);
            } else {
                const int r0 = p00 * a0 + p01 * a1;
                const int r1 = p10 * a0 + p11 * a1;
                v = (uint32_t)((((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2) & 0xffu;
            }
        }
        out |= v << (8 * j);
    }
    uint8_t* dst = ws.pyr + g.pyr_off + (size_t)frame * g.pyr_frame_stride + (size_t)brow * g.pitch;
    reinterpret_cast<uint32_t*>(dst)[word] = out;
}

cudaError_t launch_pyramid(const FrameGeom& fg, const Workspace& ws, const uint8_t* d_images, size_t frame_stride,
                           size_t pitch, int n_frames, cudaStream_t st)
{
    for (int l = 0; l < fg.nlevels; ++l) {
        const LevelGeom& g = fg.L[l];
        const int words = g.pitch / 4;
        dim3 grid((words + 127) / 128, g.rows_alloc, n_frames);
        if (l == 0) pyr_level0_kernel<<<grid, 128, 0, st>>>(fg, ws, d_images, frame_stride, pitch);
        else pyr_resize_kernel<<<grid, 128, 0, st>>>(fg, ws, l);
        count_launch();
    }
    return cudaGetLastError();
}

}  // namespace orbx
