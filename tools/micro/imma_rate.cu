// Micro-benchmark: issue rate of the legacy warp-level int8 tensor-core instruction on this GPU
// (mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 -> SASS IMMA.16832.U8.U8), the instruction a Hamming matcher on {0,1}-expanded
// descriptors would use (distance = popc(a) + popc(b) - 2 * dot(a, b)).  Prints dense int8 TOPS and the 256-bit compares/s that rate
// could carry (8 instructions per 16 x 8 tile of compares).  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_rate imma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(int iters, int* out)
{
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3u, a2 = threadIdx.x * 5u, a3 = threadIdx.x * 7u, b0 = blockIdx.x, b1 = blockIdx.x * 11u;
    int c[8][4] = {};
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[j][0]), "+r"(c[j][1]), "+r"(c[j][2]), "+r"(c[j][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0 + j), "r"(b1));
    }
    int s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int ctas = sms * 4, iters = 20000;
    int* out; cudaMalloc(&out, ctas * 256 * sizeof(int));
    k<<<ctas, 256>>>(100, out); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<<<ctas, 256>>>(iters, out); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas = (double)ctas * 8 /*warps*/ * iters * 8;
    const double ops = mmas * 16 * 8 * 32 * 2;
    printf("IMMA.16832.U8.U8: %.3f ms, %.3e mma/s, %.1f dense int8 TOPS, %.3e 256-bit compares/s if nothing else issued (%s)\n", ms, mmas / (ms * 1e-3), ops / (ms * 1e-3) / 1e12,
           mmas / (ms * 1e-3) * 16.0, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
