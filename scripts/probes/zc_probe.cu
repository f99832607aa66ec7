// Probe (GPU box): bandwidth of kernel stores into page-locked host memory (zero-copy over PCIe) against cudaMemcpyAsync D2H, for the
// block sizes PackItems writes (256-byte and 4 KiB runs, one warp per run).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a zc_probe.cu -o zc_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__global__ void copyRuns(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t runs, int vecPerRun) {
    const int lane = threadIdx.x & 31;
    for (size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < runs; r += (size_t)gridDim.x * (blockDim.x >> 5))
        for (int i = lane; i < vecPerRun; i += 32) dst[r * vecPerRun + i] = __ldg(src + r * vecPerRun + i);
}
int main() {
    const size_t bytes = 256u << 20;
    uint4 *dsrc, *hdst, *ddst;
    cudaMalloc(&dsrc, bytes); cudaMalloc(&ddst, bytes);
    cudaMemset(dsrc, 0x5a, bytes);
    cudaHostAlloc(&hdst, bytes, cudaHostAllocDefault);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); cudaMemcpyAsync(hdst, dsrc, bytes, cudaMemcpyDeviceToHost); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("cudaMemcpyAsync D2H 256 MiB: %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    }
    const int runBytes[2] = {256, 4096};
    for (int rb = 0; rb < 2; ++rb)
        for (int blocks : {16, 32, 64, 148, 296, 592, 1184}) {
            const int vec = runBytes[rb] / 16;
            const size_t runs = bytes / runBytes[rb];
            copyRuns<<<blocks, 256>>>(dsrc, hdst, runs, vec);
            cudaEventRecord(e0); copyRuns<<<blocks, 256>>>(dsrc, hdst, runs, vec); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            printf("kernel -> pinned host, %4d-byte runs, %4d blocks: %.3f ms  %.1f GB/s   (%s)\n", runBytes[rb], blocks, ms, bytes / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    cudaEventRecord(e0); copyRuns<<<592, 256>>>(dsrc, ddst, bytes / 256, 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("kernel -> device, 256-byte runs: %.3f ms  %.1f GB/s\n", ms, bytes / ms * 1e-6);
    // host check
    unsigned char* hb = (unsigned char*)hdst; size_t bad = 0; for (size_t i = 0; i < bytes; i += 4097) bad += hb[i] != 0x5a;
    printf("host bytes wrong: %zu\n", bad);
    return 0;
}
