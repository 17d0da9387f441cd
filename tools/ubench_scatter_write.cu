// Micro-benchmark: HBM write bandwidth when every warp store covers one contiguous chunk of
// CHUNK bytes but consecutive chunks are STRIDE bytes apart (the history layout
// [model][time][n*n]: at a fixed time step the 32 models of a warp are T*n*n*8 bytes apart).
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(double *buf, long long chunk_doubles, long long stride_doubles, long long nchunks, int steps) {
    // warp w handles chunks w, w + nwarps, ...; inside a chunk lanes write consecutive doubles
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (int s = 0; s < steps; ++s)
        for (long long c = warp; c < nchunks; c += nwarps) {
            double *dst = buf + c * stride_doubles + (long long)s * chunk_doubles;
            for (long long e = lane; e < chunk_doubles; e += 32) dst[e] = (double)e;
        }
}

int main() {
    const long long total = 32LL << 30;  // 32 GB buffer
    double *buf;
    if (cudaMalloc(&buf, total) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int steps = 200;
    for (long long chunk : {512LL, 1024LL, 2048LL, 4096LL}) {
        for (long long stride : {chunk * 200, 1LL << 20}) {
            if (stride < chunk * steps) stride = chunk * steps;
            long long nchunks = total / stride;
            if (nchunks > 37888) nchunks = 37888;          // 37888 models
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                k<<<148 * 8, 256>>>(buf, chunk / 8, stride / 8, nchunks, steps);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("chunk %5lld B, stride %8lld B, %lld chunks x %d steps: %.3f ms -> %.0f GB/s\n", chunk, stride, nchunks, steps,
                   ms, nchunks * chunk * steps / ms / 1e6);
        }
    }
    // contiguous reference
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        k<<<148 * 8, 256>>>(buf, (4LL << 30) / 8, (4LL << 30) / 8, 1, 1);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("contiguous 4 GB by one warp loop (reference, not parallel): %.3f ms\n", ms);
    return 0;
}
