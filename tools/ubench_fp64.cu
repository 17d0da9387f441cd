// Micro-benchmark: fp64 FMA issue rate / latency per warp on sm_100a (B200).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(int iters, double *out, long long *cyc) {
    double a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps_per_block, int blocks_per_sm) {
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 2000;
    k<CHAINS><<<148 * blocks_per_sm, 32 * warps_per_block>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    k<CHAINS><<<148 * blocks_per_sm, 32 * warps_per_block>>>(iters, out, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double n = (double)iters * 16 * CHAINS;
    printf("chains=%2d warps/block=%d blocks/SM=%d : %.2f cycles per DFMA per warp (%.2f cycles per chain step)\n",
           CHAINS, warps_per_block, blocks_per_sm, h / n, h / (double)(iters * 16));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    printf("--- one warp per SM (1 SMSP active): latency / single-warp issue rate\n");
    run<1>(1, 1); run<2>(1, 1); run<4>(1, 1); run<8>(1, 1); run<16>(1, 1);
    printf("--- 4 warps per SM (1 per SMSP)\n");
    run<1>(4, 1); run<4>(4, 1); run<8>(4, 1); run<16>(4, 1);
    printf("--- 8 warps per SM (2 per SMSP)\n");
    run<1>(8, 1); run<2>(8, 1); run<4>(8, 1); run<8>(8, 1); run<16>(8, 1);
    printf("--- 12 warps per SM (3 per SMSP)\n");
    run<2>(12, 1); run<4>(12, 1); run<8>(12, 1);
    printf("--- 16 warps per SM (4 per SMSP)\n");
    run<1>(16, 1); run<2>(16, 1); run<4>(16, 1); run<8>(16, 1);
    return 0;
}
