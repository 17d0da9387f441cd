// Micro-benchmark: rank-1 update throughput vs. register footprint (NS independent 36-entry
// triangles per thread, all updated every round).
#include <cstdio>
#include <cuda_runtime.h>

template <int NS>
__global__ void __launch_bounds__(64) k(int iters, double *out, long long *cyc, const double *in) {
    double P[NS][36], g[8], d[8];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 36; ++i) P[s][i] = in[i] + threadIdx.x + s;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = in[40 + i]; d[i] = in[50 + i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = r; c < 8; ++c) { P[s][idx] = fma(-g[r], d[c], P[s][idx]); ++idx; }
#pragma unroll
            for (int r = 0; r < 8; ++r) { g[r] = fma(g[r], 1.0000001, 1e-12); d[r] = fma(d[r], 0.9999999, 1e-12); }
        }
    }
    long long t1 = clock64();
    double sum = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int i = 0; i < 36; ++i) sum += P[s][i];
    if (sum == 123.456) out[0] = sum;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NS>
void run(int warps) {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8); cudaMalloc(&in, 8 * 128);
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int iters = 20000 / NS;
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k<NS>);
    for (int rep = 0; rep < 2; ++rep) { k<NS><<<148 * (warps / 2 > 0 ? warps / 2 : 1), warps >= 2 ? 64 : 32>>>(iters, out, cyc, in); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("NS=%d regs=%3d warps/SM=%d : %.3f cycles per DFMA per warp\n", NS, fa.numRegs, warps, c / ((double)iters * NS * 52));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
    run<1>(4); run<2>(4); run<3>(4);
    run<1>(8); run<2>(8); run<3>(8);
    return 0;
}
