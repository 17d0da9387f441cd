// Micro-benchmark: rank-1 update P -= g d^T in row order vs. diagonal order (no operand shared
// between consecutive DFMAs -> no .reuse), one warp per SMSP and two.
#include <cstdio>
#include <cuda_runtime.h>

template <int ORDER>
__global__ void k(int iters, double *out, long long *cyc, const double *in) {
    double P[36], g[8], d[8];
#pragma unroll
    for (int i = 0; i < 36; ++i) P[i] = in[i] + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = in[40 + i]; d[i] = in[50 + i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ORDER == 0) {
                int idx = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = r; c < 8; ++c) { P[idx] = fma(-g[r], d[c], P[idx]); ++idx; }
            } else {
#pragma unroll
                for (int off = 0; off < 8; ++off)
#pragma unroll
                    for (int r = 0; r + off < 8; ++r) {
                        const int c = r + off;
                        const int idx = r * 8 - (r * (r - 1)) / 2 + (c - r);
                        P[idx] = fma(-g[r], d[c], P[idx]);
                    }
            }
#pragma unroll
            for (int r = 0; r < 8; ++r) { g[r] = fma(g[r], 1.0000001, 1e-12); d[r] = fma(d[r], 0.9999999, 1e-12); }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 36; ++i) s += P[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ORDER>
void run(int warps) {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8); cudaMalloc(&in, 8 * 128);
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int iters = 5000;
    for (int rep = 0; rep < 2; ++rep) { k<ORDER><<<148, 32 * warps>>>(iters, out, cyc, in); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("order=%s warps/SM=%d : %.3f cycles per DFMA per warp\n", ORDER ? "diagonal" : "row", warps, c / ((double)iters * 4 * 52));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() { run<0>(4); run<1>(4); run<0>(8); run<1>(8); return 0; }
