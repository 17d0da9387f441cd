// Micro-benchmark: rank-1 update body P -= g d^T (36 upper-triangle DFMAs) for one warp per
// SMSP, with the loop body unrolled U times to vary its code footprint (I-cache test).
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void k(int iters, double *out, long long *cyc, const double *in) {
    double P[36], g[8], d[8];
#pragma unroll
    for (int i = 0; i < 36; ++i) P[i] = in[i] + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = in[40 + i]; d[i] = in[50 + i]; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int idx = 0;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = r; c < 8; ++c) { P[idx] = fma(-g[r], d[c], P[idx]); ++idx; }
            // perturb g, d cheaply (1 DFMA each) so the next round is not CSE'd
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

template <int U>
void run(int warps) {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8); cudaMalloc(&in, 8 * 128);
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int iters = 20000 / U;
    for (int rep = 0; rep < 2; ++rep) { k<U><<<148, 32 * warps>>>(iters, out, cyc, in); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("unroll=%3d (body ~%5.1f KB) warps/SM=%d : %.3f cycles per DFMA per warp\n", U, U * 52 * 16 / 1024.0, warps,
           c / ((double)iters * U * 52));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
    run<1>(4); run<2>(4); run<4>(4); run<8>(4); run<12>(4); run<16>(4); run<24>(4); run<32>(4); run<48>(4);
    run<1>(8); run<8>(8); run<16>(8); run<32>(8);
    return 0;
}
