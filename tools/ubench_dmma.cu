// Micro-benchmark: fp64 tensor-core MMA (mma.sync m8n8k4 f64, SASS DMMA) issue rate on sm_100a,
// alone and mixed with DFMA, to decide whether the n=32 products should use it.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE, int CH>
__global__ void k(int iters, double *out, long long *cyc) {
    double c0[CH], c1[CH], f[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; f[i] = 1.0 + i; }
    double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0 || MODE == 2) dmma(c0[i], c1[i], a, b);
                if (MODE == 1 || MODE == 2) f[i] = fma(f[i], a, b);
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += c0[i] + c1[i] + f[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE, int CH>
void run(int warps, const char *name) {
    double *out; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) { k<MODE, CH><<<148, 32 * warps>>>(iters, out, cyc); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per = c / ((double)iters * 8 * CH);
    printf("%-28s chains=%d warps/SM=%2d : %.2f cycles per loop slot per warp", name, CH, warps, per);
    if (MODE == 0) printf("  -> %.1f TFLOP/s chip (DMMA = 512 flop)", 512.0 * warps * 148 * 1.965e9 / per / 1e12);
    if (MODE == 1) printf("  -> %.1f TFLOP/s chip (DFMA = 64 flop)", 64.0 * warps * 148 * 1.965e9 / per / 1e12);
    if (MODE == 2) printf("  -> %.1f TFLOP/s chip (DMMA+DFMA = 576 flop)", 576.0 * warps * 148 * 1.965e9 / per / 1e12);
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 1>(4, "DMMA dependent chain"); run<0, 4>(4, "DMMA 4 chains"); run<0, 8>(4, "DMMA 8 chains");
    run<0, 8>(8, "DMMA 8 chains"); run<0, 8>(16, "DMMA 8 chains");
    run<1, 8>(4, "DFMA 8 chains"); run<1, 8>(8, "DFMA 8 chains");
    run<2, 8>(4, "DMMA+DFMA interleaved"); run<2, 8>(8, "DMMA+DFMA interleaved"); run<2, 8>(16, "DMMA+DFMA interleaved");
    return 0;
}
