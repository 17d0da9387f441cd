// Micro-benchmark: does a DFMA with three distinct 64-bit register operands issue at the
// same rate as one with a single register operand?  (register-file read bandwidth)
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int CH>
__global__ void k(int iters, double *out, long long *cyc, const double *in) {
    double a[CH], b[CH], c[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = in[i] + threadIdx.x; b[i] = in[CH + i]; c[i] = in[2 * CH + i]; }
    double bs = in[100], cs = in[101];
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) a[i] = fma(a[i], 1.0000001, 1e-9);      // 1 register operand
                if (MODE == 1) a[i] = fma(a[i], b[i], c[i]);            // 3 distinct register operands
                if (MODE == 2) a[i] = fma(a[i], bs, c[i]);              // shared B operand (reuse-able)
                if (MODE == 3) a[i] = fma(a[i], bs, cs);                // shared B and C
                if (MODE == 4) a[i] = a[i] * b[i];                      // DMUL 2 distinct
                if (MODE == 5) a[i] = fma(b[i], c[i], a[i]);            // accumulate form
                if (MODE == 6) a[i] = fma(bs, c[i], a[i]);              // rank-1 update form: shared A
            }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int MODE, int CH>
void run(int warps, const char *name) {
    double *out, *in; long long *cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8); cudaMalloc(&in, 8 * 128);
    double h[128]; for (int i = 0; i < 128; ++i) h[i] = 1.0 + 1e-7 * i;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) { k<MODE, CH><<<148, 32 * warps>>>(iters, out, cyc, in); cudaDeviceSynchronize(); }
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s chains=%2d warps/SM=%2d : %.2f cycles per instr per warp\n", name, CH, warps, c / ((double)iters * 8 * CH));
    cudaFree(out); cudaFree(cyc); cudaFree(in);
}

int main() {
    for (int w : {4, 8}) {
        if (w == 4) {
            run<0, 8>(4, "fma(a,imm,imm)"); run<1, 8>(4, "fma(a,b_i,c_i) 3 distinct"); run<2, 8>(4, "fma(a,bs,c_i) shared B");
            run<3, 8>(4, "fma(a,bs,cs) shared B,C"); run<4, 8>(4, "mul(a,b_i)"); run<5, 8>(4, "fma(b_i,c_i,a) accumulate");
            run<6, 8>(4, "fma(bs,c_i,a) shared A"); run<1, 16>(4, "fma(a,b_i,c_i) 3 distinct"); run<6, 16>(4, "fma(bs,c_i,a) shared A");
        } else {
            run<0, 8>(8, "fma(a,imm,imm)"); run<1, 8>(8, "fma(a,b_i,c_i) 3 distinct"); run<2, 8>(8, "fma(a,bs,c_i) shared B");
            run<4, 8>(8, "mul(a,b_i)"); run<6, 8>(8, "fma(bs,c_i,a) shared A");
        }
    }
    return 0;
}
