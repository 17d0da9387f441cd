// Micro-benchmark: history layout [model][time][64 doubles] written by ONE THREAD PER MODEL, i.e. every
// lane of a warp stores its own contiguous 512-byte record per time step (model stride T*512 B).
//   v2   : 32 x st.global.v2.f64 per record (16-byte pieces; a warp instruction touches 32 half sectors)
//   v2cs : the same with the streaming (evict-first) hint
//   smem : records transposed through shared memory so that each warp store covers 256 contiguous bytes
// Reports GB/s of payload; the question is whether L2 merges the 16-byte pieces into full sectors.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(32, 8) k(double *out, int T, int B) {
    const int lane = threadIdx.x, b = blockIdx.x * 32 + lane;
    if (b >= B) return;
    double *row = out + (size_t)b * T * 64;
    double v[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) v[e] = b + e;
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int e = 0; e < 64; ++e) v[e] = v[e] * 1.0000001 + 1e-9;   // a little arithmetic per step
        double *dst = row + (size_t)t * 64;
#pragma unroll
        for (int e = 0; e < 64; e += 2) {
            if (MODE == 0) asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(dst + e), "d"(v[e]), "d"(v[e + 1]) : "memory");
            if (MODE == 1) asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(dst + e), "d"(v[e]), "d"(v[e + 1]) : "memory");
        }
    }
}

__global__ void __launch_bounds__(32, 8) k_smem(double *out, int T, int B) {
    __shared__ double st[64 * 33];
    const int lane = threadIdx.x, b0 = blockIdx.x * 32;
    double v[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) v[e] = b0 + lane + e;
    for (int t = 0; t < T; ++t) {
#pragma unroll
        for (int e = 0; e < 64; ++e) {
            v[e] = v[e] * 1.0000001 + 1e-9;
            st[e * 33 + lane] = v[e];
        }
        __syncwarp();
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
            double *dst = out + ((size_t)(b0 + rr) * T + t) * 64;
            dst[lane] = st[lane * 33 + rr];
            dst[lane + 32] = st[(lane + 32) * 33 + rr];
        }
        __syncwarp();
    }
}

int main() {
    const int B = 37888, T = 300;
    double *out;
    cudaMalloc(&out, (size_t)B * T * 64 * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = (double)B * T * 64 * 8;
    for (int mode = 0; mode < 3; ++mode) {
        float best = 1e9f;
        for (int r = 0; r < 4; ++r) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<B / 32, 32>>>(out, T, B);
            if (mode == 1) k<1><<<B / 32, 32>>>(out, T, B);
            if (mode == 2) k_smem<<<B / 32, 32>>>(out, T, B);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        printf("%-28s %.3f ms  %.0f GB/s  %s\n", mode == 0 ? "per-thread st.v2.f64" : mode == 1 ? "per-thread st.cs.v2.f64" : "smem transposed 256-B stores",
               best, bytes / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
