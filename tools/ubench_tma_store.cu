// Micro-benchmark for the history stores of the thread-per-model filter (K1H, filter_dfm.cuh):
// every lane owns one model and produces, per time step, TWO records (predicted / filtered) of
// 8 + 64 doubles that go to four model-major arrays  mean[B][T][8], cov[B][T][64]  (twice), plus
// three per-step scalars (sigma, detf f64; cnt i32) at [B][T].
//   mode 0  legacy: records transposed through shared memory [element][33], 256-byte STG pieces
//   mode 1  TMA tensor stores: each lane writes its own record with STS.128 into a swizzled
//           staging box, lane 0 issues ONE cp.async.bulk.tensor.3d per array and emit
//           (cov: dims {16, B, 4T}, box {16, 32, 4}, SWIZZLE_128B; mean: dims {8, B, T},
//            box {8, 32, 1}, SWIZZLE_64B) -- SASS UTMASTG
//   mode 2  per-lane 1-D bulk copies (cp.async.bulk.global.shared::cta) from padded rows
// NF = number of filler DFMAs per step (the filter's arithmetic), SC = write the scalars too.
// Reports GB/s of payload.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    return (EncodeTiledFn)p;
}

__device__ __forceinline__ unsigned su32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

struct Args {
    double *pm, *pc, *fm, *fc, *sigma, *detf;
    int *cnt;
    int T, B;
};

template <int NF>
__device__ __forceinline__ void filler(double *v) {
#pragma unroll
    for (int r = 0; r < NF / 72; ++r)
#pragma unroll
        for (int e = 0; e < 72; ++e) v[e] = fma(v[e], 1.0000001, 1e-9);
}

template <int NF, int SC>
__global__ void __launch_bounds__(32, 6) k_legacy(Args a) {
    __shared__ double st[72 * 33];
    const int lane = threadIdx.x, b0 = blockIdx.x * 32, b = b0 + lane;
    double v[72];
#pragma unroll
    for (int e = 0; e < 72; ++e) v[e] = b + e;
    for (int t = 0; t < a.T; ++t) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            filler<NF / 2>(v);
            double *gm = ph ? a.fm : a.pm, *gc = ph ? a.fc : a.pc;
#pragma unroll
            for (int e = 0; e < 72; ++e) st[e * 33 + lane] = v[e];
            __syncwarp();
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
                double *dst = gc + ((size_t)(b0 + rr) * a.T + t) * 64;
                dst[lane] = st[(8 + lane) * 33 + rr];
                dst[lane + 32] = st[(8 + lane + 32) * 33 + rr];
            }
#pragma unroll
            for (int idx = lane; idx < 256; idx += 32) {
                const int rr = idx >> 3, e = idx & 7;
                gm[((size_t)(b0 + rr) * a.T + t) * 8 + e] = st[e * 33 + rr];
            }
            __syncwarp();
        }
        if (SC) {
            a.sigma[(size_t)b * a.T + t] = v[0];
            a.detf[(size_t)b * a.T + t] = v[1];
            a.cnt[(size_t)b * a.T + t] = 5;
        }
    }
}

// staging buffer of one emit: cov box [4 lines][32 models][128 B] (SWIZZLE_128B), mean box [32][64 B] (SWIZZLE_64B)
constexpr int kCovBytes = 4 * 32 * 128, kMeanBytes = 32 * 64, kBufBytes = kCovBytes + kMeanBytes;

template <int NF, int SC>
__global__ void __launch_bounds__(32, 6) k_tma(Args a, const __grid_constant__ CUtensorMap mpc,
                                               const __grid_constant__ CUtensorMap mpm,
                                               const __grid_constant__ CUtensorMap mfc,
                                               const __grid_constant__ CUtensorMap mfm) {
    extern __shared__ unsigned char raw[];
    const unsigned base = (su32(raw) + 1023u) & ~1023u;
    const int lane = threadIdx.x, b0 = blockIdx.x * 32, b = b0 + lane;
    double v[72];
#pragma unroll
    for (int e = 0; e < 72; ++e) v[e] = b + e;
    for (int t = 0; t < a.T; ++t) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            filler<NF / 2>(v);
            const unsigned buf = base + ph * kBufBytes;   // kBufBytes is a multiple of 1024
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
            // cov: element e = 16 * j + 2 * c (+1): line j, 16-byte chunk c; physical chunk c ^ (lane & 7)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const unsigned addr = buf + (unsigned)(j * 32 + lane) * 128u + (unsigned)((c ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[8 + 16 * j + 2 * c]),
                                 "d"(v[8 + 16 * j + 2 * c + 1])
                                 : "memory");
                }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const unsigned addr = buf + kCovBytes + (unsigned)lane * 64u + (unsigned)((c ^ ((lane >> 1) & 3)) << 4);
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[2 * c]), "d"(v[2 * c + 1]) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                const CUtensorMap *mc = ph ? &mfc : &mpc, *mm = ph ? &mfm : &mpm;
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(mc),
                             "r"(0), "r"(b0), "r"(4 * t), "r"(buf)
                             : "memory");
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(mm),
                             "r"(0), "r"(b0), "r"(t), "r"(buf + kCovBytes)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (SC) {
            a.sigma[(size_t)b * a.T + t] = v[0];
            a.detf[(size_t)b * a.T + t] = v[1];
            a.cnt[(size_t)b * a.T + t] = 5;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 3: as mode 1 with TS time steps staged per store (cov pieces of TS * 512 B, mean pieces of TS * 64 B);
// MEANS = 0 leaves the mean arrays out
template <int NF, int SC, int TS, int MEANS>
__global__ void __launch_bounds__(32, 1) k_tma_ts(Args a, const __grid_constant__ CUtensorMap mpc,
                                                  const __grid_constant__ CUtensorMap mpm,
                                                  const __grid_constant__ CUtensorMap mfc,
                                                  const __grid_constant__ CUtensorMap mfm) {
    extern __shared__ unsigned char raw[];
    const unsigned base = (su32(raw) + 1023u) & ~1023u;
    const int lane = threadIdx.x, b0 = blockIdx.x * 32, b = b0 + lane;
    constexpr unsigned kCov = TS * kCovBytes, kBox = TS * kBufBytes;
    double v[72];
#pragma unroll
    for (int e = 0; e < 72; ++e) v[e] = b + e;
    for (int t = 0; t < a.T; ++t) {
        const int u = t % TS;
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            filler<NF / 2>(v);
            const unsigned buf = base + ph * kBox;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const unsigned addr = buf + (unsigned)((u * 4 + j) * 32 + lane) * 128u + (unsigned)((c ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[8 + 16 * j + 2 * c]),
                                 "d"(v[8 + 16 * j + 2 * c + 1])
                                 : "memory");
                }
            if (MEANS) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const unsigned addr = buf + kCov + (unsigned)(u * 32 + lane) * 64u + (unsigned)((c ^ ((lane >> 1) & 3)) << 4);
                    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v[2 * c]), "d"(v[2 * c + 1]) : "memory");
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                if (u == TS - 1) {
                    const CUtensorMap *mc = ph ? &mfc : &mpc, *mm = ph ? &mfm : &mpm;
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(mc),
                                 "r"(0), "r"(b0), "r"(4 * (t - u)), "r"(buf)
                                 : "memory");
                    if (MEANS)
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(mm),
                                     "r"(0), "r"(b0), "r"(t - u), "r"(buf + kCov)
                                     : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (SC) {
            a.sigma[(size_t)b * a.T + t] = v[0];
            a.detf[(size_t)b * a.T + t] = v[1];
            a.cnt[(size_t)b * a.T + t] = 5;
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// mode 2: padded rows [32][528 B] (+ means [32][80 B]), one 1-D bulk copy per lane and array
constexpr int kRowC = 528, kRowM = 80, kBuf2 = 32 * (kRowC + kRowM);
template <int NF, int SC>
__global__ void __launch_bounds__(32, 6) k_bulk(Args a) {
    extern __shared__ unsigned char raw[];
    const unsigned base = (su32(raw) + 127u) & ~127u;
    const int lane = threadIdx.x, b0 = blockIdx.x * 32, b = b0 + lane;
    double v[72];
#pragma unroll
    for (int e = 0; e < 72; ++e) v[e] = b + e;
    for (int t = 0; t < a.T; ++t) {
#pragma unroll
        for (int ph = 0; ph < 2; ++ph) {
            filler<NF / 2>(v);
            const unsigned buf = base + ph * kBuf2;
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // every lane waits for its own copies
            const unsigned rc = buf + lane * kRowC, rm = buf + 32 * kRowC + lane * kRowM;
#pragma unroll
            for (int c = 0; c < 32; ++c)
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rc + 16u * c), "d"(v[8 + 2 * c]), "d"(v[8 + 2 * c + 1]) : "memory");
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rm + 16u * c), "d"(v[2 * c]), "d"(v[2 * c + 1]) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            double *gm = (ph ? a.fm : a.pm) + ((size_t)b * a.T + t) * 8, *gc = (ph ? a.fc : a.pc) + ((size_t)b * a.T + t) * 64;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 512;" ::"l"(gc), "r"(rc) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 64;" ::"l"(gm), "r"(rm) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (SC) {
            a.sigma[(size_t)b * a.T + t] = v[0];
            a.detf[(size_t)b * a.T + t] = v[1];
            a.cnt[(size_t)b * a.T + t] = 5;
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static bool make3d(EncodeTiledFn fn, CUtensorMap *m, double *p, int inner, int B, int T, int lines, CUtensorMapSwizzle sw, int ts = 1) {
    // array [B][T][lines * inner] doubles viewed as dims {inner, B, T * lines}
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)B, (cuuint64_t)T * lines};
    cuuint64_t strides[2] = {(cuuint64_t)T * lines * inner * 8, (cuuint64_t)inner * 8};
    cuuint32_t box[3] = {(cuuint32_t)inner, 32, (cuuint32_t)(lines * ts)};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode failed: %d\n", (int)r);
    return r == CUDA_SUCCESS;
}

template <int NF, int SC>
static void run(const char *name, int mode, Args a, CUtensorMap *maps, const double *host_check) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const double bytes = (double)a.B * a.T * (2.0 * 72 * 8 + (SC ? 20 : 0));
    float best = 1e9f;
    const size_t sm1 = 2 * kBufBytes + 1024, sm2 = 2 * kBuf2 + 128;
    cudaFuncSetAttribute(k_tma<NF, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    cudaFuncSetAttribute(k_bulk<NF, SC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        if (mode == 0) k_legacy<NF, SC><<<a.B / 32, 32>>>(a);
        if (mode == 1) k_tma<NF, SC><<<a.B / 32, 32, sm1>>>(a, maps[0], maps[1], maps[2], maps[3]);
        if (mode == 2) k_bulk<NF, SC><<<a.B / 32, 32, sm2>>>(a);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    // spot check: model 37, step 3, filtered cov / mean against the legacy kernel's values
    double got[72];
    cudaMemcpy(got + 8, a.fc + ((size_t)37 * a.T + 3) * 64, 64 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(got, a.fm + ((size_t)37 * a.T + 3) * 8, 8 * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    if (host_check)
        for (int e = 0; e < 72; ++e) bad += got[e] != host_check[e];
    printf("%-44s NF=%4d SC=%d  %.3f ms  %6.0f GB/s  mismatches %d  %s\n", name, NF, SC, best, bytes / best / 1e6, bad,
           cudaGetErrorString(cudaGetLastError()));
}

template <int NF, int TS, int MEANS>
static void run_ts(const char *name, Args a, EncodeTiledFn fn, int ctas_per_sm_hint) {
    CUtensorMap maps[4];
    bool ok = make3d(fn, &maps[0], a.pc, 16, a.B, a.T, 4, CU_TENSOR_MAP_SWIZZLE_128B, TS) &&
              make3d(fn, &maps[1], a.pm, 8, a.B, a.T, 1, CU_TENSOR_MAP_SWIZZLE_64B, TS) &&
              make3d(fn, &maps[2], a.fc, 16, a.B, a.T, 4, CU_TENSOR_MAP_SWIZZLE_128B, TS) &&
              make3d(fn, &maps[3], a.fm, 8, a.B, a.T, 1, CU_TENSOR_MAP_SWIZZLE_64B, TS);
    if (!ok) return;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const size_t sm = 2 * TS * kBufBytes + 1024;
    cudaFuncSetAttribute(k_tma_ts<NF, 0, TS, MEANS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const double bytes = (double)a.B * a.T * 2.0 * (64 + (MEANS ? 8 : 0)) * 8;
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k_tma_ts<NF, 0, TS, MEANS><<<a.B / 32, 32, sm>>>(a, maps[0], maps[1], maps[2], maps[3]);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tma_ts<NF, 0, TS, MEANS>, 32, sm);
    printf("%-44s NF=%4d TS=%d means=%d CTAs/SM=%d  %.3f ms  %6.0f GB/s  %s\n", name, NF, TS, MEANS, nb, best, bytes / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    const int B = 37888, T = 160;
    Args a;
    a.B = B;
    a.T = T;
    cudaMalloc(&a.pm, (size_t)B * T * 8 * 8);
    cudaMalloc(&a.fm, (size_t)B * T * 8 * 8);
    cudaMalloc(&a.pc, (size_t)B * T * 64 * 8);
    cudaMalloc(&a.fc, (size_t)B * T * 64 * 8);
    cudaMalloc(&a.sigma, (size_t)B * T * 8);
    cudaMalloc(&a.detf, (size_t)B * T * 8);
    cudaMalloc(&a.cnt, (size_t)B * T * 4);
    EncodeTiledFn fn = encode_fn();
    CUtensorMap maps[4];
    bool ok = fn != nullptr;
    ok = ok && make3d(fn, &maps[0], a.pc, 16, B, T, 4, CU_TENSOR_MAP_SWIZZLE_128B);
    ok = ok && make3d(fn, &maps[1], a.pm, 8, B, T, 1, CU_TENSOR_MAP_SWIZZLE_64B);
    ok = ok && make3d(fn, &maps[2], a.fc, 16, B, T, 4, CU_TENSOR_MAP_SWIZZLE_128B);
    ok = ok && make3d(fn, &maps[3], a.fm, 8, B, T, 1, CU_TENSOR_MAP_SWIZZLE_64B);
    printf("tensor maps: %s\n", ok ? "ok" : "FAILED");

    double ref0[72], ref500[72];
    // reference values from the legacy kernel (NF = 0 and NF = 576)
    run<0, 0>("legacy transposition (256-B STG pieces)", 0, a, maps, nullptr);
    cudaMemcpy(ref0 + 8, a.fc + ((size_t)37 * T + 3) * 64, 64 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref0, a.fm + ((size_t)37 * T + 3) * 8, 8 * 8, cudaMemcpyDeviceToHost);
    run<576, 0>("legacy transposition", 0, a, maps, nullptr);
    cudaMemcpy(ref500 + 8, a.fc + ((size_t)37 * T + 3) * 64, 64 * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(ref500, a.fm + ((size_t)37 * T + 3) * 8, 8 * 8, cudaMemcpyDeviceToHost);
    run<576, 1>("legacy transposition + scalars", 0, a, maps, ref500);
    if (ok) {
        cudaMemset(a.fc, 0, (size_t)B * T * 64 * 8);
        cudaMemset(a.fm, 0, (size_t)B * T * 8 * 8);
        run<0, 0>("TMA tensor stores (swizzled staging)", 1, a, maps, ref0);
        run<576, 0>("TMA tensor stores", 1, a, maps, ref500);
        run<576, 1>("TMA tensor stores + scalars", 1, a, maps, ref500);
    }
    if (ok) {
        run_ts<576, 1, 1>("TMA stores, TS steps per store", a, fn, 0);
        run_ts<576, 1, 0>("TMA stores, TS steps per store", a, fn, 0);
        run_ts<576, 2, 1>("TMA stores, TS steps per store", a, fn, 0);
        run_ts<576, 2, 0>("TMA stores, TS steps per store", a, fn, 0);
        run_ts<576, 4, 1>("TMA stores, TS steps per store", a, fn, 0);
        run_ts<2304, 2, 1>("TMA stores, TS steps per store", a, fn, 0);
        run_ts<2304, 1, 1>("TMA stores, TS steps per store", a, fn, 0);
    }
    cudaMemset(a.fc, 0, (size_t)B * T * 64 * 8);
    cudaMemset(a.fm, 0, (size_t)B * T * 8 * 8);
    run<0, 0>("per-lane 1-D bulk copies (padded rows)", 2, a, maps, ref0);
    run<576, 0>("per-lane 1-D bulk copies", 2, a, maps, ref500);
    run<576, 1>("per-lane 1-D bulk copies + scalars", 2, a, maps, ref500);
    return 0;
}
