// Micro-benchmark: how fast can 1184 one-warp CTAs (8 per SM) stream y[B][T][5] (model-major, fp64)
// when lane l of a warp consumes row 32*w + l, i.e. the access pattern of the thread-per-model
// filter kernels?  Compute is one DADD per element.  Schemes:
//   tma   TT x S [pf P]  2-D TMA tiles of 32 rows x TT steps, S-stage ring; optional L2 tensor
//                        prefetch of 32 rows x P steps, two super-tiles ahead
//   ldgsts TT x S        per-lane 16-byte cp.async (LDGSTS) into a private row of the ring
//   ldg  D               per-lane 128-bit loads, D two-step groups prefetched in registers
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

constexpr int M = 5;
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    return (EncodeTiledFn)p;
}
static CUtensorMap make_map(const double *y, long long B, long long T, int box_rows, int box_elems, CUtensorMapL2promotion promo) {
    alignas(64) CUtensorMap map;
    memset(&map, 0, sizeof(map));
    cuuint64_t dims[2] = {(cuuint64_t)(T * M), (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)(T * M * 8)};
    cuuint32_t box[2] = {(cuuint32_t)box_elems, (cuuint32_t)box_rows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (r != CUDA_SUCCESS) printf("tensor map failed %d\n", (int)r);
    return map;
}
__device__ __forceinline__ unsigned su32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

template <int TT, int S, int PF>
__global__ void __launch_bounds__(32, 8) k_tma(const __grid_constant__ CUtensorMap map, const __grid_constant__ CUtensorMap pmap,
                                               int T, double *out) {
    constexpr int PITCH = TT * M;
    extern __shared__ unsigned char raw[];
    unsigned char *al = raw + ((128u - (su32(raw) & 127u)) & 127u);
    const unsigned ring = su32(al), bars = ring + S * 32 * PITCH * 8;
    const int lane = threadIdx.x, b0 = blockIdx.x * 32;
    if (lane == 0) {
        for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bars + 8 * s) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const int ntiles = T / TT;
    auto issue = [&](int s, int tile) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bars + 8 * s), "r"(32 * PITCH * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         ring + s * 32 * PITCH * 8),
                     "l"(&map), "r"(tile * PITCH), "r"(b0), "r"(bars + 8 * s)
                     : "memory");
    };
    auto prefetch = [&](int step) {
        if (step < T)
            asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(&pmap), "r"(step * M), "r"(b0) : "memory");
    };
    if (elect_one()) {
        if (PF > 0) {
            prefetch(0);
            prefetch(PF);
        }
        for (int s = 0; s < S; ++s) issue(s, s);
    }
    double acc = 0.0;
    unsigned parity = 0;
    for (int tile = 0; tile < ntiles; tile += S) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int tl = tile + s;
            if (tl < ntiles) {
                unsigned ok;
                do {
                    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                                 : "=r"(ok)
                                 : "r"(bars + 8 * s), "r"(parity)
                                 : "memory");
                } while (!ok);
                double v[PITCH];
#pragma unroll
                for (int e = 0; e < PITCH / 2; ++e)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                                 : "=d"(v[2 * e]), "=d"(v[2 * e + 1])
                                 : "r"(ring + s * 32 * PITCH * 8 + lane * PITCH * 8 + 16 * e));
#pragma unroll
                for (int e = 0; e < PITCH; ++e) acc += v[e];
                // every lane's loads are consumed before the vote: orders them before the refill
                const bool any = __any_sync(0xffffffffu, acc == 123.456);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                if (elect_one()) {
                    if (tl + S < ntiles && !any) issue(s, tl + S);
                    if (PF > 0 && (tl * TT) % PF == 0) prefetch(tl * TT + 2 * PF);
                }
            }
        }
        parity ^= 1u;
    }
    out[b0 + lane] = acc;
}

template <int TT, int S>
__global__ void __launch_bounds__(32, 8) k_ldgsts(const double *y, int T, double *out) {
    constexpr int PITCH = TT * M;            // doubles per row and tile (even)
    extern __shared__ unsigned char raw[];
    unsigned char *al = raw + ((128u - (su32(raw) & 127u)) & 127u);
    const unsigned ring = su32(al);
    const int lane = threadIdx.x, b0 = blockIdx.x * 32;
    const double *row = y + (size_t)(b0 + lane) * T * M;
    const int ntiles = T / TT;
    auto issue = [&](int s, int tile) {
        const unsigned dst = ring + s * 32 * PITCH * 8 + lane * PITCH * 8;
        const double *src = row + (size_t)tile * PITCH;
#pragma unroll
        for (int e = 0; e < PITCH / 2; ++e)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * e), "l"(src + 2 * e) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int s = 0; s < S - 1; ++s) issue(s, s);
    double acc = 0.0;
    for (int tile = 0; tile < ntiles; tile += S) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const int tl = tile + s;
            if (tl < ntiles) {
                if (tl + S - 1 < ntiles) issue((s + S - 1) % S, tl + S - 1);
                else asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
                double v[PITCH];
#pragma unroll
                for (int e = 0; e < PITCH / 2; ++e)
                    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];"
                                 : "=d"(v[2 * e]), "=d"(v[2 * e + 1])
                                 : "r"(ring + s * 32 * PITCH * 8 + lane * PITCH * 8 + 16 * e));
#pragma unroll
                for (int e = 0; e < PITCH; ++e) acc += v[e];
            }
        }
    }
    out[b0 + lane] = acc;
}

template <int D>   // D groups of 2 steps (80 B = 5 x LDG.128 per lane) in flight in registers
__global__ void __launch_bounds__(32, 8) k_ldg(const double *y, int T, double *out) {
    const int lane = threadIdx.x, b0 = blockIdx.x * 32;
    const double2 *row = reinterpret_cast<const double2 *>(y + (size_t)(b0 + lane) * T * M);
    const int ngroups = T / 2;
    double2 buf[D][5];
#pragma unroll
    for (int d = 0; d < D; ++d)
#pragma unroll
        for (int e = 0; e < 5; ++e) buf[d][e] = row[d * 5 + e];
    double acc = 0.0;
    for (int g = 0; g < ngroups; g += D) {
#pragma unroll
        for (int d = 0; d < D; ++d) {
            double2 cur[5];
#pragma unroll
            for (int e = 0; e < 5; ++e) cur[e] = buf[d][e];
            if (g + d + D < ngroups) {
#pragma unroll
                for (int e = 0; e < 5; ++e) buf[d][e] = row[(size_t)(g + d + D) * 5 + e];
            }
#pragma unroll
            for (int e = 0; e < 5; ++e) acc += cur[e].x + cur[e].y;
        }
    }
    out[b0 + lane] = acc;
}

static cudaEvent_t e0, e1;
template <class F>
static void timeit(const char *name, long long bytes, F launch) {
    float best = 1e9f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-28s %.3f ms  %.0f GB/s %s\n", name, best, bytes / best / 1e6, e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

template <int TT, int S, int PF>
static void run_tma(const double *y, long long B, int T, double *out, CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    const char *tag = "") {
    CUtensorMap map = make_map(y, B, T, 32, TT * M, promo);
    CUtensorMap pmap = make_map(y, B, T, 32, (PF > 0 ? PF : TT) * M, promo);
    const int smem = 128 + S * 32 * TT * M * 8 + S * 8;
    cudaFuncSetAttribute(k_tma<TT, S, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    char name[64];
    snprintf(name, sizeof name, "tma %dx%d pf%d%s", TT, S, PF, tag);
    timeit(name, B * T * M * 8LL, [&] { k_tma<TT, S, PF><<<(unsigned)(B / 32), 32, smem>>>(map, pmap, T, out); });
}
template <int TT, int S>
static void run_ldgsts(const double *y, long long B, int T, double *out) {
    const int smem = 128 + S * 32 * TT * M * 8;
    cudaFuncSetAttribute(k_ldgsts<TT, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    char name[64];
    snprintf(name, sizeof name, "ldgsts %dx%d", TT, S);
    timeit(name, B * T * M * 8LL, [&] { k_ldgsts<TT, S><<<(unsigned)(B / 32), 32, smem>>>(y, T, out); });
}
template <int D>
static void run_ldg(const double *y, long long B, int T, double *out) {
    char name[64];
    snprintf(name, sizeof name, "ldg.128 depth %d", D);
    timeit(name, B * T * M * 8LL, [&] { k_ldg<D><<<(unsigned)(B / 32), 32>>>(y, T, out); });
}

int main(int argc, char **argv) {
    const long long B = argc > 1 ? atoll(argv[1]) : 37888;
    const int T = argc > 2 ? atoi(argv[2]) : 4800;      // multiple of every tile length used
    double *y, *out;
    cudaMalloc(&y, B * T * M * 8LL);
    cudaMalloc(&out, B * 8);
    cudaMemset(y, 0, B * T * M * 8LL);
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("B=%lld T=%d (%.2f GB), 1 warp per CTA, %lld CTAs\n", B, T, B * T * M * 8 / 1e9, B / 32);
    run_tma<2, 4, 0>(y, B, T, out);
    run_tma<2, 8, 0>(y, B, T, out);
    run_tma<4, 3, 0>(y, B, T, out);
    run_tma<4, 5, 0>(y, B, T, out);
    run_tma<6, 2, 0>(y, B, T, out);
    run_tma<6, 3, 0>(y, B, T, out);
    run_tma<8, 2, 0>(y, B, T, out);
    run_tma<8, 3, 0>(y, B, T, out);
    run_tma<12, 2, 0>(y, B, T, out);
    run_tma<6, 3, 0>(y, B, T, out, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, " promo256");
    run_tma<6, 3, 0>(y, B, T, out, CU_TENSOR_MAP_L2_PROMOTION_NONE, " promo0");
    run_tma<2, 4, 24>(y, B, T, out);
    run_tma<2, 4, 48>(y, B, T, out);
    run_tma<2, 8, 24>(y, B, T, out);
    run_tma<4, 3, 24>(y, B, T, out);
    run_tma<4, 3, 48>(y, B, T, out);
    run_tma<6, 3, 24>(y, B, T, out);
    run_tma<6, 3, 48>(y, B, T, out);
    run_ldgsts<2, 4>(y, B, T, out);
    run_ldgsts<2, 8>(y, B, T, out);
    run_ldgsts<4, 4>(y, B, T, out);
    run_ldgsts<6, 3>(y, B, T, out);
    run_ldgsts<8, 3>(y, B, T, out);
    run_ldg<2>(y, B, T, out);
    run_ldg<4>(y, B, T, out);
    run_ldg<8>(y, B, T, out);
    return 0;
}
