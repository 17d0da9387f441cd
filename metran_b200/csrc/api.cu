// Library-wide C-ABI helpers: version, error strings, deterministic reduction.
#include <string.h>

#include "mkf_common.cuh"

namespace {
thread_local char g_cuda_error[256] = "";
}

namespace mkf_host {
void set_cuda_error(cudaError_t e) {
    const char *s = cudaGetErrorString(e);
    strncpy(g_cuda_error, s ? s : "unknown", sizeof(g_cuda_error) - 1);
    g_cuda_error[sizeof(g_cuda_error) - 1] = 0;
}
}  // namespace mkf_host

extern "C" int mkf_version(void) { return MKF_VERSION; }

extern "C" const char *mkf_error_string(int code) {
    switch (code) {
        case MKF_OK: return "ok";
        case MKF_ERR_BAD_ARG: return "bad argument (null pointer, non-positive size or bad stride)";
        case MKF_ERR_UNSUPPORTED: return "shape not supported by the compiled kernels";
        case MKF_ERR_CUDA: return "CUDA runtime error (see mkf_last_cuda_error)";
        default: return "unknown error code";
    }
}

extern "C" const char *mkf_last_cuda_error(void) { return g_cuda_error; }

namespace {
// Fixed-shape tree: one CTA of 1024 threads strides over the input in a fixed order, then
// a shuffle/shared-memory tree.  Bit-reproducible run to run and independent of the SM
// count, unlike atomics, so the pooled objective fed to the all-reduce is deterministic.
__global__ void __launch_bounds__(1024) reduce_sum_kernel(const double *__restrict__ v, int64_t count,
                                                          double *__restrict__ out) {
    __shared__ double part[32];
    double s = 0.0, c = 0.0;
    for (int64_t i = threadIdx.x; i < count; i += 1024) {
        double x = v[i];
        double t = s + x;
        c += (fabs(s) >= fabs(x)) ? ((s - t) + x) : ((x - t) + s);
        s = t;
    }
    s += c;
    s = mkf::group_sum<32>(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = mkf::group_sum<32>(part[threadIdx.x]);
        if (threadIdx.x == 0) out[0] = t;
    }
}
}  // namespace

extern "C" int mkf_reduce_sum(const double *v, int64_t count, double *out, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_reduce_sum");
    if (!out || count < 0 || (count > 0 && !v)) return MKF_ERR_BAD_ARG;
    reduce_sum_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(v, count, out);
    return mkf_host::check_launch();
}

namespace {
// fp64 FMA pipe probe for the bench's roofline denominator: 8 independent DFMA chains per
// thread, `iters` rounds; result written so the compiler keeps the chain.
__global__ void __launch_bounds__(256) fp64_probe_kernel(int iters, double *__restrict__ out) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
           a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) out[0] = s;
}
}  // namespace

// Launches blocks x 256 threads, each doing iters*64 DFMAs (= iters*128 flops per thread).
extern "C" int mkf_probe_fp64(int32_t blocks, int32_t iters, double *scratch, void *stream) {
    if (blocks <= 0 || iters <= 0 || !scratch) return MKF_ERR_BAD_ARG;
    fp64_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, scratch);
    return mkf_host::check_launch();
}
