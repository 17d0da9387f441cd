// C-ABI entry points of the thread-per-model DFM filter (see filter_dfm.cuh).
#include <atomic>
#include <map>
#include <mutex>
#include <utility>

#include "filter_dfm.cuh"

namespace mkf_dfm {

// ---- hand-off scratch: grow-only, one buffer per (device, stream) ------------------------------
// Work enqueued on one stream is ordered, so consecutive calls on a stream can share one buffer;
// calls on different streams may overlap and get their own.  Growing frees the old buffer with
// cudaFree, which waits for the device: the work that used it has finished by then.
namespace {
std::mutex g_ws_mutex;
std::map<std::pair<int, cudaStream_t>, HandoffScratch> g_ws;
std::atomic<int> g_handoff{1};
std::atomic<int> g_profile{0};
std::atomic<int> g_hist_tma{1};
}  // namespace
bool hist_tma_enabled() { return g_hist_tma.load() != 0; }

HandoffScratch *handoff_scratch(cudaStream_t stream, size_t bytes) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    HandoffScratch &s = g_ws[std::make_pair(dev, stream)];
    s.profiled = false;
    if (s.bytes >= bytes) return &s;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
        (void)cudaGetLastError();
        return nullptr;            // no allocation inside a graph capture: single-kernel path
    }
    if (s.ptr) cudaFree(s.ptr);
    s.ptr = nullptr;
    s.bytes = 0;
    const size_t want = bytes + bytes / 4;
    if (cudaMalloc(&s.ptr, want) != cudaSuccess) {
        (void)cudaGetLastError();
        s.ptr = nullptr;
        return nullptr;
    }
    s.bytes = want;
    return &s;
}

bool profile_enabled() { return g_profile.load() != 0; }
bool profile_events(HandoffScratch *s) {
    if (!s->have_events) {
        for (int i = 0; i < 4; ++i)
            if (cudaEventCreate(&s->ev[i]) != cudaSuccess) {
                (void)cudaGetLastError();
                return false;
            }
        s->have_events = true;
    }
    return true;
}

bool handoff_enabled() { return g_handoff.load() != 0; }

// cuTensorMapEncodeTiled is a driver-API entry point; it is resolved through the runtime
// (cudaGetDriverEntryPoint) so the library does not link libcuda directly.
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

int make_y_tensor_map(CUtensorMap *map, const double *y, int64_t rows, int64_t row_elems, int64_t row_stride,
                      int box_rows, int box_elems, int l2_promotion_bytes) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return MKF_ERR_UNSUPPORTED;
    if (row_elems <= 0 || rows <= 0 || row_elems >= (1LL << 31) || rows >= (1LL << 31)) return MKF_ERR_UNSUPPORTED;  // int32 tile coordinates
    if ((row_stride * 8) % 16 != 0 || row_stride * 8 >= (1LL << 40) || row_stride < row_elems) return MKF_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_stride * 8};  // bytes between model rows
    cuuint32_t box[2] = {(cuuint32_t)box_elems, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(y), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    l2_promotion_bytes >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    return r == CUDA_SUCCESS ? MKF_OK : MKF_ERR_UNSUPPORTED;
}

int make_hist_map(CUtensorMap *map, double *p, int inner, int64_t B, int64_t T, int lines, int steps) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !p) return MKF_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) return MKF_ERR_UNSUPPORTED;
    if (B <= 0 || T <= 0 || B >= (1LL << 31) || T * lines >= (1LL << 31)) return MKF_ERR_UNSUPPORTED;
    const int64_t row_bytes = T * lines * inner * 8;   // one model's block
    if (row_bytes % 16 != 0 || row_bytes >= (1LL << 40)) return MKF_ERR_UNSUPPORTED;
    const CUtensorMapSwizzle sw = inner * 8 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : inner * 8 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : inner * 8 == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                    : CU_TENSOR_MAP_SWIZZLE_NONE;
    if (sw == CU_TENSOR_MAP_SWIZZLE_NONE) return MKF_ERR_UNSUPPORTED;
    // dims from the innermost: {inner doubles of one 'line', models, lines over time}; the model
    // stride is larger than the line stride on purpose: the box in shared memory is then
    // [line][model][inner], i.e. consecutive lanes (= models) are `inner * 8` bytes apart
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)B, (cuuint64_t)(T * lines)};
    cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)inner * 8};
    cuuint32_t box[3] = {(cuuint32_t)inner, 32, (cuuint32_t)(lines * steps)};
    if (lines * steps > 256) return MKF_ERR_UNSUPPORTED;
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, p, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    sw, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MKF_OK : MKF_ERR_UNSUPPORTED;
}

int make_scalar_map(CUtensorMap *map, void *p, int elem_bytes, int64_t B, int64_t T) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !p || (elem_bytes != 8 && elem_bytes != 4)) return MKF_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(p) & 15) != 0 || (T * elem_bytes) % 16 != 0) return MKF_ERR_UNSUPPORTED;
    if (B <= 0 || T <= 0 || B >= (1LL << 31) || T >= (1LL << 31)) return MKF_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[1] = {(cuuint64_t)T * elem_bytes};
    cuuint32_t box[2] = {16, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_INT32, 2, p, dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    elem_bytes == 8 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? MKF_OK : MKF_ERR_UNSUPPORTED;
}

template <> int launch_for_m<0>(int, const DfmArgs &, cudaStream_t);
template <> int launch_for_m<1>(int, const DfmArgs &, cudaStream_t);
template <> int launch_for_m<2>(int, const DfmArgs &, cudaStream_t);
template <> int launch_for_m<3>(int, const DfmArgs &, cudaStream_t);

static int run(int m, int k, DfmArgs &a, cudaStream_t s) {
    if (a.B < 0 || a.T < 0 || m <= 0 || k < 0) return MKF_ERR_BAD_ARG;
    if (!a.y || !a.mle || (k > 0 && !a.gamma)) return MKF_ERR_BAD_ARG;
    if (m > kMaxM || k > kMaxK) return MKF_ERR_UNSUPPORTED;
    if (a.B == 0) return MKF_OK;
    // TMA staging needs every model row 16-byte aligned (else: cp.async staging)
    a.use_tma = ((reinterpret_cast<uintptr_t>(a.y) & 15) == 0) && ((a.y_stride & 1) == 0);
    switch (k) {
        case 0: return launch_for_m<0>(m, a, s);
        case 1: return launch_for_m<1>(m, a, s);
        case 2: return launch_for_m<2>(m, a, s);
        case 3: return launch_for_m<3>(m, a, s);
    }
    return MKF_ERR_UNSUPPORTED;
}
}  // namespace mkf_dfm

extern "C" int mkf_dfm_loglik(int64_t B, int64_t T, int32_t m, int32_t k, const double *y, int64_t y_stride,
                              const double *phi, int64_t phi_stride, const double *q, int64_t q_stride,
                              const double *gamma, int64_t gamma_stride, const double *r, int64_t r_stride,
                              int32_t warmup, double *mle, int64_t *nobs, int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_dfm_loglik");
    if (!phi || !q) return MKF_ERR_BAD_ARG;
    mkf_dfm::DfmArgs a{};
    a.B = B; a.T = T; a.y = y; a.y_stride = y_stride;
    a.phi = phi; a.phi_stride = phi_stride; a.q = q; a.q_stride = q_stride;
    a.gamma = gamma; a.gamma_stride = gamma_stride; a.r = r; a.r_stride = r_stride;
    a.alpha = nullptr; a.dt = 1.0; a.warmup = warmup;
    a.mle = mle; a.nobs = nobs; a.status = status;
    return mkf_dfm::run(m, k, a, (cudaStream_t)stream);
}

extern "C" int mkf_dfm_loglik_alpha(int64_t B, int64_t T, int32_t m, int32_t k, const double *y,
                                    int64_t y_stride, const double *alpha, int64_t alpha_stride, double dt,
                                    const double *gamma, int64_t gamma_stride, int32_t warmup, double *mle,
                                    int64_t *nobs, int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_dfm_loglik_alpha");
    if (!alpha) return MKF_ERR_BAD_ARG;
    mkf_dfm::DfmArgs a{};
    a.B = B; a.T = T; a.y = y; a.y_stride = y_stride;
    a.gamma = gamma; a.gamma_stride = gamma_stride;
    a.alpha = alpha; a.alpha_stride = alpha_stride; a.dt = dt; a.warmup = warmup;
    a.mle = mle; a.nobs = nobs; a.status = status;
    return mkf_dfm::run(m, k, a, (cudaStream_t)stream);
}

extern "C" int mkf_dfm_loglik_compact(int64_t B, int64_t J, int32_t m, int32_t k, const double *yc, int64_t yc_stride,
                                      const int32_t *tidx, int64_t tidx_stride, const double *alpha,
                                      int64_t alpha_stride, double dt, const double *phi, int64_t phi_stride,
                                      const double *q, int64_t q_stride, const double *gamma, int64_t gamma_stride,
                                      const double *r, int64_t r_stride, int32_t warmup, double *mle, int64_t *nobs,
                                      int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_dfm_loglik_compact");
    if (!tidx || (!alpha && !(phi && q))) return MKF_ERR_BAD_ARG;
    mkf_dfm::DfmArgs a{};
    a.B = B; a.T = J; a.y = yc; a.y_stride = yc_stride;
    a.tidx = tidx; a.tidx_stride = tidx_stride;
    a.gamma = gamma; a.gamma_stride = gamma_stride; a.r = r; a.r_stride = r_stride;
    a.alpha = alpha; a.alpha_stride = alpha_stride; a.dt = alpha ? dt : 1.0;
    a.phi = phi; a.phi_stride = phi_stride; a.q = q; a.q_stride = q_stride;
    a.warmup = warmup;
    a.mle = mle; a.nobs = nobs; a.status = status;
    return mkf_dfm::run(m, k, a, (cudaStream_t)stream);
}

extern "C" int mkf_dfm_filter_history(int64_t B, int64_t T, int32_t m, int32_t k, const double *y, int64_t y_stride,
                                      const double *phi, int64_t phi_stride, const double *q, int64_t q_stride,
                                      const double *gamma, int64_t gamma_stride, const double *r, int64_t r_stride,
                                      int32_t warmup, double *sigma, double *detf, int32_t *cnt, double *fm,
                                      double *fc, double *pm, double *pc, double *mle, int64_t *nobs,
                                      int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_dfm_filter_history");
    if (!phi || !q || !mle) return MKF_ERR_BAD_ARG;
    if (!(sigma || detf || cnt || fm || fc || pm || pc)) return MKF_ERR_BAD_ARG;
    mkf_dfm::DfmArgs a{};
    a.B = B; a.T = T; a.y = y; a.y_stride = y_stride;
    a.phi = phi; a.phi_stride = phi_stride; a.q = q; a.q_stride = q_stride;
    a.gamma = gamma; a.gamma_stride = gamma_stride; a.r = r; a.r_stride = r_stride;
    a.alpha = nullptr; a.dt = 1.0; a.warmup = warmup;
    a.mle = mle; a.nobs = nobs; a.status = status;
    a.sigma = sigma; a.detf = detf; a.cnt = cnt; a.fm = fm; a.fc = fc; a.pm = pm; a.pc = pc;
    return mkf_dfm::run(m, k, a, (cudaStream_t)stream);
}

extern "C" int mkf_dfm_set_handoff(int enabled) {
    return mkf_dfm::g_handoff.exchange(enabled ? 1 : 0);
}

extern "C" int mkf_dfm_set_hist_tma(int enabled) { return mkf_dfm::g_hist_tma.exchange(enabled ? 1 : 0); }

extern "C" int mkf_dfm_set_profile(int enabled) { return mkf_dfm::g_profile.exchange(enabled ? 1 : 0); }

extern "C" int mkf_dfm_last_profile(void *stream, float *ms, int64_t *stats) {
    using namespace mkf_dfm;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return MKF_ERR_CUDA;
    HandoffScratch snap;
    {
        std::lock_guard<std::mutex> lock(g_ws_mutex);
        auto it = g_ws.find(std::make_pair(dev, (cudaStream_t)stream));
        if (it == g_ws.end() || !it->second.ptr || it->second.blocks <= 0) return MKF_ERR_BAD_ARG;
        snap = it->second;
    }
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    if (ms) {
        for (int i = 0; i < 3; ++i) {
            ms[i] = -1.0f;
            if (snap.profiled) cudaEventElapsedTime(&ms[i], snap.ev[i], snap.ev[i + 1]);
        }
    }
    if (stats) {
        int64_t *tn = new int64_t[snap.blocks];
        e = cudaMemcpy(tn, snap.tnext, sizeof(int64_t) * snap.blocks, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            delete[] tn;
            mkf_host::set_cuda_error(e);
            return MKF_ERR_CUDA;
        }
        int64_t handed = 0, steps = 0;
        for (int64_t w = 0; w < snap.blocks; ++w)
            if (tn[w] < snap.T) {
                ++handed;
                steps += snap.T - tn[w];
            }
        delete[] tn;
        stats[0] = snap.blocks;
        stats[1] = handed;
        stats[2] = steps;
    }
    (void)cudaGetLastError();
    return MKF_OK;
}

extern "C" void mkf_release_workspaces(void) {
    std::lock_guard<std::mutex> lock(mkf_dfm::g_ws_mutex);
    for (auto &kv : mkf_dfm::g_ws) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(kv.first.first);
        if (kv.second.ptr) cudaFree(kv.second.ptr);
        if (kv.second.have_events)
            for (int i = 0; i < 4; ++i) cudaEventDestroy(kv.second.ev[i]);
        cudaSetDevice(cur);
    }
    mkf_dfm::g_ws.clear();
    (void)cudaGetLastError();
}

extern "C" void mkf_limits(int32_t *n_max, int32_t *dfm_m_max, int32_t *dfm_k_max) {
    if (n_max) *n_max = 32;
    if (dfm_m_max) *dfm_m_max = mkf_dfm::kMaxM;
    if (dfm_k_max) *dfm_k_max = mkf_dfm::kMaxK;
}
