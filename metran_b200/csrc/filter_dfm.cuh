// Thread-per-model Kalman filter -> -2 log L for Metran-shaped dynamic factor models.
//
// Replaces, for a whole batch at once, the chain Metran.get_mle -> _get_matrices ->
// SPKalmanFilter.run_filter (seqkalmanfilter) -> SPKalmanFilter.get_mle of the reference
// (metran/metran.py:605-622, 246-416; metran/kalmanfilter.py:315-390, 550-567).
//
// Model structure exploited (metran.py:283-384): n = M + K states,
// A = diag(phi), Q = diag(q), H = [I_M | Gamma(M x K)], x0 = 0, P0 = I.  Then for series i
//   h = e_i + sum_j Gamma[i][j] e_{M+j}
//   d = P h      costs N*K FMAs instead of N^2
//   f = h.d      costs K FMAs
// and P stays symmetric, so only its upper triangle (N(N+1)/2 doubles) lives in registers.
// One thread owns one model for all T steps: no shuffles, no shared-memory traffic for
// the state; the fp64 pipe is the bound (kernel K1).
//
// Observations y[B][T][M] (model-major, NaN = missing) are streamed by the TMA engine:
// every lane issues one cp.async.bulk per time tile for its own model row into a
// shared-memory ring (mbarrier complete_tx), and reads it back with LDS.  The row pitch
// is padded to 2 (mod 4) doubles so the 64-bit reads of a warp are 2-way bank conflicted
// at worst.  If the rows are not 16-byte aligned the kernel stages with 8-byte cp.async.
#pragma once
#include "mkf_common.cuh"

namespace mkf_dfm {

using namespace mkf;

struct DfmArgs {
    int64_t B, T;
    const double *y;
    int64_t y_stride;
    const double *phi, *q, *gamma, *r, *alpha;
    int64_t phi_stride, q_stride, gamma_stride, r_stride, alpha_stride;
    double dt;
    int warmup;
    int use_bulk;  // rows 16-byte aligned -> cp.async.bulk staging
    double *mle;
    int64_t *nobs;
    int32_t *status;
};

constexpr int kBlock = 64;   // threads per CTA (2 independent warps)
constexpr int kTT = 8;       // time steps per tile
constexpr int kStages = 2;

__host__ __device__ constexpr int tile_pitch(int M) {
    int p = (kTT * M + 1) & ~1;       // even -> rows stay 16-byte aligned
    return (p % 4 == 0) ? p + 2 : p;  // p = 2 (mod 4): at most 2-way LDS.64 conflicts
}
__host__ __device__ constexpr size_t smem_bytes(int M) {
    return (size_t)(kBlock / 32) * (kStages * 32 * tile_pitch(M) * sizeof(double) + kStages * sizeof(uint64_t));
}

// index of (r, c), r <= c, in the packed upper triangle of an N x N matrix
template <int N>
__host__ __device__ constexpr int tri(int r, int c) {
    return r * N - (r * (r - 1)) / 2 + (c - r);
}
template <int N>
__host__ __device__ constexpr int sym(int r, int c) {
    return r <= c ? tri<N>(r, c) : tri<N>(c, r);
}

template <int M, int K>
__global__ void __launch_bounds__(kBlock) dfm_loglik_kernel(const DfmArgs a) {
    constexpr int N = M + K;
    constexpr int NT = N * (N + 1) / 2;
    constexpr int PITCH = tile_pitch(M);
    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *ring = reinterpret_cast<double *>(smem_raw) + (size_t)warp * kStages * 32 * PITCH;
    uint64_t *bars = reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smem_raw) +
                                                  (size_t)(kBlock / 32) * kStages * 32 * PITCH) +
                     warp * kStages;
    const int64_t b = (int64_t)blockIdx.x * kBlock + threadIdx.x;
    const bool active = b < a.B;
    const unsigned act_mask = __ballot_sync(0xffffffffu, active);
    if (act_mask == 0) return;
    const int nactive = __popc(act_mask);
    const bool bulk = a.use_bulk != 0;

    if (bulk) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }

    // ---- per-model constants ---------------------------------------------------------
    double phi[N], q[N], G[M * K + 1], R[M];
    const int64_t bb = active ? b : 0;
    {
        const double *gp = a.gamma + bb * a.gamma_stride;
#pragma unroll
        for (int i = 0; i < M * K; ++i) G[i] = gp[i];
#pragma unroll
        for (int i = 0; i < M; ++i) R[i] = a.r ? a.r[bb * a.r_stride + i] : 0.0;
        if (a.alpha) {
            // metran.py:262-263, 310-322
            const double *al = a.alpha + bb * a.alpha_stride;
#pragma unroll
            for (int i = 0; i < N; ++i) phi[i] = exp(-a.dt / al[i]);
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double comm = 0.0;
#pragma unroll
                for (int j = 0; j < K; ++j) comm += G[i * K + j] * G[i * K + j];
                q[i] = (1.0 - phi[i] * phi[i]) * (1.0 - comm);
            }
#pragma unroll
            for (int j = M; j < N; ++j) q[j] = 1.0 - phi[j] * phi[j];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                phi[i] = a.phi[bb * a.phi_stride + i];
                q[i] = a.q[bb * a.q_stride + i];
            }
        }
    }

    // ---- state: x0 = 0, P0 = I (kalmanfilter.py:747-750) ------------------------------
    double x[N], P[NT];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = r; c < N; ++c) P[tri<N>(r, c)] = (r == c) ? 1.0 : 0.0;

    KahanSum sum_sigma;
    double fprod = 1.0;   // running product of innovation variances, mantissa in [1,2)
    int64_t fexp = 0;     // ... and its binary exponent: sum log f = log(fprod) + fexp ln 2
    int64_t nobs = 0;
    int nsig = 0;
    bool bad = false;

    const int64_t T = a.T;
    const int64_t ntiles = (T + kTT - 1) / kTT;
    const double *yb = a.y + bb * a.y_stride;
    double *myrow = ring + (size_t)lane * PITCH;  // + stage * 32 * PITCH

    auto issue_tile = [&](int64_t tile) {
        if (tile >= ntiles) {
            if (!bulk) cp_async_commit();
            return;
        }
        const int s = (int)(tile % kStages);
        const int64_t t0 = tile * kTT;
        const int nel = (int)min((int64_t)kTT, T - t0) * M;
        double *dst = myrow + (size_t)s * 32 * PITCH;
        const double *src = yb + t0 * M;
        if (bulk) {
            const int nel16 = nel & ~1;  // bulk copies move multiples of 16 bytes
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            if (lane == 0) mbar_arrive_expect_tx(&bars[s], (unsigned)(nactive * nel16 * 8));
            __syncwarp();
            if (active) {
                if (nel16) bulk_g2s(dst, src, (unsigned)(nel16 * 8), &bars[s]);
                if (nel & 1) dst[nel - 1] = src[nel - 1];
            }
        } else {
            if (active)
                for (int e = 0; e < nel; ++e) cp_async8(dst + e, src + e);
            cp_async_commit();
        }
    };

    issue_tile(0);
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        issue_tile(tile + 1);
        const int s = (int)(tile % kStages);
        if (bulk) {
            mbar_wait(&bars[s], (unsigned)((tile / kStages) & 1));
        } else {
            cp_async_wait<kStages - 1>();
        }
        const double *ys = myrow + (size_t)s * 32 * PITCH;
        const int64_t t0 = tile * kTT;
        const int steps = (int)min((int64_t)kTT, T - t0);

        for (int tt = 0; tt < steps; ++tt) {
            double yv[M];
#pragma unroll
            for (int i = 0; i < M; ++i) yv[i] = ys[tt * M + i];

            // ---- predict: x = phi x ; P = phi P phi + Q (kalmanfilter.py:318-331) -----
#pragma unroll
            for (int i = 0; i < N; ++i) x[i] *= phi[i];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = r; c < N; ++c) {
                    double v = (phi[r] * P[tri<N>(r, c)]) * phi[c];
                    P[tri<N>(r, c)] = (r == c) ? v + q[r] : v;
                }

            // ---- sequential scalar updates (kalmanfilter.py:341-378) -------------------
            double sigma = 0.0, stepprod = 1.0;
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < M; ++i) {
                if (is_observed(yv[i])) {
                    ++cnt;
                    double v = yv[i] - x[i];
#pragma unroll
                    for (int j = 0; j < K; ++j) v = fma(-G[i * K + j], x[M + j], v);
                    double d[N];
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        double s = P[sym<N>(r, i)];
#pragma unroll
                        for (int j = 0; j < K; ++j) s = fma(G[i * K + j], P[sym<N>(r, M + j)], s);
                        d[r] = s;
                    }
                    double f = d[i] + R[i];
#pragma unroll
                    for (int j = 0; j < K; ++j) f = fma(G[i * K + j], d[M + j], f);
                    const double rf = 1.0 / f;
                    if (!(f > 0.0) || !isfinite(f)) bad = true;
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        const double g = d[r] * rf;
                        x[r] = fma(g, v, x[r]);
#pragma unroll
                        for (int c = r; c < N; ++c) P[tri<N>(r, c)] = fma(-g, d[c], P[tri<N>(r, c)]);
                    }
                    sigma = fma(v * v, rf, sigma);
                    stepprod *= f;
                }
            }
            if (cnt > 0) {
                if (nsig >= a.warmup) {
                    sum_sigma.add(sigma);
                    // fold the step's product of f into (fprod, fexp) and renormalise
                    fprod *= stepprod;
                    const int hi = __double2hiint(fprod);
                    const int ex = ((hi >> 20) & 0x7ff) - 1023;
                    fexp += ex;
                    fprod = __hiloint2double(hi - (ex << 20), __double2loint(fprod));
                }
                ++nsig;
            }
            if (t0 + tt >= a.warmup) nobs += cnt;
        }
    }
    if (!bulk) cp_async_wait<0>();

    if (active) {
        double logdet = log(fprod) + (double)fexp * kLn2;
        double mle = (double)nobs * kLog2Pi + logdet + sum_sigma.value();
        if (bad) mle = nan("");
        a.mle[b] = mle;
        if (a.nobs) a.nobs[b] = nobs;
        if (a.status) a.status[b] = bad ? MKF_STATUS_NONPOSITIVE_F : 0;
    }
}

template <int M, int K>
int launch(const DfmArgs &a, cudaStream_t stream) {
    auto kern = dfm_loglik_kernel<M, K>;
    const size_t smem = smem_bytes(M);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    const int64_t blocks = (a.B + kBlock - 1) / kBlock;
    kern<<<(unsigned)blocks, kBlock, smem, stream>>>(a);
    return mkf_host::check_launch();
}

// one translation unit per K instantiates M = 1..8 (keeps nvcc wall time parallel)
template <int K>
int launch_for_m(int m, const DfmArgs &a, cudaStream_t s);

constexpr int kMaxM = 8;
constexpr int kMaxK = 3;

}  // namespace mkf_dfm
