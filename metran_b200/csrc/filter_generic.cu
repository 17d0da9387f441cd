// Generic batched sequential-processing Kalman filter (any A/Q/H/R, n <= 32).
//
// Replaces seqkalmanfilter (reference metran/kalmanfilter.py:236-400) and the get_mle
// reduction (:550-567).  Mapping: NP = 4/8/16/32 lanes per model (NP >= n), lane r owns
// row r of the covariance P in registers and element r of the state mean; 32/NP models
// share a warp.  Per observed series (kalmanfilter.py:341-378) the lanes of a model do
//   v = y_j - h.x          (butterfly shuffle reduction)
//   d_r = P[r,:].h         (row in registers, h broadcast from shared memory)
//   f = R_j + h.d          (butterfly)
//   g_r = d_r (1/f) ; P[r,:] -= d_r g^T ; x_r += g_r v     (one branch-free reciprocal per series)
// so the only cross-lane traffic is two reductions and one n-vector broadcast.
// Observations are NaN-coded and streamed model-major through a cp.async double-buffered
// shared-memory ring; missing series are skipped with a ballot (the C0 data set has
// 94.5 % empty days).  HIST=true also writes the reference's result contract
// (filtered/predicted means and covariances, per-step sigma/detf/count).
//
// Covariances are written TRANSPOSED relative to the register layout (lane r writes
// column r) so that every store instruction covers one contiguous n*8-byte row; P is
// symmetric up to rounding (|P - P^T| <= a few ulp, as in the reference).
#include "mkf_common.cuh"

namespace {

using namespace mkf;

struct FilterArgs {
    int64_t B, T;
    int n, m;
    const double *y, *A, *Q, *H, *R, *x0, *P0;
    int64_t y_stride, A_stride, Q_stride, H_stride, R_stride, x0_stride, P0_stride;
    int warmup;
    unsigned flags;
    double *sigma, *detf;
    int32_t *cnt;
    double *fm, *fc, *pm, *pc;
    double *mle;
    int64_t *nobs;
    int32_t *status;
    int tile_steps;        // time steps per shared-memory y tile
    int smem_model;        // doubles of shared memory per model
};

__host__ __device__ inline int round_up2(int v) { return (v + 1) & ~1; }

template <int NP, bool HIST>
__global__ void __launch_bounds__(128) filter_generic_kernel(const FilterArgs a) {
    extern __shared__ __align__(16) double smem[];
    const int n = a.n, m = a.m;
    const int group_in_block = threadIdx.x / NP;
    const int r = threadIdx.x % NP;
    const int group_in_warp = (threadIdx.x & 31) / NP;
    const unsigned gmask = (NP == 32) ? 0xffffffffu : (((1u << NP) - 1u) << (group_in_warp * NP));
    const int gshift = group_in_warp * NP;
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x / NP) + group_in_block;
    if (b >= a.B) return;  // whole groups leave together; masks below are per group

    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    const bool diagQ = a.flags & MKF_FLAG_Q_DIAGONAL;
    const int TT = a.tile_steps;

    // ---- shared-memory carve-up for this model -------------------------------------
    double *sm = smem + (size_t)group_in_block * a.smem_model;
    double *sH = sm;                      sm += (size_t)m * NP;
    double *sR = sm;                      sm += round_up2(m);
    double *sa = sm;                      sm += NP;
    double *sx = sm;                      sm += NP;
    double *sy = sm;                      sm += 2 * round_up2(TT * m);
    double *sA = sm;                      if (!diagA) sm += NP * (NP + 1) + 1;
    double *sQ = sm;
    const int ytile = round_up2(TT * m);

    const double *yb = a.y + b * a.y_stride;
    const double *Ab = a.A + b * a.A_stride;
    const double *Qb = a.Q + b * a.Q_stride;
    const double *Hb = a.H + b * a.H_stride;
    const double *Rb = a.R + b * a.R_stride;

    for (int i = r; i < m * NP; i += NP) {
        int j = i / NP, c = i % NP;
        sH[i] = (c < n) ? Hb[j * n + c] : 0.0;
    }
    for (int i = r; i < m; i += NP) sR[i] = Rb[i];
    const double a_r = (r < n) ? Ab[r * n + r] : 0.0;
    sa[r] = a_r;
    if (!diagA)
        for (int i = r; i < NP * NP; i += NP) {
            int rr = i / NP, c = i % NP;
            sA[rr * (NP + 1) + c] = (rr < n && c < n) ? Ab[rr * n + c] : 0.0;
        }
    if (!diagQ)
        for (int i = r; i < NP * NP; i += NP) {
            int rr = i / NP, c = i % NP;
            sQ[rr * (NP + 1) + c] = (rr < n && c < n) ? Qb[rr * n + c] : 0.0;
        }
    const double q_r = (r < n) ? Qb[r * n + r] : 0.0;

    double x = (a.x0 && r < n) ? a.x0[b * a.x0_stride + r] : 0.0;
    double P[NP];
#pragma unroll
    for (int c = 0; c < NP; ++c) {
        double v = 0.0;
        if (r < n && c < n) v = a.P0 ? a.P0[b * a.P0_stride + r * n + c] : (r == c ? 1.0 : 0.0);
        P[c] = v;
    }

    KahanSum sum_sigma, sum_detf;
    int64_t nobs = 0;
    int nsig = 0;
    int status = 0;

    const int64_t T = a.T;
    const int64_t ntiles = (T + TT - 1) / TT;
    auto issue_tile = [&](int64_t tile) {
        if (tile < ntiles) {
            int64_t t0 = tile * TT;
            int cntd = (int)min((int64_t)TT, T - t0) * m;
            const double *src = yb + t0 * m;
            double *dst = sy + (tile & 1) * ytile;
            for (int i = r; i < cntd; i += NP) cp_async8(dst + i, src + i);
        }
        cp_async_commit();
    };
    issue_tile(0);

    const int64_t hist_base = b * T;
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        __syncwarp(gmask);  // everyone is done with the buffer tile+1 will overwrite
        issue_tile(tile + 1);
        cp_async_wait<1>();
        __syncwarp(gmask);
        const double *ys = sy + (tile & 1) * ytile;
        const int64_t t0 = tile * TT;
        const int steps = (int)min((int64_t)TT, T - t0);

        for (int tt = 0; tt < steps; ++tt) {
            const int64_t t = t0 + tt;
            // ---- predict (kalmanfilter.py:318-331) -----------------------------------
            if (diagA) {
                x = a_r * x;
            } else {
                sx[r] = x;
                __syncwarp(gmask);
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NP; ++c) s = fma(sA[r * (NP + 1) + c], sx[c], s);
                __syncwarp(gmask);
                x = s;
            }
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                double2 ac = *reinterpret_cast<const double2 *>(sa + c);
                double v0 = (a_r * P[c]) * ac.x;
                double v1 = (a_r * P[c + 1]) * ac.y;
                if (diagQ) {
                    P[c] = (c == r) ? v0 + q_r : v0;
                    P[c + 1] = (c + 1 == r) ? v1 + q_r : v1;
                } else {
                    P[c] = v0 + sQ[r * (NP + 1) + c];
                    P[c + 1] = v1 + sQ[r * (NP + 1) + c + 1];
                }
            }
            if (HIST && r < n) {
                if (a.pm) a.pm[(hist_base + t) * n + r] = x;
                if (a.pc) {
                    double *dst = a.pc + (hist_base + t) * n * n + r;
#pragma unroll
                    for (int c = 0; c < NP; ++c)
                        if (c < n) dst[(int64_t)c * n] = P[c];
                }
            }

            // ---- sequential update over observed series (kalmanfilter.py:335-382) ----
            double sigma = 0.0, fman = 1.0;
            int fexp = 0;
            int cnt = 0;
            for (int j0 = 0; j0 < m; j0 += NP) {
                const int jj = j0 + r;
                const double yv = (jj < m) ? ys[tt * m + jj] : nan("");
                unsigned bal = __ballot_sync(gmask, is_observed(yv));
                bal = (NP == 32) ? bal : ((bal >> gshift) & ((1u << NP) - 1u));
                cnt += __popc(bal);
                while (bal) {
                    const int jl = __ffs(bal) - 1;
                    bal &= bal - 1;
                    const int j = j0 + jl;
                    const double yj = __shfl_sync(gmask, yv, jl, NP);
                    const double *hrow = sH + j * NP;
                    const double h_r = hrow[r];
                    const double v = yj - group_sum<NP>(h_r * x, gmask);
                    double d;
                    if (NP >= 16) {
                        double d0 = 0.0, d1 = 0.0, d2 = 0.0, d3 = 0.0;
#pragma unroll
                        for (int c = 0; c < NP; c += 4) {
                            double2 h01 = *reinterpret_cast<const double2 *>(hrow + c);
                            double2 h23 = *reinterpret_cast<const double2 *>(hrow + c + 2);
                            d0 = fma(P[c], h01.x, d0);
                            d1 = fma(P[c + 1], h01.y, d1);
                            d2 = fma(P[c + 2], h23.x, d2);
                            d3 = fma(P[c + 3], h23.y, d3);
                        }
                        d = (d0 + d1) + (d2 + d3);
                    } else {
                        d = 0.0;
#pragma unroll
                        for (int c = 0; c < NP; c += 2) {
                            double2 h01 = *reinterpret_cast<const double2 *>(hrow + c);
                            d = fma(P[c], h01.x, d);
                            d = fma(P[c + 1], h01.y, d);
                        }
                    }
                    const double f = sR[j] + group_sum<NP>(h_r * d, gmask);
                    // one reciprocal serves k = d/f and v^2/f; f <= 0 or non-finite is flagged
                    // and poisons the result like the reference's log(f) (kalmanfilter.py:378)
                    const double rf = f_is_bad(f) ? nan("") : rcp_nr(f);
                    const double g = d * rf;
                    sx[r] = g;
                    __syncwarp(gmask);
#pragma unroll
                    for (int c = 0; c < NP; c += 2) {
                        double2 g01 = *reinterpret_cast<const double2 *>(sx + c);
                        P[c] = fma(-d, g01.x, P[c]);
                        P[c + 1] = fma(-d, g01.y, P[c + 1]);
                    }
                    __syncwarp(gmask);
                    x = fma(g, v, x);
                    sigma = fma(v * v, rf, sigma);
                    // sum log f as log(prod f): mantissa product + binary exponent, one log per step
                    fman *= f;
                    {
                        const int hi = __double2hiint(fman);
                        const int ex = ((hi >> 20) & 0x7ff) - 1023;
                        fexp += ex;
                        fman = __hiloint2double(hi - (ex << 20), __double2loint(fman));
                    }
                    if (f_is_bad(f)) status |= MKF_STATUS_NONPOSITIVE_F;
                }
            }
            const double detf = (cnt > 0) ? log(fman) + (double)fexp * kLn2 : 0.0;
            if (cnt > 0) {
                if (nsig >= a.warmup) {
                    sum_sigma.add(sigma);
                    sum_detf.add(detf);
                }
                ++nsig;
            }
            if (t >= a.warmup) nobs += cnt;

            if (HIST) {
                if (r == 0) {
                    if (a.sigma) a.sigma[hist_base + t] = sigma;
                    if (a.detf) a.detf[hist_base + t] = detf;
                    if (a.cnt) a.cnt[hist_base + t] = cnt;
                }
                if (r < n) {
                    if (a.fm) a.fm[(hist_base + t) * n + r] = x;
                    if (a.fc) {
                        double *dst = a.fc + (hist_base + t) * n * n + r;
#pragma unroll
                        for (int c = 0; c < NP; ++c)
                            if (c < n) dst[(int64_t)c * n] = P[c];
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    if (r == 0) {
        if (a.mle) a.mle[b] = (double)nobs * kLog2Pi + sum_detf.value() + sum_sigma.value();
        if (a.nobs) a.nobs[b] = nobs;
        if (a.status) a.status[b] = status;
    }
}

template <int NP, bool HIST>
int launch(const FilterArgs &args0, cudaStream_t stream) {
    FilterArgs a = args0;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    const bool diagQ = a.flags & MKF_FLAG_Q_DIAGONAL;
    int TT = 256 / a.m;
    if (TT < 1) TT = 1;
    if (TT > 32) TT = 32;
    if ((int64_t)TT > a.T && a.T > 0) TT = (int)a.T;
    a.tile_steps = TT;
    int per = a.m * NP + round_up2(a.m) + NP + NP + 2 * round_up2(TT * a.m);
    if (!diagA) per += NP * (NP + 1) + 1;
    if (!diagQ) per += NP * (NP + 1) + 1;
    per = round_up2(per);
    a.smem_model = per;
    const size_t limit = 200 * 1024;
    int threads = 128;
    while (threads > 32 && (size_t)(threads / NP) * per * sizeof(double) > limit) threads /= 2;
    size_t smem = (size_t)(threads / NP) * per * sizeof(double);
    if (smem > limit) return MKF_ERR_UNSUPPORTED;
    auto kern = filter_generic_kernel<NP, HIST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    const int mpb = threads / NP;
    const int64_t blocks = (a.B + mpb - 1) / mpb;
    kern<<<(unsigned)blocks, threads, smem, stream>>>(a);
    return mkf_host::check_launch();
}

}  // namespace

#include "filter_block.cuh"

namespace {

#ifndef MKF_BLOCK_UPDATE
#define MKF_BLOCK_UPDATE 1
#endif

template <bool HIST>
int dispatch(const FilterArgs &a, cudaStream_t s) {
    if (a.n <= 4) return launch<4, HIST>(a, s);
    if (a.n <= 8) return launch<8, HIST>(a, s);
    if (a.n <= 16) return launch<16, HIST>(a, s);
    if (a.n <= 32) {
        // wide models with many series per step: eight series per block update (filter_block.cuh)
        if (MKF_BLOCK_UPDATE && a.m >= 8) {
            const int rc = launch_block<HIST>(a, s);
            if (rc != MKF_ERR_UNSUPPORTED) return rc;
        }
        return launch<32, HIST>(a, s);
    }
    return MKF_ERR_UNSUPPORTED;
}

int validate(int64_t B, int64_t T, int n, int m, const void *y, const void *A, const void *Q, const void *H,
             const void *R) {
    if (B < 0 || T < 0 || n <= 0 || m <= 0) return MKF_ERR_BAD_ARG;
    if (!y || !A || !Q || !H || !R) return MKF_ERR_BAD_ARG;
    if (n > 32) return MKF_ERR_UNSUPPORTED;
    if (B > 0x7fffffffLL) return MKF_ERR_UNSUPPORTED;
    return MKF_OK;
}

}  // namespace

extern "C" int mkf_filter_history(int64_t B, int64_t T, int32_t n, int32_t m, const double *y, int64_t y_stride,
                                  const double *A, int64_t A_stride, const double *Q, int64_t Q_stride,
                                  const double *H, int64_t H_stride, const double *R, int64_t R_stride,
                                  const double *x0, int64_t x0_stride, const double *P0, int64_t P0_stride,
                                  int32_t warmup, uint32_t flags, double *sigma, double *detf, int32_t *cnt,
                                  double *fm, double *fc, double *pm, double *pc, double *mle, int64_t *nobs,
                                  int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_filter_history");
    int rc = validate(B, T, n, m, y, A, Q, H, R);
    if (rc != MKF_OK) return rc;
    if (B == 0) return MKF_OK;
    FilterArgs a{};
    a.B = B; a.T = T; a.n = n; a.m = m;
    a.y = y; a.A = A; a.Q = Q; a.H = H; a.R = R; a.x0 = x0; a.P0 = P0;
    a.y_stride = y_stride; a.A_stride = A_stride; a.Q_stride = Q_stride; a.H_stride = H_stride;
    a.R_stride = R_stride; a.x0_stride = x0_stride; a.P0_stride = P0_stride;
    a.warmup = warmup; a.flags = flags;
    a.sigma = sigma; a.detf = detf; a.cnt = cnt; a.fm = fm; a.fc = fc; a.pm = pm; a.pc = pc;
    a.mle = mle; a.nobs = nobs; a.status = status;
    return dispatch<true>(a, (cudaStream_t)stream);
}

extern "C" int mkf_filter_loglik(int64_t B, int64_t T, int32_t n, int32_t m, const double *y, int64_t y_stride,
                                 const double *A, int64_t A_stride, const double *Q, int64_t Q_stride,
                                 const double *H, int64_t H_stride, const double *R, int64_t R_stride,
                                 const double *x0, int64_t x0_stride, const double *P0, int64_t P0_stride,
                                 int32_t warmup, uint32_t flags, double *mle, int64_t *nobs, int32_t *status,
                                 void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_filter_loglik");
    int rc = validate(B, T, n, m, y, A, Q, H, R);
    if (rc != MKF_OK) return rc;
    if (!mle) return MKF_ERR_BAD_ARG;
    if (B == 0) return MKF_OK;
    FilterArgs a{};
    a.B = B; a.T = T; a.n = n; a.m = m;
    a.y = y; a.A = A; a.Q = Q; a.H = H; a.R = R; a.x0 = x0; a.P0 = P0;
    a.y_stride = y_stride; a.A_stride = A_stride; a.Q_stride = Q_stride; a.H_stride = H_stride;
    a.R_stride = R_stride; a.x0_stride = x0_stride; a.P0_stride = P0_stride;
    a.warmup = warmup; a.flags = flags;
    a.mle = mle; a.nobs = nobs; a.status = status;
    return dispatch<false>(a, (cudaStream_t)stream);
}
