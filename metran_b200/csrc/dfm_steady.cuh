// Steady-state continuation of the thread-per-model DFM filter (kernel K1S).
//
// Reference path: the same chain as filter_dfm.cuh -- Metran.get_mle -> SPKalmanFilter.run_filter
// (seqkalmanfilter, metran/kalmanfilter.py:315-390) -> get_mle (:550-567) -- for the part of a
// series in which every step is observed completely and the factor covariance has converged.
//
// dfm_loglik_kernel (filter_dfm.cuh) carries a warp of 32 models through the transient: full-state
// steps, factor-only steps, until every lane holds its converged C and one more step has captured
// the (now constant) reciprocal pivots, elimination multipliers and gains.  At that point it writes
// a hand-off record per model (c, held C, the 3 x 15 constants, phi, the running sums) and exits.
// This kernel picks the warp up at the next time step and applies collapsed_frozen_step
// (dfm_collapsed.cuh: the SAME function, hence bit-identical results) to the rest of the series:
//   v = y_t - phi_s*y_{t-1} + D c;  c = phi_c*c;  sigma += v_i^2/f_i;  v_l -= m_il v_i;  c_j += g_ji v_i
// 58 fp64 instructions per step and model at (M, K) = (5, 3) against 40 bytes of observations, so
// the bound is the HBM stream of y -- provided nothing else is in the way.  What was in the way
// inside the 255-register full kernel: 45 LDS per step for the constants (shared-memory bandwidth:
// 11.5 KB per warp-step), a 2-stage ring of 6-step tiles (one tile in flight per warp), and the
// register allocation of the full-state code.  Here the constants live in registers, the ring is
// MKF_ST_STAGES deep with short tiles (all stages in flight while one tile sits in registers), and
// every step is validated (all series observed in all lanes, one warp vote) before it is applied; the
// first step with a missing value hands the warp back to dfm_loglik_kernel in resume mode, which
// expands (c, C, y(t-1)) to the full posterior exactly as it does at a gap of its own.
#pragma once
#include "filter_dfm.cuh"

namespace mkf_dfm {

#ifndef MKF_ST_TT
#define MKF_ST_TT 2        // time steps per tile of the steady kernel (M*TT even: 16-byte rows)
#endif
#ifndef MKF_ST_STAGES
#define MKF_ST_STAGES 6    // ring depth: all stages are in flight while one is being consumed
#endif
#ifndef MKF_ST_MINB
#define MKF_ST_MINB 12     // resident one-warp CTAs per SM the register allocation must allow
#endif

__host__ __device__ constexpr int steady_tile_steps(int M) { return (M % 2 == 1 && MKF_ST_TT % 2 == 1) ? MKF_ST_TT + 1 : MKF_ST_TT; }
__host__ __device__ constexpr size_t steady_smem_bytes(int M) {
    return 128 + (size_t)MKF_ST_STAGES * 32 * steady_tile_steps(M) * M * sizeof(double) + MKF_ST_STAGES * sizeof(uint64_t);
}

template <int M, int K>
__global__ void __launch_bounds__(32, MKF_ST_MINB) dfm_steady_kernel(const DfmArgs a, const __grid_constant__ CUtensorMap ymap) {
    using HO = Handoff<M, K>;
    constexpr int N = M + K;
    constexpr int TT = steady_tile_steps(M);
    constexpr int PITCH = TT * M;
    constexpr int S = MKF_ST_STAGES;
    constexpr int NC = HO::NCONST;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem_al = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    double *ring = reinterpret_cast<double *>(smem_al);
    uint64_t *bars = reinterpret_cast<uint64_t *>(ring + (size_t)S * 32 * PITCH);

    const int lane = threadIdx.x;
    const int64_t w = blockIdx.x;
    if (a.ho_state[w] != HO_STEADY) return;
    const int64_t b0 = w * 32;
    const int64_t b = b0 + lane;
    const bool active = b < a.B;
    const bool shared_y = a.y_stride == 0;
    const int64_t T = a.T;
    const int64_t t_begin = a.ho_tnext[w];
    const double *wsb = a.ho_ws + b;          // field f of this lane: wsb[f * Bp]
    const int64_t Bp = a.ho_pitch;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // ---- tiles on the absolute grid of TT steps; the first and the last may be partial -------------
    const int64_t tile0 = t_begin / TT;
    const int64_t ntiles = (T + TT - 1) / TT - tile0;      // >= 1: the full kernel hands off only with steps left
    const double *myrow = ring + ((shared_y || !active) ? 0 : (size_t)lane * PITCH);
    constexpr unsigned kTileBytes = 32u * PITCH * 8u;
    auto issue = [&](int64_t rel) {   // caller: lane 0 only, stage free, rel < ntiles
        const int s = (int)(rel % S);
        mbar_arrive_expect_tx(&bars[s], shared_y ? (unsigned)(PITCH * 8) : kTileBytes);
        tma_load_2d(ring + (size_t)s * 32 * PITCH, &ymap, (int)((tile0 + rel) * TT * M), shared_y ? 0 : (int)b0, &bars[s]);
    };
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (s < ntiles) issue(s);
    }

    // ---- hand-off record -> registers (overlaps the first tiles' flight) ----------------------------
    double cst[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) cst[i] = wsb[(HO::CONST + i) * Bp];
    double x[N], phi[N];
#pragma unroll
    for (int i = 0; i < N; ++i) phi[i] = wsb[(HO::PHI + i) * Bp];
#pragma unroll
    for (int i = 0; i < M; ++i) x[i] = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) x[M + j] = wsb[(HO::C + j) * Bp];
    KahanSum sum_sigma;
    sum_sigma.s = wsb[HO::SUM_S * Bp];
    sum_sigma.c = wsb[HO::SUM_C * Bp];
    double fprod = wsb[HO::FPROD * Bp];
    double pend_sigma = wsb[HO::PEND_SIGMA * Bp];
    double pend_prod = wsb[HO::PEND_PROD * Bp];
    const double frozen_prod = wsb[HO::FROZEN_PROD * Bp];
    int fexp = (int)wsb[HO::FEXP * Bp];
    int nobs = (int)wsb[HO::NOBS * Bp];
    int nsig = (int)wsb[HO::NSIG * Bp];
    const int fl = (int)wsb[HO::FLAGS * Bp];
    const bool bad = (fl & 1) != 0;
    bool pend_on = (fl & 2) != 0;

    // y(t_begin - 1): straight from global memory, once (t_begin >= 1 by construction)
    double yp[M];
    {
        const double *yb = a.y + (active ? b : b0) * a.y_stride + (t_begin - 1) * M;
#pragma unroll
        for (int i = 0; i < M; ++i) yp[i] = yb[i];
    }

    auto fold = [&]() {
        sum_sigma.add(pend_on ? pend_sigma : 0.0);
        fprod *= pend_on ? pend_prod : 1.0;
        const int hi = __double2hiint(fprod);
        const int ex = ((hi >> 20) & 0x7ff) - 1023;
        fexp += ex;
        fprod = __hiloint2double(hi - (ex << 20), __double2loint(fprod));
    };
    // one steady step at absolute time t (the arithmetic and its order are dfm_loglik_kernel's)
    auto step = [&](const double *yc, int64_t t) {
        fold();
        double sigma = 0.0;
        collapsed_frozen_step<M, K>(x, phi, yp, yc, [&](int i) { return cst[i]; }, sigma);
        pend_on = nsig >= a.warmup;
        pend_sigma = sigma;
        pend_prod = frozen_prod;
        nsig += 1;
        nobs += (t >= a.warmup) ? M : 0;
#pragma unroll
        for (int i = 0; i < M; ++i) yp[i] = yc[i];
    };

    int64_t t = t_begin;           // next step to do
    int64_t rel = 0;
    bool gap = false;
    for (; rel < ntiles; ++rel) {
        const int s = (int)(rel % S);
        mbar_wait(&bars[s], (unsigned)((rel / S) & 1));
        const int64_t ts = (tile0 + rel) * TT;           // first step of the tile
        const int lo = (int)(t - ts);                     // > 0 only in the first tile
        const int hi = (int)min((int64_t)TT, T - ts);     // < TT only in the last tile
        // the whole tile -> registers (128-bit LDS: rows are 16-byte aligned, PITCH is even)
        double yt[PITCH];
        {
            const double2 *src = reinterpret_cast<const double2 *>(myrow + (size_t)s * 32 * PITCH);
#pragma unroll
            for (int e = 0; e < PITCH / 2; ++e) {
                const double2 v = src[e];
                yt[2 * e] = v.x;
                yt[2 * e + 1] = v.y;
            }
        }
        // stage s is free again: refill it with tile rel + S (WAR across proxies: fence, then issue)
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncwarp();
        if (lane == 0 && rel + S < ntiles) issue(rel + S);
        // validate before applying: step u of the tile counts when every lane observed all M series
        // (one vote per step, taken for the whole tile up front)
        bool full = true;        // warp-uniform: all steps [lo, hi) of the tile are complete
        int nok = hi;            // first incomplete step (warp-uniform)
#pragma unroll
        for (int u = TT - 1; u >= 0; --u) {
            unsigned ok = 1u;
#pragma unroll
            for (int i = 0; i < M; ++i) ok &= is_observed(yt[u * M + i]) ? 1u : 0u;
            const bool all = __all_sync(0xffffffffu, ok != 0u);
            if (!all && u >= lo && u < hi) {
                full = false;
                nok = u;
            }
        }
        if (full && lo == 0 && hi == TT) {
#pragma unroll
            for (int u = 0; u < TT; ++u) step(yt + u * M, ts + u);
        } else {
#pragma unroll
            for (int u = 0; u < TT; ++u)
                if (u >= lo && u < nok) step(yt + u * M, ts + u);
        }
        if (!full) {
            t = ts + nok;
            gap = true;
            break;
        }
        t = ts + hi;
    }

    if (gap) {
        // drain the loads still in flight (their destination is this CTA's shared memory), then hand
        // the warp back: c, the running sums and the position; held C, phi and the constants are
        // still in the record
        for (int64_t r2 = rel + 1; r2 < min(rel + S, ntiles); ++r2) mbar_wait(&bars[r2 % S], (unsigned)((r2 / S) & 1));
        double *wo = a.ho_ws + b;
#pragma unroll
        for (int j = 0; j < K; ++j) wo[(HO::C + j) * Bp] = x[M + j];
        wo[HO::SUM_S * Bp] = sum_sigma.s;
        wo[HO::SUM_C * Bp] = sum_sigma.c;
        wo[HO::FPROD * Bp] = fprod;
        wo[HO::PEND_SIGMA * Bp] = pend_sigma;
        wo[HO::PEND_PROD * Bp] = pend_prod;
        wo[HO::FEXP * Bp] = (double)fexp;
        wo[HO::NOBS * Bp] = (double)nobs;
        wo[HO::NSIG * Bp] = (double)nsig;
        wo[HO::FLAGS * Bp] = (double)((bad ? 1 : 0) | (pend_on ? 2 : 0));
        __syncwarp();
        if (lane == 0) {
            a.ho_tnext[w] = t;
            a.ho_state[w] = HO_RESUME;
        }
        return;
    }

    // ---- end of the series: the reduction of SPKalmanFilter.get_mle (kalmanfilter.py:563-566) ----
    sum_sigma.add(pend_on ? pend_sigma : 0.0);
    fprod *= pend_on ? pend_prod : 1.0;
    if (active) {
        const double logdet = log(fprod) + (double)fexp * kLn2;
        double mle = (double)nobs * kLog2Pi + logdet + sum_sigma.value();
        if (bad) mle = nan("");
        a.mle[b] = mle;
        if (a.nobs) a.nobs[b] = (int64_t)nobs;
        if (a.status) a.status[b] = bad ? MKF_STATUS_NONPOSITIVE_F : 0;
    }
    if (lane == 0) a.ho_state[w] = HO_DONE;
}

}  // namespace mkf_dfm
