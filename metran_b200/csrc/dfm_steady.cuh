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

// Tile shape of the steady kernel, from tools/ubench_rowstream.cu (pure streaming of y with this
// access pattern, 8 one-warp CTAs per SM, profiles/r02_rowstream.txt): the TMA engine wants few, long
// row pieces -- 32 rows x 6 steps (240 B per row) x 3 stages streams at 6.9 TB/s with 256-byte L2
// promotion, 2-step tiles x 8 stages only 4.1 TB/s; L2 prefetch, LDGSTS and plain LDG.128 stay
// below 4.7 TB/s.
#ifndef MKF_ST_TT
#define MKF_ST_TT 6        // time steps per tile of the steady kernel (M*TT even: 16-byte rows)
#endif
#ifndef MKF_ST_STAGES
#define MKF_ST_STAGES 3    // ring depth: all stages are in flight while one tile sits in registers
#endif
#ifndef MKF_ST_PROXY_FENCE
#define MKF_ST_PROXY_FENCE 1   // fence.proxy.async between a tile's LDS and the TMA refill of its stage
#endif
#ifndef MKF_ST_MINB
#define MKF_ST_MINB 8      // resident one-warp CTAs per SM the register allocation must allow
#endif

__host__ __device__ constexpr int steady_tile_steps(int M) { return (M % 2 == 1 && MKF_ST_TT % 2 == 1) ? MKF_ST_TT + 1 : MKF_ST_TT; }
__host__ __device__ constexpr size_t steady_smem_bytes(int M) {
    return 128 + (size_t)MKF_ST_STAGES * 32 * steady_tile_steps(M) * M * sizeof(double) + MKF_ST_STAGES * sizeof(uint64_t);
}

// One elected lane of the (converged) warp: SASS ELECT, no divergence loop around the TMA issue.
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
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
    // time indices fit 32 bits: the tensor map of y needs T * M < 2^31 (make_y_tensor_map)
    const int T = (int)a.T;
    const int t_begin = (int)a.ho_tnext[w];
    const double *wsb = a.ho_ws + b;          // field f of this lane: wsb[f * Bp]
    const int64_t Bp = a.ho_pitch;

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    // ---- tiles on the absolute grid of TT steps; the first and the last may be partial -------------
    const int tile0 = t_begin / TT;
    const int ntiles = (T + TT - 1) / TT - tile0;      // >= 1: the full kernel hands off only with steps left
    const unsigned row_off = (shared_y || !active) ? 0u : (unsigned)lane * PITCH * 8u;
    const unsigned ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
    constexpr unsigned kStageBytes = 32u * PITCH * 8u;
    const unsigned tx_bytes = shared_y ? (unsigned)(PITCH * 8) : kStageBytes;
    const int row0 = shared_y ? 0 : (int)b0;
    // refill of stage s with tile `rel` (one elected lane; the stage must be free)
    auto issue = [&](int s, int rel) {
        const unsigned bar = bars_u32 + 8u * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(tx_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
                ring_u32 + kStageBytes * s),
            "l"(&ymap), "r"((tile0 + rel) * (TT * M)), "r"(row0), "r"(bar)
            : "memory");
    };
    if (elect_one()) {
#pragma unroll
        for (int s = 0; s < S; ++s)
            if (s < ntiles) issue(s, s);
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
    const int warmup = a.warmup;

    // y(t_begin - 1): straight from global memory, once (t_begin >= 1 by construction)
    double yp[M];
    {
        const double *yb = a.y + (active ? b : b0) * a.y_stride + (int64_t)(t_begin - 1) * M;
#pragma unroll
        for (int i = 0; i < M; ++i) yp[i] = yb[i];
    }

    // The exponent of the running product of the innovation variances is peeled off once per tile
    // (an exact scaling: the mantissa bits and the summed exponent do not depend on when it happens).
    auto peel = [&]() {
        const int hi = __double2hiint(fprod);
        const int ex = ((hi >> 20) & 0x7ff) - 1023;
        fexp += ex;
        fprod = __hiloint2double(hi - (ex << 20), __double2loint(fprod));
    };
    // one steady step at absolute time t (the arithmetic and its order are dfm_loglik_kernel's)
    auto step = [&](const double *yc, int t) {
        sum_sigma.add(pend_on ? pend_sigma : 0.0);
        fprod *= pend_on ? pend_prod : 1.0;
        double sigma = 0.0;
        collapsed_frozen_step<M, K>(x, phi, yp, yc, [&](int i) { return cst[i]; }, sigma);
        pend_on = nsig >= warmup;
        pend_sigma = sigma;
        pend_prod = frozen_prod;
        nsig += 1;
        nobs += (t >= warmup) ? M : 0;
#pragma unroll
        for (int i = 0; i < M; ++i) yp[i] = yc[i];
    };

    int t = t_begin;             // next step to do
    bool gap = false;
    // wait for stage s and pull its tile into registers (128-bit LDS: rows are 16-byte aligned)
    auto load_tile = [&](int s, unsigned parity, double *yt) {
        const unsigned bar = bars_u32 + 8u * s;
        unsigned ok;
        do {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
                "selp.u32 %0, 1, 0, p;\n"
                "}\n"
                : "=r"(ok)
                : "r"(bar), "r"(parity)
                : "memory");
        } while (!ok);
        const unsigned src = ring_u32 + kStageBytes * s + row_off;
#pragma unroll
        for (int e = 0; e < PITCH / 2; ++e)
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(yt[2 * e]), "=d"(yt[2 * e + 1]) : "r"(src + 16u * e));
    };
    // One FULL tile (all TT steps inside [t, T)): validate every step (all M series observed in all
    // 32 lanes; one vote over the whole tile), refill the stage, apply the steps.  The vote depends
    // on every lane's loads, so it also orders those loads before the refill; the proxy fence makes
    // that ordering visible to the async (TMA) proxy.  A tile with an incomplete step is left to
    // slow_tile: it is NOT refilled, so S - 1 loads are then in flight, the ones the exit path waits
    // for.  Called with a literal s in the unrolled main loop, where all addressing folds to
    // immediates.
    auto fast_tile = [&](int s, unsigned parity, int rel) {
        double yt[PITCH];
        load_tile(s, parity, yt);
        unsigned ok = 1u;
#pragma unroll
        for (int e = 0; e < PITCH; ++e) ok &= is_observed(yt[e]) ? 1u : 0u;
        if (!__all_sync(0xffffffffu, ok != 0u)) {
            gap = true;
            return;
        }
#if MKF_ST_PROXY_FENCE
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
#endif
        if (rel + S < ntiles) {
            if (elect_one()) issue(s, rel + S);
        }
        const int ts = (tile0 + rel) * TT;
#pragma unroll
        for (int u = 0; u < TT; ++u) step(yt + u * M, ts + u);
        peel();
        t = ts + TT;
    };
    // Any tile: partial on the left (first tile), on the right (last tile), or with an incomplete
    // step; applies the complete steps in front of the first incomplete one and reports the gap.
    auto slow_tile = [&](int s, unsigned parity, int rel) {
        const int ts = (tile0 + rel) * TT;
        const int lo = max(t - ts, 0), hi = min(TT, T - ts);
        double yt[PITCH];
        load_tile(s, parity, yt);
        int nok = hi;            // first incomplete step in [lo, hi) (warp-uniform)
#pragma unroll
        for (int u = TT - 1; u >= 0; --u) {
            unsigned ok = 1u;
#pragma unroll
            for (int i = 0; i < M; ++i) ok &= is_observed(yt[u * M + i]) ? 1u : 0u;
            const bool all = __all_sync(0xffffffffu, ok != 0u);
            if (!all && u >= lo && u < hi) nok = u;
        }
        if (nok == hi) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            if (rel + S < ntiles) {
                if (elect_one()) issue(s, rel + S);
            }
        }
#pragma unroll
        for (int u = 0; u < TT; ++u)
            if (u >= lo && u < nok) step(yt + u * M, ts + u);
        peel();
        t = ts + nok;
        gap = nok != hi;
    };

    // the first S tiles (run-time stage; the first may start in the middle), then groups of S full
    // tiles with compile-time stages, then the rest (the last tile may be partial)
    int rel = 0;
    const int nfull = (T / TT) - tile0;          // tiles [1, nfull) are complete on both sides
#pragma unroll 1
    for (; !gap && rel < min(S, ntiles); ++rel) slow_tile(rel, 0u, rel);
    unsigned parity = 1u;
    while (!gap && rel + S <= nfull) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (!gap) fast_tile(s, parity, rel + s);
        }
        if (gap) {
            rel = t / TT - tile0;                // the tile fast_tile refused: still in its stage
            slow_tile(rel % S, parity, rel);
            break;
        }
        rel += S;
        parity ^= 1u;
    }
    if (!gap) {
        // rel is a multiple of S; fewer than S full tiles and possibly a partial one remain
#pragma unroll 1
        for (int s = 0; !gap && rel < ntiles; ++s, ++rel) slow_tile(s, parity, rel);
    }

    if (gap) {
        // drain the loads still in flight (their destination is this CTA's shared memory), then hand
        // the warp back: c, the running sums and the position; held C, phi and the constants are
        // still in the record
        const int rel_gap = t / TT - tile0;            // the tile that held the incomplete step
        for (int r2 = rel_gap + 1; r2 < min(rel_gap + S, ntiles); ++r2) mbar_wait(&bars[r2 % S], (unsigned)((r2 / S) & 1));
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
