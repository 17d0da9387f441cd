// TEST INFRASTRUCTURE: runs the __host__ __device__ arithmetic of
// metran_b200/csrc/dfm_collapsed.cuh on the CPU for one model, with the same mode logic as
// the CUDA kernel (collapsed step when this and the previous step are fully observed,
// expand + ordinary scalar updates otherwise), so that `pytest -m "not gpu"` can compare it
// with the oracle.  Not part of the product library.
//
// use_collapse: 0 = full-state recursion only (the kernel's general path, structured updates),
//               1 = + factor-only step in runs of fully observed steps,
//               3 = + steady-state switch (convergence monitor, hold, capture, data-only step).
//               bit 2 (5, 7) = the ordinary factor-only steps in information form (collapsed_info_step);
//               the capture step stays in the elimination form.
// n_collapsed returns the number of factor-only steps (modes 0, 1) or of steady-state steps (3).
#include <math.h>
#include <stdint.h>

#include "../../metran_b200/csrc/dfm_collapsed.cuh"

using namespace mkf_dfm;

static bool observed(double v) { return isfinite(v) && v != -1e10; }

template <int M, int K>
static double run(int64_t T, const double *y, const double *phi, const double *q, const double *G, int warmup,
                  int use_collapse, int64_t *n_collapsed) {
    constexpr int N = M + K;
    constexpr int NT = N * (N + 1) / 2;
    double cst[collapsed_const_count(M, K) + 1];
    collapsed_consts<M, K>(phi, q, G, [&](int i, double v) { cst[i] = v; });
    const bool info = (use_collapse & 4) != 0;
    use_collapse &= 3;
    double ics[collapsed_info_const_count(M, K) + 1];
    collapsed_info_consts<M, K>(phi, q, G, [&](int i, double v) { ics[i] = v; });
    double x[N], P[NT];
    for (int i = 0; i < N; ++i) x[i] = 0.0;
    for (int r = 0; r < N; ++r)
        for (int c = r; c < N; ++c) P[tri<N>(r, c)] = (r == c) ? 1.0 : 0.0;
    double yp[M] = {0};
    constexpr int MT = M * (M + 1) / 2;
    constexpr int kCheckEvery = 6;
    const double kFreezeTol = ldexp(1.0, -50);
    bool frozen = false, capture_next = false;
    int since_check = 0;
    double frozen_prod = 1.0, cap[MT + K * M + 1], cchk = -1.0, Chold[K * (K + 1) / 2 + 1];
    int nconv = 0;
    bool collapsed = false, bad = false;
    long double sum_sigma = 0.0L, sum_det = 0.0L;
    int64_t nobs = 0, nsig = 0;
    *n_collapsed = 0;
    for (int64_t t = 0; t < T; ++t) {
        const double *yv = y + t * M;
        bool all_obs = true, any_obs = false;
        for (int i = 0; i < M; ++i) {
            const bool o = observed(yv[i]);
            all_obs = all_obs && o;
            any_obs = any_obs || o;
        }
        double sigma = 0.0, stepprod = 1.0;
        int cnt = 0;
        if (use_collapse && collapsed && all_obs) {
            cnt = M;
            if (frozen) {
                // steady state: gains frozen at the values captured below
                collapsed_frozen_step<M, K>(x, phi, yp, yv, [&](int i) { return cst[i]; }, sigma);
                stepprod = frozen_prod;
                ++*n_collapsed;   // (mode 3 reports frozen steps only, see below)
            } else if (capture_next) {
                collapsed_step_impl<M, K, true>(x, P, phi, q + M, yp, yv, [&](int i) { return cst[i]; },
                                                [&](int i, double v) { cap[i] = v; }, sigma, stepprod, bad);
                for (int i = 0; i < MT + K * M; ++i) cst[M * K + i] = cap[i];   // over S0 and QG
                for (int j = 0; j < K; ++j)
                    for (int l = j; l < K; ++l) P[tri<N>(M + j, M + l)] = Chold[tri<K>(j, l)];   // C is held
                frozen_prod = stepprod;
                frozen = true;
                capture_next = false;
            } else {
                if (info)
                    collapsed_info_step<M, K>(x, P, phi, yp, yv, [&](int i) { return ics[i]; }, sigma, stepprod, bad);
                else
                    collapsed_step<M, K>(x, P, phi, q + M, yp, yv, [&](int i) { return cst[i]; }, sigma, stepprod, bad);
                if (use_collapse != 3) ++*n_collapsed;
                if (use_collapse == 3 && ++since_check == kCheckEvery) {
                    // the kernel's test (there: once per tile and for all lanes of the warp)
                    since_check = 0;
                    const double cs = collapsed_checksum<M, K>(P);
                    const bool same = fabs(cs - cchk) <= kFreezeTol * fabs(cs);
                    cchk = cs;
                    nconv = same ? nconv + 1 : 0;
                    capture_next = nconv >= 2;
                    if (capture_next)   // from here on the lane holds C (a warp may wait for its slowest lane)
                        for (int j = 0; j < K; ++j)
                            for (int l = j; l < K; ++l) Chold[tri<K>(j, l)] = P[tri<N>(M + j, M + l)];
                }
            }
        } else {
            if (collapsed) {
                if (frozen) collapsed_consts<M, K>(phi, q, G, [&](int i, double v) { cst[i] = v; });   // S0, QG back
                collapsed_expand<M, K>(x, P, yp, [&](int i) { return G[i]; });
            }
            collapsed = false;
            frozen = false;
            capture_next = false;
            since_check = 0;
            cchk = -1.0;
            nconv = 0;
            for (int i = 0; i < N; ++i) x[i] *= phi[i];
            for (int r = 0; r < N; ++r)
                for (int c = r; c < N; ++c) {
                    double v = (phi[r] * P[tri<N>(r, c)]) * phi[c];
                    P[tri<N>(r, c)] = (r == c) ? v + q[r] : v;
                }
            for (int i = 0; i < M; ++i) {
                if (!observed(yv[i])) continue;
                ++cnt;
                double v = yv[i] - x[i];
                for (int j = 0; j < K; ++j) v = fma(-G[i * K + j], x[M + j], v);
                double d[N];
                for (int r = 0; r < N; ++r) {
                    double s = P[sym<N>(r, i)];
                    for (int j = 0; j < K; ++j) s = fma(G[i * K + j], P[sym<N>(r, M + j)], s);
                    d[r] = s;
                }
                double f = d[i];
                for (int j = 0; j < K; ++j) f = fma(G[i * K + j], d[M + j], f);
                const double rf = 1.0 / f;
                bad = bad || !(f > 0.0);
                for (int r = 0; r < N; ++r) {
                    const double g = d[r] * rf;
                    x[r] = fma(g, v, x[r]);
                    for (int c = r; c < N; ++c) P[tri<N>(r, c)] = fma(-g, d[c], P[tri<N>(r, c)]);
                }
                sigma = fma(v * v, rf, sigma);
                stepprod *= f;
            }
            if (all_obs && K > 0) collapsed = true;
        }
        if (cnt > 0 && nsig >= warmup) {
            sum_sigma += sigma;
            sum_det += log(stepprod);
        }
        nsig += cnt > 0;
        nobs += (t >= warmup) ? cnt : 0;
        for (int i = 0; i < M; ++i) yp[i] = yv[i];
    }
    if (bad) return NAN;
    return (double)((long double)nobs * 1.8378770664093454835606594728112L + sum_det + sum_sigma);
}

template <int K>
static double run_m(int m, int64_t T, const double *y, const double *phi, const double *q, const double *G, int w,
                    int uc, int64_t *nc) {
    switch (m) {
        case 1: return run<1, K>(T, y, phi, q, G, w, uc, nc);
        case 2: return run<2, K>(T, y, phi, q, G, w, uc, nc);
        case 3: return run<3, K>(T, y, phi, q, G, w, uc, nc);
        case 4: return run<4, K>(T, y, phi, q, G, w, uc, nc);
        case 5: return run<5, K>(T, y, phi, q, G, w, uc, nc);
        case 6: return run<6, K>(T, y, phi, q, G, w, uc, nc);
        case 7: return run<7, K>(T, y, phi, q, G, w, uc, nc);
        case 8: return run<8, K>(T, y, phi, q, G, w, uc, nc);
    }
    return NAN;
}

extern "C" double host_dfm_loglik(int m, int k, int64_t T, const double *y, const double *phi, const double *q,
                                  const double *G, int warmup, int use_collapse, int64_t *n_collapsed) {
    switch (k) {
        case 1: return run_m<1>(m, T, y, phi, q, G, warmup, use_collapse, n_collapsed);
        case 2: return run_m<2>(m, T, y, phi, q, G, warmup, use_collapse, n_collapsed);
        case 3: return run_m<3>(m, T, y, phi, q, G, warmup, use_collapse, n_collapsed);
    }
    return NAN;
}
