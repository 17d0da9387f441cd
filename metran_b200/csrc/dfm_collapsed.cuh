// Collapsed (factor-only) form of the Metran filter step for runs of fully observed steps.
//
// Reference path: seqkalmanfilter, metran/kalmanfilter.py:315-390, on the matrices of
// Metran._get_matrices (metran/metran.py:283-384): A = diag(phi), Q = diag(q),
// H = [I_M | Gamma], R = 0.
//
// With R = 0, a step in which all M series are observed leaves H P = 0 and H x = y exactly,
// i.e. the posterior is determined by the K common-factor states alone:
//     x_s = y - Gamma c,   P_sc = -Gamma C,   P_ss = Gamma C Gamma^T        (c = x_c, C = P_cc)
// If the NEXT step is fully observed too, its innovations and their covariances follow from
// (c, C, y_prev) without ever forming the M+K state:
//     v   = y - phi_s*y_prev + D c                 D  = Phi_s Gamma - Gamma Phi_c     (M x K)
//     S   = D C D^T + S0                           S0 = Gamma Q_c Gamma^T + Q_s       (M x M)
//     T   = QG - Phi_c C D^T  = cov(c_pred, v)     QG = Q_c Gamma^T                   (K x M)
//     C-  = Phi_c C Phi_c + Q_c,  c- = Phi_c c
// and the reference's series-by-series scalar updates become an LDL^T elimination of
// [S T^T; T C-] pivot by pivot: the pivots ARE the reference's innovation variances f_i, the
// eliminated v_i its innovations, so sigma = sum v_i^2/f_i and detf = sum log f_i are the same
// numbers (to rounding) at (M=5, K=3) for 310 instead of 538 fp64 instructions per step.
// D, S0, QG are per-model constants.  A step with a missing value first expands (c, C, y_prev)
// back to the full (x, P) with the three identities above and then runs the ordinary path.
//
// The arithmetic lives in __host__ __device__ functions so that tests/native/host_collapsed.cu
// can run exactly this code on the CPU against the oracle.
#pragma once
#include <string.h>

#include "mkf_common.cuh"

namespace mkf_dfm {

// index of (r, c), r <= c, in the packed upper triangle of an N x N matrix
template <int N>
__host__ __device__ constexpr int tri(int r, int c) {
    return r * N - (r * (r - 1)) / 2 + (c - r);
}
template <int N>
__host__ __device__ constexpr int sym(int r, int c) {
    return r <= c ? tri<N>(r, c) : tri<N>(c, r);
}

// index of the multiplier (i, l), i < l, among the M(M-1)/2 strictly-upper pairs
template <int M>
__host__ __device__ constexpr int frozen_pair(int i, int l) {
    return i * M - (i * (i + 1)) / 2 + (l - i - 1);
}

__host__ __device__ __forceinline__ double cs_rcp(double f) {
#ifdef __CUDA_ARCH__
    return mkf::rcp_nr(f);
#else
    return 1.0 / f;
#endif
}
__host__ __device__ __forceinline__ bool cs_bad(double f) {
#ifdef __CUDA_ARCH__
    return mkf::f_is_bad(f);
#else
    return !(f > 0.0) || !(f < 1.7e308) || f < 2.3e-308;
#endif
}

// per-model constants of the collapsed step: [D (M*K)] [S0 (M(M+1)/2, packed upper)] [QG (K*M)]
__host__ __device__ constexpr int collapsed_const_count(int M, int K) { return M * K + M * (M + 1) / 2 + K * M; }

template <int M, int K, class Store>
__host__ __device__ __forceinline__ void collapsed_consts(const double *phi, const double *q, const double *G,
                                                          Store store) {
    constexpr int MT = M * (M + 1) / 2;
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) store(i * K + j, G[i * K + j] * (phi[i] - phi[M + j]));
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int l = i; l < M; ++l) {
            double s = (i == l) ? q[i] : 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) s = fma(G[i * K + j] * q[M + j], G[l * K + j], s);
            store(M * K + tri<M>(i, l), s);
        }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) store(M * K + MT + j * M + i, q[M + j] * G[i * K + j]);
}

// One fully observed step from the collapsed posterior of the previous (fully observed) step.
// In/out: x[M..N) = c, P[cc block] = C (packed upper triangle of the N x N matrix).
// The other entries of x and P are neither read nor written.
// With CAP the reciprocal pivots, the elimination multipliers and the gains of this step are
// handed to cap(slot, value) (slots relative to constant M*K, layout of collapsed_frozen_step).
template <int M, int K, bool CAP, class Load, class Cap>
__host__ __device__ __forceinline__ void collapsed_step_impl(double *x, double *P, const double *phi,
                                                             const double *qc, const double *yp, const double *yv,
                                                             Load ld, Cap cap, double &sigma, double &stepprod,
                                                             bool &bad) {
    constexpr int N = M + K;
    constexpr int MT = M * (M + 1) / 2;
    double D[M * K];
#pragma unroll
    for (int i = 0; i < M * K; ++i) D[i] = ld(i);
    // innovations before any update of this step
    double v[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double s = fma(-phi[i], yp[i], yv[i]);
#pragma unroll
        for (int j = 0; j < K; ++j) s = fma(D[i * K + j], x[M + j], s);
        v[i] = s;
    }
    // W = D C
    double W[M * K];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int l = 0; l < K; ++l) {
            double s = D[i * K] * P[sym<N>(M, M + l)];
#pragma unroll
            for (int j = 1; j < K; ++j) s = fma(D[i * K + j], P[sym<N>(M + j, M + l)], s);
            W[i * K + l] = s;
        }
    // Tm[j][i] = cov(c_pred[j], v[i])
    double Tm[K * M];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int i = 0; i < M; ++i) Tm[j * M + i] = fma(-phi[M + j], W[i * K + j], ld(M * K + MT + j * M + i));
    // S = S0 + W D^T (upper triangle)
    double S[MT];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int l = i; l < M; ++l) {
            double s = ld(M * K + tri<M>(i, l));
#pragma unroll
            for (int j = 0; j < K; ++j) s = fma(W[i * K + j], D[l * K + j], s);
            S[tri<M>(i, l)] = s;
        }
    // predict the factor block
#pragma unroll
    for (int j = 0; j < K; ++j) x[M + j] *= phi[M + j];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = j; l < K; ++l) {
            const double t = (phi[M + j] * P[tri<N>(M + j, M + l)]) * phi[M + l];
            P[tri<N>(M + j, M + l)] = (j == l) ? t + qc[j] : t;
        }
    // pivot by pivot = series by series
#pragma unroll
    for (int i = 0; i < M; ++i) {
        const double f = S[tri<M>(i, i)];
        const double rf = cs_rcp(f);
        bad = bad || cs_bad(f);
        stepprod *= f;
        sigma = fma(v[i] * v[i], rf, sigma);
        if (CAP) cap(i, rf);
#pragma unroll
        for (int l = i + 1; l < M; ++l) {
            const double mlt = S[tri<M>(i, l)] * rf;
            if (CAP) cap(M + frozen_pair<M>(i, l), mlt);
            v[l] = fma(-mlt, v[i], v[l]);
#pragma unroll
            for (int l2 = l; l2 < M; ++l2) S[tri<M>(l, l2)] = fma(-mlt, S[tri<M>(i, l2)], S[tri<M>(l, l2)]);
#pragma unroll
            for (int j = 0; j < K; ++j) Tm[j * M + l] = fma(-mlt, Tm[j * M + i], Tm[j * M + l]);
        }
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double g = Tm[j * M + i] * rf;
            if (CAP) cap(MT + j * M + i, g);
            x[M + j] = fma(g, v[i], x[M + j]);
#pragma unroll
            for (int l = j; l < K; ++l) P[tri<N>(M + j, M + l)] = fma(-g, Tm[l * M + i], P[tri<N>(M + j, M + l)]);
        }
    }
}

template <int M, int K, class Load>
__host__ __device__ __forceinline__ void collapsed_step(double *x, double *P, const double *phi, const double *qc,
                                                        const double *yp, const double *yv, Load ld, double &sigma,
                                                        double &stepprod, bool &bad) {
    collapsed_step_impl<M, K, false>(x, P, phi, qc, yp, yv, ld, [](int, double) {}, sigma, stepprod, bad);
}

// Steady state.  For constant parameters the factor covariance C of consecutive fully observed
// steps converges (Riccati iteration); once it has stopped changing, S, its LDL^T factors and
// the gains are constants and a step is only
//     v = y - phi_s*y_prev + D c;  c = phi_c c;  per series i: sigma += v_i^2/f_i,
//     v_l -= m_il v_i (l > i),  c_j += g_ji v_i
// i.e. exactly the operations of collapsed_step that involve the data, in the same order, with
// 1/f_i, m_il, g_ji captured from the last full step (slots M*K.. of the constant block, over S0
// and QG, which the frozen step does not need): M + M(M-1)/2 + K*M = M(M+1)/2 + K*M values.
template <int M, int K, class Load>
__host__ __device__ __forceinline__ void collapsed_frozen_step(double *x, const double *phi, const double *yp,
                                                               const double *yv, Load ld, double &sigma) {
    constexpr int MT = M * (M + 1) / 2;
    double v[M];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double s = fma(-phi[i], yp[i], yv[i]);
#pragma unroll
        for (int j = 0; j < K; ++j) s = fma(ld(i * K + j), x[M + j], s);
        v[i] = s;
    }
#pragma unroll
    for (int j = 0; j < K; ++j) x[M + j] *= phi[M + j];
#pragma unroll
    for (int i = 0; i < M; ++i) {
        sigma = fma(v[i] * v[i], ld(M * K + i), sigma);
#pragma unroll
        for (int l = i + 1; l < M; ++l) v[l] = fma(-ld(M * K + M + frozen_pair<M>(i, l)), v[i], v[l]);
#pragma unroll
        for (int j = 0; j < K; ++j) x[M + j] = fma(ld(M * K + MT + j * M + i), v[i], x[M + j]);
    }
}

// Convergence monitor of the steady-state switch: one number per model that moves whenever an
// entry of C moves (weights 1, sqrt(2)-ish ... chosen without simple ratios).
template <int M, int K>
__host__ __device__ __forceinline__ double collapsed_checksum(const double *P) {
    constexpr int N = M + K;
    constexpr double w[6] = {1.0, 0.7071067811865476, 0.5773502691896258, 1.618033988749895, 0.36787944117144233,
                             1.2599210498948732};
    double s = 0.0;
    int n = 0;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = j; l < K; ++l) s = fma(w[n++ % 6], P[tri<N>(M + j, M + l)], s);
    return s;
}

// Rebuild the full posterior (x, P) of the last fully observed step from (c, C, y_prev).
template <int M, int K, class Gam>
__host__ __device__ __forceinline__ void collapsed_expand(double *x, double *P, const double *yp, Gam gam) {
    constexpr int N = M + K;
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double s = yp[i];
#pragma unroll
        for (int j = 0; j < K; ++j) s = fma(-gam(i * K + j), x[M + j], s);
        x[i] = s;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double c = 0.0;
#pragma unroll
            for (int l = 0; l < K; ++l) c = fma(-gam(i * K + l), P[sym<N>(M + l, M + j)], c);
            P[tri<N>(i, M + j)] = c;
        }
    }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int l = i; l < M; ++l) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) s = fma(-P[tri<N>(i, M + j)], gam(l * K + j), s);
            P[tri<N>(i, l)] = s;
        }
}

}  // namespace mkf_dfm
