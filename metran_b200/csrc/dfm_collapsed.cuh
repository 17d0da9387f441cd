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

// ---- information form of the same step ------------------------------------------------------------
// The elimination above works on the (M + K) x (M + K) joint matrix [S T^T; T C-] with M pivots.  The
// same step can be taken in the K x K factor space (K <= 3 < M for Metran): with S0 = L0 L0^T,
//     vt = L0^-1 v = L0^-1 z + Dt c                 z = y - phi_s*y_prev,  Dt = L0^-1 D          (M x K)
//     v  = -D e + eps,  e ~ N(0, C) the factor error, eps ~ N(0, S0), cov(w_c, eps) = QG
//     N  = Dt^T Dt = R^T R  (R upper triangular),   b = Dt^T vt
//     C' = (C^-1 + N)^-1 = C - (R C)^T (I + R C R^T)^-1 (R C)          posterior covariance of e
//     e' = -C' b                                                        ... and its mean
//     sigma = v^T S^-1 v = vt.vt + b.e',     det S = det S0 * det(I + R C R^T)
//     c+ = Phi_c c + Jt vt + Ft e',   C+ = Ft C' Ft^T + Qt
//     Jt = QG L0^-T (K x M),  Ft = Phi_c + Jt Dt (K x K),  Qt = Q_c - Jt Jt^T (K x K)
// (w_c = Jt (vt + Dt e) + w~ with w~ independent of eps: the correlated part of the process noise is
// known once v is).  E = I + R C R^T is K x K: K pivots instead of M, 227 instead of 312 fp64
// instructions at (5, 3).  The pivots are NOT the reference's innovation variances any more -- only
// sigma and det S of the whole step are reproduced (to rounding), which is all -2 log L needs; the
// step that captures the steady-state constants (1/f_i, multipliers, gains) is taken in the
// elimination form.
// Constants: [L0^-1 (packed lower, M(M+1)/2)] [Dt (M*K)] [Jt (K*M)] [R (packed upper, K(K+1)/2)]
//            [Ft (K*K)] [Qt (packed upper, K(K+1)/2)] [det S0].
__host__ __device__ constexpr int collapsed_info_const_count(int M, int K) {
    return M * (M + 1) / 2 + 2 * M * K + K * (K + 1) + K * K + 1;
}
// the M-sized part (L0^-1, Dt, Jt): same size as the elimination form's constants
__host__ __device__ constexpr int collapsed_info_smem_count(int M, int K) { return M * (M + 1) / 2 + 2 * M * K; }
template <int N>
__host__ __device__ constexpr int ltri(int r, int c) {   // packed lower triangle, c <= r
    return r * (r + 1) / 2 + c;
}

template <int M, int K, class Store>
__host__ __device__ __forceinline__ void collapsed_info_consts(const double *phi, const double *q, const double *G,
                                                               Store store) {
    constexpr int ML = M * (M + 1) / 2, KT = K * (K + 1) / 2;
    // S0 = Gamma Q_c Gamma^T + Q_s (lower), D, QG
    double L[ML], D[M * K], QG[K * M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int l = 0; l <= i; ++l) {
            double s = (i == l) ? q[i] : 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) s = fma(G[i * K + j] * q[M + j], G[l * K + j], s);
            L[ltri<M>(i, l)] = s;
        }
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            D[i * K + j] = G[i * K + j] * (phi[i] - phi[M + j]);
            QG[j * M + i] = q[M + j] * G[i * K + j];
        }
    // Cholesky S0 = L0 L0^T in place, then W = L0^-1 in place (S0 is positive definite: q_s > 0)
    double inv[M], det = 1.0;
#pragma unroll
    for (int j = 0; j < M; ++j) {
        const double pjj = L[ltri<M>(j, j)];
        det *= pjj;
        const double iv = 1.0 / sqrt(pjj);
        inv[j] = iv;
#pragma unroll
        for (int r = j + 1; r < M; ++r) L[ltri<M>(r, j)] *= iv;
#pragma unroll
        for (int r = j + 1; r < M; ++r)
#pragma unroll
            for (int c = j + 1; c <= r; ++c) L[ltri<M>(r, c)] = fma(-L[ltri<M>(r, j)], L[ltri<M>(c, j)], L[ltri<M>(r, c)]);
    }
#pragma unroll
    for (int j = 0; j < M; ++j) {
        L[ltri<M>(j, j)] = inv[j];
#pragma unroll
        for (int i = j + 1; i < M; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int k = j; k < i; ++k) acc = fma(L[ltri<M>(i, k)], L[ltri<M>(k, j)], acc);
            L[ltri<M>(i, j)] = -inv[i] * acc;
        }
    }
    // Dt = W D, Jt = QG W^T
    double Dt[M * K], Jt[K * M];
#pragma unroll
    for (int i = 0; i < M; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int l = 0; l <= i; ++l) {
                a = fma(L[ltri<M>(i, l)], D[l * K + j], a);
                b = fma(L[ltri<M>(i, l)], QG[j * M + l], b);
            }
            Dt[i * K + j] = a;
            Jt[j * M + i] = b;
        }
    // N = Dt^T Dt = R^T R: Cholesky of N (lower factor = R^T); a direction D does not see (rank of D
    // below K, e.g. equal decay rates) gives a zero pivot: that row of R is zero
    double Nl[KT];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = 0; l <= j; ++l) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < M; ++i) s = fma(Dt[i * K + j], Dt[i * K + l], s);
            Nl[ltri<K>(j, l)] = s;
        }
    double tr = 0.0;
#pragma unroll
    for (int j = 0; j < K; ++j) tr += Nl[ltri<K>(j, j)];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double pjj = Nl[ltri<K>(j, j)];
        const bool ok = pjj > 1e-28 * tr && pjj > 0.0;
        const double iv = ok ? 1.0 / sqrt(pjj) : 0.0;
        Nl[ltri<K>(j, j)] = ok ? pjj * iv : 0.0;
#pragma unroll
        for (int r = j + 1; r < K; ++r) Nl[ltri<K>(r, j)] *= iv;
#pragma unroll
        for (int r = j + 1; r < K; ++r)
#pragma unroll
            for (int c = j + 1; c <= r; ++c) Nl[ltri<K>(r, c)] = fma(-Nl[ltri<K>(r, j)], Nl[ltri<K>(c, j)], Nl[ltri<K>(r, c)]);
    }
    int o = 0;
#pragma unroll
    for (int i = 0; i < ML; ++i) store(o + i, L[i]);
    o += ML;
#pragma unroll
    for (int i = 0; i < M * K; ++i) store(o + i, Dt[i]);
    o += M * K;
#pragma unroll
    for (int i = 0; i < K * M; ++i) store(o + i, Jt[i]);
    o += K * M;
    // R (upper, row-major packed: (j, l), l >= j) = transpose of the lower factor
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = j; l < K; ++l) store(o + tri<K>(j, l), Nl[ltri<K>(l, j)]);
    o += KT;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = 0; l < K; ++l) {
            double s = (j == l) ? phi[M + j] : 0.0;
#pragma unroll
            for (int i = 0; i < M; ++i) s = fma(Jt[j * M + i], Dt[i * K + l], s);
            store(o + j * K + l, s);
        }
    o += K * K;
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = j; l < K; ++l) {
            double s = (j == l) ? q[M + j] : 0.0;
#pragma unroll
            for (int i = 0; i < M; ++i) s = fma(-Jt[j * M + i], Jt[l * M + i], s);
            store(o + tri<K>(j, l), s);
        }
    o += KT;
    store(o, det);
}

// One fully observed step in information form; same in/out contract as collapsed_step_impl (x[M..N) = c,
// P[cc block] = C).  stepprod is multiplied by det S.
template <int M, int K, class Load>
__host__ __device__ __forceinline__ void collapsed_info_step(double *x, double *P, const double *phi, const double *yp,
                                                             const double *yv, Load ld, double &sigma, double &stepprod,
                                                             bool &bad) {
    constexpr int N = M + K;
    constexpr int ML = M * (M + 1) / 2, KT = K * (K + 1) / 2;
    constexpr int oD = ML, oJ = oD + M * K, oR = oJ + K * M, oF = oR + KT, oQ = oF + K * K, oDet = oQ + KT;
    // vt = L0^-1 z + Dt c
    double z[M], vt[M];
#pragma unroll
    for (int i = 0; i < M; ++i) z[i] = fma(-phi[i], yp[i], yv[i]);
#pragma unroll
    for (int i = 0; i < M; ++i) {
        double s = ld(ltri<M>(i, 0)) * z[0];
#pragma unroll
        for (int l = 1; l <= i; ++l) s = fma(ld(ltri<M>(i, l)), z[l], s);
#pragma unroll
        for (int j = 0; j < K; ++j) s = fma(ld(oD + i * K + j), x[M + j], s);
        vt[i] = s;
    }
    double s0 = 0.0, b[K], jv[K];
#pragma unroll
    for (int i = 0; i < M; ++i) s0 = fma(vt[i], vt[i], s0);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double sb = 0.0, sj = 0.0;
#pragma unroll
        for (int i = 0; i < M; ++i) {
            sb = fma(ld(oD + i * K + j), vt[i], sb);
            sj = fma(ld(oJ + j * M + i), vt[i], sj);
        }
        b[j] = sb;
        jv[j] = sj;
    }
    // Gm = R C (K x K), E = I + Gm R^T (packed lower)
    double R[KT], Gm[K * K], E[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) R[i] = ld(oR + i);
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = 0; l < K; ++l) {
            double s = R[tri<K>(j, j)] * P[sym<N>(M + j, M + l)];
#pragma unroll
            for (int k = j + 1; k < K; ++k) s = fma(R[tri<K>(j, k)], P[sym<N>(M + k, M + l)], s);
            Gm[j * K + l] = s;
        }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = 0; l <= j; ++l) {
            double s = (j == l) ? 1.0 : 0.0;
#pragma unroll
            for (int k = l; k < K; ++k) s = fma(Gm[j * K + k], R[tri<K>(l, k)], s);   // R[l][k] = 0 for k < l
            E[ltri<K>(j, l)] = s;
        }
    // E = Le diag(e) Le^T; U = Le^-1 Gm; C' = C - U^T diag(1/e) U
    double re[K], det = ld(oDet);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const double e = E[ltri<K>(j, j)];
        bad = bad || cs_bad(e);
        det *= e;
        re[j] = cs_rcp(e);
#pragma unroll
        for (int r = j + 1; r < K; ++r) {
            const double mlt = E[ltri<K>(r, j)] * re[j];
#pragma unroll
            for (int c = j + 1; c <= r; ++c) E[ltri<K>(r, c)] = fma(-mlt, E[ltri<K>(c, j)], E[ltri<K>(r, c)]);
#pragma unroll
            for (int l = 0; l < K; ++l) Gm[r * K + l] = fma(-mlt, Gm[j * K + l], Gm[r * K + l]);
        }
    }
    stepprod *= det;
    double Cp[KT];   // packed upper of the K x K block, index tri<K>
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = j; l < K; ++l) {
            double s = P[tri<N>(M + j, M + l)];
#pragma unroll
            for (int k = 0; k < K; ++k) s = fma(-(Gm[k * K + j] * re[k]), Gm[k * K + l], s);
            Cp[tri<K>(j, l)] = s;
        }
    // e' = -C' b; sigma; c+ = Phi_c c + Jt vt + Ft e'
    double ep[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double s = 0.0;
#pragma unroll
        for (int l = 0; l < K; ++l) s = fma(-Cp[sym<K>(j, l)], b[l], s);
        ep[j] = s;
    }
    double sg = s0;
#pragma unroll
    for (int j = 0; j < K; ++j) sg = fma(b[j], ep[j], sg);
    sigma += sg;
    double Ft[K * K];
#pragma unroll
    for (int i = 0; i < K * K; ++i) Ft[i] = ld(oF + i);
#pragma unroll
    for (int j = 0; j < K; ++j) {
        double s = fma(phi[M + j], x[M + j], jv[j]);
#pragma unroll
        for (int l = 0; l < K; ++l) s = fma(Ft[j * K + l], ep[l], s);
        x[M + j] = s;
    }
    // C+ = Ft C' Ft^T + Qt
    double Tm[K * K];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = 0; l < K; ++l) {
            double s = Ft[j * K] * Cp[sym<K>(0, l)];
#pragma unroll
            for (int k = 1; k < K; ++k) s = fma(Ft[j * K + k], Cp[sym<K>(k, l)], s);
            Tm[j * K + l] = s;
        }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int l = j; l < K; ++l) {
            double s = ld(oQ + tri<K>(j, l));
#pragma unroll
            for (int k = 0; k < K; ++k) s = fma(Tm[j * K + k], Ft[l * K + k], s);
            P[tri<N>(M + j, M + l)] = s;
        }
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
