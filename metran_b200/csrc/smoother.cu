// Batched Rauch-Tung-Striebel smoother and state projections (n <= 32).
//
// mkf_smoother replaces kalmansmoother (reference metran/kalmanfilter.py:403-476):
//   G_t    = P_f[t] A^T P_p[t+1]^-1
//   x_s[t] = x_f[t] + G_t (x_s[t+1] - x_p[t+1])
//   P_s[t] = P_f[t] + G_t (P_s[t+1] - P_p[t+1]) G_t^T
// The reference forms pinv(P_p[t+1]) by SVD (:455); P_p = A P A^T + Q is symmetric
// positive definite for every Metran model (q_i > 0), so this kernel solves
// P_p G^T = A P_f with a Cholesky factorisation instead (cond(P_p) <= ~1e2 on
// Metran-like models, SURVEY.md s7: the two agree to ~1e-14).
//
// Mapping: NP = 4/8/16/32 lanes per model, lane r owns row r.  Per backward step
//   1. right-looking Cholesky of P_p, one column per iteration, column broadcast through
//      shared memory;
//   2. lane c solves L L^T g = (A P_f)[:,c] entirely in registers with L broadcast from
//      shared memory -> lane c holds row c of G;
//   3. two n x n x n products with the shared-memory operand broadcast to all lanes.
// Covariances are read/written column-per-lane (coalesced); they are symmetric up to
// rounding, see filter_generic.cu.
//
// mkf_project replaces the per-time-step Python loop of SPKalmanFilter.simulate
// (metran/kalmanfilter.py:596-602).
#include <type_traits>

#include "mkf_common.cuh"

namespace {

using namespace mkf;

struct SmootherArgs {
    int64_t B, T;
    int n;
    const double *fm, *fc, *pm, *pc, *A;
    int64_t A_stride;
    unsigned flags;
    double *sm, *sc;
    int32_t *status;
    int smem_model;
    // fused epilogue (mkf_smoother_project): projections of the smoothed moments, written while the
    // covariance of the step is still on chip
    int p;                      // projected series (0 = none)
    const double *Hp;           // [p][n] per model
    int64_t Hp_stride;
    double *pmean, *pvar;       // [B][T][p]
    double *svar;               // [B][T][n] diag of the smoothed covariance (may be NULL)
};

__device__ __forceinline__ int64_t base_of(int64_t b, int64_t T) { return b * T; }

template <int NP>
__global__ void __launch_bounds__(NP == 32 ? 64 : 128) smoother_kernel(const SmootherArgs a) {
    extern __shared__ __align__(16) double smem[];
    constexpr int LD = NP + 2;  // padded leading dimension (even: rows stay double2-aligned)
    const int n = a.n;
    const int group_in_block = threadIdx.x / NP;
    const int r = threadIdx.x % NP;
    const int group_in_warp = (threadIdx.x & 31) / NP;
    const unsigned gmask = (NP == 32) ? 0xffffffffu : (((1u << NP) - 1u) << (group_in_warp * NP));
    const int64_t b = (int64_t)blockIdx.x * (blockDim.x / NP) + group_in_block;
    if (b >= a.B) return;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;

    // Only one n-vector of registers is live at a time (L, then z, then M); every matrix
    // that must survive a phase lives in shared memory, rows owned by lanes.
    double *sm_ = smem + (size_t)group_in_block * a.smem_model;
    double *sL = sm_;   sm_ += NP * LD;   // Cholesky factor, row-major
    double *sD = sm_;   sm_ += NP * LD;   // P_s[t+1] - P_p[t+1]
    double *sG = sm_;   sm_ += NP * LD;   // smoothing gain G_t
    double *sPs = sm_;  sm_ += NP * LD;   // P_s[t+1] (rows), carried between steps
    double *sPf = sm_;  sm_ += NP * LD;   // P_f[t]
    double *sv = sm_;   sm_ += NP;        // broadcast vector
    double *sa = sm_;   sm_ += NP;        // diag(A)
    double *sinv = sm_; sm_ += NP;        // 1 / L_ii
    double *sHp = sm_;  sm_ += (size_t)a.p * NP;   // projection matrix rows (fused epilogue)
    double *sA = sm_;                     // full A (only when not diagonal)

    const double *Ab = a.A + b * a.A_stride;
    sa[r] = (r < n) ? Ab[r * n + r] : 0.0;
    if (!diagA)
        for (int i = r; i < NP * NP; i += NP) {
            int rr = i / NP, c = i % NP;
            sA[rr * LD + c] = (rr < n && c < n) ? Ab[rr * n + c] : 0.0;
        }
    for (int i = r; i < a.p * NP; i += NP) {
        const int j = i / NP, c = i % NP;
        sHp[i] = (c < n) ? a.Hp[b * a.Hp_stride + j * n + c] : 0.0;
    }
    // Fused projections (SPKalmanFilter.simulate, kalmanfilter.py:596-602): H x_s[t] and
    // max(diag(H P_s[t] H^T), 0) from the row of P_s this lane holds in shared memory.
    auto project = [&](int64_t t, double xs_r) {
        if (a.p <= 0) return;
        if (a.svar && r < n) a.svar[(base_of(b, a.T) + t) * n + r] = sPs[r * LD + r];
        for (int j = 0; j < a.p; ++j) {
            const double *h = sHp + j * NP;
            double u = 0.0;
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                const double2 pr = *reinterpret_cast<const double2 *>(sPs + r * LD + c);
                const double2 hc = *reinterpret_cast<const double2 *>(h + c);
                u = fma(pr.x, hc.x, u);
                u = fma(pr.y, hc.y, u);
            }
            const double hr = h[r];
            const double var = group_sum<NP>(hr * u, gmask);
            const double mu = group_sum<NP>(hr * xs_r, gmask);
            if (r == 0) {
                a.pmean[(base_of(b, a.T) + t) * a.p + j] = mu;
                a.pvar[(base_of(b, a.T) + t) * a.p + j] = fmax(var, 0.0);
            }
        }
    };

    const int64_t T = a.T;
    const int64_t base = b * T;
    const bool rv = r < n;
    int status = 0;

    // t = T-1: smoothed = filtered (kalmanfilter.py:450-451)
    double xs = rv ? a.fm[(base + T - 1) * n + r] : 0.0;
    {
        const double *src = a.fc + (base + T - 1) * n * n + r;
        double *dst = a.sc + (base + T - 1) * n * n + r;
#pragma unroll
        for (int c = 0; c < NP; ++c) {
            const double v = (rv && c < n) ? src[(int64_t)c * n] : 0.0;
            sPs[r * LD + c] = v;
            if (rv && c < n && a.sc) dst[(int64_t)c * n] = v;
        }
        if (rv) a.sm[(base + T - 1) * n + r] = xs;
    }
    __syncwarp(gmask);
    project(T - 1, xs);

    // For NP <= 8 the inputs of step t-1 are fetched into registers while step t computes
    // (global-load latency off the critical path); at NP >= 16 the extra registers spill.
    constexpr bool PREFETCH = NP <= 16;
    constexpr int NPF = PREFETCH ? NP : 1;
    double nL[NPF], nF[NPF], nxp = 0.0, nxf = 0.0;
    auto fetch = [&](int64_t t, double *Lr, double *Fr, double &xp_, double &xf_) {
        const double *srcp = a.pc + (base + t + 1) * n * n + r;
        const double *srcf = a.fc + (base + t) * n * n + r;
#pragma unroll
        for (int c = 0; c < NPF; ++c) {
            Lr[c] = (rv && c < n) ? srcp[(int64_t)c * n] : ((c == r) ? 1.0 : 0.0);
            Fr[c] = (rv && c < n) ? srcf[(int64_t)c * n] : 0.0;
        }
        xp_ = rv ? a.pm[(base + t + 1) * n + r] : 0.0;
        xf_ = rv ? a.fm[(base + t) * n + r] : 0.0;
    };
    if (PREFETCH && T >= 2) fetch(T - 2, nL, nF, nxp, nxf);

    for (int64_t t = T - 2; t >= 0; --t) {
        // ---- P_p[t+1] row, D = P_s[t+1] - P_p[t+1], delta ------------------------------------
        double L[NP];  // row r of P_p[t+1], becomes row r of its Cholesky factor
        double Fcur[NPF];
        double xp, xf;
        if (PREFETCH) {
#pragma unroll
            for (int c = 0; c < NPF; ++c) {
                L[c] = nL[c];
                Fcur[c] = nF[c];
            }
            xp = nxp;
            xf = nxf;
            if (t > 0) fetch(t - 1, nL, nF, nxp, nxf);
        } else {
            const double *src = a.pc + (base + t + 1) * n * n + r;
#pragma unroll
            for (int c = 0; c < NP; ++c) L[c] = (rv && c < n) ? src[(int64_t)c * n] : ((c == r) ? 1.0 : 0.0);
            xp = rv ? a.pm[(base + t + 1) * n + r] : 0.0;
            xf = rv ? a.fm[(base + t) * n + r] : 0.0;
        }
#pragma unroll
        for (int c = 0; c < NP; c += 2) {
            const double2 ps = *reinterpret_cast<const double2 *>(sPs + r * LD + c);
            const double p0 = (rv && c < n) ? L[c] : 0.0, p1 = (rv && c + 1 < n) ? L[c + 1] : 0.0;
            *reinterpret_cast<double2 *>(sD + r * LD + c) = make_double2(ps.x - p0, ps.y - p1);
        }
        const double delta = xs - xp;
        double d0 = 0.0;                         // original diagonal entry of this lane's row
#pragma unroll
        for (int c = 0; c < NP; ++c) d0 = (c == r) ? L[c] : d0;
        double dmax = rv ? d0 : 0.0;             // largest diagonal entry: scale of pinv's cut-off
#pragma unroll
        for (int off = NP / 2; off > 0; off >>= 1) dmax = fmax(dmax, __shfl_xor_sync(gmask, dmax, off, NP));

        // ---- Cholesky P_p = L L^T (padded rows/cols are the identity) ---------------------
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const double pjj = __shfl_sync(gmask, L[j], j, NP);
            // A pivot that is not positive relative to its original diagonal entry belongs to a direction in
            // which P_p is (numerically) singular -- a state without process noise that the data determine
            // exactly (communality 1 with R = 0).  The reference's pinv (kalmanfilter.py:455) assigns such a
            // direction zero gain (singular values below 1e-15 of the largest); zeroing the column of L and
            // its reciprocal does the same here, and the model is flagged instead of returning NaN.
            const double d0j = __shfl_sync(gmask, d0, j, NP);
            const bool okp = pjj > 1.5e-14 * d0j && pjj > 1e-15 * dmax && pjj > 0.0;
            if (!okp) status |= MKF_STATUS_CHOL_FAILED;
            const double inv = okp ? rsqrt(pjj) : 0.0;
            const double ljj = pjj * inv;
            const double lrj = (r == j) ? ljj : ((r > j) ? L[j] * inv : 0.0);
            L[j] = lrj;
            if (r == j) sinv[j] = inv;
            if (j + 1 < NP) {
                sv[r] = lrj;
                __syncwarp(gmask);
#pragma unroll
                for (int c = j + 1; c < NP; ++c) L[c] = fma(-lrj, sv[c], L[c]);
                __syncwarp(gmask);
            }
        }
#pragma unroll
        for (int c = 0; c < NP; c += 2)
            *reinterpret_cast<double2 *>(sL + r * LD + c) =
                make_double2((c <= r) ? L[c] : 0.0, (c + 1 <= r) ? L[c + 1] : 0.0);
        sv[r] = delta;

        // ---- lane c: solve (L L^T) g = (A P_f)[:, c]  ->  g = row c of G ----------------
        double z[NP];
        {
            const double *src = a.fc + (base + t) * n * n + r;
#pragma unroll
            for (int i = 0; i < NP; ++i)
                z[i] = PREFETCH ? Fcur[PREFETCH ? i : 0] : ((rv && i < n) ? src[(int64_t)i * n] : 0.0);
#pragma unroll
            for (int i = 0; i < NP; i += 2)
                *reinterpret_cast<double2 *>(sPf + r * LD + i) = make_double2(z[i], z[i + 1]);
        }
        __syncwarp(gmask);
        if (diagA) {
#pragma unroll
            for (int i = 0; i < NP; ++i) z[i] *= sa[i];
        } else {
#pragma unroll
            for (int i = 0; i < NP; ++i) z[i] = 0.0;
            for (int k = 0; k < n; ++k) {
                const double pk = sPf[r * LD + k];
#pragma unroll
                for (int i = 0; i < NP; ++i) z[i] = fma(sA[i * LD + k], pk, z[i]);
            }
        }
        // The factor is read through a volatile pointer in the (fully unrolled) substitutions:
        // ptxas otherwise hoists hundreds of shared-memory loads ahead of their use and spills
        // (4.5 KB of local memory per thread at NP = 32).
        typedef typename std::conditional<(NP >= 16), const volatile double, const double>::type LRead;
        LRead *vL = sL;
#pragma unroll
        for (int k = 0; k < NP; ++k) {  // forward substitution, column oriented
            z[k] *= sinv[k];
#pragma unroll
            for (int i = k + 1; i < NP; ++i) z[i] = fma(-vL[i * LD + k], z[k], z[i]);
        }
#pragma unroll
        for (int k = NP - 1; k >= 0; --k) {  // backward substitution with L^T
            z[k] *= sinv[k];
#pragma unroll
            for (int i = 0; i < k; ++i) z[i] = fma(-vL[k * LD + i], z[k], z[i]);
        }
        // rows/cols >= n: P_f row is zero -> z = 0 there.

        // ---- x_s[t] = x_f + G delta -------------------------------------------------------
        double xnew = xf;
#pragma unroll
        for (int i = 0; i < NP; i += 2) {
            double2 dv = *reinterpret_cast<const double2 *>(sv + i);
            xnew = fma(z[i], dv.x, xnew);
            xnew = fma(z[i + 1], dv.y, xnew);
        }
#pragma unroll
        for (int i = 0; i < NP; i += 2)
            *reinterpret_cast<double2 *>(sG + r * LD + i) = make_double2(z[i], z[i + 1]);
        __syncwarp(gmask);

        // ---- M = G D (row r) -----------------------------------------------------------------
        double Mrow[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) Mrow[j] = 0.0;
        for (int k = 0; k < n; ++k) {
            const double zk = sG[r * LD + k];
#pragma unroll
            for (int j = 0; j < NP; j += 2) {
                double2 dk = *reinterpret_cast<const double2 *>(sD + k * LD + j);
                Mrow[j] = fma(zk, dk.x, Mrow[j]);
                Mrow[j + 1] = fma(zk, dk.y, Mrow[j + 1]);
            }
        }
        __syncwarp(gmask);  // all reads of sPs/sD done before sPs is rewritten
        // ---- P_s[t] = P_f + M G^T : element (r, j), stored column-per-lane ----------------
        double *dst = a.sc + (base + t) * n * n + r;
        for (int j = 0; j < n; ++j) {
            double s0 = sPf[r * LD + j], s1 = 0.0;
#pragma unroll
            for (int k = 0; k < NP; k += 4) {
                double2 g0 = *reinterpret_cast<const double2 *>(sG + j * LD + k);
                double2 g1 = *reinterpret_cast<const double2 *>(sG + j * LD + k + 2);
                s0 = fma(Mrow[k], g0.x, s0);
                s1 = fma(Mrow[k + 1], g0.y, s1);
                s0 = fma(Mrow[k + 2], g1.x, s0);
                s1 = fma(Mrow[k + 3], g1.y, s1);
            }
            const double s = s0 + s1;
            sPs[r * LD + j] = s;
            if (rv && a.sc) dst[(int64_t)j * n] = s;
        }
        xs = xnew;
        if (rv) a.sm[(base + t) * n + r] = xs;
        __syncwarp(gmask);  // shared buffers are rewritten next step
        project(t, xs);
    }
    if (r == 0 && a.status) a.status[b] = status;
}

template <int NP>
int launch_smoother(SmootherArgs a, cudaStream_t stream) {
    constexpr int LD = NP + 2;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    int per = 5 * NP * LD + 3 * NP + a.p * NP + (diagA ? 0 : NP * LD);
    per = (per + 1) & ~1;
    a.smem_model = per;
    int threads = (NP == 32) ? 64 : 128;
    size_t smem = (size_t)(threads / NP) * per * sizeof(double);
    auto kern = smoother_kernel<NP>;
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

#include "smoother_wide.cuh"
#include "smoother_thread.cuh"

namespace {

// ---- projections ---------------------------------------------------------------------
// one warp per (model, time step): stage x_t and P_t in shared memory, lane i computes
// output series i (strided when p > 32).
__global__ void __launch_bounds__(128) project_kernel(int64_t B, int64_t T, int n, int p,
                                                      const double *__restrict__ means,
                                                      const double *__restrict__ covs,
                                                      const double *__restrict__ Hp, int64_t Hp_stride,
                                                      double *__restrict__ mout, double *__restrict__ vout) {
    extern __shared__ __align__(16) double smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *sP = smem + (size_t)warp * (n * n + n);
    double *sx = sP + n * n;
    const int64_t total = B * T;
    for (int64_t bt = (int64_t)blockIdx.x * 4 + warp; bt < total; bt += (int64_t)gridDim.x * 4) {
        const int64_t b = bt / T;
        __syncwarp();
        for (int i = lane; i < n * n; i += 32) sP[i] = covs ? covs[bt * n * n + i] : 0.0;
        for (int i = lane; i < n; i += 32) sx[i] = means[bt * n + i];
        __syncwarp();
        const double *Hb = Hp + b * Hp_stride;
        for (int i = lane; i < p; i += 32) {
            const double *h = Hb + (int64_t)i * n;
            double mu = 0.0, var = 0.0;
            for (int c = 0; c < n; ++c) {
                const double hc = h[c];
                mu = fma(hc, sx[c], mu);
                double s = 0.0;
                for (int k = 0; k < n; ++k) s = fma(sP[c * n + k], h[k], s);
                var = fma(hc, s, var);
            }
            if (mout) mout[bt * p + i] = mu;
            if (vout) vout[bt * p + i] = fmax(var, 0.0);  // kalmanfilter.py:601-602
        }
    }
}

}  // namespace

static int smoother_entry(int64_t B, int64_t T, int32_t n, const double *fm, const double *fc, const double *pm,
                          const double *pc, const double *A, int64_t A_stride, uint32_t flags, int32_t p,
                          const double *Hp, int64_t Hp_stride, double *sm, double *sc, double *proj_mean,
                          double *proj_var, double *state_var, int32_t *status, void *stream) {
    if (B < 0 || T < 0 || n <= 0 || p < 0) return MKF_ERR_BAD_ARG;
    if (!fm || !fc || !pm || !pc || !A || !sm) return MKF_ERR_BAD_ARG;
    if (p == 0 && !sc) return MKF_ERR_BAD_ARG;
    if (p > 0 && (!Hp || !proj_mean || !proj_var)) return MKF_ERR_BAD_ARG;
    if (n > 32 || p > 256) return MKF_ERR_UNSUPPORTED;
    if (B == 0 || T == 0) return MKF_OK;
    SmootherArgs a{};
    a.B = B; a.T = T; a.n = n; a.fm = fm; a.fc = fc; a.pm = pm; a.pc = pc;
    a.A = A; a.A_stride = A_stride; a.flags = flags; a.sm = sm; a.sc = sc; a.status = status;
    a.p = p; a.Hp = Hp; a.Hp_stride = Hp_stride; a.pmean = proj_mean; a.pvar = proj_var; a.svar = state_var;
    cudaStream_t s = (cudaStream_t)stream;
    if (p == 0 && (n == 4 || n == 8)) {     // big batches of small diagonal-A models: one thread per model
        const int rc = launch_smoother_thread(a, s);
        if (rc != MKF_ERR_UNSUPPORTED) return rc;
    }
    if (n <= 4) return launch_smoother<4>(a, s);
    if (n <= 8) return launch_smoother<8>(a, s);
    if (n <= 16) return launch_smoother<16>(a, s);
#ifndef MKF_SMOOTHER_WIDE
#define MKF_SMOOTHER_WIDE 1
#endif
    if (MKF_SMOOTHER_WIDE) return launch_smoother_wide(a, s);      // four warps per model (smoother_wide.cuh)
    return launch_smoother<32>(a, s);
}

extern "C" int64_t mkf_smoother_set_thread_min_batch(int64_t min_batch) {
    return (int64_t)g_smt_min_b.exchange(min_batch < 0 ? (long long)MKF_SMT_MIN_B : (long long)min_batch);
}

extern "C" int mkf_smoother(int64_t B, int64_t T, int32_t n, const double *fm, const double *fc,
                            const double *pm, const double *pc, const double *A, int64_t A_stride,
                            uint32_t flags, double *sm, double *sc, int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_smoother");
    return smoother_entry(B, T, n, fm, fc, pm, pc, A, A_stride, flags, 0, nullptr, 0, sm, sc, nullptr, nullptr,
                          nullptr, status, stream);
}

extern "C" int mkf_smoother_project(int64_t B, int64_t T, int32_t n, const double *fm, const double *fc,
                                    const double *pm, const double *pc, const double *A, int64_t A_stride,
                                    uint32_t flags, int32_t p, const double *Hp, int64_t Hp_stride, double *sm,
                                    double *sc, double *proj_mean, double *proj_var, double *state_var,
                                    int32_t *status, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_smoother_project");
    if (p <= 0) return MKF_ERR_BAD_ARG;
    return smoother_entry(B, T, n, fm, fc, pm, pc, A, A_stride, flags, p, Hp, Hp_stride, sm, sc, proj_mean, proj_var,
                          state_var, status, stream);
}

extern "C" int mkf_project(int64_t B, int64_t T, int32_t n, int32_t p, const double *means,
                           const double *covs, const double *Hp, int64_t Hp_stride, double *means_out,
                           double *var_out, void *stream) {
    mkf_host::NvtxRange nvtx_range("mkf_project");
    if (B < 0 || T < 0 || n <= 0 || p <= 0 || !means || !Hp) return MKF_ERR_BAD_ARG;
    if (var_out && !covs) return MKF_ERR_BAD_ARG;
    if (n > 64) return MKF_ERR_UNSUPPORTED;
    if (B == 0 || T == 0) return MKF_OK;
    const int64_t total = B * T;
    int64_t blocks = (total + 3) / 4;
    if (blocks > 148 * 16) blocks = 148 * 16;
    size_t smem = 4 * (size_t)(n * n + n) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    project_kernel<<<(unsigned)blocks, 128, smem, (cudaStream_t)stream>>>(B, T, n, p, means, covs, Hp, Hp_stride,
                                                                         means_out, var_out);
    return mkf_host::check_launch();
}
