// RTS smoother for wide models (16 < n <= 32): one CTA of four warps per model.
//
// Reference: kalmansmoother, metran/kalmanfilter.py:453-474 (see smoother.cu for the formulas and
// for the Cholesky-instead-of-pinv argument).  smoother_kernel<32> runs a model on one warp: a
// 32-column Cholesky with a broadcast and two warp barriers per column, 32 + 32 dependent
// substitution steps per lane and two 32^3 products, all serial in one instruction stream
// (C3: ~36 000 cycles per time step, the global loads of the step not even prefetched).
//
// Here 128 threads share the step.  Thread (w, l): warp w = 0..3, lane l.
//   0. P_p[t+1], P_f[t] arrive by cp.async (issued one step ahead) and are spread into padded
//      shared-memory matrices; D = P_s[t+1] - P_p[t+1].
//   1. P_p^-1 in place by Gauss-Jordan elimination without pivoting: thread (w, l) keeps its 8 entries
//      of row l in registers, column k and row k of the current matrix travel through two small
//      double-buffered vectors, two pivots per block barrier (what the first does to the second's row and
//      column every thread works out itself): 16 barriers and 32 x 8 + 16 x 9 FMAs per thread (round 2;
//      the blocked Cholesky + 64-step substitutions of the first version took twice as long).
//   2. G = (A P_f)^T P_p^-1 as a 32 x 32 x 32 product with the operands in shared memory.
//   3. x_s[t], M = G D, P_s[t] = P_f + M G^T: thread (w, l) produces row l, columns [8w, 8w+8).
#pragma once

namespace {

#ifndef MKF_SW_PW
#define MKF_SW_PW 8          // columns of a row per thread: 8 = four warps per model, 4 = eight warps
#endif
constexpr int kSwWarps = 32 / MKF_SW_PW;
constexpr int kSwThreads = 32 * kSwWarps;

__global__ void __launch_bounds__(kSwThreads, 2) smoother_wide_kernel(const SmootherArgs a) {
    constexpr int NP = 32, LD = NP + 2, PW = MKF_SW_PW;       // PW: columns per warp
    extern __shared__ __align__(16) double smem[];
    const int n = a.n;
    const int tid = threadIdx.x, l = tid & 31, w = tid >> 5, c0 = w * PW;
    const int64_t b = blockIdx.x;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;

    double *sm_ = smem;
    double *sL = sm_;    sm_ += NP * LD;   // P_p[t+1] -> its Cholesky factor (lower)
    double *sPf = sm_;   sm_ += NP * LD;   // P_f[t]
    double *sD = sm_;    sm_ += NP * LD;   // P_s[t+1] - P_p[t+1]
    double *sG = sm_;    sm_ += NP * LD;   // smoothing gain, row-major
    double *sM = sm_;    sm_ += NP * LD;   // G D
    double *sPs = sm_;   sm_ += NP * LD;   // P_s[t+1], carried between steps
    double *sdel = sm_;  sm_ += NP;        // x_s[t+1] - x_p[t+1]
    double *sa = sm_;    sm_ += NP;        // diag(A)
    double *sxs = sm_;   sm_ += NP;        // x_s[t]
    double *sdg = sm_;   sm_ += NP;        // original diagonal of P_p[t+1] (singularity test of the pivots)
    double *scol = sm_;  sm_ += 4 * NP;    // Gauss-Jordan: columns k, k+1 of the current matrix (double-buffered)
    double *srow = sm_;  sm_ += 4 * NP;    // ... and its rows k, k+1
    double *spart = sm_; sm_ += kSwWarps * NP;    // partial sums of G delta, one per warp
    double *stage = sm_; sm_ += 2 * 2 * NP * LD;   // [buffer][P_p | P_f][row][LD]
    double *sHp = sm_;   sm_ += (size_t)a.p * NP;  // projection matrix rows (fused epilogue)
    double *sA = sm_;                      // full A (only when not diagonal)

    const double *Ab = a.A + b * a.A_stride;
    if (tid < NP) sa[tid] = (tid < n) ? Ab[tid * n + tid] : 0.0;
    if (!diagA)
        for (int i = tid; i < NP * NP; i += kSwThreads) {
            const int rr = i / NP, c = i % NP;
            sA[rr * LD + c] = (rr < n && c < n) ? Ab[rr * n + c] : 0.0;
        }
    const int64_t T = a.T;
    const int64_t base = b * T;
    const int nn = n * n;
    int status = 0;
    for (int i = tid; i < a.p * NP; i += kSwThreads) {
        const int j = i / NP, c = i % NP;
        sHp[i] = (c < n) ? a.Hp[b * a.Hp_stride + j * n + c] : 0.0;
    }
    // fused projections of the smoothed moments (kalmanfilter.py:596-602) from sPs / sxs: warp w takes
    // the series w, w + 4, ...; call between two block barriers
    auto project = [&](int64_t t) {
        if (a.p <= 0) return;
        if (a.svar && tid < n) a.svar[(base + t) * n + tid] = sPs[tid * LD + tid];
        for (int j = w; j < a.p; j += kSwWarps) {
            const double *h = sHp + j * NP;
            double u = 0.0;
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                const double2 pr = *reinterpret_cast<const double2 *>(sPs + l * LD + c);
                const double2 hc = *reinterpret_cast<const double2 *>(h + c);
                u = fma(pr.x, hc.x, u);
                u = fma(pr.y, hc.y, u);
            }
            const double hr = h[l];
            const double var = group_sum<32>(hr * u);
            const double mu = group_sum<32>(hr * sxs[l]);
            if (l == 0) {
                a.pmean[(base + t) * a.p + j] = mu;
                a.pvar[(base + t) * a.p + j] = fmax(var, 0.0);
            }
        }
    };

    // The staged matrices keep the padded row pitch LD of the working matrices: read back row-per-lane
    // at the global pitch n = 32 every lane would hit the same bank.
    const int srow0 = tid / n, scol0 = tid % n, srow_inc = kSwThreads / n, scol_inc = kSwThreads % n;
    auto issue_step = [&](int64_t t, int buf) {     // inputs of backward step t: P_p[t+1], P_f[t]
        if (t >= 0) {
            const double *srcp = a.pc + (base + t + 1) * nn;
            const double *srcf = a.fc + (base + t) * nn;
            double *dst = stage + (size_t)buf * 2 * NP * LD;
            int rr = srow0, cc = scol0;
            for (int i = tid; i < nn; i += kSwThreads) {
                cp_async8(dst + rr * LD + cc, srcp + i);
                cp_async8(dst + NP * LD + rr * LD + cc, srcf + i);
                rr += srow_inc;
                cc += scol_inc;
                if (cc >= n) {
                    cc -= n;
                    ++rr;
                }
            }
        }
        cp_async_commit();
    };

    // t = T-1: smoothed = filtered (kalmanfilter.py:450-451)
    for (int i = tid; i < NP * NP; i += kSwThreads) {
        const int rr = i / NP, c = i % NP;
        double v = 0.0;
        if (rr < n && c < n) {
            v = a.fc[(base + T - 1) * nn + rr * n + c];
            if (a.sc) a.sc[(base + T - 1) * nn + rr * n + c] = v;
        }
        sPs[rr * LD + c] = v;
    }
    if (tid < NP) {
        const double v = (tid < n) ? a.fm[(base + T - 1) * n + tid] : 0.0;
        sxs[tid] = v;
        if (tid < n) a.sm[(base + T - 1) * n + tid] = v;
    }
    issue_step(T - 2, 0);
    __syncthreads();
    project(T - 1);

    int buf = 0;
    for (int64_t t = T - 2; t >= 0; --t, buf ^= 1) {
        issue_step(t - 1, buf ^ 1);
        cp_async_wait<1>();
        __syncthreads();
        // ---- 0. spread the staged matrices; D; delta -----------------------------------------
        {
            const double *sp = stage + (size_t)buf * 2 * NP * LD, *sf = sp + NP * LD;
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                const int cc = c0 + c;
                const bool in = l < n && cc < n;
                const double pp = in ? sp[l * LD + cc] : ((l == cc) ? 1.0 : 0.0);   // padding: identity
                const double pf = in ? sf[l * LD + cc] : 0.0;
                sL[l * LD + cc] = pp;
                if (l == cc) sdg[l] = pp;
                sPf[l * LD + cc] = pf;
                sD[l * LD + cc] = in ? sPs[l * LD + cc] - pp : 0.0;
            }
            if (tid < NP) sdel[tid] = (tid < n) ? sxs[tid] - a.pm[(base + t + 1) * n + tid] : 0.0;
        }
        __syncthreads();

        // ---- 1. P_p^-1 in place: Gauss-Jordan elimination without pivoting (P_p is symmetric positive
        // definite; the pivots are the squares of the Cholesky pivots, so the singularity rule of
        // smoother.cu carries over unchanged).  Thread (w, l) keeps its 8 entries of row l in registers
        // for all 32 pivots; per pivot only column k (from the warp that owns it) and row k (from lane k of
        // every warp) are published through two small double-buffered vectors: ONE block barrier per
        // pivot, 8 FMAs per thread, no triangular solves afterwards -- the blocked Cholesky with its
        // redundant 8 x 8 factorisations, three barriers per panel and 64 dependent substitution steps per
        // right-hand side (four lanes per column, a shuffle per step) cost three times as much.
        double dmax = (l < n) ? sdg[l] : 0.0;     // largest diagonal entry: scale of pinv's cut-off
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, off));
        double pr[PW];
#pragma unroll
        for (int c = 0; c < PW; ++c) pr[c] = sL[l * LD + c0 + c];
        // TWO pivots per barrier: rows k, k+1 and columns k, k+1 are published as they are before pivot k;
        // what pivot k does to row k+1 and to column k+1 every thread works out for itself (8 + 1 extra
        // FMAs), so that pivot k+1 follows without another exchange -- the same operations on the same
        // numbers as one pivot per round (bit-identical results), half the barriers.
#pragma unroll
        for (int k = 0; k < NP; k += 2) {
            double *colk = scol + ((k >> 1) & 1) * 2 * NP, *rowk = srow + ((k >> 1) & 1) * 2 * NP;
            if (w == k / PW) {               // (k and k + 1 belong to the same warp: PW is even)
                colk[l] = pr[k % PW];
                colk[NP + l] = pr[(k + 1) % PW];
            }
            if (l == k || l == k + 1) {
                double *dst = rowk + (l - k) * NP + c0;
#pragma unroll
                for (int c = 0; c < PW; c += 2) *reinterpret_cast<double2 *>(dst + c) = make_double2(pr[c], pr[c + 1]);
            }
            __syncthreads();
            // ---- pivot k.  Row k becomes (row k) * pv, every other row l: a_lj - (a_lk pv) a_kj; with
            // clp = a_lk pv (= -pv in row k, whose own entries then start from zero) both are ONE fma per
            // entry, and the entry in column k is -clp in either case.
            const double d0 = rowk[k];
            // numerically singular direction: zero gain, as the reference's pinv (see smoother.cu)
            const bool ok0 = d0 > 1.5e-14 * sdg[k] && d0 > 1e-15 * dmax && d0 > 0.0;
            if (!ok0) status |= MKF_STATUS_CHOL_FAILED;
            const double pv0 = ok0 ? rcp_nr(d0) : 0.0;
            const bool mine0 = l == k;
            const double clp0 = mine0 ? -pv0 : colk[l] * pv0;
            const double c1p = colk[k + 1] * pv0;          // clp of row k + 1
            double r1[PW];                                 // row k + 1 after pivot k, this thread's columns
#pragma unroll
            for (int c = 0; c < PW; c += 2) {
                const double2 a0 = *reinterpret_cast<const double2 *>(rowk + c0 + c);
                const double2 a1 = *reinterpret_cast<const double2 *>(rowk + NP + c0 + c);
                pr[c] = fma(-clp0, a0.x, mine0 ? 0.0 : pr[c]);
                pr[c + 1] = fma(-clp0, a0.y, mine0 ? 0.0 : pr[c + 1]);
                r1[c] = fma(-c1p, a0.x, a1.x);
                r1[c + 1] = fma(-c1p, a0.y, a1.y);
            }
            if (w == k / PW) {
                pr[k % PW] = -clp0;
                r1[k % PW] = -c1p;
            }
            // column k + 1 of this thread's row after pivot k, and the pivot itself
            const double a_k_k1 = rowk[k + 1];
            const double cl1 = fma(-clp0, a_k_k1, mine0 ? 0.0 : colk[NP + l]);
            const double d1 = fma(-c1p, a_k_k1, rowk[NP + k + 1]);
            // ---- pivot k + 1
            const bool ok1 = d1 > 1.5e-14 * sdg[k + 1] && d1 > 1e-15 * dmax && d1 > 0.0;
            if (!ok1) status |= MKF_STATUS_CHOL_FAILED;
            const double pv1 = ok1 ? rcp_nr(d1) : 0.0;
            const bool mine1 = l == k + 1;
            const double clp1 = mine1 ? -pv1 : cl1 * pv1;
#pragma unroll
            for (int c = 0; c < PW; ++c) pr[c] = fma(-clp1, r1[c], mine1 ? 0.0 : pr[c]);
            if (w == (k + 1) / PW) pr[(k + 1) % PW] = -clp1;
        }
#pragma unroll
        for (int c = 0; c < PW; c += 2) *reinterpret_cast<double2 *>(sL + l * LD + c0 + c) = make_double2(pr[c], pr[c + 1]);
        if (!diagA) {
            // F = A P_f (row l, columns c0..c0+7) into sM; it is consumed before step 3 rewrites sM
            double fr[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) fr[c] = 0.0;
#pragma unroll 8
            for (int k = 0; k < NP; ++k) {          // (padding rows / columns are zero)
                const double av = sA[l * LD + k];
#pragma unroll
                for (int c = 0; c < PW; c += 2) {
                    const double2 pf = *reinterpret_cast<const double2 *>(sPf + k * LD + c0 + c);
                    fr[c] = fma(av, pf.x, fr[c]);
                    fr[c + 1] = fma(av, pf.y, fr[c + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c < PW; c += 2) *reinterpret_cast<double2 *>(sM + l * LD + c0 + c) = make_double2(fr[c], fr[c + 1]);
        }
        __syncthreads();

        // ---- 2. G = (A P_f)^T P_p^-1: row l, columns c0..c0+7; x_s[t] = x_f + G delta --------------------
        {
            double g[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) g[c] = 0.0;
#pragma unroll 8
            for (int k = 0; k < NP; ++k) {
                const double f = diagA ? sa[k] * sPf[k * LD + l] : sM[k * LD + l];
#pragma unroll
                for (int c = 0; c < PW; c += 2) {
                    const double2 pi = *reinterpret_cast<const double2 *>(sL + k * LD + c0 + c);
                    g[c] = fma(f, pi.x, g[c]);
                    g[c + 1] = fma(f, pi.y, g[c + 1]);
                }
            }
            double part = 0.0;
#pragma unroll
            for (int c = 0; c < PW; c += 2) {
                *reinterpret_cast<double2 *>(sG + l * LD + c0 + c) = make_double2(g[c], g[c + 1]);
                part = fma(g[c], sdel[c0 + c], part);
                part = fma(g[c + 1], sdel[c0 + c + 1], part);
            }
            spart[w * NP + l] = part;
        }
        __syncthreads();
        if (tid < NP) {
            const double xf = (tid < n) ? a.fm[(base + t) * n + tid] : 0.0;
            double xg = (spart[tid] + spart[NP + tid]) + (spart[2 * NP + tid] + spart[3 * NP + tid]);
            if (kSwWarps == 8)
                xg += (spart[4 * NP + tid] + spart[5 * NP + tid]) + (spart[6 * NP + tid] + spart[7 * NP + tid]);
            const double xn = xf + xg;
            sxs[tid] = xn;                         // read again only after the next barriers
            if (tid < n) a.sm[(base + t) * n + tid] = xn;
        }

        // ---- 3. M = G D (row l, columns c0..c0+7) -------------------------------------------------
        {
            double Mrow[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) Mrow[c] = 0.0;
#pragma unroll 8
            for (int k = 0; k < NP; ++k) {
                const double g = sG[l * LD + k];
#pragma unroll
                for (int c = 0; c < PW; c += 2) {
                    const double2 d = *reinterpret_cast<const double2 *>(sD + k * LD + c0 + c);
                    Mrow[c] = fma(g, d.x, Mrow[c]);
                    Mrow[c + 1] = fma(g, d.y, Mrow[c + 1]);
                }
            }
#pragma unroll
            for (int c = 0; c < PW; c += 2) *reinterpret_cast<double2 *>(sM + l * LD + c0 + c) = make_double2(Mrow[c], Mrow[c + 1]);
        }
        __syncthreads();
        // ---- P_s[t] = P_f + M G^T ------------------------------------------------------------------
        {
            double out[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) out[c] = sPf[l * LD + c0 + c];
#pragma unroll 4
            for (int k = 0; k < NP; k += 2) {
                const double2 mk = *reinterpret_cast<const double2 *>(sM + l * LD + k);
#pragma unroll
                for (int c = 0; c < PW; ++c) {
                    const double2 g = *reinterpret_cast<const double2 *>(sG + (c0 + c) * LD + k);
                    out[c] = fma(mk.x, g.x, out[c]);
                    out[c] = fma(mk.y, g.y, out[c]);
                }
            }
            double *dst = a.sc + (base + t) * nn + (int64_t)l * n + c0;
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                sPs[l * LD + c0 + c] = out[c];
                if (a.sc && l < n && c0 + c < n) dst[c] = out[c];
            }
        }
        __syncthreads();
        project(t);        // reads sPs / sxs, which the next step rewrites only after its own barriers
    }
    cp_async_wait<0>();
    // any thread that saw a non-positive pivot reports it (the factorisation is redundant: all do)
    if (tid == 0 && a.status) a.status[b] = status;
}

int launch_smoother_wide(SmootherArgs a, cudaStream_t stream) {
    constexpr int NP = 32, LD = NP + 2;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    size_t per = 6 * NP * LD + (12 + kSwWarps) * NP + 4 * NP * LD + (size_t)a.p * NP + (diagA ? 0 : NP * LD);
    const size_t smem = (per + 2) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(smoother_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    smoother_wide_kernel<<<(unsigned)a.B, kSwThreads, smem, stream>>>(a);
    return mkf_host::check_launch();
}

}  // namespace
