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
//   1. Blocked Cholesky of P_p in shared memory, 4 panels of 8 columns: every thread factorises
//      the 8 x 8 diagonal block redundantly in registers (8 dependent pivots, the only serial
//      part), warp 0 solves the panel rows below it, all threads update the trailing matrix
//      (three block barriers per panel).
//   2. Substitutions  (L L^T) g_c = (A P_f)[:, c]: column c is owned by FOUR adjacent lanes that
//      hold rows q, q+4, ... of it; the solved entry of step k is passed among the four by one
//      shuffle.  32 columns x 4 lanes = 128 threads, no block barrier inside.
//   3. x_s[t], M = G D, P_s[t] = P_f + M G^T: thread (w, l) produces row l, columns [8w, 8w+8).
#pragma once

namespace {

constexpr int kSwThreads = 128;

__global__ void __launch_bounds__(kSwThreads, 2) smoother_wide_kernel(const SmootherArgs a) {
    constexpr int NP = 32, LD = NP + 2, PW = 8;       // PW: panel width = columns per warp
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
    double *sinv = sm_;  sm_ += NP;        // 1 / L_ii
    double *sa = sm_;    sm_ += NP;        // diag(A)
    double *sxs = sm_;   sm_ += NP;        // x_s[t]
    double *sdg = sm_;   sm_ += NP;        // original diagonal of P_p[t+1] (singularity test of the pivots)
    double *stage = sm_; sm_ += 2 * 2 * NP * NP;   // [buffer][P_p | P_f][n*n] as they lie in global memory
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
        for (int j = w; j < a.p; j += 4) {
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

    auto issue_step = [&](int64_t t, int buf) {     // inputs of backward step t: P_p[t+1], P_f[t]
        if (t >= 0) {
            const double *srcp = a.pc + (base + t + 1) * nn;
            const double *srcf = a.fc + (base + t) * nn;
            double *dst = stage + (size_t)buf * 2 * NP * NP;
            for (int i = tid; i < nn; i += kSwThreads) {
                cp_async8(dst + i, srcp + i);
                cp_async8(dst + NP * NP + i, srcf + i);
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
            const double *sp = stage + (size_t)buf * 2 * NP * NP, *sf = sp + NP * NP;
#pragma unroll
            for (int c = 0; c < PW; ++c) {
                const int cc = c0 + c;
                const bool in = l < n && cc < n;
                const double pp = in ? sp[l * n + cc] : ((l == cc) ? 1.0 : 0.0);   // padding: identity
                const double pf = in ? sf[l * n + cc] : 0.0;
                sL[l * LD + cc] = pp;
                if (l == cc) sdg[l] = pp;
                sPf[l * LD + cc] = pf;
                sD[l * LD + cc] = in ? sPs[l * LD + cc] - pp : 0.0;
            }
            if (tid < NP) sdel[tid] = (tid < n) ? sxs[tid] - a.pm[(base + t + 1) * n + tid] : 0.0;
        }
        __syncthreads();

        // ---- 1. blocked Cholesky P_p = L L^T in sL (lower triangle) -----------------------------
        double dmax = 0.0;                        // largest diagonal entry: scale of pinv's cut-off
        for (int i = 0; i < n; ++i) dmax = fmax(dmax, sdg[i]);
#pragma unroll 1
        for (int p = 0; p < NP / PW; ++p) {
            const int j0 = p * PW;
            // diagonal block, redundantly in every thread
            double Ld[PW * (PW + 1) / 2], inv[PW];
            {
                int e = 0;
#pragma unroll
                for (int i = 0; i < PW; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) Ld[e++] = sL[(j0 + i) * LD + j0 + j];
            }
            auto lt = [](int i, int j) { return i * (i + 1) / 2 + j; };       // j <= i
            __syncthreads();      // every thread has the block in registers before warp 0 overwrites it
#pragma unroll
            for (int j = 0; j < PW; ++j) {
                const double pjj = Ld[lt(j, j)];
                // numerically singular direction: zero gain, as the reference's pinv (see smoother.cu)
                const bool okp = pjj > 1.5e-14 * sdg[j0 + j] && pjj > 1e-15 * dmax && pjj > 0.0;
                if (!okp) status |= MKF_STATUS_CHOL_FAILED;
                const double iv = okp ? rsqrt(pjj) : 0.0;
                inv[j] = iv;
                Ld[lt(j, j)] = pjj * iv;
#pragma unroll
                for (int i = j + 1; i < PW; ++i) Ld[lt(i, j)] *= iv;
#pragma unroll
                for (int i = j + 1; i < PW; ++i)
#pragma unroll
                    for (int k = j + 1; k <= i; ++k) Ld[lt(i, k)] = fma(-Ld[lt(i, j)], Ld[lt(k, j)], Ld[lt(i, k)]);
            }
            // warp 0: rows of the diagonal block and the panel rows below it
            if (w == 0) {
                if (l >= j0 && l < j0 + PW) {
                    const int i = l - j0;
#pragma unroll
                    for (int j = 0; j < PW; ++j) {
                        double v = 0.0;
#pragma unroll
                        for (int ii = 0; ii < PW; ++ii)
                            if (ii == i && j <= ii) v = Ld[lt(ii, j)];
                        sL[l * LD + j0 + j] = v;
                    }
                } else if (l >= j0 + PW) {
                    // row l of the panel: solve x Ld^T = A[l][j0..j0+8)
                    double xr[PW];
#pragma unroll
                    for (int j = 0; j < PW; ++j) xr[j] = sL[l * LD + j0 + j];
#pragma unroll
                    for (int j = 0; j < PW; ++j) {
                        xr[j] *= inv[j];
#pragma unroll
                        for (int k = j + 1; k < PW; ++k) xr[k] = fma(-xr[j], Ld[lt(k, j)], xr[k]);
                    }
#pragma unroll
                    for (int j = 0; j < PW; ++j) sL[l * LD + j0 + j] = xr[j];
                }
#pragma unroll
                for (int j = 0; j < PW; ++j)
                    if (l == j) sinv[j0 + j] = inv[j];
            }
            __syncthreads();
            // trailing update: A[r][c] -= sum_k L[r][j0+k] L[c][j0+k], rows/cols beyond the panel
            if (l >= j0 + PW && c0 >= j0 + PW) {
                double lr[PW];
#pragma unroll
                for (int k = 0; k < PW; ++k) lr[k] = sL[l * LD + j0 + k];
#pragma unroll
                for (int c = 0; c < PW; ++c) {
                    const int cc = c0 + c;
                    double acc = sL[l * LD + cc];
#pragma unroll
                    for (int k = 0; k < PW; k += 2) {
                        const double2 lc = *reinterpret_cast<const double2 *>(sL + cc * LD + j0 + k);
                        acc = fma(-lr[k], lc.x, acc);
                        acc = fma(-lr[k + 1], lc.y, acc);
                    }
                    sL[l * LD + cc] = acc;       // (entries above the diagonal are never read)
                }
            }
            __syncthreads();
        }

        // ---- 2. column c of G^T: (L L^T) g = (A P_f)[:, c]; four lanes per column -------------
        const int col = w * PW + (l >> 2), q = l & 3;
        double z[NP / 4];                         // rows q, q + 4, ..., q + 28 of the column
        if (diagA) {
#pragma unroll
            for (int j = 0; j < NP / 4; ++j) z[j] = sa[q + 4 * j] * sPf[(q + 4 * j) * LD + col];
        } else {
#pragma unroll
            for (int j = 0; j < NP / 4; ++j) {
                double s = 0.0;
                for (int k = 0; k < n; ++k) s = fma(sA[(q + 4 * j) * LD + k], sPf[k * LD + col], s);
                z[j] = s;
            }
        }
#pragma unroll
        for (int k = 0; k < NP; ++k) {            // forward substitution with L
            const double mine = z[k >> 2] * sinv[k];
            const double zk = __shfl_sync(0xffffffffu, mine, (l & ~3) | (k & 3));
            if (q == (k & 3)) z[k >> 2] = zk;
#pragma unroll
            for (int j = 0; j < NP / 4; ++j) {
                const int i = q + 4 * j;
                if (4 * j + 3 > k) {              // some lane of the group has a row below k here
                    const double lik = sL[i * LD + k];
                    if (i > k) z[j] = fma(-lik, zk, z[j]);
                }
            }
        }
#pragma unroll
        for (int k = NP - 1; k >= 0; --k) {       // backward substitution with L^T
            const double mine = z[k >> 2] * sinv[k];
            const double zk = __shfl_sync(0xffffffffu, mine, (l & ~3) | (k & 3));
            if (q == (k & 3)) z[k >> 2] = zk;
#pragma unroll
            for (int j = 0; j < NP / 4; ++j) {
                const int i = q + 4 * j;
                if (4 * j < k) {
                    const double lki = sL[k * LD + i];
                    if (i < k) z[j] = fma(-lki, zk, z[j]);
                }
            }
        }
        // z = row `col` of G at columns q + 4j; x_s[t][col] = x_f + G[col][:] . delta
        {
            double part = 0.0;
#pragma unroll
            for (int j = 0; j < NP / 4; ++j) {
                sG[col * LD + q + 4 * j] = z[j];
                part = fma(z[j], sdel[q + 4 * j], part);
            }
            part += __shfl_xor_sync(0xffffffffu, part, 1);
            part += __shfl_xor_sync(0xffffffffu, part, 2);
            if (q == 0) {
                const double xf = (col < n) ? a.fm[(base + t) * n + col] : 0.0;
                const double xn = xf + part;
                sxs[col] = xn;                     // read again only after the next barriers
                if (col < n) a.sm[(base + t) * n + col] = xn;
            }
        }
        __syncthreads();

        // ---- 3. M = G D (row l, columns c0..c0+7) -------------------------------------------------
        {
            double Mrow[PW];
#pragma unroll
            for (int c = 0; c < PW; ++c) Mrow[c] = 0.0;
            for (int k = 0; k < n; ++k) {
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
            for (int k = 0; k < n; k += 2) {
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
    size_t per = 6 * NP * LD + 5 * NP + 4 * NP * NP + (size_t)a.p * NP + (diagA ? 0 : NP * LD);
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
