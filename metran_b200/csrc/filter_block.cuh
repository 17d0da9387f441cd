// Block form of the sequential-processing update for wide models (16 < n <= 32): one CTA of four
// warps per model.
//
// Reference: seqkalmanfilter, metran/kalmanfilter.py:341-378 -- per observed series i
//   v_i = y_i - h_i.x,  d_i = P h_i,  f_i = h_i.d_i + r_i,  k_i = d_i / f_i,
//   P <- P - k_i k_i^T f_i,  x <- x + k_i v_i,  sigma += v_i^2 / f_i,  detf += log f_i.
// filter_generic_kernel does exactly that, one series after the other: two butterfly reductions, a
// reciprocal and a shared-memory broadcast per series, all on the critical path (C3: 64 series per
// step -> ~770 cycles per series for a warp that is alone on its SMSP).
//
// Here BS = 8 observed series are taken together.  With D = P H_b^T (n x 8, from the covariance
// BEFORE the block), S = H_b D + R_b and v0 = y_b - H_b x:
//   S = L F L^T  (unit lower L, F = diag f)   -- the pivots f_i ARE the reference's innovation
//                                                variances, w = L^-1 v0 its innovations
//   U = D L^-T                                -- column i of U is the reference's d_i
//   x <- x + U F^-1 w,   P <- P - U F^-1 U^T,   sigma += sum w_i^2 / f_i,   detf += sum log f_i
// Same numbers as eight sequential updates up to rounding (the elimination performs the same
// subtractions in a different association).
//
// Mapping: thread (w, l), w = warp 0..3, l = lane: row l of P, columns [8w, 8w + 8) in registers;
// x_l replicated in the four warps.  Per block
//   1. partial D over the thread's 8 columns (64 FMA), to shared memory            -- barrier A
//   2. D (row l) = sum of the four partials; warp w transpose-reduces 16 of the 64 lane products
//      (36 of S, 8 of H_b x, padding) over its lanes and publishes the totals      -- barrier B
//   3. every thread factorises S redundantly (8 pivots; the only sequential part), forms its
//      row of U, the innovations w and the gain columns; the warp that owns rows [8w, 8w + 8) as
//      COLUMNS of the other threads' update publishes them                          -- barrier C
//   4. rank-8 update of the thread's 8 covariance entries (64 FMA).
// The small exchange buffers are double buffered by block parity, so three barriers per block are
// all the synchronisation there is.  Fewer than 8 observed series left: the block is padded with
// null series (h = 0, r = 1, y = 0: f = 1, v = 0, no effect).
#pragma once

namespace {

// Sum over the 32 lanes of 16 per-lane values with 16 shuffles: four halving steps (offsets 16..2)
// leave one value per lane, a final exchange with the neighbour completes it.  Afterwards lanes 2e
// and 2e + 1 both hold the total of entry e in v[0].
__device__ __forceinline__ void transpose_reduce16(double (&v)[16], int lane) {
#pragma unroll
    for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int e = 0; e < half; ++e) {
            const double keep = upper ? v[e + half] : v[e];
            const double send = upper ? v[e] : v[e + half];
            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

constexpr int kBlkThreads = 128;

template <bool HIST>
__global__ void __launch_bounds__(kBlkThreads, 2) filter_block_kernel(const FilterArgs a) {
    constexpr int NP = 32, BS = 8, NS = BS * (BS + 1) / 2, CW = 8;   // CW: columns per warp
    extern __shared__ __align__(16) double smem[];
    const int n = a.n, m = a.m;
    const int r = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c0 = w * CW;
    const int64_t b = blockIdx.x;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    const bool diagQ = a.flags & MKF_FLAG_Q_DIAGONAL;
    const int TT = a.tile_steps;
    const int ytile = round_up2(TT * m);

    double *sm = smem;
    double *sH = sm;                      sm += (size_t)(m + 1) * NP;     // row m: the null series (zeros)
    double *sR = sm;                      sm += round_up2(m + 1);
    double *sa = sm;                      sm += NP;
    double *sx = sm;                      sm += NP;
    double *sy = sm;                      sm += 2 * ytile;
    double *sDp = sm;                     sm += 2 * 4 * BS * NP;          // partial D [parity][warp][series][row]
    double *sG = sm;                      sm += 2 * NP * BS;              // gain columns [parity][row][series]
    double *sS = sm;                      sm += 2 * 64;                   // reduced S (upper) and H_b x [parity][64]
    double *sYo = sm;                     sm += round_up2(m + BS);        // observed values of the step, compacted
    int *sIdx = reinterpret_cast<int *>(sm); sm += round_up2((m + BS + 1) / 2);
    double *sA = sm;                      if (!diagA) sm += NP * (NP + 1) + 1;
    double *sQ = sm;

    const double *yb = a.y + b * a.y_stride;
    const double *Ab = a.A + b * a.A_stride;
    const double *Qb = a.Q + b * a.Q_stride;
    const double *Hb = a.H + b * a.H_stride;
    const double *Rb = a.R + b * a.R_stride;
    for (int i = threadIdx.x; i < (m + 1) * NP; i += kBlkThreads) {
        const int j = i / NP, c = i % NP;
        sH[i] = (j < m && c < n) ? Hb[j * n + c] : 0.0;
    }
    for (int i = threadIdx.x; i <= m; i += kBlkThreads) sR[i] = (i < m) ? Rb[i] : 1.0;
    const double a_r = (r < n) ? Ab[r * n + r] : 0.0;
    if (w == 0) sa[r] = a_r;
    if (!diagA)
        for (int i = threadIdx.x; i < NP * NP; i += kBlkThreads) {
            const int rr = i / NP, c = i % NP;
            sA[rr * (NP + 1) + c] = (rr < n && c < n) ? Ab[rr * n + c] : 0.0;
        }
    if (!diagQ)
        for (int i = threadIdx.x; i < NP * NP; i += kBlkThreads) {
            const int rr = i / NP, c = i % NP;
            sQ[rr * (NP + 1) + c] = (rr < n && c < n) ? Qb[rr * n + c] : 0.0;
        }
    const double q_r = (r < n) ? Qb[r * n + r] : 0.0;

    double x = (a.x0 && r < n) ? a.x0[b * a.x0_stride + r] : 0.0;     // replicated in the four warps
    double P[CW];                                                     // P[r][c0 + c]
#pragma unroll
    for (int c = 0; c < CW; ++c) {
        const int cc = c0 + c;
        double v = 0.0;
        if (r < n && cc < n) v = a.P0 ? a.P0[b * a.P0_stride + r * n + cc] : (r == cc ? 1.0 : 0.0);
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
            const int64_t t0 = tile * TT;
            const int cntd = (int)min((int64_t)TT, T - t0) * m;
            const double *src = yb + t0 * m;
            double *dst = sy + (tile & 1) * ytile;
            for (int i = threadIdx.x; i < cntd; i += kBlkThreads) cp_async8(dst + i, src + i);
        }
        cp_async_commit();
    };
    issue_tile(0);
    const int64_t hist_base = b * T;
    __syncthreads();

    int par = 0;       // parity of the exchange buffers
    for (int64_t tile = 0; tile < ntiles; ++tile) {
        __syncthreads();
        issue_tile(tile + 1);
        cp_async_wait<1>();
        __syncthreads();
        const double *ys = sy + (tile & 1) * ytile;
        const int64_t t0 = tile * TT;
        const int steps = (int)min((int64_t)TT, T - t0);
        for (int tt = 0; tt < steps; ++tt) {
            const int64_t t = t0 + tt;
            // ---- predict (kalmanfilter.py:318-331) ---------------------------------------------
            if (diagA) {
                x = a_r * x;
            } else {
                if (w == 0) sx[r] = x;
                __syncthreads();
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NP; ++c) s = fma(sA[r * (NP + 1) + c], sx[c], s);
                __syncthreads();
                x = s;
            }
#pragma unroll
            for (int c = 0; c < CW; c += 2) {
                const double2 ac = *reinterpret_cast<const double2 *>(sa + c0 + c);
                const double v0 = (a_r * P[c]) * ac.x;
                const double v1 = (a_r * P[c + 1]) * ac.y;
                if (diagQ) {
                    P[c] = (c0 + c == r) ? v0 + q_r : v0;
                    P[c + 1] = (c0 + c + 1 == r) ? v1 + q_r : v1;
                } else {
                    P[c] = v0 + sQ[r * (NP + 1) + c0 + c];
                    P[c + 1] = v1 + sQ[r * (NP + 1) + c0 + c + 1];
                }
            }
            if (HIST && r < n) {
                if (a.pm && w == 0) a.pm[(hist_base + t) * n + r] = x;
                if (a.pc) {
                    // row r, columns c0..c0+7 (P is symmetric to rounding, as the reference's own)
                    double *dst = a.pc + (hist_base + t) * n * n + (int64_t)r * n + c0;
#pragma unroll
                    for (int c = 0; c < CW; ++c)
                        if (c0 + c < n) dst[c] = P[c];
                }
            }

            // ---- observed series of this step, in ascending order (kalmanfilter.py:667-674) ----
            int cnt = 0;
            for (int j0 = 0; j0 < m; j0 += NP) {
                const int jj = j0 + r;
                const double yv = (jj < m) ? ys[tt * m + jj] : nan("");
                const bool ob = is_observed(yv);
                const unsigned bal = __ballot_sync(0xffffffffu, ob);
                if (ob && w == 0) {
                    const int pos = cnt + __popc(bal & ((1u << r) - 1u));
                    sIdx[pos] = jj;
                    sYo[pos] = yv;
                }
                cnt += __popc(bal);
            }
            if (w == 0 && r < BS) {          // padding of the last block: the null series
                sIdx[cnt + r] = m;
                sYo[cnt + r] = 0.0;
            }
            __syncthreads();

            double sigma = 0.0, fman = 1.0;
            int fexp = 0;
            for (int blk = 0; blk < cnt; blk += BS, par ^= 1) {
                const double *hrow[BS];
                double rr[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    const int j = sIdx[blk + i];
                    hrow[i] = sH + j * NP;
                    rr[i] = sR[j];
                }
                // 1. partial D over this thread's columns
                double *dp = sDp + (size_t)par * 4 * BS * NP;
                {
                    double Dp[BS];
#pragma unroll
                    for (int i = 0; i < BS; ++i) Dp[i] = 0.0;
#pragma unroll
                    for (int c = 0; c < CW; c += 2) {
#pragma unroll
                        for (int i = 0; i < BS; ++i) {
                            const double2 h = *reinterpret_cast<const double2 *>(hrow[i] + c0 + c);
                            Dp[i] = fma(P[c], h.x, Dp[i]);
                            Dp[i] = fma(P[c + 1], h.y, Dp[i]);
                        }
                    }
#pragma unroll
                    for (int i = 0; i < BS; ++i) dp[(w * BS + i) * NP + r] = Dp[i];
                }
                __syncthreads();                                                    // barrier A
                // 2. D (row r); lane products of S = H_b D and of H_b x, 16 entries per warp
                double D[BS], hr[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    D[i] = (dp[(0 * BS + i) * NP + r] + dp[(1 * BS + i) * NP + r]) +
                           (dp[(2 * BS + i) * NP + r] + dp[(3 * BS + i) * NP + r]);
                    hr[i] = hrow[i][r];
                }
                double *ss = sS + par * 64;
                {
                    // entry e of the 64: e < 36: S(i, l) in row-major upper order; 36..43: (H_b x)_i
                    double prod[48];
                    {
                        int e = 0;
#pragma unroll
                        for (int i = 0; i < BS; ++i)
#pragma unroll
                            for (int l = i; l < BS; ++l) prod[e++] = hr[i] * D[l];
#pragma unroll
                        for (int i = 0; i < BS; ++i) prod[NS + i] = hr[i] * x;
#pragma unroll
                        for (int i = NS + BS; i < 48; ++i) prod[i] = 0.0;
                    }
                    double red[16];
                    if (w == 0) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) red[e] = prod[e];
                    } else if (w == 1) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) red[e] = prod[16 + e];
                    } else {
#pragma unroll
                        for (int e = 0; e < 16; ++e) red[e] = prod[32 + e];
                    }
                    if (w < 3) {
                        transpose_reduce16(red, r);
                        if ((r & 1) == 0) ss[w * 16 + (r >> 1)] = red[0];
                    }
                }
                __syncthreads();                                                    // barrier B
                // 3. every thread factorises S = L F L^T redundantly (8 pivots: the innovation variances)
                double S[NS], wv[BS];
                {
                    int e = 0;
#pragma unroll
                    for (int i = 0; i < BS; ++i)
#pragma unroll
                        for (int l = i; l < BS; ++l, ++e) S[e] = ss[e] + ((l == i) ? rr[i] : 0.0);
#pragma unroll
                    for (int i = 0; i < BS; ++i) wv[i] = sYo[blk + i] - ss[NS + i];
                }
                auto tri8 = [](int i, int l) { return i * BS - (i * (i - 1)) / 2 + (l - i); };
                double U[BS], G[BS], gw = 0.0;
#pragma unroll
                for (int i = 0; i < BS; ++i) U[i] = D[i];
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    const double f = S[tri8(i, i)];
                    const bool badf = f_is_bad(f);
                    const double rf = badf ? nan("") : rcp_nr(f);
                    if (badf) status |= MKF_STATUS_NONPOSITIVE_F;
                    sigma = fma(wv[i] * wv[i], rf, sigma);
                    fman *= f;
                    G[i] = U[i] * rf;                          // gain column i, row r
                    gw = fma(G[i], wv[i], gw);
#pragma unroll
                    for (int l = i + 1; l < BS; ++l) {
                        const double mlt = S[tri8(i, l)] * rf;
                        wv[l] = fma(-mlt, wv[i], wv[l]);
                        U[l] = fma(-mlt, U[i], U[l]);
#pragma unroll
                        for (int l2 = l; l2 < BS; ++l2) S[tri8(l, l2)] = fma(-mlt, S[tri8(i, l2)], S[tri8(l, l2)]);
                    }
                }
                {
                    const int hi = __double2hiint(fman);
                    const int ex = ((hi >> 20) & 0x7ff) - 1023;
                    fexp += ex;
                    fman = __hiloint2double(hi - (ex << 20), __double2loint(fman));
                }
                x += gw;
                // gain columns of all rows ([row][series]); rows are identical in the four warps
                double *sg = sG + par * NP * BS;
                if (w == (r >> 3)) {
#pragma unroll
                    for (int i = 0; i < BS; i += 2) *reinterpret_cast<double2 *>(sg + r * BS + i) = make_double2(G[i], G[i + 1]);
                }
                __syncthreads();                                                    // barrier C
                // 4. P[r][c0 + c] -= sum_i U[r][i] G[c0 + c][i]
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    double acc = P[c];
#pragma unroll
                    for (int i = 0; i < BS; i += 2) {
                        const double2 g = *reinterpret_cast<const double2 *>(sg + (c0 + c) * BS + i);
                        acc = fma(-U[i], g.x, acc);
                        acc = fma(-U[i + 1], g.y, acc);
                    }
                    P[c] = acc;
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
                if (threadIdx.x == 0) {
                    if (a.sigma) a.sigma[hist_base + t] = sigma;
                    if (a.detf) a.detf[hist_base + t] = detf;
                    if (a.cnt) a.cnt[hist_base + t] = cnt;
                }
                if (r < n) {
                    if (a.fm && w == 0) a.fm[(hist_base + t) * n + r] = x;
                    if (a.fc) {
                        double *dst = a.fc + (hist_base + t) * n * n + (int64_t)r * n + c0;
#pragma unroll
                        for (int c = 0; c < CW; ++c)
                            if (c0 + c < n) dst[c] = P[c];
                    }
                }
            }
            __syncthreads();      // sIdx / sYo are rewritten by the next step
        }
    }
    cp_async_wait<0>();
    if (threadIdx.x == 0) {
        if (a.mle) a.mle[b] = (double)nobs * kLog2Pi + sum_detf.value() + sum_sigma.value();
        if (a.nobs) a.nobs[b] = nobs;
        if (a.status) a.status[b] = status;
    }
}

template <bool HIST>
int launch_block(const FilterArgs &args0, cudaStream_t stream) {
    constexpr int NP = 32, BS = 8;
    FilterArgs a = args0;
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    const bool diagQ = a.flags & MKF_FLAG_Q_DIAGONAL;
    int TT = 256 / a.m;
    if (TT < 1) TT = 1;
    if (TT > 32) TT = 32;
    if ((int64_t)TT > a.T && a.T > 0) TT = (int)a.T;
    a.tile_steps = TT;
    size_t per = (size_t)(a.m + 1) * NP + round_up2(a.m + 1) + NP + NP + 2 * round_up2(TT * a.m) + 2 * 4 * BS * NP +
                 2 * NP * BS + 2 * 64 + round_up2(a.m + BS) + round_up2((a.m + BS + 1) / 2);
    if (!diagA) per += NP * (NP + 1) + 1;
    if (!diagQ) per += NP * (NP + 1) + 1;
    const size_t smem = (per + 2) * sizeof(double);
    if (smem > 200 * 1024) return MKF_ERR_UNSUPPORTED;
    auto kern = filter_block_kernel<HIST>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    kern<<<(unsigned)a.B, kBlkThreads, smem, stream>>>(a);
    return mkf_host::check_launch();
}

}  // namespace
