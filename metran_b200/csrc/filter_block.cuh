// Block form of the sequential-processing update for wide models (16 < n <= 32), one warp per model.
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
// subtractions in a different association).  Per block the lanes exchange data twice (one
// transpose-reduction of the 36 + 8 lane products of S and v0, one broadcast of the 8 gain
// columns) instead of 24 times, and the 2 x 256 FMAs per lane of D and of the rank-8 update are
// independent instruction streams.  Fewer than 8 observed series left: the block is padded with
// null series (h = 0, r = 1, y = 0: f = 1, v = 0, no effect).
#pragma once

namespace {

// Sum over the 32 lanes of 64 per-lane values with 62 shuffles (recursive halving): afterwards lane
// l holds the totals of entries 2l and 2l + 1 in v[0], v[1].
__device__ __forceinline__ void transpose_reduce64(double (&v)[64], int lane) {
#pragma unroll
    for (int half = 32, off = 16; half >= 2; half >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int e = 0; e < half; ++e) {
            // keep the half that matches this lane's bit, send the other one to the partner
            const double keep = upper ? v[e + half] : v[e];
            const double send = upper ? v[e] : v[e + half];
            v[e] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

template <bool HIST>
__global__ void __launch_bounds__(32) filter_block_kernel(const FilterArgs a) {
    constexpr int NP = 32, BS = 8, NS = BS * (BS + 1) / 2;
    extern __shared__ __align__(16) double smem[];
    const int n = a.n, m = a.m;
    const int r = threadIdx.x;
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
    double *sG = sm;                      sm += NP * BS;                  // gain columns [row][series]
    double *sS = sm;                      sm += 64;                       // reduced S (upper) and v0
    double *sYo = sm;                     sm += round_up2(m + BS);        // observed values of the step, compacted
    int *sIdx = reinterpret_cast<int *>(sm); sm += round_up2((m + BS + 1) / 2);
    double *sA = sm;                      if (!diagA) sm += NP * (NP + 1) + 1;
    double *sQ = sm;

    const double *yb = a.y + b * a.y_stride;
    const double *Ab = a.A + b * a.A_stride;
    const double *Qb = a.Q + b * a.Q_stride;
    const double *Hb = a.H + b * a.H_stride;
    const double *Rb = a.R + b * a.R_stride;
    for (int i = r; i < (m + 1) * NP; i += NP) {
        const int j = i / NP, c = i % NP;
        sH[i] = (j < m && c < n) ? Hb[j * n + c] : 0.0;
    }
    for (int i = r; i <= m; i += NP) sR[i] = (i < m) ? Rb[i] : 1.0;
    const double a_r = (r < n) ? Ab[r * n + r] : 0.0;
    sa[r] = a_r;
    if (!diagA)
        for (int i = r; i < NP * NP; i += NP) {
            const int rr = i / NP, c = i % NP;
            sA[rr * (NP + 1) + c] = (rr < n && c < n) ? Ab[rr * n + c] : 0.0;
        }
    if (!diagQ)
        for (int i = r; i < NP * NP; i += NP) {
            const int rr = i / NP, c = i % NP;
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
            const int64_t t0 = tile * TT;
            const int cntd = (int)min((int64_t)TT, T - t0) * m;
            const double *src = yb + t0 * m;
            double *dst = sy + (tile & 1) * ytile;
            for (int i = r; i < cntd; i += NP) cp_async8(dst + i, src + i);
        }
        cp_async_commit();
    };
    issue_tile(0);
    const int64_t hist_base = b * T;
    __syncwarp();

    for (int64_t tile = 0; tile < ntiles; ++tile) {
        __syncwarp();
        issue_tile(tile + 1);
        cp_async_wait<1>();
        __syncwarp();
        const double *ys = sy + (tile & 1) * ytile;
        const int64_t t0 = tile * TT;
        const int steps = (int)min((int64_t)TT, T - t0);
        for (int tt = 0; tt < steps; ++tt) {
            const int64_t t = t0 + tt;
            // ---- predict (kalmanfilter.py:318-331) ---------------------------------------------
            if (diagA) {
                x = a_r * x;
            } else {
                sx[r] = x;
                __syncwarp();
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < NP; ++c) s = fma(sA[r * (NP + 1) + c], sx[c], s);
                __syncwarp();
                x = s;
            }
#pragma unroll
            for (int c = 0; c < NP; c += 2) {
                const double2 ac = *reinterpret_cast<const double2 *>(sa + c);
                const double v0 = (a_r * P[c]) * ac.x;
                const double v1 = (a_r * P[c + 1]) * ac.y;
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

            // ---- observed series of this step, in ascending order (kalmanfilter.py:667-674) ----
            int cnt = 0;
            for (int j0 = 0; j0 < m; j0 += NP) {
                const int jj = j0 + r;
                const double yv = (jj < m) ? ys[tt * m + jj] : nan("");
                const bool ob = is_observed(yv);
                const unsigned bal = __ballot_sync(0xffffffffu, ob);
                if (ob) {
                    const int pos = cnt + __popc(bal & ((1u << r) - 1u));
                    sIdx[pos] = jj;
                    sYo[pos] = yv;
                }
                cnt += __popc(bal);
            }
            if (r < BS) {          // padding of the last block: the null series
                sIdx[cnt + r] = m;
                sYo[cnt + r] = 0.0;
            }
            __syncwarp();

            double sigma = 0.0, fman = 1.0;
            int fexp = 0;
            for (int blk = 0; blk < cnt; blk += BS) {
                const double *hrow[BS];
                double rr[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    const int j = sIdx[blk + i];
                    hrow[i] = sH + j * NP;
                    rr[i] = sR[j];
                }
                // D = P H_b^T (this lane: row r), from the covariance before the block
                double D[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i) D[i] = 0.0;
#pragma unroll
                for (int c = 0; c < NP; c += 2) {
#pragma unroll
                    for (int i = 0; i < BS; ++i) {
                        const double2 h = *reinterpret_cast<const double2 *>(hrow[i] + c);
                        D[i] = fma(P[c], h.x, D[i]);
                        D[i] = fma(P[c + 1], h.y, D[i]);
                    }
                }
                // lane products of S = H_b D (upper triangle) and of H_b x, summed over the lanes
                double red[64];
                double hr[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i) hr[i] = hrow[i][r];
                {
                    int e = 0;
#pragma unroll
                    for (int i = 0; i < BS; ++i)
#pragma unroll
                        for (int l = i; l < BS; ++l) red[e++] = hr[i] * D[l];
#pragma unroll
                    for (int i = 0; i < BS; ++i) red[NS + i] = hr[i] * x;
#pragma unroll
                    for (int i = NS + BS; i < 64; ++i) red[i] = 0.0;
                }
                transpose_reduce64(red, r);
                *reinterpret_cast<double2 *>(sS + 2 * r) = make_double2(red[0], red[1]);
                __syncwarp();
                // every lane factorises S = L F L^T redundantly (8 pivots: the innovation variances)
                double S[NS], w[BS];
                {
                    int e = 0;
#pragma unroll
                    for (int i = 0; i < BS; ++i)
#pragma unroll
                        for (int l = i; l < BS; ++l, ++e) S[e] = sS[e] + ((l == i) ? rr[i] : 0.0);
#pragma unroll
                    for (int i = 0; i < BS; ++i) w[i] = sYo[blk + i] - sS[NS + i];
                }
                __syncwarp();          // sS may be rewritten by the next block
                auto tri8 = [](int i, int l) { return i * BS - (i * (i - 1)) / 2 + (l - i); };
                double U[BS], gw = 0.0;
#pragma unroll
                for (int i = 0; i < BS; ++i) U[i] = D[i];
                double G[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    const double f = S[tri8(i, i)];
                    const bool badf = f_is_bad(f);
                    const double rf = badf ? nan("") : rcp_nr(f);
                    if (badf) status |= MKF_STATUS_NONPOSITIVE_F;
                    sigma = fma(w[i] * w[i], rf, sigma);
                    fman *= f;
                    G[i] = U[i] * rf;                          // gain column i, row r
                    gw = fma(G[i], w[i], gw);
#pragma unroll
                    for (int l = i + 1; l < BS; ++l) {
                        const double mlt = S[tri8(i, l)] * rf;
                        w[l] = fma(-mlt, w[i], w[l]);
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
                // P <- P - U G^T: gain columns of all rows through shared memory ([row][series])
#pragma unroll
                for (int i = 0; i < BS; i += 2) *reinterpret_cast<double2 *>(sG + r * BS + i) = make_double2(G[i], G[i + 1]);
                __syncwarp();
#pragma unroll
                for (int c = 0; c < NP; ++c) {
                    double acc = P[c];
#pragma unroll
                    for (int i = 0; i < BS; i += 2) {
                        const double2 g = *reinterpret_cast<const double2 *>(sG + c * BS + i);
                        acc = fma(-U[i], g.x, acc);
                        acc = fma(-U[i + 1], g.y, acc);
                    }
                    P[c] = acc;
                }
                __syncwarp();
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
    size_t per = (size_t)(a.m + 1) * NP + round_up2(a.m + 1) + NP + NP + 2 * round_up2(TT * a.m) + NP * BS + 64 +
                 round_up2(a.m + BS) + round_up2((a.m + BS + 1) / 2);
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
    kern<<<(unsigned)a.B, 32, smem, stream>>>(a);
    return mkf_host::check_launch();
}

}  // namespace
