// RTS smoother, one THREAD per model, for big batches of small diagonal-A models (n = 4, 8: Metran's
// own shape at BASELINE's state_dim = 8).  Kernel K3T.
//
// Reference: kalmansmoother, metran/kalmanfilter.py:453-474 (formulas and the Cholesky-instead-of-
// pinv argument: see smoother.cu).  smoother_kernel<8> gives a model 8 lanes (lane r = row r): every
// operand that crosses rows goes through shared memory, 1450 instructions and ~250 shared-memory
// loads per lane and step, L1TEX pipe 85 % busy, 2.6 TB/s (profiles/r02k_*).  With one model per
// lane nothing crosses lanes: P_p^-1 (36 doubles, from its Cholesky factor), D = P_s[t+1] - P_p[t+1] (36) and the
// vectors of the step live in registers; per step and model ~1700 fp64 instructions and ~9 warp-level
// shared-memory instructions, i.e. the kernel can run at the speed the step's 1728 bytes move.
//
// Data movement (all by the TMA engine, tensor maps of filter_dfm.cu):
//   in   P_p[t+1], P_f[t] as {16 doubles, 32 models, n*n/16 lines} boxes under the 128-byte swizzle,
//        x_p[t+1], x_f[t] as {n, 32 models} boxes under the n*8-byte swizzle: every lane reads ITS
//        row of the box with conflict-free 16-byte loads (the mirror image of the history kernel's
//        stores).  P_p has one box: it is consumed at the top of the step (D, Cholesky) and the load
//        of the next step's matrix is issued right after.  P_f, x_f, x_p: one box each, refilled at
//        the end of the step that read them (the load flies during the next step's factorisation);
//        MKF_SMT_LEAN=0 double-buffers them (101 KB instead of 73 KB per 32 models: two instead of
//        three CTAs per SM, measured 16 % slower).
//   out  P_s[t] is assembled in a box of the same layout -- which is also where the next step reads
//        P_s[t+1] from -- and stored by one tensor store per step; x_s in boxes of 2 steps.
// The gain G (n x n, all of it is needed for P_f + (G D) G^T) is the one matrix that does not fit the
// register file next to P_p^-1 and D: row c is written to a per-lane shared-memory row when it is
// computed and read back by the rows after it (P_s is produced triangle-wise, row c needs G rows
// j <= c; the last row is never read back and not stored).
// Shared memory, not registers or bandwidth, decides how many warps an SM holds, and a lone warp runs
// this 2700-instruction step at 0.29 IPC: MKF_SMT_WARPS = 2 warps share one set of boxes (both
// factorise, the rows of the step alternate between them, one CTA barrier per round of rows).
#pragma once
#include <cuda.h>
#include <string.h>

#include <atomic>
#include <type_traits>

namespace mkf_dfm {
int make_hist_map(CUtensorMap *map, double *p, int inner, int64_t B, int64_t T, int lines, int steps);
bool hist_tma_enabled();
}  // namespace mkf_dfm

namespace {

struct SmThreadMaps {
    CUtensorMap pc, fc, pm, fm, sc, sm;
};

#ifndef MKF_SMT_MIN_B
#define MKF_SMT_MIN_B 6144      // below this the lane-per-row kernel is faster: more warps to hide latency with
                                // (measured crossover between 4096 and 8192 models, profiles/r02t_smoother8.txt)
#endif
std::atomic<long long> g_smt_min_b{MKF_SMT_MIN_B};
inline int64_t smoother_thread_min_b() { return (int64_t)g_smt_min_b.load(); }
#ifndef MKF_SMT_LEAN
#define MKF_SMT_LEAN 1          // 1: P_f / x_f / x_p single-buffered, x_s boxes of 2 steps, n - 1 gain rows kept:
#endif                          //    73 KB instead of 101 KB per 32 models = three CTAs (six warps) per SM
constexpr bool kSmtLean = MKF_SMT_LEAN != 0;
constexpr int kSmtTSM = kSmtLean ? 2 : 4;      // steps per tensor store of the smoothed means

template <int N>
struct SmtLayout {
    static constexpr int LINES = N * N / 16;
    static constexpr unsigned COV = LINES * 32u * 128u;   // one covariance box
    static constexpr unsigned MEAN = 32u * N * 8u;        // one mean box (one step)
    // offsets from the 1 KB aligned base
    static constexpr unsigned PC = 0;
    static constexpr unsigned FC0 = PC + COV, FC1 = kSmtLean ? FC0 : FC0 + COV;
    static constexpr unsigned SC = FC1 + COV;
    static constexpr unsigned G = SC + COV;                // gain rows [c][lane][N], N*8-byte swizzle
    static constexpr unsigned GBYTES = (kSmtLean ? N - 1 : N) * MEAN;   // (the last row is never read back)
    static constexpr unsigned PM0 = G + GBYTES, PM1 = kSmtLean ? PM0 : PM0 + MEAN;
    static constexpr unsigned FM0 = PM1 + MEAN, FM1 = kSmtLean ? FM0 : FM0 + MEAN;
    static constexpr unsigned SMO = FM1 + MEAN;            // smoothed means, kSmtTSM steps
    static constexpr unsigned AD = SMO + kSmtTSM * MEAN;   // diag(A) [i][lane]
    static constexpr unsigned BARS = AD + N * 32u * 8u;    // mbarriers: pc, stage 0, stage 1
    static constexpr unsigned TOTAL = BARS + 64u;
};

template <int I, int E, class F>
__device__ __forceinline__ void smt_static_for(F &&f) {
    if constexpr (I < E) {
        f(std::integral_constant<int, I>{});
        smt_static_for<I + 1, E>(f);
    }
}

#ifndef MKF_SMT_WARPS
#define MKF_SMT_WARPS 2      // warps that share one set of boxes (32 models): row c of the step goes to warp c % NW
#endif

template <int N>
__global__ void __launch_bounds__(32 * MKF_SMT_WARPS, MKF_SMT_LEAN ? 3 : 2) smoother_thread_kernel(const SmootherArgs a, const __grid_constant__ SmThreadMaps maps) {
    using LY = SmtLayout<N>;
    constexpr int NT = N * (N + 1) / 2;
    constexpr int NW = MKF_SMT_WARPS;
    static_assert(N % NW == 0, "rows must divide among the warps");
    extern __shared__ unsigned char smt_raw[];
    const unsigned base = (smem_u32(smt_raw) + 1023u) & ~1023u;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool leader = threadIdx.x == 0;          // issues every TMA operation of the CTA
    const int64_t b0 = (int64_t)blockIdx.x * 32, b = b0 + lane;
    const bool active = b < a.B;
    const int64_t T = a.T;
    const unsigned bar_pc = base + LY::BARS, bar_st0 = bar_pc + 8, bar_st1 = bar_pc + 16;
    auto cta_sync = [&]() {
        if (NW == 1) __syncwarp();
        else __syncthreads();
    };

    // ---- addressing inside the swizzled boxes -------------------------------------------------
    // covariance element (r, c): line (r*N + c) / 16, 16-byte chunk ((r*N + c) % 16) / 2 ^ (lane & 7)
    const unsigned lrow = (unsigned)lane * 128u, lsw = (unsigned)(lane & 7);
    auto cov_pair = [&](unsigned box, int r, int c) -> unsigned {   // address of the pair (r, c), (r, c + 1), c even
        const int e = r * N + c;
        return box + (unsigned)(e / 16) * 4096u + lrow + ((((unsigned)(e % 16) >> 1) ^ lsw) << 4);
    };
    const unsigned mrow = (unsigned)lane * (N * 8u);
    const unsigned msw = (N == 8) ? (unsigned)((lane >> 1) & 3) : (unsigned)((lane >> 2) & 1);
    auto vec_pair = [&](unsigned box, int i) -> unsigned {          // pair (i, i + 1) of a lane's n-vector, i even
        return box + mrow + ((((unsigned)i >> 1) ^ msw) << 4);
    };
    auto lds2 = [](unsigned addr, double &v0, double &v1) {
        asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(addr));
    };
    auto sts2 = [](unsigned addr, double v0, double v1) {
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(v0), "d"(v1) : "memory");
    };
    auto sts1 = [](unsigned addr, double v0) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v0) : "memory"); };
    auto lds1 = [](unsigned addr) -> double {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
        return v;
    };
    auto bar_wait = [](unsigned bar, unsigned parity) {
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
    };
    auto load3d = [&](unsigned dst, const CUtensorMap *map, int c2, unsigned bar) {
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(dst),
            "l"(map), "r"(0), "r"((int)b0), "r"(c2), "r"(bar)
            : "memory");
    };
    auto store3d = [&](const CUtensorMap *map, int c2, unsigned src) {
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(0),
                     "r"((int)b0), "r"(c2), "r"(src)
                     : "memory");
    };
    auto expect = [](unsigned bar, unsigned bytes) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
    };
    // loads of one backward step t: P_f[t], x_f[t], x_p[t+1] into stage (t & 1); P_p[t+1] separately
    auto issue_stage = [&](int64_t t) {
        const unsigned s = kSmtLean ? 0u : (unsigned)(t & 1), bar = s ? bar_st1 : bar_st0;
        expect(bar, LY::COV + 2u * LY::MEAN);
        load3d(base + (s ? LY::FC1 : LY::FC0), &maps.fc, (int)t * LY::LINES, bar);
        load3d(base + (s ? LY::FM1 : LY::FM0), &maps.fm, (int)t, bar);
        load3d(base + (s ? LY::PM1 : LY::PM0), &maps.pm, (int)t + 1, bar);
    };
    auto issue_pc = [&](int64_t t) {   // P_p[t + 1]
        expect(bar_pc, LY::COV);
        load3d(base + LY::PC, &maps.pc, (int)(t + 1) * LY::LINES, bar_pc);
    };

    if (leader) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_pc) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_st0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(bar_st1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    // diag(A) of this lane's model: [i][lane]
    if (w == 0) {
        const double *Ab = a.A + (active ? b : b0) * a.A_stride;
#pragma unroll
        for (int i = 0; i < N; ++i) sts1(base + LY::AD + (unsigned)(i * 32 + lane) * 8u, Ab[i * N + i]);
    }
    cta_sync();

    // ---- t = T - 1: smoothed = filtered (kalmanfilter.py:450-451) ---------------------------------
    // P_f[T-1] and x_f[T-1] are fetched as "stage T-1" (the x_p box of that stage is not used)
    unsigned ph_pc = 0, ph_st0 = 0, ph_st1 = 0;
    if (leader) {
        const unsigned s = kSmtLean ? 0u : (unsigned)((T - 1) & 1), bar = s ? bar_st1 : bar_st0;
        expect(bar, LY::COV + LY::MEAN);
        load3d(base + (s ? LY::FC1 : LY::FC0), &maps.fc, (int)(T - 1) * LY::LINES, bar);
        load3d(base + (s ? LY::FM1 : LY::FM0), &maps.fm, (int)(T - 1), bar);
        if (T >= 2) {
            if (!kSmtLean) issue_stage(T - 2);
            issue_pc(T - 2);
        }
    }
    {
        const unsigned s = kSmtLean ? 0u : (unsigned)((T - 1) & 1);
        bar_wait(s ? bar_st1 : bar_st0, 0u);
        if (s) ph_st1 ^= 1u; else ph_st0 ^= 1u;
        const unsigned fcb = base + (s ? LY::FC1 : LY::FC0), fmb = base + (s ? LY::FM1 : LY::FM0);
#pragma unroll
        for (int r = 0; r < N; ++r)
            if (r % NW == w) {
#pragma unroll
                for (int c = 0; c < N; c += 2) {
                    double v0, v1;
                    lds2(cov_pair(fcb, r, c), v0, v1);
                    sts2(cov_pair(base + LY::SC, r, c), v0, v1);
                }
            }
        if (w == 0) {
            const unsigned slot = base + LY::SMO + (unsigned)((T - 1) % kSmtTSM) * LY::MEAN;
#pragma unroll
            for (int i = 0; i < N; i += 2) {
                double v0, v1;
                lds2(vec_pair(fmb, i), v0, v1);
                sts2(vec_pair(slot, i), v0, v1);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cta_sync();
        if (leader) {
            store3d(&maps.sc, (int)(T - 1) * LY::LINES, base + LY::SC);
            if ((T - 1) % kSmtTSM == 0) store3d(&maps.sm, (int)(T - 1), base + LY::SMO);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (kSmtLean && T >= 2) issue_stage(T - 2);     // the one stage is free again
        }
    }

    int status = 0;
    for (int64_t t = T - 2; t >= 0; --t) {
        const unsigned s = kSmtLean ? 0u : (unsigned)(t & 1);
        // two stages: stage (t - 1) & 1 == (t + 1) & 1 was read by step t + 1 (or by the prologue), which has
        // finished -- proxy fence and barrier in front of its stores: refill it for step t - 1.  (One stage:
        // refilled at the end of the step that read it; the load flies during the next step's factorisation.)
        if (!kSmtLean && leader && t >= 1) issue_stage(t - 1);

        // ---- P_p[t+1] -> registers; D = P_s[t+1] - P_p[t+1] (every warp: all of it) --------------------
        bar_wait(bar_pc, ph_pc);
        ph_pc ^= 1u;
        double L[NT], D[NT];   // packed lower triangle, (r, c), c <= r: index r(r+1)/2 + c
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c <= r; c += 2) {
                double p0, p1, s0, s1;
                lds2(cov_pair(base + LY::PC, r, c), p0, p1);
                lds2(cov_pair(base + LY::SC, r, c), s0, s1);
                L[r * (r + 1) / 2 + c] = p0;
                D[r * (r + 1) / 2 + c] = s0 - p0;
                if (c + 1 <= r) {
                    L[r * (r + 1) / 2 + c + 1] = p1;
                    D[r * (r + 1) / 2 + c + 1] = s1 - p1;
                }
            }
        // x_s[t+1]: the previous step left it in its slot of the output box
        double xs[N];
        {
            const unsigned prev = base + LY::SMO + (unsigned)((t + 1) % kSmtTSM) * LY::MEAN;
#pragma unroll
            for (int i = 0; i < N; i += 2) lds2(vec_pair(prev, i), xs[i], xs[i + 1]);
        }
        // the P_p box is free (and nobody reads the old P_s any more): fetch the next step's matrix while
        // this step computes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cta_sync();
        if (leader && t >= 1) issue_pc(t - 1);

        // ---- Cholesky of P_p (in place); a pivot that is not positive relative to its original diagonal entry
        // belongs to a direction in which P_p is numerically singular: zero gain there, as the reference's pinv
        // (kalmanfilter.py:455) -- the rule of smoother_kernel
        double dmax = 0.0, inv[N], d0[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            d0[j] = L[j * (j + 1) / 2 + j];
            dmax = fmax(dmax, d0[j]);
        }
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double pjj = L[j * (j + 1) / 2 + j];
            const bool okp = pjj > 1.5e-14 * d0[j] && pjj > 1e-15 * dmax && pjj > 0.0;
            if (!okp) status |= MKF_STATUS_CHOL_FAILED;
            const double iv = okp ? rsqrt(pjj) : 0.0;
            inv[j] = iv;
#pragma unroll
            for (int r = j + 1; r < N; ++r) L[r * (r + 1) / 2 + j] *= iv;
#pragma unroll
            for (int r = j + 1; r < N; ++r)
#pragma unroll
                for (int c = j + 1; c <= r; ++c)
                    L[r * (r + 1) / 2 + c] = fma(-L[r * (r + 1) / 2 + j], L[c * (c + 1) / 2 + j], L[r * (r + 1) / 2 + c]);
        }
        // ---- P_p^-1 = L^-T L^-1 explicitly: the n solves (L L^T) g = (A P_f)[:, c] become matrix-vector
        // products with n independent accumulators each instead of 2n dependent substitution steps.  W = L^-1
        // in place, column by column; a rejected pivot (inv = 0) leaves its row and column of W, hence of
        // P_p^-1, zero: zero gain in that direction, as before.
#pragma unroll
        for (int j = 0; j < N; ++j) {
            L[j * (j + 1) / 2 + j] = inv[j];
#pragma unroll
            for (int i = j + 1; i < N; ++i) {
                double acc = 0.0;
#pragma unroll
                for (int k = j; k < i; ++k) acc = fma(L[i * (i + 1) / 2 + k], L[k * (k + 1) / 2 + j], acc);
                L[i * (i + 1) / 2 + j] = -inv[i] * acc;
            }
        }
        double PI[NT];   // P_p^-1, packed lower triangle
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) {
                double acc = L[r * (r + 1) / 2 + r] * L[r * (r + 1) / 2 + c];
#pragma unroll
                for (int k = r + 1; k < N; ++k) acc = fma(L[k * (k + 1) / 2 + r], L[k * (k + 1) / 2 + c], acc);
                PI[r * (r + 1) / 2 + c] = acc;
            }

        // ---- the step's vectors ---------------------------------------------------------------------
        bar_wait(s ? bar_st1 : bar_st0, s ? ph_st1 : ph_st0);
        if (s) ph_st1 ^= 1u; else ph_st0 ^= 1u;
        const unsigned fcb = base + (s ? LY::FC1 : LY::FC0), fmb = base + (s ? LY::FM1 : LY::FM0),
                       pmb = base + (s ? LY::PM1 : LY::PM0);
        double delta[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            double p0, p1;
            lds2(vec_pair(pmb, i), p0, p1);
            delta[i] = xs[i] - p0;
            delta[i + 1] = xs[i + 1] - p1;
        }
        const unsigned slot = base + LY::SMO + (unsigned)(t % kSmtTSM) * LY::MEAN;

        // ---- rows c = NW * round + w: G[c] = (A P_f)[:, c]^T P_p^-1, x_s[c], M[c] = G[c] D, P_s[c][j <= c] ----
        smt_static_for<0, N / NW>([&](auto rdc) {
            constexpr int RD = decltype(rdc)::value;
            double pf[N], g[N];
            auto row_body = [&](auto cc) {
                constexpr int C = decltype(cc)::value;
#pragma unroll
                for (int i = 0; i < N; i += 2) lds2(cov_pair(fcb, C, i), pf[i], pf[i + 1]);   // row C of P_f (= column C)
                double z[N];
#pragma unroll
                for (int i = 0; i < N; ++i) z[i] = pf[i] * lds1(base + LY::AD + (unsigned)(i * 32 + lane) * 8u);   // (A P_f)[:, C]
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    double acc = 0.0;
#pragma unroll
                    for (int k = 0; k < N; ++k) {
                        const int hi = k > i ? k : i, lo = k > i ? i : k;
                        acc = fma(PI[hi * (hi + 1) / 2 + lo], z[k], acc);
                    }
                    g[i] = acc;
                }
                if (!kSmtLean || C < N - 1) {
#pragma unroll
                    for (int i = 0; i < N; i += 2) sts2(vec_pair(base + LY::G + (unsigned)C * LY::MEAN, i), g[i], g[i + 1]);
                }
            };
            auto row_tail = [&](auto cc) {
                constexpr int C = decltype(cc)::value;
                // x_s[t][C] = x_f[C] + G[C] . delta
                double xn = lds1(vec_pair(fmb, C & ~1) + (unsigned)(C & 1) * 8u);
#pragma unroll
                for (int i = 0; i < N; ++i) xn = fma(g[i], delta[i], xn);
                sts1(vec_pair(slot, C & ~1) + (unsigned)(C & 1) * 8u, xn);
                // M[C] = G[C] D
                double mrow[N];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    double acc = 0.0;
#pragma unroll
                    for (int k2 = 0; k2 < N; ++k2) {
                        const int hi = k2 > j ? k2 : j, lo = k2 > j ? j : k2;
                        acc = fma(g[k2], D[hi * (hi + 1) / 2 + lo], acc);
                    }
                    mrow[j] = acc;
                }
                // P_s[t][C][j] = P_f[C][j] + M[C] . G[j], j <= C; mirrored to (j, C)
                double prow[N];
#pragma unroll
                for (int j = 0; j <= C; ++j) {
                    double gj[N];
                    if (j < C) {
#pragma unroll
                        for (int i = 0; i < N; i += 2) lds2(vec_pair(base + LY::G + (unsigned)j * LY::MEAN, i), gj[i], gj[i + 1]);
                    } else {
#pragma unroll
                        for (int i = 0; i < N; ++i) gj[i] = g[i];
                    }
                    double s0 = pf[j], s1 = 0.0;
#pragma unroll
                    for (int k2 = 0; k2 < N; k2 += 2) {
                        s0 = fma(mrow[k2], gj[k2], s0);
                        s1 = fma(mrow[k2 + 1], gj[k2 + 1], s1);
                    }
                    prow[j] = s0 + s1;
                }
#pragma unroll
                for (int j = 0; j <= C; j += 2) {
                    if (j + 1 <= C) sts2(cov_pair(base + LY::SC, C, j), prow[j], prow[j + 1]);
                    else sts1(cov_pair(base + LY::SC, C, j), prow[j]);                       // (C, C) with C even: first half of its pair
                }
#pragma unroll
                for (int j = 0; j < C; ++j) sts1(cov_pair(base + LY::SC, j, C & ~1) + (unsigned)(C & 1) * 8u, prow[j]);   // mirror (j, C)
            };
            // the row index decides register indices and addresses: one unrolled copy per (round, warp)
#define MKF_SMT_ROW(FN)                                                                   \
            do {                                                                          \
                if (NW == 1 || w == 0) FN(std::integral_constant<int, RD * NW>{});        \
                else if (NW >= 2 && w == 1) FN(std::integral_constant<int, (RD * NW + 1) % N>{}); \
                else if (NW >= 4 && w == 2) FN(std::integral_constant<int, (RD * NW + 2) % N>{}); \
                else FN(std::integral_constant<int, (RD * NW + 3) % N>{});                \
            } while (0)
            MKF_SMT_ROW(row_body);
            // the box of P_s is about to be rewritten: its store (issued at the end of the previous step) must
            // have read it, and so must the store of the x_s box when this step starts a new one
            if (RD == 0 && leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            cta_sync();     // G rows of this round visible to every warp; (round 0) the boxes are free
            MKF_SMT_ROW(row_tail);
#undef MKF_SMT_ROW
        });

        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        cta_sync();
        if (leader) {
            if (kSmtLean && t >= 1) issue_stage(t - 1);
            store3d(&maps.sc, (int)t * LY::LINES, base + LY::SC);
            if (t % kSmtTSM == 0) store3d(&maps.sm, (int)t, base + LY::SMO);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (w == 0 && active && a.status) a.status[b] = status;
}

// Returns MKF_ERR_UNSUPPORTED when the call is not this kernel's (the caller then takes smoother_kernel).
template <int N>
int launch_smoother_thread_n(const SmootherArgs &a, cudaStream_t stream) {
    using LY = SmtLayout<N>;
    alignas(64) SmThreadMaps maps;
    memset(&maps, 0, sizeof(maps));
    using mkf_dfm::make_hist_map;
    double *pc = const_cast<double *>(a.pc), *fc = const_cast<double *>(a.fc), *pm = const_cast<double *>(a.pm),
           *fm = const_cast<double *>(a.fm);
    if (make_hist_map(&maps.pc, pc, 16, a.B, a.T, LY::LINES, 1) != MKF_OK ||
        make_hist_map(&maps.fc, fc, 16, a.B, a.T, LY::LINES, 1) != MKF_OK ||
        make_hist_map(&maps.pm, pm, N, a.B, a.T, 1, 1) != MKF_OK ||
        make_hist_map(&maps.fm, fm, N, a.B, a.T, 1, 1) != MKF_OK ||
        make_hist_map(&maps.sc, a.sc, 16, a.B, a.T, LY::LINES, 1) != MKF_OK ||
        make_hist_map(&maps.sm, a.sm, N, a.B, a.T, 1, kSmtTSM) != MKF_OK)
        return MKF_ERR_UNSUPPORTED;
    auto kern = smoother_thread_kernel<N>;
    const size_t smem = LY::TOTAL + 1024;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    const int64_t blocks = (a.B + 31) / 32;
    kern<<<(unsigned)blocks, 32 * MKF_SMT_WARPS, smem, stream>>>(a, maps);
    return mkf_host::check_launch();
}

int launch_smoother_thread(const SmootherArgs &a, cudaStream_t stream) {
    const bool diagA = a.flags & MKF_FLAG_A_DIAGONAL;
    if (!diagA || a.p != 0 || !a.sc || !a.sm || a.B < smoother_thread_min_b() || !mkf_dfm::hist_tma_enabled())
        return MKF_ERR_UNSUPPORTED;
    if (a.T * (int64_t)a.n * a.n / 16 >= (1LL << 31)) return MKF_ERR_UNSUPPORTED;
    if (a.n == 8) return launch_smoother_thread_n<8>(a, stream);
    if (a.n == 4) return launch_smoother_thread_n<4>(a, stream);
    return MKF_ERR_UNSUPPORTED;
}

}  // namespace
