// Thread-per-model Kalman filter -> -2 log L for Metran-shaped dynamic factor models.
//
// Replaces, for a whole batch at once, the chain Metran.get_mle -> _get_matrices ->
// SPKalmanFilter.run_filter (seqkalmanfilter) -> SPKalmanFilter.get_mle of the reference
// (metran/metran.py:605-622, 246-416; metran/kalmanfilter.py:315-390, 550-567).
//
// Model structure exploited (metran.py:283-384): n = M + K states,
// A = diag(phi), Q = diag(q), H = [I_M | Gamma(M x K)], x0 = 0, P0 = I.  Then for series i
//   h = e_i + sum_j Gamma[i][j] e_{M+j}
//   d = P h      costs N*K FMAs instead of N^2
//   f = h.d      costs K FMAs
// and P stays symmetric, so only its upper triangle (N(N+1)/2 doubles) lives in registers.
// One thread owns one model for all T steps: no shuffles and no shared-memory traffic for
// the state; the fp64 FMA pipe is the bound (kernel K1).
//
// Observations y[B][T][M] (model-major, NaN = missing) are streamed by the TMA engine:
// one elected lane issues ONE cp.async.bulk.tensor.2d per time tile -- a box of 32 model
// rows x (TT*M) contiguous doubles described by a CUtensorMap over y viewed as [B][T*M] --
// into a two-stage shared-memory ring (mbarrier complete_tx); lanes read their own row
// back with LDS.  When the rows are not 16-byte aligned (odd T*M) the kernel stages with
// 8-byte cp.async instead.  phi (and q, Gamma when M*K <= 16, n <= 8) stay in registers; for
// larger shapes q and Gamma are read from shared memory ([index][lane]) so that the register file is left to
// P, x and the instruction scheduler.  A one-series look-ahead variant of the dense step
// (MKF_LOOKAHEAD=1: d_{i+1} = P h_{i+1} - g_i (d_i.h_{i+1}) computed while 1/f_i is in
// flight) is kept for reference; it costs 44 more fp64 instructions per step and measured
// 6 % slower than the plain order on B200 (profiles/r01_microbench_fp64.md).
//
// Three forms of the step, chosen by warp votes on the observations (R = 0, no history):
//   full state   538 fp64 instructions at (M, K) = (5, 3): any pattern of missing values
//   factor only  312: this and the previous step observed completely in every lane; the
//                posterior is (c, C) = (x_c, P_cc), see dfm_collapsed.cuh
//   steady state  58: additionally every lane's C has stopped changing; reciprocal pivots,
//                multipliers and gains are constants captured from one factor-only step
// The time loop is flattened over (tile, step) and written as a full-state loop with a nested
// factor-only loop, so that P_ss, P_sc, x_s are dead registers in the latter.
#pragma once
#include <cuda.h>
#include <string.h>

#include "dfm_collapsed.cuh"
#include "mkf_common.cuh"

namespace mkf_dfm {

using namespace mkf;

struct DfmArgs {
    int64_t B, T;
    const double *y;
    int64_t y_stride;
    const double *phi, *q, *gamma, *r, *alpha;
    int64_t phi_stride, q_stride, gamma_stride, r_stride, alpha_stride;
    double dt;
    int warmup;
    int use_tma;  // rows 16-byte aligned -> 2-D tensor-map staging
    double *mle;
    int64_t *nobs;
    int32_t *status;
    // history mode (reference result contract, kalmanfilter.py:392-400); NULL pointers are skipped
    double *sigma, *detf;
    int32_t *cnt;
    double *fm, *fc, *pm, *pc;
    // compacted time grid (mkf_dfm_loglik_compact): y holds the observed steps only, tidx[j] is the
    // original time index of compacted step j (ascending; padding repeats the last index).  NULL = off
    const int32_t *tidx;
    int64_t tidx_stride;
    // hand-off between this kernel and the steady-state kernel (dfm_steady.cuh); ho_mode 0 = off
    int ho_mode;          // HO_MODE_*
    double *ho_ws;        // records, field-major: ho_ws[field * ho_pitch + model]
    int64_t ho_pitch;     // models rounded up to a multiple of 32
    int64_t *ho_tnext;    // per warp: next time step to do
    int32_t *ho_state;    // per warp: HO_*
    // history stores (HIST builds): bit mask of the arrays written by TMA tensor stores
    // (HIST_TMA_*; 0 = the shared-memory transposition with 256-byte STG pieces), and whether the
    // per-step scalars can be written as whole 32-byte sectors (bases 32-byte aligned)
    int hist_tma;
    int hist_vec;
};

enum { HIST_TMA_PC = 1, HIST_TMA_PM = 2, HIST_TMA_FC = 4, HIST_TMA_FM = 8, HIST_TMA_STATE = 15,
       HIST_TMA_SIGMA = 16, HIST_TMA_DETF = 32, HIST_TMA_CNT = 64, HIST_TMA_SCALARS = 112 };
// Tensor maps of the history arrays (kernel parameter of the HIST builds only).
template <bool HIST>
struct HistMaps {
    CUtensorMap pc, pm, fc, fm;      // [B][T][n][n] / [B][T][n]
    CUtensorMap sigma, detf, cnt;    // [B][T]
};
template <>
struct HistMaps<false> {};

enum { HO_MODE_OFF = 0, HO_MODE_HANDOFF = 1, HO_MODE_RESUME = 2 };
enum { HO_DONE = 0, HO_STEADY = 1, HO_RESUME = 2 };
// a warp is handed to the steady-state kernel only when at least this many steps remain
constexpr int64_t kMinSteadySteps = 96;

// Hand-off record of one model (doubles; the integers are small and exact in a double).
template <int M, int K>
struct Handoff {
    static constexpr int N = M + K;
    static constexpr int C = 0;                                   // factor means c[K]
    static constexpr int CHOLD = C + K;                           // held factor covariance, packed upper [K(K+1)/2]
    static constexpr int CONST = CHOLD + K * (K + 1) / 2;         // D[M*K], 1/f[M], multipliers, gains (dfm_collapsed.cuh)
    static constexpr int NCONST = collapsed_const_count(M, K);
    static constexpr int PHI = CONST + NCONST;                    // phi[N]
    static constexpr int SUM_S = PHI + N;
    static constexpr int SUM_C = SUM_S + 1;
    static constexpr int FPROD = SUM_C + 1;
    static constexpr int PEND_SIGMA = FPROD + 1;
    static constexpr int PEND_PROD = PEND_SIGMA + 1;
    static constexpr int FROZEN_PROD = PEND_PROD + 1;
    static constexpr int FEXP = FROZEN_PROD + 1;
    static constexpr int NOBS = FEXP + 1;
    static constexpr int NSIG = NOBS + 1;
    static constexpr int FLAGS = NSIG + 1;                        // bit 0: bad, bit 1: pend_on
    static constexpr int NF = FLAGS + 1;
};

constexpr int kBlock = 32;  // one warp per CTA: CTAs spread evenly over the 592 SMSPs
#ifndef MKF_STAGES
#define MKF_STAGES 2   // ring depth; tiles are fetched MKF_STAGES - 1 ahead
#endif
constexpr int kStages = MKF_STAGES;
#ifndef MKF_LOOKAHEAD
#define MKF_LOOKAHEAD 0
#endif
constexpr bool kLookAhead = MKF_LOOKAHEAD != 0;

// time steps per tile: TT*M even (16-byte rows) and, where possible, TT*M = 2 (mod 4) so the
// 64-bit LDS of a warp (row pitch TT*M doubles) is only 2-way bank conflicted
#ifndef MKF_TT_ODD
#define MKF_TT_ODD 6
#endif
// Information-form factor-only step (dfm_collapsed.cuh).  Its M-sized constants (L0^-1, Dt, Jt) take
// exactly the room of the elimination form's (D, S0, QG); the K-sized ones (R, Ft, Qt, det S0: 22
// doubles at K = 3) are loaded into registers at every entry of the factor-only loop.
#ifndef MKF_INFO
#define MKF_INFO 0
#endif
#ifndef MKF_COMPACT_POW
#define MKF_COMPACT_POW 1    // compacted grid: phi^k by binary powering for gaps below 256 steps (else exp(k ln phi))
#endif
// history mode is bound by its stores and wants many resident warps: short tiles there
__host__ __device__ constexpr int tile_steps(int M, bool hist = false, bool info = false) {
    return hist ? ((M % 2 == 1) ? 2 : 1) : ((M % 2 == 1) ? MKF_TT_ODD : ((M % 4 == 2) ? 5 : 6));
}
// q and the loadings live in registers when they fit (then no shared-memory copy is needed)
__host__ __device__ constexpr bool consts_in_regs(int M, int K) { return (M * K <= 16) && (M + K <= 8); }
__host__ __device__ constexpr int const_count(int M, int K) {
    return consts_in_regs(M, K) ? 0 : (M + K) + M * K;
}
// Collapsed (factor-only) dense step, dfm_collapsed.cuh: its constants sit in shared memory next to
// q / Gamma.  Enabled for the shapes whose CTA still fits 8 to an SM (2 warps per SMSP at <= 255
// registers): (228 KB - 8 x 1 KB reserved) / 8.
#ifndef MKF_COLLAPSE
#define MKF_COLLAPSE 1
#endif
__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr bool fits_8_ctas(int ring_steps, int M, int K, int consts) {
    return 128 + (kStages * 32 * ring_steps * M + 32 * const_count(M, K) + 32 * consts) * 8 + kStages * 8 <= 28160;
}
// the information form needs room for ITS constants and for the elimination form's (the capture step
// of the steady-state switch and the frozen gains use that layout, in the same place)
__host__ __device__ constexpr int info_region(int M, int K) {
    return cmax(collapsed_info_smem_count(M, K), collapsed_const_count(M, K));
}
__host__ __device__ constexpr bool info_shape(int M, int K) {
    return MKF_INFO != 0 && MKF_COLLAPSE != 0 && K >= 1 && fits_8_ctas(tile_steps(M, false, true), M, K, info_region(M, K));
}
__host__ __device__ constexpr bool collapse_shape(int M, int K) {
    return MKF_COLLAPSE != 0 && K >= 1 &&
           (info_shape(M, K) || fits_8_ctas(tile_steps(M, false, false), M, K, collapsed_const_count(M, K)));
}
__host__ __device__ constexpr int collapse_region(int M, int K) {
    return info_shape(M, K) ? info_region(M, K) : collapsed_const_count(M, K);
}
// (collapse: the R = 0, no-history build of a collapse-capable shape)
__host__ __device__ constexpr int ring_steps(int M, int K, bool hist, bool collapse) {
    return tile_steps(M, hist, collapse && !hist && info_shape(M, K));
}
__host__ __device__ constexpr int warp_doubles(int M, int K, bool hist = false, bool collapse = false) {
    return kStages * 32 * ring_steps(M, K, hist, collapse && collapse_shape(M, K)) * M + 32 * const_count(M, K) +
           ((collapse && collapse_shape(M, K)) ? 32 * collapse_region(M, K) : 0);
}
// History staging.  Legacy: one record set [N + N*N][33] transposed through shared memory.  Shapes
// with N*N a multiple of 16 doubles and N*8 a swizzle span (N = 4, 8) stage each lane's own record in
// a swizzled box that ONE cp.async.bulk.tensor.3d per array writes out (MKF_HIST_NBUF boxes, 1 KB
// aligned).  Both are followed by the staging of the step scalars: three boxes of 16 steps x 32
// models for tensor stores (sigma, detf: 4 KB each, cnt: 2 KB; 1 KB aligned), or -- in the same
// place, when the rows do not allow tensor stores -- the per-lane sector staging sigma[4][32],
// detf[4][32] (doubles), cnt[8][32] (int32).
#ifndef MKF_HIST_NBUF
#define MKF_HIST_NBUF 2    // staging boxes per warp: 2 = predicted and filtered records in separate boxes
#endif
// Time steps staged per tensor store.  The memory system, not the SM, bounds the history kernel, and
// it cares about the size of the contiguous pieces (tools/ubench_tma_store.cu, pure stores of this
// pattern): 64-byte mean pieces next to 512-byte covariance pieces 4.9 TB/s, without the means 5.9,
// 128-byte + 1-KB pieces 5.9, 1 KB alone 6.3 TB/s.
#ifndef MKF_HIST_TSC
#define MKF_HIST_TSC 1     // covariances: steps per store (N*N*8 bytes each)
#endif
#ifndef MKF_HIST_TSM
#define MKF_HIST_TSM 4     // means: steps per store (N*8 bytes each)
#endif
__host__ __device__ constexpr bool hist_tma_shape(int M, int K) { return (M + K) == 4 || (M + K) == 8; }
__host__ __device__ constexpr int hist_scalar_doubles() { return 1280 + 128; }
__host__ __device__ constexpr int hist_doubles(int M, int K) {
    const int n = M + K, legacy = (n + n * n) * 33,
              boxes = MKF_HIST_NBUF * (MKF_HIST_TSC * n * n + MKF_HIST_TSM * n) * 32 + 128;
    return (hist_tma_shape(M, K) && boxes > legacy ? boxes : legacy) + hist_scalar_doubles();
}
__host__ __device__ constexpr size_t smem_bytes(int M, int K, bool hist = false, bool collapse = false) {
    return 128 + (size_t)(kBlock / 32) *
                     ((warp_doubles(M, K, hist, collapse) + (hist ? hist_doubles(M, K) : 0)) * sizeof(double) +
                      kStages * sizeof(uint64_t));
}

#ifndef MKF_FREEZE
#define MKF_FREEZE 1
#endif
#ifndef MKF_CAPTURE_SPLIT
#define MKF_CAPTURE_SPLIT 0       // 1: the capture step is a separate copy of the factor-only step
#endif
#ifndef MKF_FREEZE_NCONV
#define MKF_FREEZE_NCONV 2        // consecutive unchanged checks before a lane holds its covariance
#endif
#ifndef MKF_FREEZE_EVERY_STEP
#define MKF_FREEZE_EVERY_STEP 0   // 1: check every step instead of once per tile (untested variant)
#endif
constexpr bool kFreeze = MKF_FREEZE != 0;
constexpr double kFreezeTol = 8.8817841970012523e-16;   // 2^-50
__device__ __forceinline__ void sts_f64(double *p, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(smem_u32(p)), "d"(v) : "memory");
}
__device__ __forceinline__ double lds_f64(const double *p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int M, int K, bool HASR, bool HIST>
__global__ void __launch_bounds__(kBlock) dfm_loglik_kernel(const DfmArgs a, const __grid_constant__ CUtensorMap ymap,
                                                            const __grid_constant__ HistMaps<HIST> hmaps) {
    constexpr int N = M + K;
    constexpr int NT = N * (N + 1) / 2;
    // factor-only dense steps need R = 0 and no history output
    constexpr bool COLLAPSE = !HASR && !HIST && collapse_shape(M, K);
    constexpr bool INFO = COLLAPSE && info_shape(M, K);    // ... taken in information form (dfm_collapsed.cuh)
    constexpr int TT = tile_steps(M, HIST, INFO);
    constexpr int PITCH = TT * M;  // doubles per model row of a tile
    constexpr int WARP_DOUBLES = warp_doubles(M, K, HIST, COLLAPSE) + (HIST ? hist_doubles(M, K) : 0);
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem_al =
        smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);  // TMA destinations: 128 B aligned

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *ring = reinterpret_cast<double *>(smem_al) + (size_t)warp * WARP_DOUBLES;
    double *cst = ring + kStages * 32 * PITCH + lane;  // constant i of this lane: cst[i * 32]
    double *ccs = cst + 32 * const_count(M, K);        // collapsed-step constant i: ccs[i * 32]
    double *hst = ring + warp_doubles(M, K, HIST, COLLAPSE);  // history staging [N + N*N][33] (HIST only)
    uint64_t *bars =
        reinterpret_cast<uint64_t *>(reinterpret_cast<double *>(smem_al) + (size_t)(kBlock / 32) * WARP_DOUBLES) +
        warp * kStages;
    const int64_t b0 = (int64_t)blockIdx.x * kBlock + (threadIdx.x & ~31);  // first model of the warp
    const int64_t b = b0 + lane;
    const bool active = b < a.B;
    if (b0 >= a.B) return;
    // hand-off (only the R = 0, no-history builds of collapse-capable shapes take part)
    constexpr bool HO_CAPABLE = !HASR && !HIST && collapse_shape(M, K) && kFreeze;
    using HO = Handoff<M, K>;
    const int64_t wid = b0 >> 5;
    const int ho_mode = HO_CAPABLE ? a.ho_mode : HO_MODE_OFF;
    if (ho_mode == HO_MODE_RESUME && a.ho_state[wid] != HO_RESUME) return;
    const bool tma = a.use_tma != 0;
    const bool shared_y = a.y_stride == 0;

    if (tma) {
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncwarp();
    }

    // ---- per-model constants ---------------------------------------------------------
    double phi[N];
    double R[M];
    constexpr bool GREG = consts_in_regs(M, K);   // q and the loadings in registers when they fit
    double qreg[GREG ? M + K : 1];
    double Greg[GREG ? M * K + 1 : 1];
    const int64_t bb = active ? b : b0;  // idle lanes of a ragged last warp shadow its first model
    {
        double q[N], G[M * K + 1];
        const double *gp = a.gamma + bb * a.gamma_stride;
#pragma unroll
        for (int i = 0; i < M * K; ++i) G[i] = gp[i];
#pragma unroll
        for (int i = 0; i < M; ++i) R[i] = (HASR && a.r) ? a.r[bb * a.r_stride + i] : 0.0;
        if (a.alpha) {
            // metran.py:262-263, 310-322
            const double *al = a.alpha + bb * a.alpha_stride;
#pragma unroll
            for (int i = 0; i < N; ++i) phi[i] = exp(-a.dt / al[i]);
#pragma unroll
            for (int i = 0; i < M; ++i) {
                double comm = 0.0;
#pragma unroll
                for (int j = 0; j < K; ++j) comm += G[i * K + j] * G[i * K + j];
                q[i] = (1.0 - phi[i] * phi[i]) * (1.0 - comm);
            }
#pragma unroll
            for (int j = M; j < N; ++j) q[j] = 1.0 - phi[j] * phi[j];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                phi[i] = a.phi[bb * a.phi_stride + i];
                q[i] = a.q[bb * a.q_stride + i];
            }
        }
        if (GREG) {
#pragma unroll
            for (int i = 0; i < N; ++i) qreg[GREG ? i : 0] = q[i];
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) cst[i * 32] = q[i];
#pragma unroll
            for (int i = 0; i < M * K; ++i) cst[(N + i) * 32] = G[i];
        }
        if (GREG) {
#pragma unroll
            for (int i = 0; i < M * K; ++i) Greg[GREG ? i : 0] = G[i];
        }
        if constexpr (COLLAPSE && !INFO) collapsed_consts<M, K>(phi, q, G, [&](int i, double v) { ccs[i * 32] = v; });
        __syncwarp();
    }
#define MKF_QQ(i) (GREG ? qreg[GREG ? (i) : 0] : lds_f64(cst + (i) * 32))
#define MKF_GG(i) (GREG ? Greg[GREG ? (i) : 0] : lds_f64(cst + (N + (i)) * 32))

    // Compacted time grid (only the R-capable, no-history build; never together with the factor-only
    // step).  Between two observed steps tidx[j-1] < tidx[j] lie k - 1 steps without any observation:
    // the reference runs k predicts (kalmanfilter.py:318-331, the update branch :335 not taken), which
    // for diagonal A, Q have the closed form  x <- phi^k x,  P_ij <- phi_i^k phi_j^k P_ij +
    // delta_ij q_i (1 - phi_i^2k) / (1 - phi_i^2), with phi^k = exp(k ln phi); a warp whose lanes all have
    // k = 1 uses phi and q themselves (bit-identical to the full-grid kernel).
    constexpr bool CAN_COMPACT = HASR && !HIST;
    const bool compact = CAN_COMPACT && a.tidx != nullptr;
    double lnphi[CAN_COMPACT ? N : 1], cq[CAN_COMPACT ? N : 1];
    if constexpr (CAN_COMPACT) {
        if (compact) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                lnphi[i] = a.alpha ? -a.dt / a.alpha[bb * a.alpha_stride + i] : log(phi[i]);
                const double den = -expm1(2.0 * lnphi[i]);          // 1 - phi^2
                cq[i] = den > 0.0 ? MKF_QQ(i) / den : 0.0;
            }
        }
    }
    const int32_t *tix = compact ? a.tidx + bb * a.tidx_stride : nullptr;
    int t_prev = -1, t_cur = 0;

    // ---- state: x0 = 0, P0 = I (kalmanfilter.py:747-750) ------------------------------
    double x[N], P[NT];
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 0.0;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = r; c < N; ++c) P[tri<N>(r, c)] = (r == c) ? 1.0 : 0.0;

    // History mode.  Legacy path (any shape): the warp transposes its 32 per-thread records through
    // shared memory ([element][lane], pitch 33: conflict-free writes, 2-way reads) so that every store
    // instruction covers one contiguous 256-byte piece of a model's [n] / [n][n] block.
    // TMA path (N = 4, 8; hist_tma): every lane writes its OWN record with 16-byte shared-memory stores
    // into a staging box laid out as the tensor maps' boxes -- covariances {16 doubles, 32 models,
    // N*N/16 lines} under the 128-byte swizzle, means {N doubles, 32 models} under the N*8-byte swizzle
    // (both conflict-free for consecutive lanes) -- and lane 0 issues one cp.async.bulk.tensor.3d
    // shared -> global per array (SASS UTMASTG): no LDS/STG transposition, the store drains while
    // the next phase computes, rows beyond B are clipped by the tensor map.
    constexpr bool TST = HIST && hist_tma_shape(M, K);
    constexpr int kLines = (N * N) / 16;
    constexpr int TSC = MKF_HIST_TSC, TSM = MKF_HIST_TSM;   // steps per store
    constexpr unsigned kCovB = (unsigned)kLines * 32u * 128u, kMeanB = 32u * N * 8u;   // one step's records
    constexpr unsigned kBoxB = TSC * kCovB + TSM * kMeanB;
    const int htm = HIST ? a.hist_tma : 0;
    const int tst = TST ? (htm & HIST_TMA_STATE) : 0;
    const unsigned hbox = (smem_u32(hst) + 1023u) & ~1023u;
    double *sc_base = hst + (hist_doubles(M, K) - hist_scalar_doubles());
    const unsigned sbox = (smem_u32(sc_base) + 1023u) & ~1023u;     // sigma box; detf at + 4096, cnt at + 8192
    double *sc_sigma = sc_base + lane;   // per-lane sector staging: [slot * 32]
    double *sc_detf = sc_sigma + 128;
    int *sc_cnt = reinterpret_cast<int *>(sc_base + 256) + lane;
    auto emit = [&](const int which, int64_t t) {   // which: 0 = predicted, 1 = filtered
        if constexpr (HIST) {
            double *gmean = which ? a.fm : a.pm, *gcov = which ? a.fc : a.pc;
            if constexpr (TST) {
                const int bit_c = which ? HIST_TMA_FC : HIST_TMA_PC, bit_m = which ? HIST_TMA_FM : HIST_TMA_PM;
                if (tst & (bit_c | bit_m)) {
                    const unsigned box = hbox + (MKF_HIST_NBUF > 1 ? (unsigned)which * kBoxB : 0u);
                    // the store that last read this box must have drained it
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(MKF_HIST_NBUF - 1) : "memory");
                    __syncwarp();
                    const int uc = (int)(t % TSC), um = (int)(t % TSM);
                    const bool last = t == a.T - 1;
                    if (tst & bit_c) {
                        const unsigned rowb = box + (unsigned)uc * kCovB + (unsigned)lane * 128u, sw = (unsigned)(lane & 7);
#pragma unroll
                        for (int e = 0; e < N * N; e += 2) {
                            const int r = e / N, c = e % N, line = e / 16, ch = (e % 16) / 2;
                            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rowb + (unsigned)line * 4096u +
                                                                                 (((unsigned)ch ^ sw) << 4)),
                                         "d"(P[sym<N>(r, c)]), "d"(P[sym<N>(r, c + 1)])
                                         : "memory");
                        }
                    }
                    if (tst & bit_m) {
                        const unsigned rowb = box + TSC * kCovB + (unsigned)um * kMeanB + (unsigned)lane * (N * 8u);
                        const unsigned sw = (N == 8) ? (unsigned)((lane >> 1) & 3) : (unsigned)((lane >> 2) & 1);
#pragma unroll
                        for (int e = 0; e < N; e += 2)
                            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(rowb + (((unsigned)(e / 2) ^ sw) << 4)),
                                         "d"(x[e]), "d"(x[e + 1])
                                         : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        // a box goes out when its last step is in (or the series ends: the steps beyond T
                        // are clipped by the tensor map); one bulk group per emit, possibly empty, so
                        // that "all but the latest group" above always means "the previous use of this box"
                        if ((tst & bit_c) && (uc == TSC - 1 || last))
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                                             which ? &hmaps.fc : &hmaps.pc),
                                         "r"(0), "r"((int)b0), "r"((int)(t - uc) * kLines), "r"(box)
                                         : "memory");
                        if ((tst & bit_m) && (um == TSM - 1 || last))
                            asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(
                                             which ? &hmaps.fm : &hmaps.pm),
                                         "r"(0), "r"((int)b0), "r"((int)(t - um)), "r"(box + TSC * kCovB)
                                         : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (tst & bit_c) gcov = nullptr;
                    if (tst & bit_m) gmean = nullptr;
                }
            }
            if (!gcov && !gmean) return;
#pragma unroll
            for (int i = 0; i < N; ++i) hst[i * 33 + lane] = x[i];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c = 0; c < N; ++c) hst[(N + r * N + c) * 33 + lane] = P[sym<N>(r, c)];
            __syncwarp();
            const int rows = (int)min((int64_t)32, a.B - b0);
            if (gcov) {
                if (rows == 32) {   // full warp: unrolled so that several LDS/STG pairs are in flight
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        double *dst = gcov + ((b0 + rr) * a.T + t) * (N * N);
#pragma unroll
                        for (int e = lane; e < N * N; e += 32) dst[e] = hst[(N + e) * 33 + rr];
                    }
                } else {
                    for (int rr = 0; rr < rows; ++rr) {
                        double *dst = gcov + ((b0 + rr) * a.T + t) * (N * N);
#pragma unroll
                        for (int e = lane; e < N * N; e += 32) dst[e] = hst[(N + e) * 33 + rr];
                    }
                }
            }
            if (gmean) {
#pragma unroll
                for (int idx = lane; idx < 32 * N; idx += 32) {
                    const int rr = idx / N, e = idx % N;
                    if (rr < rows) gmean[((b0 + rr) * a.T + t) * N + e] = hst[e * 33 + rr];
                }
            }
            __syncwarp();
        }
    };
    // Step scalars (sigma, detf: f64; cnt: i32) at [b][t]: one value per lane and step, 8 (4) bytes 32
    // model rows apart -- written directly that is 32 partial sectors per store instruction, and the
    // read-modify-write traffic it causes costs a quarter of the history kernel's throughput
    // (tools/ubench_tma_store.cu).  Each lane therefore collects the values of one 32-byte sector
    // of ITS row in shared memory and writes the sector whole (STG.256) when its last element arrives;
    // the ragged head and tail of a row (first / last sector not covered completely) go out as scalars.
    // When the rows allow it (T even / a multiple of 4: 16-byte row pitch) the scalars of 16 steps are
    // staged as a {16 steps, 32 models} box (128-byte rows under the 128-byte swizzle; cnt: 64-byte rows
    // under the 64-byte swizzle) and written by one 2-D tensor store per array: whole lines instead of
    // sectors.  All 32 lanes take part (rows beyond B are clipped).
    auto put_scalars = [&](int64_t t, double sigma, double detf, int cnt) {
        if constexpr (HIST) {
            if (htm & HIST_TMA_SCALARS) {
                const int u = (int)(t & 15);
                const bool last = t == a.T - 1;
                if (u == 0) {
                    // the boxes' previous store has drained: with state stores in flight it is older than
                    // the two groups of this step's emits, otherwise it is the latest group
                    if (lane == 0) {
                        if (tst) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    __syncwarp();
                }
                const unsigned o8 = (unsigned)lane * 128u + ((((unsigned)u >> 1) ^ (unsigned)(lane & 7)) << 4) + ((unsigned)u & 1u) * 8u;
                if (htm & HIST_TMA_SIGMA) asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbox + o8), "d"(sigma) : "memory");
                if (htm & HIST_TMA_DETF) asm volatile("st.shared.f64 [%0], %1;" ::"r"(sbox + 4096u + o8), "d"(detf) : "memory");
                if (htm & HIST_TMA_CNT) {
                    const unsigned o4 = (unsigned)lane * 64u + ((((unsigned)u >> 2) ^ (unsigned)((lane >> 1) & 3)) << 4) + ((unsigned)u & 3u) * 4u;
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbox + 8192u + o4), "r"(cnt) : "memory");
                }
                if (u == 15 || last) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        const int c0 = (int)(t - u), c1 = (int)b0;
                        if (htm & HIST_TMA_SIGMA)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&hmaps.sigma),
                                         "r"(c0), "r"(c1), "r"(sbox)
                                         : "memory");
                        if (htm & HIST_TMA_DETF)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&hmaps.detf),
                                         "r"(c0), "r"(c1), "r"(sbox + 4096u)
                                         : "memory");
                        if (htm & HIST_TMA_CNT)
                            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&hmaps.cnt),
                                         "r"(c0), "r"(c1), "r"(sbox + 8192u)
                                         : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                return;
            }
            if (!active) return;
            const int64_t o = b * a.T + t;
            if (!a.hist_vec) {
                if (a.sigma) a.sigma[o] = sigma;
                if (a.detf) a.detf[o] = detf;
                if (a.cnt) a.cnt[o] = cnt;
                return;
            }
            const int s4 = (int)(o & 3), s8 = (int)(o & 7);
            const bool last = t == a.T - 1;
            if (a.sigma) sc_sigma[s4 * 32] = sigma;
            if (a.detf) sc_detf[s4 * 32] = detf;
            if (a.cnt) sc_cnt[s8 * 32] = cnt;
            if (s4 == 3 || last) {
                const int first = (int)max((int64_t)0, (int64_t)s4 - t);   // slots [first, s4] hold steps of this series
                if (first == 0 && s4 == 3) {
                    if (a.sigma)
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.sigma + o - 3), "d"(sc_sigma[0]),
                                     "d"(sc_sigma[32]), "d"(sc_sigma[64]), "d"(sc_sigma[96])
                                     : "memory");
                    if (a.detf)
                        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(a.detf + o - 3), "d"(sc_detf[0]),
                                     "d"(sc_detf[32]), "d"(sc_detf[64]), "d"(sc_detf[96])
                                     : "memory");
                } else {
                    for (int s = first; s <= s4; ++s) {
                        if (a.sigma) a.sigma[o - s4 + s] = sc_sigma[s * 32];
                        if (a.detf) a.detf[o - s4 + s] = sc_detf[s * 32];
                    }
                }
            }
            if (a.cnt && (s8 == 7 || last)) {
                const int first = (int)max((int64_t)0, (int64_t)s8 - t);
                if (first == 0 && s8 == 7) {
                    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(a.cnt + o - 7),
                                 "r"(sc_cnt[0]), "r"(sc_cnt[32]), "r"(sc_cnt[64]), "r"(sc_cnt[96]), "r"(sc_cnt[128]),
                                 "r"(sc_cnt[160]), "r"(sc_cnt[192]), "r"(sc_cnt[224])
                                 : "memory");
                } else {
                    for (int s = first; s <= s8; ++s) a.cnt[o - s8 + s] = sc_cnt[s * 32];
                }
            }
        }
    };

    KahanSum sum_sigma;
    double fprod = 1.0;  // running product of innovation variances, mantissa in [1,2)
    int fexp = 0;        // ... and its binary exponent: sum log f = log(fprod) + fexp ln 2
    int nobs = 0;
    int nsig = 0;
    bool bad = false;
    bool pend_on = false;  // previous step's contribution, folded one step later
    double pend_sigma = 0.0, pend_prod = 1.0;

    const int64_t T = a.T;
    const double *yb = a.y + bb * a.y_stride;
    // Resume mode: the steady-state kernel met a step with a missing value at time t_res and handed
    // the warp back.  (c, held C, y(t_res - 1)) expand to the full posterior exactly as at a gap
    // inside this kernel; the time loop then starts in the middle of tile t_res / TT.
    int64_t t_res = 0;
    if (ho_mode == HO_MODE_RESUME) {
        if constexpr (HO_CAPABLE) {
            const double *wsb = a.ho_ws + b;
            const int64_t Bp = a.ho_pitch;
            t_res = a.ho_tnext[wid];
#pragma unroll
            for (int j = 0; j < K; ++j) x[M + j] = wsb[(HO::C + j) * Bp];
#pragma unroll
            for (int j = 0; j < K; ++j)
#pragma unroll
                for (int l = j; l < K; ++l) P[tri<N>(M + j, M + l)] = wsb[(HO::CHOLD + tri<K>(j, l)) * Bp];
            sum_sigma.s = wsb[HO::SUM_S * Bp];
            sum_sigma.c = wsb[HO::SUM_C * Bp];
            fprod = wsb[HO::FPROD * Bp];
            pend_sigma = wsb[HO::PEND_SIGMA * Bp];
            pend_prod = wsb[HO::PEND_PROD * Bp];
            fexp = (int)wsb[HO::FEXP * Bp];
            nobs = (int)wsb[HO::NOBS * Bp];
            nsig = (int)wsb[HO::NSIG * Bp];
            const int fl = (int)wsb[HO::FLAGS * Bp];
            bad = (fl & 1) != 0;
            pend_on = (fl & 2) != 0;
            double yq[M];
#pragma unroll
            for (int i = 0; i < M; ++i) yq[i] = yb[(t_res - 1) * M + i];
            collapsed_expand<M, K>(x, P, yq, [&](int i) { return MKF_GG(i); });
        }
    }
    // (a box must start 16-byte aligned in global memory or the TMA engine faults with "illegal
    // instruction": TT*M is even for every M, see tile_steps)
    const int64_t tile_base = t_res / TT;                  // tiles are counted from here (mbarrier phases start at 0)
    const int64_t ntiles = (T + TT - 1) / TT - tile_base;
    // shared y: the tile holds ONE row that every lane reads (broadcast)
    double *myrow = ring + ((shared_y && tma) || !active ? 0 : (size_t)lane * PITCH);  // + stage * 32 * PITCH
    constexpr unsigned kTileBytes = 32u * PITCH * 8u;

    auto issue_tile = [&](int64_t tile) {
        if (tile >= ntiles) {
            if (!tma) cp_async_commit();
            return;
        }
        const int s = (int)(tile % kStages);
        const int64_t tb = (tile_base + tile) * TT;   // first step of the box
        double *stage = ring + (size_t)s * 32 * PITCH;
        if (tma) {
            // WAR across proxies: the lanes' earlier LDS of this stage vs. the async write
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_expect_tx(&bars[s], shared_y ? (unsigned)(PITCH * 8) : kTileBytes);
                tma_load_2d(stage, &ymap, (int)(tb * M), shared_y ? 0 : (int)b0, &bars[s]);
            }
        } else {
            const int nel = (int)min((int64_t)TT, T - tb) * M;
            double *dst = stage + (size_t)lane * PITCH;
            const double *src = yb + tb * M;
            if (active)
                for (int e = 0; e < nel; ++e) cp_async8(dst + e, src + e);
            cp_async_commit();
        }
    };

    // ---- time loop, flattened over (tile, step) ----------------------------------------
    // Two loops advance through the same sequence of steps with next_step(): the full-state
    // loop and, for COLLAPSE builds, the factor-only loop (dfm_collapsed.cuh) that takes over
    // after a step every lane observed completely.  Inside the factor-only loop P_ss, P_sc and
    // x_s are dead, which leaves their registers to the instruction scheduler.
    int64_t tile = 0, t0 = 0;
    int tt = 0, steps = 0;
    const double *ys = myrow;
    double yv[M], yn[M];
    auto open_tile = [&]() {   // start fetching tile + kStages - 1, wait for `tile`
        issue_tile(tile + kStages - 1);
        const int s = (int)(tile % kStages);
        if (tma) {
            mbar_wait(&bars[s], (unsigned)((tile / kStages) & 1));
        } else {
            cp_async_wait<kStages - 1>();
        }
        ys = myrow + (size_t)s * 32 * PITCH;
        t0 = (tile_base + tile) * TT;
        steps = (int)min((int64_t)TT, T - t0);
        tt = 0;
    };
    // prefetch the next step's observations of this tile (LDS latency off the path)
    auto prefetch_next = [&]() {
        const int tn = (tt + 1 < steps) ? tt + 1 : tt;
#pragma unroll
        for (int i = 0; i < M; ++i) yn[i] = ys[tn * M + i];
    };
    auto advance = [&]() -> bool {
        if (++tt < steps) return true;
        if (++tile >= ntiles) return false;
        open_tile();
        return true;
    };
    auto next_step = [&]() -> bool {   // full-state loop: yv = observations of the step to do
        const bool same_tile = tt + 1 < steps;
        const bool m = advance();
#pragma unroll
        for (int i = 0; i < M; ++i) yv[i] = same_tile ? yn[i] : ys[i];
        return m;
    };
    // fold the PREVIOUS step's sigma / product of f into the running sums, in the shadow of
    // the current step's arithmetic (branch-free: neutral elements when it does not count)
    auto fold = [&]() {
        sum_sigma.add(pend_on ? pend_sigma : 0.0);
        fprod *= pend_on ? pend_prod : 1.0;
        const int hi = __double2hiint(fprod);
        const int ex = ((hi >> 20) & 0x7ff) - 1023;
        fexp += ex;
        fprod = __hiloint2double(hi - (ex << 20), __double2loint(fprod));
    };
    auto account = [&](int cnt, double sigma, double stepprod) {
        if (HIST) put_scalars(t0 + tt, sigma, (cnt > 0) ? log(stepprod) : 0.0, cnt);
        // quirk Q4 (kalmanfilter.py:563-566): the first `warmup` OBSERVED steps are
        // dropped from the sums, nobs counts raw steps t >= warmup
        pend_on = (cnt > 0) && (nsig >= a.warmup);
        pend_sigma = sigma;
        pend_prod = stepprod;
        nsig += (cnt > 0) ? 1 : 0;
        nobs += ((compact ? (int64_t)t_cur : t0 + tt) >= a.warmup) ? cnt : 0;
    };

#pragma unroll
    for (int s0 = 0; s0 < kStages - 1; ++s0) issue_tile(s0);
    bool more = ntiles > 0;
    if (more) {
        open_tile();
        tt = (int)(t_res - t0);          // 0 unless resuming in the middle of a tile
#pragma unroll
        for (int i = 0; i < M; ++i) yv[i] = ys[tt * M + i];
    }
    while (more) {
        bool warp_all;
        {
            prefetch_next();
            unsigned all_obs = 1u, any_obs = 0u;   // bitwise: no short-circuit branches
#pragma unroll
            for (int i = 0; i < M; ++i) {
                const unsigned o = is_observed(yv[i]) ? 1u : 0u;
                all_obs &= o;
                any_obs |= o;
            }
            const bool warp_any = __any_sync(0xffffffffu, any_obs != 0u);
            warp_all = __all_sync(0xffffffffu, all_obs != 0u);
            fold();
            double sigma = 0.0, stepprod = 1.0;
            int cnt = 0;
            // ---- predict: x = phi x ; P = phi P phi + Q (kalmanfilter.py:318-331) -----
            bool kstep = false;
            if constexpr (CAN_COMPACT) {
                if (compact) {
                    t_cur = tix[t0 + tt];     // (fetching it one step ahead measured no faster: 0.374 vs 0.367 ms at C2)
                    kstep = !__all_sync(0xffffffffu, t_cur - t_prev == 1);
                }
            }
            if (!kstep) {
#pragma unroll
                for (int i = 0; i < N; ++i) x[i] *= phi[i];
#pragma unroll
                for (int r = 0; r < N; ++r)
#pragma unroll
                    for (int c = r; c < N; ++c) {
                        double v = (phi[r] * P[tri<N>(r, c)]) * phi[c];
                        P[tri<N>(r, c)] = (r == c) ? v + MKF_QQ(r) : v;
                    }
            } else {
                if constexpr (CAN_COMPACT) {
                    // k = t_cur - t_prev predicts in one (k = 0 on padding steps: identity; lanes with
                    // k = 1 inside a warp that took this branch get phi^1 and q to rounding)
                    const int ki = t_cur - t_prev;
                    const double kk = (double)ki;
                    double pk[N], qk[N];
                    // phi^k: binary powering from phi itself while every lane's gap is short (up to 8 squarings and
                    // 8 conditional multiplies per state: a third of the instructions of eight exp calls, and the
                    // same rounded phi the reference multiplies by k times); exp(k ln phi) for long gaps
                    if (MKF_COMPACT_POW && __all_sync(0xffffffffu, ki < 256)) {
                        double bs[N];
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            bs[i] = phi[i];
                            pk[i] = 1.0;
                        }
                        const int kmax = __reduce_max_sync(0xffffffffu, ki);
                        for (int bit = 0; (kmax >> bit) != 0; ++bit) {
                            const bool on = ((ki >> bit) & 1) != 0;
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                pk[i] = on ? pk[i] * bs[i] : pk[i];
                                bs[i] *= bs[i];
                            }
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < N; ++i) pk[i] = exp(kk * lnphi[i]);
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        // 1 - phi^2k by one FMA: its ABSOLUTE error is 2^-53, times cq <= q / (1 - phi^2)
                        // (= 1 - communality or 1 for Metran) -- far below the 1e-9 contract even where the
                        // difference cancels (huge alpha), and expm1 per state and step is not free
                        const double om = fma(-pk[i], pk[i], 1.0);
                        qk[i] = cq[i] > 0.0 ? cq[i] * om : kk * MKF_QQ(i);
                    }
#pragma unroll
                    for (int i = 0; i < N; ++i) x[i] *= pk[i];
#pragma unroll
                    for (int r = 0; r < N; ++r)
#pragma unroll
                        for (int c = r; c < N; ++c) {
                            double v = (pk[r] * P[tri<N>(r, c)]) * pk[c];
                            P[tri<N>(r, c)] = (r == c) ? v + qk[r] : v;
                        }
                }
            }
            t_prev = t_cur;
            emit(0, t0 + tt);

            // ---- sequential scalar updates (kalmanfilter.py:341-378) -------------------
            if (!warp_any) {
                // nobody in the warp observes anything at this step (irregular sampling on a
                // daily grid: 94.5 % of the steps of the reference's example data): predict only
            } else if (warp_all) {
                // Dense step, straight-line code with one-series look-ahead: while the
                // reciprocal of f_i is in flight, e = P h_{i+1} and s = h_{i+1}.d_i are
                // computed from the NOT yet updated P; afterwards
                //   d_{i+1} = (P - g d^T) h_{i+1} = e - g s
                // which is the same vector the reference gets from the updated P.
                cnt = M;
                if (kLookAhead) {
                double d[N], e[N], gc[K + 1], gn[K + 1];
#pragma unroll
                for (int j = 0; j < K; ++j) gc[j] = MKF_GG(j);
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    double s = P[sym<N>(r, 0)];
#pragma unroll
                    for (int j = 0; j < K; ++j) s = fma(gc[j], P[sym<N>(r, M + j)], s);
                    d[r] = s;
                }
                double f = HASR ? d[0] + R[0] : d[0];
#pragma unroll
                for (int j = 0; j < K; ++j) f = fma(gc[j], d[M + j], f);
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    const double rf = rcp_nr(f);
                    bad = bad || f_is_bad(f);
                    stepprod *= f;
                    double snext = 0.0;
                    if (i + 1 < M) {
#pragma unroll
                        for (int j = 0; j < K; ++j) gn[j] = MKF_GG((i + 1) * K + j);
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            double s = P[sym<N>(r, i + 1)];
#pragma unroll
                            for (int j = 0; j < K; ++j) s = fma(gn[j], P[sym<N>(r, M + j)], s);
                            e[r] = s;
                        }
                        snext = d[i + 1];
#pragma unroll
                        for (int j = 0; j < K; ++j) snext = fma(gn[j], d[M + j], snext);
                    }
                    double v = yv[i] - x[i];
#pragma unroll
                    for (int j = 0; j < K; ++j) v = fma(-gc[j], x[M + j], v);
                    double g[N];
#pragma unroll
                    for (int r = 0; r < N; ++r) g[r] = d[r] * rf;
#pragma unroll
                    for (int r = 0; r < N; ++r) {
                        x[r] = fma(g[r], v, x[r]);
#pragma unroll
                        for (int c = r; c < N; ++c) P[tri<N>(r, c)] = fma(-g[r], d[c], P[tri<N>(r, c)]);
                    }
                    sigma = fma(v * v, rf, sigma);
                    if (i + 1 < M) {
#pragma unroll
                        for (int r = 0; r < N; ++r) d[r] = fma(-g[r], snext, e[r]);
                        f = HASR ? d[i + 1] + R[i + 1] : d[i + 1];
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            f = fma(gn[j], d[M + j], f);
                            gc[j] = gn[j];
                        }
                    }
                }
                } else {
                    // same arithmetic without the look-ahead (44 fewer fp64 instructions per
                    // step; the reciprocal chain is hidden by the second warp of the SMSP only)
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        double gc[K + 1];
#pragma unroll
                        for (int j = 0; j < K; ++j) gc[j] = MKF_GG(i * K + j);
                        double v = yv[i] - x[i];
#pragma unroll
                        for (int j = 0; j < K; ++j) v = fma(-gc[j], x[M + j], v);
                        double d[N];
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            double s = P[sym<N>(r, i)];
#pragma unroll
                            for (int j = 0; j < K; ++j) s = fma(gc[j], P[sym<N>(r, M + j)], s);
                            d[r] = s;
                        }
                        double f = HASR ? d[i] + R[i] : d[i];
#pragma unroll
                        for (int j = 0; j < K; ++j) f = fma(gc[j], d[M + j], f);
                        const double rf = rcp_nr(f);
                        bad = bad || f_is_bad(f);
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            const double g = d[r] * rf;
                            x[r] = fma(g, v, x[r]);
#pragma unroll
                            for (int c = r; c < N; ++c) P[tri<N>(r, c)] = fma(-g, d[c], P[tri<N>(r, c)]);
                        }
                        sigma = fma(v * v, rf, sigma);
                        stepprod *= f;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < M; ++i) {
                    if (is_observed(yv[i])) {
                        ++cnt;
                        double gc[K + 1];
#pragma unroll
                        for (int j = 0; j < K; ++j) gc[j] = MKF_GG(i * K + j);
                        double v = yv[i] - x[i];
#pragma unroll
                        for (int j = 0; j < K; ++j) v = fma(-gc[j], x[M + j], v);
                        double d[N];
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            double s = P[sym<N>(r, i)];
#pragma unroll
                            for (int j = 0; j < K; ++j) s = fma(gc[j], P[sym<N>(r, M + j)], s);
                            d[r] = s;
                        }
                        double f = HASR ? d[i] + R[i] : d[i];
#pragma unroll
                        for (int j = 0; j < K; ++j) f = fma(gc[j], d[M + j], f);
                        const double rf = rcp_nr(f);
                        bad = bad || f_is_bad(f);
#pragma unroll
                        for (int r = 0; r < N; ++r) {
                            const double g = d[r] * rf;
                            x[r] = fma(g, v, x[r]);
#pragma unroll
                            for (int c = r; c < N; ++c) P[tri<N>(r, c)] = fma(-g, d[c], P[tri<N>(r, c)]);
                        }
                        sigma = fma(v * v, rf, sigma);
                        stepprod *= f;
                    }
                }
            }
            emit(1, t0 + tt);
            account(cnt, sigma, stepprod);
        }
        if (!(COLLAPSE && warp_all)) more = next_step();
        if constexpr (COLLAPSE) {
            // the step just done was observed completely in every lane: (c, C, y_prev) carry the
            // whole posterior from here until a step with a missing value turns up.  In this loop
            // yv is NOT rotated: it keeps the observations of the last step of the previous tile
            // (needed as y(t-1) at tt = 0, when that tile's stage is already being refilled).
            if (warp_all) {
                // Steady state (dfm_collapsed.cuh).  Each lane watches its own C; when it has stopped
                // changing the lane HOLDS it (every later step restarts from the held value), so
                // that its results do not depend on when the rest of the warp converges.  Once all
                // lanes hold, one more full step stores their (now constant) reciprocal pivots,
                // multipliers and gains and the warp switches to the data-only step.
                constexpr int KT = K * (K + 1) / 2;
                bool frozen = false, capture_next = false, lane_holds = false, handoff_now = false;
                double frozen_prod = 1.0;
                double cchk = -1.0;   // weighted sum of the entries of C at the last check
                int nconv = 0;        // consecutive checks without change
                volatile double Chold[KT];   // local memory on purpose: read back once per step
                int since_chk = 0;           // INFO builds: steps since the last convergence check
                constexpr int kInfoSmem = collapsed_info_smem_count(M, K);
                double kreg[INFO ? collapsed_info_const_count(M, K) - kInfoSmem : 1];   // R, Ft, Qt, det S0
                if constexpr (INFO) {
                    // the constants of the information form (the region holds the elimination form's after a
                    // capture): once per entry of this loop, i.e. once per run of fully observed steps
                    double qq[N], gg[M * K];
#pragma unroll
                    for (int i = 0; i < N; ++i) qq[i] = MKF_QQ(i);
#pragma unroll
                    for (int i = 0; i < M * K; ++i) gg[i] = MKF_GG(i);
                    collapsed_info_consts<M, K>(phi, qq, gg, [&](int i, double v) {
                        if (i < kInfoSmem) sts_f64(ccs + i * 32, v);
                        else kreg[INFO ? i - kInfoSmem : 0] = v;
                    });
                    __syncwarp();
                }
                more = advance();
                while (more) {
                    // y(t) and y(t-1) straight from the ring; across a tile boundary y(t-1) is the
                    // register copy in yv (its stage is being refilled)
                    const double *yr = ys + (tt > 0 ? tt - 1 : 0) * M;
                    double yc[M], yq[M];
                    unsigned all_obs = 1u;
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        yc[i] = ys[tt * M + i];
                        all_obs &= is_observed(yc[i]) ? 1u : 0u;
                    }
#pragma unroll
                    for (int i = 0; i < M; ++i) {
                        const double prev = yr[i];
                        yq[i] = (tt > 0) ? prev : yv[i];
                    }
                    if (!__all_sync(0xffffffffu, all_obs != 0u)) {
                        if (kFreeze && frozen && !INFO) {
                            // S0 and QG were overwritten by the frozen gains: put them back
                            double qq[N], gg[M * K];
#pragma unroll
                            for (int i = 0; i < N; ++i) qq[i] = MKF_QQ(i);
#pragma unroll
                            for (int i = 0; i < M * K; ++i) gg[i] = MKF_GG(i);
                            collapsed_consts<M, K>(phi, qq, gg, [&](int i, double v) { sts_f64(ccs + i * 32, v); });
                            __syncwarp();
                        }
                        collapsed_expand<M, K>(x, P, yq, [&](int i) { return MKF_GG(i); });
#pragma unroll
                        for (int i = 0; i < M; ++i) yv[i] = yc[i];
                        break;   // the full-state loop handles this step
                    }
                    fold();
                    double sigma = 0.0, stepprod = 1.0;
                    if (kFreeze && frozen) {
                        collapsed_frozen_step<M, K>(x, phi, yq, yc, [&](int i) { return lds_f64(ccs + i * 32); }, sigma);
                        stepprod = frozen_prod;
                    } else {
                        double qc[K];
#pragma unroll
                        for (int j = 0; j < K; ++j) qc[j] = MKF_QQ(M + j);
                        if constexpr (INFO) {
                            // ordinary steps in information form; the one step that captures the steady-state
                            // constants (capture_next is warp-uniform) in the elimination form, whose D, S0, QG
                            // replace the information-form constants first
                            if (kFreeze && capture_next) {
                                double qq[N], gg[M * K];
#pragma unroll
                                for (int i = 0; i < N; ++i) qq[i] = MKF_QQ(i);
#pragma unroll
                                for (int i = 0; i < M * K; ++i) gg[i] = MKF_GG(i);
                                collapsed_consts<M, K>(phi, qq, gg, [&](int i, double v) { sts_f64(ccs + i * 32, v); });
                                __syncwarp();
                                collapsed_step_impl<M, K, true>(
                                    x, P, phi, qc, yq, yc, [&](int i) { return lds_f64(ccs + i * 32); },
                                    [&](int i, double v) { sts_f64(ccs + (M * K + i) * 32, v); }, sigma, stepprod, bad);
                            } else {
                                collapsed_info_step<M, K>(
                                    x, P, phi, yq, yc,
                                    [&](int i) { return i < kInfoSmem ? lds_f64(ccs + i * 32) : kreg[INFO ? i - kInfoSmem : 0]; },
                                    sigma, stepprod, bad);
                            }
                        } else {
#if MKF_CAPTURE_SPLIT
                        // the one step that stores its reciprocal pivots, multipliers and gains over S0 / QG
                        // (capture_next is warp-uniform) is a separate copy of the code: the 30 stores and
                        // their predicates stay out of the loop every other step runs
                        if (kFreeze && capture_next) {
                            collapsed_step_impl<M, K, true>(
                                x, P, phi, qc, yq, yc, [&](int i) { return lds_f64(ccs + i * 32); },
                                [&](int i, double v) { sts_f64(ccs + (M * K + i) * 32, v); }, sigma, stepprod, bad);
                        } else {
                            collapsed_step_impl<M, K, false>(
                                x, P, phi, qc, yq, yc, [&](int i) { return lds_f64(ccs + i * 32); },
                                [](int, double) {}, sigma, stepprod, bad);
                        }
#else
                        // one code path: when capture_next is set (warp-uniform) the step also stores
                        // its reciprocal pivots, multipliers and gains over S0 / QG (predicated stores)
                        collapsed_step_impl<M, K, kFreeze>(
                            x, P, phi, qc, yq, yc, [&](int i) { return lds_f64(ccs + i * 32); },
                            [&](int i, double v) {
                                if (capture_next) sts_f64(ccs + (M * K + i) * 32, v);
                            },
                            sigma, stepprod, bad);
#endif
                        }
                        if (kFreeze) {
                            if (lane_holds) {
#pragma unroll
                                for (int j = 0; j < K; ++j)
#pragma unroll
                                    for (int l = j; l < K; ++l) P[tri<N>(M + j, M + l)] = Chold[tri<K>(j, l)];
                            }
                            if (capture_next) {
                                __syncwarp();
                                frozen_prod = stepprod;
                                frozen = true;
                                handoff_now = true;
                                capture_next = false;
                            } else if (INFO ? (++since_chk == 6) : (MKF_FREEZE_EVERY_STEP || tt == 0)) {
                                since_chk = 0;
                                // once per tile: a weighted sum of the entries of C (weights without
                                // simple ratios, so that opposite drifts do not cancel) must not have
                                // moved by more than 2^-50 of itself over two consecutive tiles
                                if (!lane_holds) {
                                    const double cs = collapsed_checksum<M, K>(P);
                                    const bool same = fabs(cs - cchk) <= kFreezeTol * fabs(cs);
                                    cchk = cs;
                                    nconv = same ? nconv + 1 : 0;
                                    if (nconv >= MKF_FREEZE_NCONV) {
                                        lane_holds = true;
#pragma unroll
                                        for (int j = 0; j < K; ++j)
#pragma unroll
                                            for (int l = j; l < K; ++l) Chold[tri<K>(j, l)] = P[tri<N>(M + j, M + l)];
                                    }
                                }
                                capture_next = __all_sync(0xffffffffu, lane_holds);
                            }
                        }
                    }
                    account(M, sigma, stepprod);
                    if constexpr (HO_CAPABLE) {
                        // every lane holds C and this step captured the constants: the rest of the
                        // run (as long as it stays fully observed) is the steady-state kernel's
                        if (ho_mode == HO_MODE_HANDOFF && handoff_now && T - (t0 + tt + 1) >= kMinSteadySteps) {
                            double *wo = a.ho_ws + b;
                            const int64_t Bp = a.ho_pitch;
#pragma unroll
                            for (int j = 0; j < K; ++j) wo[(HO::C + j) * Bp] = x[M + j];
#pragma unroll
                            for (int i = 0; i < KT; ++i) wo[(HO::CHOLD + i) * Bp] = Chold[i];
#pragma unroll
                            for (int i = 0; i < HO::NCONST; ++i) wo[(HO::CONST + i) * Bp] = lds_f64(ccs + i * 32);
#pragma unroll
                            for (int i = 0; i < N; ++i) wo[(HO::PHI + i) * Bp] = phi[i];
                            wo[HO::SUM_S * Bp] = sum_sigma.s;
                            wo[HO::SUM_C * Bp] = sum_sigma.c;
                            wo[HO::FPROD * Bp] = fprod;
                            wo[HO::PEND_SIGMA * Bp] = pend_sigma;
                            wo[HO::PEND_PROD * Bp] = pend_prod;
                            wo[HO::FROZEN_PROD * Bp] = frozen_prod;
                            wo[HO::FEXP * Bp] = (double)fexp;
                            wo[HO::NOBS * Bp] = (double)nobs;
                            wo[HO::NSIG * Bp] = (double)nsig;
                            wo[HO::FLAGS * Bp] = (double)((bad ? 1 : 0) | (pend_on ? 2 : 0));
                            if (lane == 0) {
                                a.ho_tnext[wid] = t0 + tt + 1;
                                a.ho_state[wid] = HO_STEADY;
                            }
                            // the loads still in flight target this CTA's shared memory: wait for them
                            if (tma)
                                for (int64_t r2 = tile + 1; r2 < min(tile + (int64_t)kStages, ntiles); ++r2)
                                    mbar_wait(&bars[r2 % kStages], (unsigned)((r2 / kStages) & 1));
                            return;
                        }
                        handoff_now = false;
                    }
                    if (tt + 1 == steps) {
#pragma unroll
                        for (int i = 0; i < M; ++i) yv[i] = yc[i];
                    }
                    more = advance();
                }
            }
        }
    }
    if (!tma) cp_async_wait<0>();
    if constexpr (HIST) {
        // the staging boxes are this CTA's shared memory: the last tensor stores must have read them
        if (htm && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
#undef MKF_QQ
#undef MKF_GG

    sum_sigma.add(pend_on ? pend_sigma : 0.0);
    fprod *= pend_on ? pend_prod : 1.0;
    if (active) {
        double logdet = log(fprod) + (double)fexp * kLn2;
        double mle = (double)nobs * kLog2Pi + logdet + sum_sigma.value();
        if (bad) mle = nan("");
        a.mle[b] = mle;
        if (a.nobs) a.nobs[b] = (int64_t)nobs;
        if (a.status) a.status[b] = bad ? MKF_STATUS_NONPOSITIVE_F : 0;
    }
    if (HO_CAPABLE && ho_mode != HO_MODE_OFF && lane == 0) {
        a.ho_state[wid] = HO_DONE;
        if (ho_mode == HO_MODE_HANDOFF) a.ho_tnext[wid] = T;
    }
}

// Builds the tensor map of y viewed as [rows][row_elems] doubles (defined in filter_dfm.cu).
int make_y_tensor_map(CUtensorMap *map, const double *y, int64_t rows, int64_t row_elems, int64_t row_stride,
                      int box_rows, int box_elems, int l2_promotion_bytes = 128);
// Tensor map of a history array [B][T][lines * inner] doubles viewed as dims {inner, B, T * lines}, box
// {inner, 32, lines * steps}, swizzle span = inner * 8 bytes (32, 64 or 128).  Defined in filter_dfm.cu.
int make_hist_map(CUtensorMap *map, double *p, int inner, int64_t B, int64_t T, int lines, int steps);
// Tensor map of a per-step scalar array [B][T] (8-byte doubles or 4-byte integers) viewed as dims {T, B},
// box {16, 32}, swizzle span = the 16 steps of a row.
int make_scalar_map(CUtensorMap *map, void *p, int elem_bytes, int64_t B, int64_t T);
bool hist_tma_enabled();
// Grow-only device scratch of the hand-off records, one per (device, stream); NULL when the allocation
// fails (the caller then stays on the single-kernel path).  Defined in filter_dfm.cu.  The transient
// kernel leaves tnext = T for a warp it finished itself, so sum(T - tnext) counts the warp-steps done
// after the hand-off.
struct HandoffScratch {
    void *ptr = nullptr;
    size_t bytes = 0;
    // the most recent three-launch call on this stream (mkf_dfm_last_profile)
    int64_t blocks = 0, T = 0;
    int64_t *tnext = nullptr;
    cudaEvent_t ev[4] = {};
    bool have_events = false, profiled = false;
};
HandoffScratch *handoff_scratch(cudaStream_t stream, size_t bytes);
bool handoff_enabled();
bool profile_enabled();
bool profile_events(HandoffScratch *s);

// one translation unit per K instantiates M = 1..8 (keeps nvcc wall time parallel)
template <int K>
int launch_for_m(int m, const DfmArgs &a, cudaStream_t s);

constexpr int kMaxM = 8;
constexpr int kMaxK = 3;

}  // namespace mkf_dfm
