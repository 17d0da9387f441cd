// Host-side launch of the thread-per-model DFM kernels: dfm_loglik_kernel alone, or the
// three-launch sequence  transient (hand-off on) -> steady state -> resume  (see dfm_steady.cuh).
#pragma once
#include "dfm_steady.cuh"

namespace mkf_dfm {

template <int M, int K, bool HASR, bool HIST>
int launch_r(const DfmArgs &a0, cudaStream_t stream) {
    DfmArgs a = a0;
    alignas(64) CUtensorMap map;
    memset(&map, 0, sizeof(map));
    constexpr int TT = tile_steps(M, HIST, !HASR && !HIST && info_shape(M, K));
    const bool shared_y = a.y_stride == 0;
    if (a.use_tma) {
        int rc = make_y_tensor_map(&map, a.y, shared_y ? 1 : a.B, a.T * M, shared_y ? a.T * M : a.y_stride,
                                   shared_y ? 1 : 32, TT * M);
        if (rc != MKF_OK) a.use_tma = 0;  // cp.async staging instead
    }
    auto kern = dfm_loglik_kernel<M, K, HASR, HIST>;
    const size_t smem = smem_bytes(M, K, HIST, !HASR && !HIST);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        mkf_host::set_cuda_error(e);
        return MKF_ERR_CUDA;
    }
    const int64_t blocks = (a.B + kBlock - 1) / kBlock;
    a.ho_mode = HO_MODE_OFF;
    // history stores: tensor maps of the four state arrays (all requested arrays or none), whole-sector
    // stores of the step scalars when their bases allow it
    HistMaps<HIST> hmaps;
    a.hist_tma = 0;
    a.hist_vec = 0;
    if constexpr (HIST) {
        memset(&hmaps, 0, sizeof(hmaps));
        constexpr int N = M + K;
        auto al32 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; };
        a.hist_vec = al32(a.sigma) && al32(a.detf) && al32(a.cnt);
        if (hist_tma_shape(M, K) && hist_tma_enabled()) {
            int mask = 0;
            bool ok = true;
            if (a.pc) { ok = ok && make_hist_map(&hmaps.pc, a.pc, 16, a.B, a.T, N * N / 16, MKF_HIST_TSC) == MKF_OK; mask |= HIST_TMA_PC; }
            if (a.pm) { ok = ok && make_hist_map(&hmaps.pm, a.pm, N, a.B, a.T, 1, MKF_HIST_TSM) == MKF_OK; mask |= HIST_TMA_PM; }
            if (a.fc) { ok = ok && make_hist_map(&hmaps.fc, a.fc, 16, a.B, a.T, N * N / 16, MKF_HIST_TSC) == MKF_OK; mask |= HIST_TMA_FC; }
            if (a.fm) { ok = ok && make_hist_map(&hmaps.fm, a.fm, N, a.B, a.T, 1, MKF_HIST_TSM) == MKF_OK; mask |= HIST_TMA_FM; }
            a.hist_tma = ok ? mask : 0;
        }
        // the step scalars as {16 steps, 32 models} boxes: all requested ones or none (a mix would put
        // tensor stores and per-lane stores into the same staging area)
        if (hist_tma_enabled() && (a.sigma || a.detf || a.cnt)) {
            int mask = 0;
            bool ok = true;
            if (a.sigma) { ok = ok && make_scalar_map(&hmaps.sigma, a.sigma, 8, a.B, a.T) == MKF_OK; mask |= HIST_TMA_SIGMA; }
            if (a.detf) { ok = ok && make_scalar_map(&hmaps.detf, a.detf, 8, a.B, a.T) == MKF_OK; mask |= HIST_TMA_DETF; }
            if (a.cnt) { ok = ok && make_scalar_map(&hmaps.cnt, a.cnt, 4, a.B, a.T) == MKF_OK; mask |= HIST_TMA_CNT; }
            if (ok) a.hist_tma |= mask;
        }
    }

    // Hand-off path: R = 0, no history, a collapse-capable shape, TMA staging and a series long
    // enough for the transient to end well before T.
    constexpr bool HO_OK = !HASR && !HIST && collapse_shape(M, K) && kFreeze;
    if constexpr (HO_OK) {
        using HO = Handoff<M, K>;
        alignas(64) CUtensorMap smap;
        memset(&smap, 0, sizeof(smap));
        bool handoff = a.use_tma && handoff_enabled() && a.T >= 4 * kMinSteadySteps;
        if (handoff)
            handoff = make_y_tensor_map(&smap, a.y, shared_y ? 1 : a.B, a.T * M, shared_y ? a.T * M : a.y_stride,
                                        shared_y ? 1 : 32, steady_tile_steps(M) * M, 256) == MKF_OK;
        if (handoff) {
            const int64_t Bp = blocks * 32;
            const size_t rec_bytes = (size_t)HO::NF * Bp * sizeof(double);
            const size_t bytes = rec_bytes + (size_t)blocks * (sizeof(int64_t) + sizeof(int32_t)) + 64;
            HandoffScratch *hs = handoff_scratch(stream, bytes);
            if (hs) {
                unsigned char *ws = static_cast<unsigned char *>(hs->ptr);
                a.ho_ws = reinterpret_cast<double *>(ws);
                a.ho_pitch = Bp;
                a.ho_tnext = reinterpret_cast<int64_t *>(ws + rec_bytes);
                a.ho_state = reinterpret_cast<int32_t *>(ws + rec_bytes + (size_t)blocks * sizeof(int64_t));
                auto skern = dfm_steady_kernel<M, K>;
                const size_t ssmem = steady_smem_bytes(M);
                e = cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem);
                if (e != cudaSuccess) {
                    mkf_host::set_cuda_error(e);
                    return MKF_ERR_CUDA;
                }
                hs->blocks = blocks;
                hs->T = a.T;
                hs->tnext = a.ho_tnext;
                const bool prof = profile_enabled() && profile_events(hs);
                a.ho_mode = HO_MODE_HANDOFF;   // every warp writes its ho_state: no memset needed
                if (prof) cudaEventRecord(hs->ev[0], stream);
                kern<<<(unsigned)blocks, kBlock, smem, stream>>>(a, map, hmaps);
                if (prof) cudaEventRecord(hs->ev[1], stream);
                skern<<<(unsigned)blocks, 32, ssmem, stream>>>(a, smap);
                if (prof) cudaEventRecord(hs->ev[2], stream);
                a.ho_mode = HO_MODE_RESUME;
                kern<<<(unsigned)blocks, kBlock, smem, stream>>>(a, map, hmaps);
                if (prof) cudaEventRecord(hs->ev[3], stream);
                hs->profiled = prof;
                return mkf_host::check_launch();
            }
        }
    }
    kern<<<(unsigned)blocks, kBlock, smem, stream>>>(a, map, hmaps);
    return mkf_host::check_launch();
}

template <int M, int K>
int launch(const DfmArgs &a, cudaStream_t stream) {
    const bool hist = a.sigma || a.detf || a.cnt || a.fm || a.fc || a.pm || a.pc;
    if (hist) return launch_r<M, K, true, true>(a, stream);
    if (a.tidx) return launch_r<M, K, true, false>(a, stream);   // compacted grid: the build without factor-only steps
    return a.r ? launch_r<M, K, true, false>(a, stream) : launch_r<M, K, false, false>(a, stream);
}

}  // namespace mkf_dfm
