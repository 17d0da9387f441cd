// Instantiates the thread-per-model DFM kernels for K = 0 common factors, M = 1..8 series.
#include "dfm_launch.cuh"

namespace mkf_dfm {
template <>
int launch_for_m<0>(int m, const DfmArgs &a, cudaStream_t s) {
    switch (m) {
#ifdef MKF_DFM_ONLY_M5
        case 5: return launch<5, 0>(a, s);
#else
        case 1: return launch<1, 0>(a, s);
        case 2: return launch<2, 0>(a, s);
        case 3: return launch<3, 0>(a, s);
        case 4: return launch<4, 0>(a, s);
        case 5: return launch<5, 0>(a, s);
        case 6: return launch<6, 0>(a, s);
        case 7: return launch<7, 0>(a, s);
        case 8: return launch<8, 0>(a, s);
#endif
        default: return MKF_ERR_UNSUPPORTED;
    }
}
}  // namespace mkf_dfm
