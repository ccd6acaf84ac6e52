// Translation unit of the five-warp tensor-core K1 (k1_tc5.cuh): instantiations + launch.
#include "k1_tc5.cuh"

namespace bann {

template <int H, int S, int D, int ACT, bool DEFER>
static int launch_tc5(K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st) {
    using C = TcShape<H, S, D>;
    constexpr bool kNct7 = ACT == BANN_TANH;      // the 49..56-marker specialisation exists for the benchmarked activation only
    static bool configured = false;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(k1_tc5<H, S, D, ACT, true, 0, DEFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        if (kNct7) BANN_CUDA(cudaFuncSetAttribute(k1_tc5<H, S, D, ACT, true, kNct7 ? 7 : 0, DEFER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        configured = true;
    }
    dim3 grid(a.nchunk, nlist);
    if (kNct7 && a.nc_uniform == 7) BANN_CUDA(launch_pdl(k1_tc5<H, S, D, ACT, true, kNct7 ? 7 : 0, DEFER>, grid, dim3(kTc5Threads), smem, st, a));
    else BANN_CUDA(launch_pdl(k1_tc5<H, S, D, ACT, true, 0, DEFER>, grid, dim3(kTc5Threads), smem, st, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

template <int H, int S, int D, bool DEFER>
static int launch_tc5_defer(int act, K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st) {
    switch (act) {
        case BANN_TANH: return launch_tc5<H, S, D, BANN_TANH, DEFER>(a, nlist, smem, st);
        case BANN_RELU: return launch_tc5<H, S, D, BANN_RELU, DEFER>(a, nlist, smem, st);
        case BANN_LEAKY_RELU: return launch_tc5<H, S, D, BANN_LEAKY_RELU, DEFER>(a, nlist, smem, st);
        case BANN_SILU: return launch_tc5<H, S, D, BANN_SILU, DEFER>(a, nlist, smem, st);
        default: return launch_tc5<H, S, D, BANN_IDENTITY, DEFER>(a, nlist, smem, st);
    }
}
// DPLAIN: also instantiate the variant with the sums inside part 1 (BANN_TC_FIVE_WARPS_PLAIN; A/B of the benchmarked architecture)
template <int H, int S, int D, bool DPLAIN>
static int launch_tc5_act(int act, K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st) {
    if constexpr (DPLAIN) {
        if (a.tc_variant == BANN_TC_FIVE_WARPS_PLAIN) return launch_tc5_defer<H, S, D, false>(act, a, nlist, smem, st);
    }
    return launch_tc5_defer<H, S, D, true>(act, a, nlist, smem, st);
}

int launch_one_tc5(int H, int S, int D, int act, K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st, bool* launched) {
    *launched = false;
#define BANN_TRY_TC5(HH, SS, DD, PLAIN)                                   \
    if (!*launched && H == HH && S == SS && D == DD) {                    \
        int rc = launch_tc5_act<HH, SS, DD, PLAIN>(act, a, nlist, smem, st); \
        if (rc) return rc;                                                \
        *launched = true;                                                 \
    }
    // the architectures with at most one hidden layer (two hidden layers need more accumulators than 128 registers hold)
    BANN_TRY_TC5(5, 5, 1, true)
    BANN_TRY_TC5(2, 2, 1, false)
    BANN_TRY_TC5(4, 3, 1, false)
    BANN_TRY_TC5(3, 3, 1, false)
    BANN_TRY_TC5(4, 4, 1, false)
    BANN_TRY_TC5(2, 2, 0, false)
    BANN_TRY_TC5(5, 5, 0, false)
#undef BANN_TRY_TC5
    return 0;
}

}  // namespace bann
