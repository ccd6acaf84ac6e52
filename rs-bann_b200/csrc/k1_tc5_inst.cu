// Translation unit of the five-warp tensor-core K1 (k1_tc5.cuh): instantiations + launch.
#include "k1_tc5.cuh"

namespace bann {

template <int H, int S, int D, int ACT>
static int launch_tc5(K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st) {
    using C = TcShape<H, S, D>;
    constexpr bool kNct7 = ACT == BANN_TANH;      // the 49..56-marker specialisation exists for the benchmarked activation only
    static bool configured = false;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(k1_tc5<H, S, D, ACT, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        if (kNct7) BANN_CUDA(cudaFuncSetAttribute(k1_tc5<H, S, D, ACT, true, kNct7 ? 7 : 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem(8)));
        configured = true;
    }
    dim3 grid(a.nchunk, nlist);
    if (kNct7 && a.nc_uniform == 7) BANN_CUDA(launch_pdl(k1_tc5<H, S, D, ACT, true, kNct7 ? 7 : 0>, grid, dim3(kTc5Threads), smem, st, a));
    else BANN_CUDA(launch_pdl(k1_tc5<H, S, D, ACT, true, 0>, grid, dim3(kTc5Threads), smem, st, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

template <int H, int S, int D>
static int launch_tc5_act(int act, K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st) {
    switch (act) {
        case BANN_TANH: return launch_tc5<H, S, D, BANN_TANH>(a, nlist, smem, st);
        case BANN_RELU: return launch_tc5<H, S, D, BANN_RELU>(a, nlist, smem, st);
        case BANN_LEAKY_RELU: return launch_tc5<H, S, D, BANN_LEAKY_RELU>(a, nlist, smem, st);
        case BANN_SILU: return launch_tc5<H, S, D, BANN_SILU>(a, nlist, smem, st);
        default: return launch_tc5<H, S, D, BANN_IDENTITY>(a, nlist, smem, st);
    }
}

int launch_one_tc5(int H, int S, int D, int act, K1Args& a, uint32_t nlist, size_t smem, cudaStream_t st, bool* launched) {
    *launched = false;
#define BANN_TRY_TC5(HH, SS, DD)                                   \
    if (!*launched && H == HH && S == SS && D == DD) {             \
        int rc = launch_tc5_act<HH, SS, DD>(act, a, nlist, smem, st); \
        if (rc) return rc;                                         \
        *launched = true;                                          \
    }
    BANN_TRY_TC5(5, 5, 1)
#undef BANN_TRY_TC5
    return 0;
}

}  // namespace bann
