// Translation unit of the persistent per-branch HMC kernel (k1_tcp.cuh): instantiations + the cooperative launch.
#include "k1_tcp.cuh"

namespace bann {

template <int H, int S, int D, int ACT>
static int launch_tcp_one(const TcpArgs& a, uint32_t P, int num_sms, cudaStream_t st, bool* launched, float** part_io, bann_net* net) {
    using TS = TcpShape<H, S, D>;
    auto kern = k_hmc_persistent<H, S, D, ACT>;
    const size_t smem = TS::smem(a.ncb, P);
    static bool configured = false;
    static int per_sm = 0;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS::smem(8, 4096)));
        configured = true;
    }
    BANN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    const uint32_t maxcoop = (uint32_t)std::max(0, per_sm) * (uint32_t)num_sms;
    if (maxcoop == 0) return 0;
    const uint32_t tpc = (a.nst + maxcoop - 1) / maxcoop;
    if (tpc == 0 || tpc > (uint32_t)kTcpMaxTiles) return 0;          // too many rows to keep resident: the launch-per-step path runs
    const uint32_t grid = (a.nst + tpc - 1) / tpc;
    float* part = bann_net_partials(net, (size_t)grid * a.pstride);
    if (!part) return -2;
    *part_io = part;
    TcpArgs args = a;
    args.part = part;
    void* params[] = {&args};
    BANN_CUDA(cudaMemsetAsync(args.bar, 0, sizeof(unsigned int), st));
    BANN_CUDA(cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(128), params, smem, st));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    *launched = true;
    return 0;
}

// The whole HMC trajectory of one branch in one cooperative launch.  *launched stays false when the branch / net is not
// eligible (tensor-core store, <= 64 markers, 3 * W0 <= 16, an instantiated tanh architecture, few enough super-tiles to
// keep resident); the caller then runs the launch-per-step path.
int launch_hmc_persistent(const BranchDesc& d0, int act, TcpArgs& a, int num_sms, cudaStream_t st, bool* launched, bann_net* net) {
    *launched = false;
    if (act != BANN_TANH || !a.store_tc || d0.m > (uint32_t)kTcMaxMarkers || a.nst == 0) return 0;
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    a.ncb = (d0.m + 7) / 8;
    float* part = nullptr;
#define BANN_TRY_TCP(HH, SS, DD)                                                                                          \
    if (!*launched && H == HH && S == SS && D == DD) {                                                                   \
        int rc = launch_tcp_one<HH, SS, DD, BANN_TANH>(a, d0.P, num_sms, st, launched, &part, net);                      \
        if (rc) return rc;                                                                                               \
    }
    BANN_TRY_TCP(5, 5, 1)
    BANN_TRY_TCP(2, 2, 1)
    BANN_TRY_TCP(3, 3, 1)
    BANN_TRY_TCP(4, 3, 1)
    BANN_TRY_TCP(4, 4, 1)
#undef BANN_TRY_TCP
    return 0;
}

}  // namespace bann
