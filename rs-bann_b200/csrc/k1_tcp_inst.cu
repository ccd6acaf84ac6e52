// Translation unit of the persistent per-branch HMC kernel (k1_tcp.cuh): instantiations + the cooperative launch.
#include <cstdio>
#include <cstdlib>

#include "k1_tcp.cuh"

namespace bann {

template <int H, int S, int D, int ACT>
static int launch_tcp_one(const TcpArgs& a, uint32_t P, int num_sms, cudaStream_t st, bool* launched) {
    using TS = TcpShape<H, S, D>;
    auto kern = k_hmc_persistent<H, S, D, ACT>;
    static const bool dbg = getenv("BANN_DEBUG_TCP") != nullptr;          // debugging aid: per-phase clocks of one CTA, read once
    // One CTA per SM is what a cooperative launch of a tensor-memory kernel is granted (k1_tcp.cuh): as few super-tiles per CTA
    // as make the grid fit the SM count; more than kTcpMaxTiles would not stay resident -> launch-per-step path.
    cudaError_t err = cudaErrorCooperativeLaunchTooLarge;
    for (uint32_t tpc = 2; tpc <= (uint32_t)kTcpMaxTiles && err == cudaErrorCooperativeLaunchTooLarge; tpc += 2) {   // one / two tiles per warpgroup
        const uint32_t grid = (a.nst + tpc - 1) / tpc;
        if (grid > (uint32_t)num_sms) continue;
        const size_t smem = std::max<size_t>(TS::smem(a.ncb, P, tpc), 120 * 1024);   // one CTA per SM, whatever the driver would allow
        BANN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        TcpArgs args = a;
        args.tpc = tpc;
        static unsigned long long* d_timing = nullptr;
        if (dbg) {
            if (!d_timing) { cudaMalloc(&d_timing, 8 * sizeof(unsigned long long)); }
            cudaMemsetAsync(d_timing, 0, 8 * sizeof(unsigned long long), st);
            args.timing = d_timing;
            if (const char* e = getenv("BANN_DEBUG_TCP_CTA")) args.timing_cta = (uint32_t)atoi(e) % grid;
        }
        void* params[] = {&args};
        err = cudaLaunchCooperativeKernel((void*)kern, dim3(grid), dim3(kTcpThreads), params, smem, st);
        if (dbg) {
            int per_sm = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTcpThreads, smem);
            fprintf(stderr, "[tcp] cooperative launch grid %u (tiles per CTA %u, %zu B shared memory, occupancy %d): %s\n", grid, tpc, smem,
                    per_sm, cudaGetErrorString(err));
        }
        if (dbg && err == cudaSuccess) {
            unsigned long long h[8];
            cudaStreamSynchronize(st);
            cudaMemcpy(h, d_timing, sizeof(h), cudaMemcpyDeviceToHost);
            const double per = 1.0 / ((a.L + 1) * 1.965e3);      // us per evaluation at 1965 MHz
            fprintf(stderr, "[tcp] us per evaluation (one CTA): staging %.2f  forward MMAs %.2f  tails + backward MMAs %.2f  partial sums %.2f  "
                            "slice reduction (awaiting partials) %.2f  gather (awaiting sums) %.2f  update %.2f\n",
                    h[0] * per, h[1] * per, h[2] * per, h[3] * per, h[5] * per, h[6] * per, h[7] * per);
        }
        if (err == cudaErrorCooperativeLaunchTooLarge) cudaGetLastError();      // clear, try more tiles per CTA
    }
    if (err == cudaErrorCooperativeLaunchTooLarge) return 0;
    BANN_CUDA(err);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    *launched = true;
    return 0;
}

// The whole HMC trajectory of one branch in one cooperative launch.  *launched stays false when the branch / net is not
// eligible (tensor-core store, <= 64 markers, 3 * W0 <= 16, an instantiated architecture, few enough super-tiles to
// keep resident); the caller then runs the launch-per-step path.
int launch_hmc_persistent(const BranchDesc& d0, int act, TcpArgs& a, int num_sms, cudaStream_t st, bool* launched) {
    *launched = false;
    if (!a.store_tc || d0.m > (uint32_t)kTcMaxMarkers || a.nst == 0) return 0;
    if (num_sms > kTcpMaxGrid || d0.P + 1 > (uint32_t)kTcpMaxValues) return 0;
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    a.ncb = (d0.m + 7) / 8;
#define BANN_TRY_TCP(HH, SS, DD)                                                                                          \
    if (!*launched && H == HH && S == SS && D == DD) {                                                                   \
        int rc;                                                                                                          \
        switch (act) {                                                                                                   \
            case BANN_TANH: rc = launch_tcp_one<HH, SS, DD, BANN_TANH>(a, d0.P, num_sms, st, launched); break;           \
            case BANN_RELU: rc = launch_tcp_one<HH, SS, DD, BANN_RELU>(a, d0.P, num_sms, st, launched); break;           \
            case BANN_LEAKY_RELU: rc = launch_tcp_one<HH, SS, DD, BANN_LEAKY_RELU>(a, d0.P, num_sms, st, launched); break; \
            case BANN_SILU: rc = launch_tcp_one<HH, SS, DD, BANN_SILU>(a, d0.P, num_sms, st, launched); break;           \
            default: rc = launch_tcp_one<HH, SS, DD, BANN_IDENTITY>(a, d0.P, num_sms, st, launched); break;              \
        }                                                                                                                \
        if (rc) return rc;                                                                                               \
    }
    BANN_TRY_TCP(5, 5, 1)
    BANN_TRY_TCP(2, 2, 1)
    BANN_TRY_TCP(3, 3, 1)
    BANN_TRY_TCP(4, 3, 1)
    BANN_TRY_TCP(4, 4, 1)
    BANN_TRY_TCP(4, 3, 2)
    BANN_TRY_TCP(5, 3, 2)
    BANN_TRY_TCP(2, 2, 0)
    BANN_TRY_TCP(5, 5, 2)
#undef BANN_TRY_TCP
    return 0;
}

}  // namespace bann
