// Tuned K1 for narrow branches (placeholder: the generic kernel is used until this lands).
#pragma once
#include "kernels.cuh"

struct bann_net;
float* bann_net_partials(bann_net* net, size_t need);
float* bann_net_gsum(bann_net* net);
uint32_t bann_net_pstride(bann_net* net);

namespace bann {
inline int launch_k1_small(const std::vector<BranchDesc>&, int, K1Args&, uint32_t, int, cudaStream_t, bool* launched,
                           uint32_t*, float**, bann_net*) {
    *launched = false;
    return 0;
}
}  // namespace bann
