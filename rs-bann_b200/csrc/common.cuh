// Shared declarations of libbann_b200: device-side descriptors, error plumbing, Philox RNG.
// Layout contracts follow the reference (see include/bann.h for the citations).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>
#include <utility>
#include <vector>

#include "../../include/bann.h"

namespace bann {

constexpr int kMaxLayers = BANN_MAX_LAYERS;
constexpr int kTileRows = 128;   // individuals per row tile (32 PLINK bytes per marker)
constexpr int kTileQuads = 32;   // 4 individuals per byte

// status codes of a branch inside an HMC transition (device side)
enum : int { ST_IDLE = -1, ST_RUNNING = 3, ST_REJECTED_EARLY = BANN_HMC_REJECTED_EARLY,
             ST_REJECTED = BANN_HMC_REJECTED, ST_ACCEPTED = BANN_HMC_ACCEPTED };

// One branch: architecture + offsets into the arenas.  param_vec order (params.rs:700-715).
struct BranchDesc {
    uint32_t m;                 // markers in the branch
    uint32_t m_pad4;            // bytes per row-quad in a tile: m rounded up to whole 32-bit words, odd word count
    uint32_t nl;                // layers incl. output
    uint32_t P;                 // parameters
    uint32_t nprec;             // precisions (weight precs, bias precs, error prec)
    uint32_t sumw;              // sum of widths of the activated layers (l < nl-1)
    uint32_t widths[kMaxLayers];
    uint32_t in_dim[kMaxLayers];
    uint32_t w_off[kMaxLayers]; // offset of W_l inside the branch's param vec
    uint32_t b_off[kMaxLayers]; // offset of b_l (l < nl-1)
    uint32_t a_off[kMaxLayers]; // offset of layer l inside a per-row activation record
    uint32_t wp_off[kMaxLayers];// offset of the weight precision(s) of layer l in the precision vec
    uint32_t wp_len[kMaxLayers];// in_l for ARD layers l<last, else 1
    uint32_t bp_off[kMaxLayers];
    uint32_t ep_off;            // error precision
    uint64_t param_off;         // offset (floats) into theta / mom / grad / eps arenas
    uint64_t prec_off;          // offset (floats) into the precision arena
    uint64_t tile_off;          // byte offset of the branch's first tile in the store
    uint64_t col_off;           // offset into the per-branch gathered mu / sd arrays
    uint64_t tc_off;            // byte offset of the branch in the tensor-core store (k1_tc.cuh)
    uint32_t nc;                // 8-marker chunks per row in the tensor-core store: ceil(m / 8)
    uint32_t dense_off;         // offset (floats) of the branch in a DENSE concatenation of param vecs (host-facing layout)
};

// GlobalParams + OutputBias + LPD + TrainingStats, device resident (net/params.rs:13-56,
// net/net.rs:29-36, net/log_posterior_density.rs:9-25, net/train_stats.rs:23-32).
struct NetGlobals {
    float error_precision;
    float output_layer_precision;
    float ow_reg_sum;      // over ALL branches
    float ow_num_params;
    float output_bias;
    float lpd_rss;
    float lpd_out_w;
    float resid_ss;        // sum r^2 of the current residual (all ranks)
    float resid_sum;
    unsigned long long num_samples, num_accepted, num_early_rejected;
    unsigned long long visit_counter;
};

// Per-branch HMC state.
struct BranchState {
    int   status;
    int   steps_done;
    int   u_turn_step;
    float neg_h_init;
    float neg_h_cur;
    float log_density;
    float rss;
    float log_acc;
};

struct Hyper6 { float v[6]; };  // dense(shape,scale) summary(shape,scale) output(shape,scale)

__host__ __device__ inline void layer_prior(const Hyper6& h, int l, int nl, float& shape, float& scale) {
    // params.rs:146-163
    int k = (l == nl - 1) ? 2 : (l == nl - 2 ? 1 : 0);
    shape = h.v[2 * k];
    scale = h.v[2 * k + 1];
}

// ---------------------------------------------------------------- error plumbing
void set_error(const std::string& s);
#define BANN_FAIL(msg)                                                         \
    do {                                                                       \
        ::bann::set_error(std::string(msg) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
        return -1;                                                             \
    } while (0)
#define BANN_CUDA(expr)                                                        \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) {                                               \
            ::bann::set_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " in " #expr " (" + \
                              __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
            return -2;                                                         \
        }                                                                      \
    } while (0)
#define BANN_CHECK(expr)                                                       \
    do {                                                                       \
        int _r = (expr);                                                       \
        if (_r != 0) return _r;                                                \
    } while (0)

extern unsigned long long g_launch_count;
#define BANN_LAUNCHED() (++::bann::g_launch_count)

// ---------------------------------------------------------------- programmatic dependent launch (sm_90+)
// The kernels on the critical path of a leapfrog step of the sequential schedule (K1 -> KR -> K2 -> K1 ...) are launched with
// programmatic stream serialisation: a kernel may be scheduled -- and run the part of its prologue that touches no global data
// of its predecessors -- while the previous kernel is still running; pdl_wait() returns once the previous kernel has completed
// and its writes are visible.  EVERY path of such a kernel executes pdl_wait() before it exits or touches dependent data, so that
// "kernel N complete" implies "kernel N - 1 complete".  Launched the ordinary way both calls are no-ops.
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#if defined(__CUDACC__)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
#endif

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
    uint32_t key[2];
    uint32_t ctr[4];
    __host__ __device__ Philox(uint64_t seed, uint64_t stream, uint64_t offset) {
        key[0] = (uint32_t)seed;
        key[1] = (uint32_t)(seed >> 32);
        ctr[0] = (uint32_t)offset;
        ctr[1] = (uint32_t)(offset >> 32);
        ctr[2] = (uint32_t)stream;
        ctr[3] = (uint32_t)(stream >> 32);
    }
    __host__ __device__ static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
        uint64_t p = (uint64_t)a * b;
        hi = (uint32_t)(p >> 32);
        lo = (uint32_t)p;
    }
    __host__ __device__ inline void next(uint32_t out[4]) {
        uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
        uint32_t k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0, lo0, hi1, lo1;
            mulhilo(0xD2511F53u, c0, hi0, lo0);
            mulhilo(0xCD9E8D57u, c2, hi1, lo1);
            uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
        if (++ctr[0] == 0) ++ctr[1];
    }
};

__host__ __device__ inline float u01_open(uint32_t x) {  // (0,1]
    return ((x >> 8) + 1u) * (1.0f / 16777216.0f);
}
__host__ __device__ inline float u01_half_open(uint32_t x) {  // [0,1)
    return (x >> 8) * (1.0f / 16777216.0f);
}

}  // namespace bann
