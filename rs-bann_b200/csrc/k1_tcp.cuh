// Persistent per-branch HMC transition for the sequential-exact schedule (VERDICT r1 "next" #3).
//
// `bann_visit_branch` runs L leapfrog steps of ONE branch; with separate launches every step is K1 (fused forward + backward
// over ~300 CTAs) -> KR (chunk reduction) -> K2 (update), three dependent launches, and every K1 launch re-does what does not
// change during a trajectory: tensor-memory allocation, barrier set-up, the bulk copy of the branch's packed genotypes from
// HBM and their expansion into tcgen05 operands (84 instructions per thread and super-tile).  19.2 us per leapfrog at
// N = 100k, against 0.25 us of roofline time.
//
// Here the whole transition is ONE cooperative launch.  Every CTA (two warpgroups, each running the FP32 tails of its own
// tiles) owns up to four 256-row super-tiles of the branch for the whole trajectory: the packed words are loaded and expanded ONCE -- the forward A operand stays in tensor memory, the
// backward A operand in shared memory -- and the targets stay in registers.  Per leapfrog step a CTA stages the three bf16 pieces
// of W' = W0 / sd from ITS OWN copy of the parameters, issues the forward MMAs, runs the FP32 tail (TcTail, k1_tc.cuh), issues
// the backward MMAs, publishes its partial sums; every CTA reduces a slice of the P + 1 values over all partials in a fixed
// order and publishes the sums; every CTA applies the SAME parameter / momentum update (gradient under the prior, half steps,
// position step, Hamiltonian, early-reject and U-turn checks: k2_step's arithmetic) to its own copy -- replicated,
// deterministic, nothing is broadcast.  CTA 0 writes the branch state and the final parameters.
//
// There is no grid barrier.  Every published value travels as ONE 64-bit word {float bits, tag}, tag = launch base + evaluation
// index + 1 (st / ld.relaxed b64: single-copy atomic), in a buffer of four slots (parity of the launch count x parity of the
// evaluation); a reader spins on the words it needs until their tag is this evaluation's.  A slot of evaluation e is
// overwritten at e + 2, which a CTA reaches only after it has read every sum of e + 1, i.e. after every reducer has finished
// reading the partials of e + 1 (and of e before that) -- so no reader can still need the old word; the launch parity keeps a
// rank that already runs the next transition from touching what a slower peer still reads of this one.  The all-reduce costs
// two L2 round trips per evaluation instead of two grid barriers plus two read passes (profiles/r2_seq_rate.log; sleeping
// between polls only costs time).
// With the rows sharded over GPUs the same kernel runs on every rank: the reducers push the rank-level sums as tagged words
// into every peer's table over NVLink, the gather adds the world entries in rank order -- compute and collective in one kernel,
// no host, no separate exchange launch; every rank takes the same decisions because every rank adds the same numbers in
// the same order.
// Spinning relies on the co-residency a cooperative launch guarantees; a clock limit turns a lost CTA or rank into an error
// flag instead of a hang.
#pragma once
#include <cooperative_groups.h>

#include "k1_tc.cuh"

namespace bann {

// super-tiles a CTA keeps resident.  The co-residency a cooperative launch is granted for a kernel that allocates tensor
// memory is ONE CTA per SM on this driver (the runtime cannot know how many columns tcgen05.alloc will ask for; ncu reports
// a theoretical occupancy of 3 for the same kernel) -- so the grid is at most the SM count and a CTA takes up to four super-tiles:
// 148 x 4 x 256 = 151k rows per GPU.
constexpr int kTcpMaxTiles = 4;
constexpr long long kTcpTimeoutClk = 6ll * 1000ll * 1000ll * 1000ll;      // ~3 s of SM clock (clock64: reading %globaltimer inside the poll loop costs more than the poll)

struct TcpArgs {
    const uint32_t* store_tc;
    const BranchDesc* descs;
    uint32_t b;
    const float* mu;
    const float* sd;
    uint32_t n, nst, ncb;
    uint32_t tpc;              // super-tiles per CTA (1 .. kTcpMaxTiles): sizes the resident operand images and the tensor-memory allocation
    uint32_t L;                // leapfrog steps
    // HMC state (arenas, as K2Args)
    BranchState* state;
    float* theta;
    const float* theta0;
    float* mom;
    float* grad;
    const float* eps;
    const float* prec;
    int model;
    float max_h_err;
    // targets: t = resid + own prediction at the first evaluation (net.rs:279-280) when resid != NULL, else the shared target tgt
    const float* resid;
    const float* tgt;
    float* tgt_out;            // t (resid mode), may be NULL
    float* prev_out;           // own prediction at the first evaluation, may be NULL
    float* ynew_out;           // own prediction at the last evaluation, may be NULL
    // scratch
    uint2* part;               // [4][gridDim.x][pstride] tagged partial sums {float bits, tag}
    uint2* sums;               // [4][kTcpSumCopies][pstride] tagged reduced sums (copies: a CTA polls copy blockIdx.x % kTcpSumCopies)
    uint32_t launch_par;       // parity of the launch count (slot = 2 launch_par + evaluation parity)
    uint32_t tag_base;         // tags of this launch are tag_base + 1 .. tag_base + L + 1; every older tag in the buffers is smaller
    // rows sharded over ranks (world > 1): the rank-level sums go to EVERY rank's [4][8][pstride] table over NVLink peer memory
    // (slot [parity][source rank]), each rank adds the world entries in rank order -- bit-identical everywhere
    uint32_t world, rank;
    uint2* rank_sums[8];
    float* gsum;               // [pstride]: the reduced sums of the last evaluation (rss at [P]), written by CTA 0 at the end
    uint32_t pstride;
    int* error_flag;
    unsigned long long* timing;   // NULL, or 8 phase accumulators (clock64 ticks of CTA timing_cta, BANN_DEBUG_TCP)
    uint32_t timing_cta;
};

template <int H, int S, int D>
struct TcpShape {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    // [expanded genotypes: tpc images][weight pieces][packed words x 2][delta pieces x 2][tail params + b0p][reduction scratch][barriers][theta, mom, eps, theta0,
    //  grad (maxP each)][reduced sums][prior precisions][this CTA's partial sums]
    static size_t smem(uint32_t ncb, uint32_t P, uint32_t tpc) {
        const size_t img = (size_t)ncb * kTcChunkStride;
        const size_t misc = (size_t)(2 * ((T::n_tail() + 3) & ~3) + 2 * T::W0P + ((8 * T::NTACC + 3) & ~3) + 16) * 4 + 64;
        size_t used = (size_t)tpc * img + C::SW + (size_t)2 * ncb * 512 + 2 * C::SD + misc + (size_t)(8 * ((P + 4) & ~3u)) * 4 + 512;
        const size_t need = (size_t)(ncb + 8) * kTcChunkStride + (tpc - 1) * img;     // the M = 64 backward operand reads 8 chunks of the LAST image
        return (used > need ? used : need) + 256;
    }
};

constexpr int kTcpSumCopies = 8;          // every sum is published kTcpSumCopies times so that <= ~19 CTAs spin on one line
constexpr int kTcpMaxGrid = 160;          // lanes of a reducer warp take <= 5 partials each
constexpr int kTcpMaxValues = 512;        // a thread gathers <= 2 reduced sums

__device__ __forceinline__ void tg_store(uint2* p, float v, uint32_t tag) {
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 tg_load(const uint2* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}
// the same across GPUs: the word is written by a peer over NVLink into this GPU's memory
__device__ __forceinline__ void tg_store_sys(uint2* p, float v, uint32_t tag) {
    const unsigned long long w = (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ uint2 tg_load_sys(const uint2* p) {
    unsigned long long w;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return make_uint2((uint32_t)w, (uint32_t)(w >> 32));
}
// NV words p[i * stride] (i < NV, i-th word wanted iff bit i of `want`), each awaited until its tag is `tag`.  All loads are in
// flight together; only late words are re-read.  false: timed out (error flag 3)
template <int NV, bool SYS = false>
__device__ __forceinline__ bool tg_await(const uint2* p, size_t stride, uint32_t want, uint32_t tag, uint2 (&v)[NV], int* error_flag) {
    uint32_t pending = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        v[i] = make_uint2(0u, tag);
        if (want >> i & 1u) {
            v[i] = SYS ? tg_load_sys(p + i * stride) : tg_load(p + i * stride);
            if (v[i].y != tag) pending |= 1u << i;
        }
    }
    if (pending) {
        const long long t0 = clock64();
        while (pending) {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (pending >> i & 1u) {
                    v[i] = SYS ? tg_load_sys(p + i * stride) : tg_load(p + i * stride);
                    if (v[i].y == tag) pending &= ~(1u << i);
                }
            if (pending && clock64() - t0 > kTcpTimeoutClk) {
                atomicExch(error_flag, 3);
                return false;
            }
        }
    }
    return true;
}

constexpr int kTcpThreads = 256;          // two warpgroups: group g owns the CTA's tiles g, g + 2 and runs their tails concurrently with the other's

__device__ __forceinline__ void group_sync(uint32_t grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1u) : "memory"); }

template <int H, int S, int D, int ACT>
__global__ void __launch_bounds__(kTcpThreads, 1) k_hmc_persistent(TcpArgs a) {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    using TT = TcTail<H, S, D, ACT>;
    constexpr int W0 = T::W0, W0P = T::W0P, NN = C::NN, NT = kTcpThreads, NTACC = T::NTACC;
    constexpr int kGroupTiles = kTcpMaxTiles / 2;
    constexpr float cA = TT::cA;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t grp = warp >> 2, wq = warp & 3u, gt = tid & 127u;       // warpgroup, warp within it (= its tensor-memory lane quarter), thread within it
    const uint32_t cta = blockIdx.x, ncta = gridDim.x;
    const BranchDesc& d = a.descs[a.b];
    const uint32_t m = d.m, NC = d.nc, NKS = (NC + 1) >> 1, NCB = a.ncb, P = d.P;
    const uint32_t Pp = (P + 4) & ~3u;
    // ---- shared memory carve-up
    uint8_t* sA = smraw + ((128u - (umma::smem_u32(smraw) & 127u)) & 127u);
    const uint32_t sa_bytes = NCB * kTcChunkStride;
    uint8_t* sW = sA + (size_t)a.tpc * sa_bytes;
    uint32_t* sG = reinterpret_cast<uint32_t*>(sW + C::SW);                  // packed words being expanded: one buffer per group
    uint8_t* sD = reinterpret_cast<uint8_t*>(sG) + (size_t)2 * NCB * 512;    // delta pieces: one buffer per group
    float* wp = reinterpret_cast<float*>(__builtin_assume_aligned(sD + 2 * C::SD, 16));
    float* b0p = wp + ((T::n_tail() + 3) & ~3);
    float* red = b0p + ((T::n_tail() + 3) & ~3) + 2 * W0P;                   // [8 warps][NTACC]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + ((8 * NTACC + 3) & ~3));   // [0] forward MMAs, [1 + g] backward MMAs of group g, [4 + g] its words landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 6);
    float* s_th = reinterpret_cast<float*>(tmem_slot + 4);           // this CTA's copy of the branch's parameters ...
    float* s_p = s_th + Pp;                                          // ... momenta
    float* s_eps = s_p + Pp;
    float* s_th0 = s_eps + Pp;
    float* s_g = s_th0 + Pp;                                         // gradient under the prior
    float* s_sum = s_g + Pp;                                         // reduced raw sums [P + 1]
    float* s_lam = s_sum + Pp;                                       // prior precision of each parameter (constant over the trajectory); < 0: a bias
    float* s_part = s_lam + Pp;                                      // this CTA's partial sums [P + 1] before they are published

    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;
    const float* pr = a.prec + d.prec_off;
    const uint32_t tpc = a.tpc;                                      // 1 .. kTcpMaxTiles (host)
    const uint32_t t_begin = min(a.nst, cta * tpc);
    const uint32_t t_end = min(a.nst, t_begin + tpc);
    const uint32_t ntile = t_end - t_begin;                          // 0 .. kTcpMaxTiles (the host guarantees the bound)
    const uint32_t gtiles = ntile > grp ? (ntile - grp + 1) / 2 : 0; // tiles of this group: local k = 2 kk + grp

    // ---- one-time setup
    {
        const uint32_t nz = (uint32_t)((sD + 2 * C::SD - sA) / 16);
        for (uint32_t k = tid; k < nz; k += NT) reinterpret_cast<uint4*>(sA)[k] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 1);
        umma::mbar_init(&mbar[1], 1);
        umma::mbar_init(&mbar[2], 1);
        umma::mbar_init(&mbar[4], 1);
        umma::mbar_init(&mbar[5], 1);
        umma::fence_mbar_init();
    }
    // per tile: 2 x NN forward accumulators + 64 columns forward A operand; + 2 x 32 for the groups' backward accumulators
    const uint32_t kTmemCols = tpc <= 2 ? 256u : 512u;
    if (warp == 0) umma::tmem_alloc(tmem_slot, kTmemCols);
    for (uint32_t k = tid; k < P; k += NT) {
        s_th[k] = a.theta[d.param_off + k];
        s_p[k] = a.mom[d.param_off + k];
        s_eps[k] = a.eps[d.param_off + k];
        s_th0[k] = a.theta0[d.param_off + k];
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        s_lam[k] = isb ? -1.f : param_prior_precision(d, pr, a.model, l, row, false);
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((wq * 32u) << 16);
    // tensor-memory columns: tile k: forward accumulators [96 k, 96 k + 32), forward A operand [96 k + 32, 96 k + 96);
    // backward accumulator of group g: [96 tpc + 32 g, + NN)
    auto tD = [&](uint32_t k) { return 96u * k; };
    auto tA = [&](uint32_t k) { return 96u * k + 32u; };
    const uint32_t tB = 96u * tpc + 32u * grp;
    uint8_t* sDg = sD + grp * C::SD;
    uint32_t* sGg = sG + grp * NCB * 128;
    const uint32_t sA_u = umma::smem_u32(sA), sD_u = umma::smem_u32(sDg), sW_u = umma::smem_u32(sW);
    constexpr uint32_t idesc_f = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 0, 0, 128, NN);
    constexpr uint32_t idesc_b = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 1, 1, 64, NN);
    const uint64_t dW_f = umma::make_desc(sW_u, NN * 16, 128);
    const uint64_t dA_b = umma::make_desc(sA_u, 128, kTcChunkStride), dD_b = umma::make_desc(sD_u, 128, kTcChunkStride);

    // ---- the CTA's super-tiles: packed words -> operands, once for the whole trajectory (each group expands its own tiles)
    const uint32_t* gwords = a.store_tc + (d.tc_off >> 2);
    for (uint32_t kk = 0; kk < gtiles; ++kk) {
        const uint32_t k = 2 * kk + grp;
        if (wq == 0) {
            if (umma::elect_one()) umma::bulk_load(sGg, gwords + (size_t)(t_begin + k) * NC * 128, NC * 512u, &mbar[4 + grp]);
            __syncwarp();
        }
        umma::mbar_wait(&mbar[4 + grp], kk & 1u);
        uint8_t* rowA = sA + k * sa_bytes + gt * 16;
        const uint32_t ta = tlane + tA(k);
        for (uint32_t c = 0; c < 64; c += 4) umma::tmem_st4(ta + c, 0u, 0u, 0u, 0u);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((uint32_t)i < NC) {
                const uint32_t x = sGg[i * 128 + gt], y = x >> 8;
                const uint4 oa = make_uint4(x & 0x00030003u, x & 0x000C000Cu, x & 0x00300030u, x & 0x00C000C0u);
                const uint4 ob = make_uint4(y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride) = oa;
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride + 128 * 16) = ob;
                umma::tmem_st4(ta + 4 * i, oa.x, oa.y, oa.z, oa.w);
                umma::tmem_st4(ta + 32 + 4 * i, ob.x, ob.y, ob.z, ob.w);
            }
        umma::tmem_st_wait();
        group_sync(grp);                     // every thread of the group has read the staged words before the next copy overwrites them
    }
    // targets of the group's rows, in registers for the whole trajectory
    const f2 zero2 = dup2(0.f);
    f2 tg[kGroupTiles], valid[kGroupTiles];
    const float* tsrc = a.resid ? a.resid : a.tgt;
#pragma unroll
    for (int kk = 0; kk < kGroupTiles; ++kk) {
        const uint32_t rA = (t_begin + 2 * kk + grp) * kTcRows + gt, rB = rA + 128;
        const bool vA = (uint32_t)kk < gtiles && rA < a.n, vB = (uint32_t)kk < gtiles && rB < a.n;
        valid[kk] = mk2(vA ? 1.f : 0.f, vB ? 1.f : 0.f);
        tg[kk] = mk2(vA ? tsrc[rA] : 0.f, vB ? tsrc[rB] : 0.f);
    }

    const float lam_e = pr[d.ep_off];
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    uint32_t nbwd = 0;                                 // backward commits of this group so far (phase parity of mbar[1 + grp])
    __shared__ float s_gb0[W0];
    __shared__ float s_red3[3][8];
    __shared__ int s_status;
    int steps_done = 0, u_turn_step = -1;
    float neg_h_init = 0.f;

    long long tick = clock64();
    auto lap = [&](int phase) {                        // debug timing: one CTA, thread 0 only
        if (a.timing && cta == a.timing_cta && tid == 0) {
            const long long now = clock64();
            a.timing[phase] += (unsigned long long)(now - tick);
            tick = now;
        }
    };
    for (uint32_t ev = 0; ev <= a.L; ++ev) {          // evaluation ev at the current parameters: ev = 0 is the initial one
        lap(7);
        // ---- stage the tail parameters and W' = W0 / sd (three bf16 pieces) from this CTA's copy of the parameters
        TT::stage_tail(s_th + m * W0, wp, tid, NT);
        float* wtmp = s_g;                             // W' in fp32 for the mean fold (the gradient buffer is dead here)
        for (uint32_t k = tid; k < m * W0; k += NT) {
            const uint32_t j = k / W0, c = k % W0;
            const float w = __fdiv_rn(s_th[c * m + j], sd[j]);
            wtmp[k] = w;
            const float v = w * pow2f(100 - 2 * (int)((j & 7u) >> 1));
            const float p0 = bf16_round(v), r1 = v - p0, p1 = bf16_round(r1), p2 = bf16_round(r1 - p1);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sW + (j >> 3) * (NN * 16) + (j & 7u) * 2);
            dst[(0 * W0 + c) * 8] = __float2bfloat16_rn(p0);
            dst[(1 * W0 + c) * 8] = __float2bfloat16_rn(p1);
            dst[(2 * W0 + c) * 8] = __float2bfloat16_rn(p2);
        }
        __syncthreads();
        // mean fold b0' = b0 - sum_j mu_j W'_j: one warp per unit, lanes over the markers in a fixed tree
        for (uint32_t c = warp; c < (uint32_t)W0P; c += NT / 32) {
            float acc = 0.f;
            if (c < (uint32_t)W0)
                for (uint32_t j = lane; j < m; j += 32) acc = fmaf(mu[j], wtmp[j * W0 + c], acc);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) b0p[c] = c < (uint32_t)W0 ? (s_th[m * W0 + T::b_off(0) + c] - acc) * cA : 0.f;
        }
        umma::fence_async_smem();
        umma::fence_before_sync();
        __syncthreads();
        lap(0);      // staging
        // ---- forward contractions of the CTA's tiles (A operand resident in tensor memory)
        if (warp == 0) {
            umma::fence_after_sync();
            if (umma::elect_one()) {
                for (uint32_t k = 0; k < ntile; ++k)
#pragma unroll
                    for (uint32_t h = 0; h < 2; ++h)
#pragma unroll
                        for (uint32_t ks = 0; ks < 4; ++ks)
                            if (ks < NKS)
                                umma::mma_f16_ts(tmem + tD(k) + h * NN, tmem + tA(k) + h * 32 + ks * 8, dW_f + ((ks * 2u * (NN * 16)) >> 4),
                                                 idesc_f, ks > 0);
                umma::commit(&mbar[0]);
            }
            __syncwarp();
        }
        umma::mbar_wait(&mbar[0], ev & 1u);
        umma::fence_after_sync();
        lap(1);      // forward MMAs
        // ---- tail per tile, the two groups side by side; a group's backward contractions accumulate over its tiles
        typename TT::Acc A;
        A.clear();
#pragma unroll
        for (int kk = 0; kk < kGroupTiles; ++kk) {
            if ((uint32_t)kk < gtiles) {
                const uint32_t k = 2 * kk + grp;
                float accA[16], accB[16];
                umma::tmem_ld16x2(tlane + tD(k), tlane + tD(k) + NN, accA, accB);
                umma::fence_before_sync();
                f2 yh, sg0[W0], ef0;
                f2 t = tg[kk];
                TT::part1(accA, accB, wp, b0p, t, a.resid != nullptr && ev == 0, valid[kk], true, A, yh, sg0, ef0);
                const uint32_t rA = (t_begin + k) * kTcRows + gt, rB = rA + 128;
                if (ev == 0) {
                    tg[kk] = t;                         // t = resid + own prediction, fixed for the trajectory (net.rs:280)
                    if (a.resid && a.tgt_out) { if (rA < a.n) a.tgt_out[rA] = lo2(t); if (rB < a.n) a.tgt_out[rB] = hi2(t); }
                    if (a.prev_out) { if (rA < a.n) a.prev_out[rA] = lo2(yh); if (rB < a.n) a.prev_out[rB] = hi2(yh); }
                }
                if (a.ynew_out) { if (rA < a.n) a.ynew_out[rA] = lo2(yh); if (rB < a.n) a.ynew_out[rB] = hi2(yh); }
                {
                    f2 v[W0];
                    TT::delta0(sg0, ef0, A, v);
                    TT::store_pieces(v, sDg + gt * 16);
                }
                umma::fence_async_smem();
                group_sync(grp);
                if (wq == 0) {
                    umma::fence_after_sync();
                    if (umma::elect_one()) {
                        const uint64_t base = dA_b + ((k * sa_bytes) >> 4);
#pragma unroll
                        for (uint32_t ks = 0; ks < kTcRows / 16; ++ks)
                            umma::mma_f16(tmem + tB, base + ks * 16u, dD_b + ks * 16u, idesc_b, ((uint32_t)kk | ks) != 0);
                        umma::commit(&mbar[1 + grp]);
                    }
                    __syncwarp();
                }
                // the group's delta buffer is reused by its next tile: wait for this tile's backward MMAs
                umma::mbar_wait(&mbar[1 + grp], nbwd & 1u);
                ++nbwd;
                umma::fence_after_sync();
            }
        }
        umma::fence_before_sync();
        lap(2);      // tails + backward MMAs
        // ---- partial sums of this CTA
        TT::template reduce_and_store<true, NT / 32>(A, red, warp, lane, tid, s_part, m, P, s_gb0);
        umma::fence_after_sync();
        if (warp < 4) {                                 // accumulator row j lives in lane j % 16 of warp j / 16 (M = 64 layout); both groups' sums
            const uint32_t j = warp * 16 + lane;
            float s0[16], s1[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) s0[q] = s1[q] = 0.f;
            if (ntile > 0) umma::tmem_ld16(tlane + 96u * tpc, s0);          // warp-wide (.sync.aligned)
            if (ntile > 1) umma::tmem_ld16(tlane + 96u * tpc + 32u, s1);
            if (lane < 16 && j < m) {
                const float unscale = pow2f(33 - 2 * (int)((j & 7u) >> 1));
#pragma unroll
                for (int c = 0; c < W0; ++c) {
                    const float s = ((s0[c] + (s0[W0 + c] + s0[2 * W0 + c])) + (s1[c] + (s1[W0 + c] + s1[2 * W0 + c]))) * unscale;
                    s_part[c * m + j] = __fdiv_rn(s - mu[j] * s_gb0[c], sd[j]);
                }
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        const uint32_t tag = a.tag_base + ev + 1u, par = (a.launch_par << 1) | (ev & 1u);   // four slots: a rank that is already in the next transition cannot overwrite what a slower peer still reads
        {
            uint2* mine = a.part + ((size_t)par * ncta + cta) * a.pstride;
            for (uint32_t k = tid; k <= P; k += NT) tg_store(mine + k, s_part[k], tag);
        }
        lap(3);      // partial sums, published
        // ---- values cta, cta + ncta, ...: one warp per value, lanes stride over the CTAs' partials (awaited by tag), summed in
        //      double in a fixed order (lane's partials ascending, then a fixed shuffle tree): deterministic
        bool ok = true;
        {
            const uint2* theirs = a.part + (size_t)par * ncta * a.pstride;
            uint2* out = a.sums + ((size_t)par * kTcpSumCopies + (lane & (kTcpSumCopies - 1))) * a.pstride;
            uint32_t want = 0;
#pragma unroll
            for (int i = 0; i < kTcpMaxGrid / 32; ++i)
                if (lane + 32u * i < ncta) want |= 1u << i;
            for (uint32_t k = cta + warp * ncta; k <= P; k += (NT / 32) * ncta) {
                uint2 v[kTcpMaxGrid / 32];
                ok = tg_await(theirs + (size_t)lane * a.pstride + k, (size_t)32 * a.pstride, want, tag, v, a.error_flag) && ok;
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < kTcpMaxGrid / 32; ++i)
                    if (want >> i & 1u) s += (double)__uint_as_float(v[i].x);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                const bool all_ok = __all_sync(0xffffffffu, ok);
                if (a.world > 1) {          // this rank's sum -> slot [par][rank] of every rank's table (its own included)
                    if (lane < a.world && all_ok) tg_store_sys(a.rank_sums[lane] + ((size_t)par * 8 + a.rank) * a.pstride + k, (float)s, tag);
                } else if (lane < (uint32_t)kTcpSumCopies && all_ok) tg_store(out + k, (float)s, tag);
            }
        }
        lap(5);      // slice reduction
        // ---- every CTA gathers all the sums
        if (a.world > 1) {
            // the world rank-level sums of every value, from this rank's own table, added in rank order
            const uint2* in = a.rank_sums[a.rank] + (size_t)par * 8 * a.pstride;
            const uint32_t want = (1u << a.world) - 1u;
            for (uint32_t k = tid; k <= P; k += NT) {
                uint2 v[8];
                ok = tg_await<8, true>(in + k, a.pstride, want, tag, v, a.error_flag) && ok;
                float s = __uint_as_float(v[0].x);
#pragma unroll
                for (int r = 1; r < 8; ++r)
                    if ((uint32_t)r < a.world) s += __uint_as_float(v[r].x);
                s_sum[k] = s;
            }
        } else {
            const uint2* in = a.sums + ((size_t)par * kTcpSumCopies + (cta & (kTcpSumCopies - 1))) * a.pstride;
            constexpr int NG = (kTcpMaxValues + NT - 1) / NT;
            uint32_t want = 0;
#pragma unroll
            for (int i = 0; i < NG; ++i)
                if (tid + (uint32_t)NT * i <= P) want |= 1u << i;
            uint2 v[NG];
            ok = tg_await(in + tid, NT, want, tag, v, a.error_flag) && ok;
#pragma unroll
            for (int i = 0; i < NG; ++i)
                if (want >> i & 1u) s_sum[tid + (uint32_t)NT * i] = __uint_as_float(v[i].x);
        }
        if (__syncthreads_or(!ok)) break;
        lap(6);      // gather
        // ---- the update, replicated in every CTA (k2_step's arithmetic; branch_sampler.rs:1239-1284)
        const bool is_init = ev == 0, is_last = ev == a.L;
        float kin = 0.f, prior = 0.f, uturn = 0.f;
        for (uint32_t k = tid; k < P; k += NT) {
            const float w = s_th[k], lam = s_lam[k];
            float g;
            if (lam < 0.f) {                                          // a bias (branch_sampler.rs:322-331)
                g = -(lam_e * s_sum[k]);
                if (a.model == BANN_STD_NORMAL) prior -= 0.5f * w * w;
            } else if (a.model == BANN_STD_NORMAL) { g = -(lam_e * s_sum[k] + w); prior -= 0.5f * w * w; }
            else if (lasso) {
                const float sg = (w > 0.f) ? 1.f : (w < 0.f ? -1.f : 0.f);
                g = -(lam_e * s_sum[k] + lam * sg);
                prior -= lam * fabsf(w);
            } else { g = -(lam_e * s_sum[k] + lam * w); prior -= 0.5f * lam * w * w; }
            s_g[k] = g;
            float pk = s_p[k];
            if (!is_init) { pk = pk + (0.5f * s_eps[k]) * g; s_p[k] = pk; }
            kin = fmaf(pk, pk, kin);
            uturn = fmaf(w - s_th0[k], pk, uturn);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            kin += __shfl_xor_sync(0xffffffffu, kin, o);
            prior += __shfl_xor_sync(0xffffffffu, prior, o);
            uturn += __shfl_xor_sync(0xffffffffu, uturn, o);
        }
        if (lane == 0) { s_red3[0][warp] = kin; s_red3[1][warp] = prior; s_red3[2][warp] = uturn; }
        __syncthreads();
        if (tid == 0) {
            auto sum8 = [&](const float* r) { return ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7])); };
            kin = sum8(s_red3[0]);
            prior = sum8(s_red3[1]);
            uturn = sum8(s_red3[2]);
            const float rss = s_sum[P];
            const float ld = prior + (-1.0f * lam_e * (rss / 2.0f));
            const float negh = ld - 0.5f * kin;
            int status = ST_RUNNING;
            if (is_init) neg_h_init = negh;
            else {
                const int step = steps_done;
                steps_done = step + 1;
                if (fabsf(negh - neg_h_init) > a.max_h_err) status = ST_REJECTED_EARLY;
                else if (u_turn_step < 0 && uturn < 0.f) u_turn_step = step;
            }
            s_status = status;
            if (cta == 0) {
                BranchState& st = *a.state;
                st.rss = rss;
                st.log_density = ld;
                st.neg_h_cur = negh;
                if (is_init) st.neg_h_init = negh;
                st.steps_done = steps_done;
                st.u_turn_step = u_turn_step;
                st.status = status;
            }
        }
        __syncthreads();
        if (s_status == ST_REJECTED_EARLY || is_last) {
            if (cta == 0) {
                const float* th_end = s_status == ST_REJECTED_EARLY ? s_th0 : s_th;
                for (uint32_t k = tid; k < P; k += NT) { a.theta[d.param_off + k] = th_end[k]; a.grad[d.param_off + k] = s_g[k]; a.mom[d.param_off + k] = s_p[k]; }
                for (uint32_t k = tid; k <= P; k += NT) a.gsum[k] = s_sum[k];
            }
            break;
        }
        for (uint32_t k = tid; k < P; k += NT) {
            const float e = s_eps[k];
            const float pk = s_p[k] + (0.5f * e) * s_g[k];
            s_p[k] = pk;
            s_th[k] = s_th[k] + e * pk;
        }
        __syncthreads();
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, kTmemCols);
}

}  // namespace bann
