// K1 on the 5th-generation tensor cores (tcgen05 + tensor memory), for branches with <= 64 markers.
//
// The two big contractions of a branch -- z0 = X W0 (forward) and S = X^T delta_0 (backward) -- run as
// tcgen05.mma kind::f16 with bf16 operands and f32 accumulators in tensor memory; everything in
// between (remaining layers, tanh, error, deltas, cross-row sums of the small layers) stays in FP32
// registers, one thread per individual, exactly as in k1_small.
//
// Genotypes never get converted arithmetically.  The tensor-core store (genotypes.cu, k_build_tc)
// keeps, per 256-row super-tile and 8-marker chunk, one 32-bit word per ROW PAIR (t, 128 + t) whose
// 2-bit fields sit where a single AND turns them into bf16 numbers: a genotype code g at bits
// [2p+1:2p] (p = 0..3) of a 16-bit half is the bf16 SUBNORMAL g * 4^p * 2^-133 (exact; checked on
// hardware by tests/probe/probe_umma.cu).  One LOP3 yields two operand elements; the per-marker power
// of two 4^p(j) is folded into the staged weights (forward) and undone on the accumulator (backward).
// The expanded tile [chunk][row] x 16 B is written once per super-tile and read by BOTH contractions:
// as the K-major A operand (rows x markers) and as the MN-major A operand (markers x rows).
//
// f32 weights and deltas enter as three bf16 pieces each (hi + mid + lo = the f32 value, 24 significand
// bits), side by side in the N dimension, so products are exact and only the f32 accumulation rounds.
// Pieces are pre-scaled by 2^100 (exact) so that subnormal * piece stays a normal f32.
//
// Per CTA (128 threads, one branch, a range of super-tiles): expand -> MMA fwd (2 x M128 N16 K16*ks)
// -> tcgen05.ld -> per-row tail -> delta pieces to shared memory -> MMA bwd (M64 N16 K16 x 16,
// accumulating in tensor memory over ALL super-tiles of the CTA) -> one read of the gradient at the end.
#pragma once
#include <cuda_bf16.h>

#include "k1_small.cuh"
#include "umma.cuh"

namespace bann {

constexpr int kTcMaxMarkers = 64;     // one M = 64 backward accumulator, four K = 16 forward steps
constexpr int kTcRows = 256;          // rows per super-tile: thread t owns rows t and 128 + t
constexpr uint32_t kTcChunkStride = kTcRows * 16;   // bytes between 8-marker chunks of the expanded tile

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float pow2f(int e) { return __int_as_float((127 + e) << 23); }

template <int H, int S, int D>
struct TcShape {
    using T = TailShape<H, S, D>;
    static constexpr int W0 = T::W0;
    static constexpr int NN = 16;                       // accumulator columns: 3 pieces x W0 units, padded
    static_assert(3 * W0 <= NN, "first-layer width too large for one N = 16 accumulator");
    static constexpr int TMEM_COLS = 64;                // 2 x NN forward (row halves) + NN backward
    static constexpr size_t SA = 8 * kTcChunkStride;    // expanded genotypes, 8 chunks x 256 rows x 16 B
    static constexpr size_t SD = (NN / 8) * kTcChunkStride;   // delta pieces  [n-chunk][row] x 16 B
    static constexpr size_t SW = 8 * NN * 16;           // weight pieces [k-chunk][n] x 16 B
    static constexpr int NRED = 4 * (T::NTACC > 64 ? T::NTACC : 64);
    static constexpr size_t SMEM = SA + SD + SW + (size_t)(((T::n_tail() + 3) & ~3) + T::W0P + NRED + 8) * 4 + 64 + 128;
};

template <int H, int S, int D>
__global__ void __launch_bounds__(128, 4) k1_tc(K1Args a) {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    constexpr int NLA = T::NLA, W0 = T::W0, W0P = T::W0P, MW = T::MW, NTACC = T::NTACC, NN = C::NN;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    if (a.states && a.states[b].status != ST_RUNNING) return;
    const BranchDesc& d = a.descs[b];
    const uint32_t m = d.m, NC = d.nc, NKS = (NC + 1) >> 1;
    // ---- shared memory carve-up
    uint8_t* sA = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smraw) + 127) & ~(uintptr_t)127);   // whole core matrices
    uint8_t* sD = sA + C::SA;
    uint8_t* sW = sD + C::SD;
    float* sp = reinterpret_cast<float*>(sW + C::SW);          // tail parameters [n_tail]
    float* b0p = sp + ((T::n_tail() + 3) & ~3);                // [W0P] first-layer bias with the means folded in
    float* red = b0p + W0P;                                    // [NRED] reduction scratch
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + C::NRED);   // [0] forward done, [1] backward done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;

    // ---- one-time setup: zero operand buffers, barriers, tensor memory
    for (uint32_t k = tid; k < (C::SA + C::SD + C::SW) / 16; k += 128) reinterpret_cast<uint4*>(sA)[k] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 1);
        umma::mbar_init(&mbar[1], 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, C::TMEM_COLS);
    for (uint32_t k = tid; k < (uint32_t)T::n_tail(); k += 128) sp[k] = th[m * W0 + k];
    __syncthreads();
    // ---- stage W' = W0 / sd (f32, in the delta buffer for the bias fold) and its three bf16 pieces
    float* wtmp = reinterpret_cast<float*>(sD);                // [m][W0], transient
    for (uint32_t k = tid; k < m * W0; k += 128) {
        const uint32_t j = k / W0, c = k % W0;
        const float w = __fdiv_rn(th[c * m + j], sd[j]);       // bed.rs:354 folded into the first layer
        wtmp[k] = w;
        const float v = w * pow2f(100 - 2 * (int)((j & 7u) >> 1));   // exact: undoes 4^p of the operand, lifts out of the subnormals
        const float p0 = bf16_round(v), r1 = v - p0, p1 = bf16_round(r1), p2 = bf16_round(r1 - p1);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sW + (j >> 3) * (NN * 16) + (j & 7u) * 2);
        dst[(0 * W0 + c) * 8] = __float2bfloat16_rn(p0);
        dst[(1 * W0 + c) * 8] = __float2bfloat16_rn(p1);
        dst[(2 * W0 + c) * 8] = __float2bfloat16_rn(p2);
    }
    __syncthreads();
    if (tid < W0P) {
        float acc = 0.f;
        if (tid < W0) {
            for (uint32_t j = 0; j < m; ++j) acc = fmaf(mu[j], wtmp[j * W0 + tid], acc);
            acc = sp[T::b_off(0) + tid] - acc;
        }
        b0p[tid] = acc;
    }
    __syncthreads();
    for (uint32_t k = tid; k < m * W0; k += 128) wtmp[k] = 0.f;   // the pad columns of the delta operand must stay zero
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);
    const uint32_t sA_u = umma::smem_u32(sA), sD_u = umma::smem_u32(sD), sW_u = umma::smem_u32(sW);
    constexpr uint32_t idesc_f = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 0, 0, 128, NN);
    constexpr uint32_t idesc_b = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 1, 1, 64, NN);

    // ---- persistent per-thread accumulators (cross-row sums of the layers >= 1)
    float gb0[W0], gWo[S], rss = 0.f;
    float gWt[NLA > 1 ? NLA - 1 : 1][MW][MW], gbt[NLA > 1 ? NLA - 1 : 1][MW];
#pragma unroll
    for (int c = 0; c < W0; ++c) gb0[c] = 0.f;
#pragma unroll
    for (int c = 0; c < S; ++c) gWo[c] = 0.f;
#pragma unroll
    for (int l = 0; l < (NLA > 1 ? NLA - 1 : 1); ++l)
#pragma unroll
        for (int i = 0; i < MW; ++i) {
            gbt[l][i] = 0.f;
#pragma unroll
            for (int c = 0; c < MW; ++c) gWt[l][i][c] = 0.f;
        }

    const size_t eoff = a.out_per_entry ? (size_t)li * a.n : 0;
    const size_t toff = (a.target_mode == TGT_PER_ENTRY) ? (size_t)li * a.n : 0;
    const uint32_t t_begin = chunk * a.st_per_chunk;
    const uint32_t t_end = min(a.nst, t_begin + a.st_per_chunk);
    const uint32_t* gbase = a.store_tc + (d.tc_off >> 2) + tid;
    const float* tsrc = (a.target_mode == TGT_RESID_PLUS_PRED) ? a.resid : (a.tgt ? a.tgt + toff : nullptr);

    uint32_t wreg[8];
    auto load_words = [&](uint32_t st) {
        const uint32_t* src = gbase + (size_t)st * NC * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((uint32_t)i < NC) wreg[i] = __ldg(src + i * 128);
    };
    if (t_begin < t_end) load_words(t_begin);
    uint32_t it = 0;
    for (uint32_t st = t_begin; st < t_end; ++st, ++it) {
        // the previous backward contraction still reads both operand buffers
        if (it > 0 && !a.fwd_only) umma::mbar_wait(&mbar[1], (it - 1) & 1u);
        // ---- expand: one AND per two operand elements, 16-byte conflict-free stores
        {
            uint8_t* rowA = sA + tid * 16;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if ((uint32_t)i < NC) {
                    const uint32_t x = wreg[i], y = x >> 8;
                    *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride) =
                        make_uint4(x & 0x00030003u, x & 0x000C000Cu, x & 0x00300030u, x & 0x00C000C0u);
                    *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride + 128 * 16) =
                        make_uint4(y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
                }
        }
        if (st + 1 < t_end) load_words(st + 1);
        const uint32_t rowA_g = st * kTcRows + tid, rowB_g = rowA_g + 128;
        float tg2[2] = {0.f, 0.f};
        if (tsrc) {
            if (rowA_g < a.n) tg2[0] = __ldg(tsrc + rowA_g);
            if (rowB_g < a.n) tg2[1] = __ldg(tsrc + rowB_g);
        }
        umma::fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            umma::fence_after_sync();
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h)
                for (uint32_t ks = 0; ks < NKS; ++ks) {
                    const uint64_t ad = umma::make_desc(sA_u + h * 2048u + ks * 2u * kTcChunkStride, kTcChunkStride, 128);
                    const uint64_t bd = umma::make_desc(sW_u + ks * 2u * (NN * 16), NN * 16, 128);
                    umma::mma_f16(tmem + h * NN, ad, bd, idesc_f, ks > 0);
                }
            umma::commit(&mbar[0]);
        }
        umma::mbar_wait(&mbar[0], it & 1u);
        umma::fence_after_sync();
        float acc[2][16];
        umma::tmem_ld16x2(tlane, tlane + NN, acc[0], acc[1]);

        // ---- tail: remaining layers, error, backward deltas, one row per half
        float yh2[2], dl0[2][W0];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool valid = (h ? rowB_g : rowA_g) < a.n;
            float act[NLA][MW];
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float z = acc[h][c] + (acc[h][W0 + c] + acc[h][2 * W0 + c]);
                act[0][c] = fast_tanh(fmaf(z, 8589934592.f /* 2^33 */, b0p[c]));
            }
#pragma unroll
            for (int l = 1; l < NLA; ++l) {
#pragma unroll
                for (int c = 0; c < MW; ++c) {
                    if (c < T::width(l)) {
                        float zz = sp[T::b_off(l) + c];
#pragma unroll
                        for (int i = 0; i < MW; ++i)
                            if (i < T::in_w(l)) zz = fmaf(act[l - 1][i], sp[T::w_off(l) + c * T::in_w(l) + i], zz);
                        act[l][c] = fast_tanh(zz);
                    }
                }
            }
            float yh = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i) yh = fmaf(act[NLA - 1][i], sp[T::w_off(NLA) + i], yh);
            yh2[h] = yh;
            float tg = tg2[h];
            if (a.target_mode == TGT_RESID_PLUS_PRED) { tg = tg + yh; tg2[h] = tg; }   // net.rs:280
            const float e = valid ? yh - tg : 0.f;                                      // branch_sampler.rs:821
            rss = fmaf(e, e, rss);
            float delta[MW];
#pragma unroll
            for (int i = 0; i < S; ++i) {
                gWo[i] = fmaf(act[NLA - 1][i], e, gWo[i]);
                delta[i] = (1.f - act[NLA - 1][i] * act[NLA - 1][i]) * (e * sp[T::w_off(NLA) + i]);
            }
#pragma unroll
            for (int l = NLA - 1; l >= 1; --l) {
                float nd[MW];
#pragma unroll
                for (int i = 0; i < MW; ++i) nd[i] = 0.f;
#pragma unroll
                for (int c = 0; c < MW; ++c) {
                    if (c < T::width(l)) {
                        gbt[l - 1][c] += delta[c];
#pragma unroll
                        for (int i = 0; i < MW; ++i)
                            if (i < T::in_w(l)) {
                                gWt[l - 1][i][c] = fmaf(act[l - 1][i], delta[c], gWt[l - 1][i][c]);
                                nd[i] = fmaf(delta[c], sp[T::w_off(l) + c * T::in_w(l) + i], nd[i]);
                            }
                    }
                }
#pragma unroll
                for (int i = 0; i < MW; ++i)
                    if (i < T::in_w(l)) delta[i] = (1.f - act[l - 1][i] * act[l - 1][i]) * nd[i];
            }
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                gb0[c] += delta[c];
                dl0[h][c] = delta[c];
            }
        }
        // ---- per-row outputs
        {
            auto put = [&](float* dst, uint32_t row, float v, int accumulate) {
                if (!dst || row >= a.n) return;
                float* p = dst + eoff + row;
                if (accumulate > 0) *p += v;
                else if (accumulate < 0) *p -= v;
                else *p = v;
            };
            if (a.target_mode == TGT_RESID_PLUS_PRED) {
                put(a.tgt_out, rowA_g, tg2[0], 0); put(a.tgt_out, rowB_g, tg2[1], 0);
                put(a.prev_out, rowA_g, yh2[0], 0); put(a.prev_out, rowB_g, yh2[1], 0);   // net.rs:279
            }
            put(a.yhat_out, rowA_g, yh2[0], a.yhat_accumulate);
            put(a.yhat_out, rowB_g, yh2[1], a.yhat_accumulate);
        }
        umma::fence_before_sync();      // the accumulator reads above precede the next forward MMA
        if (!a.fwd_only) {
            // ---- delta_0 -> three bf16 pieces per unit (n = piece * W0 + unit), MN-major B operand
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float pv[NN];
#pragma unroll
                for (int n = 0; n < NN; ++n) pv[n] = 0.f;
#pragma unroll
                for (int c = 0; c < W0; ++c) {
                    const float v = dl0[h][c] * 1.2676506002282294e30f;   // 2^100, exact
                    const float p0 = bf16_round(v), r1 = v - p0, p1 = bf16_round(r1), p2 = bf16_round(r1 - p1);
                    pv[c] = p0; pv[W0 + c] = p1; pv[2 * W0 + c] = p2;
                }
                uint8_t* dst = sD + (h * 128 + tid) * 16;
#pragma unroll
                for (int q = 0; q < NN / 8; ++q)
                    *reinterpret_cast<uint4*>(dst + q * kTcChunkStride) =
                        make_uint4(pack_bf16(pv[8 * q], pv[8 * q + 1]), pack_bf16(pv[8 * q + 2], pv[8 * q + 3]),
                                   pack_bf16(pv[8 * q + 4], pv[8 * q + 5]), pack_bf16(pv[8 * q + 6], pv[8 * q + 7]));
            }
            umma::fence_async_smem();
            __syncthreads();
            if (tid == 0) {
                umma::fence_after_sync();
#pragma unroll
                for (uint32_t ks = 0; ks < kTcRows / 16; ++ks) {
                    const uint64_t ad = umma::make_desc(sA_u + ks * 256u, 128, kTcChunkStride);
                    const uint64_t bd = umma::make_desc(sD_u + ks * 256u, 128, kTcChunkStride);
                    umma::mma_f16(tmem + 2 * NN, ad, bd, idesc_b, (it | ks) != 0);
                }
                umma::commit(&mbar[1]);
            }
        } else {
            __syncthreads();
        }
    }
    const bool has_bwd = !a.fwd_only && a.part && it > 0;
    float sacc[16];
    if (has_bwd) {
        umma::mbar_wait(&mbar[1], (it - 1) & 1u);
        umma::fence_after_sync();
        umma::tmem_ld16(tlane + 2 * NN, sacc);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, C::TMEM_COLS);
    if (a.fwd_only || !a.part) return;

    // ---- CTA epilogue: fixed-order reduction over lanes and warps, unfold the standardisation
    float* pp = a.part + ((size_t)li * a.nchunk + chunk) * a.pstride;
    const uint32_t P = d.P;
    {
        float* rw = red + warp * NTACC;
        int idx = 0;
        auto put = [&](float v) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) rw[idx] = v;
            ++idx;
        };
        put(rss);
#pragma unroll
        for (int c = 0; c < S; ++c) put(gWo[c]);
#pragma unroll
        for (int c = 0; c < W0; ++c) put(gb0[c]);
#pragma unroll
        for (int l = 1; l < NLA; ++l) {
#pragma unroll
            for (int c = 0; c < MW; ++c) put(gbt[l - 1][c]);
#pragma unroll
            for (int i = 0; i < MW; ++i)
#pragma unroll
                for (int c = 0; c < MW; ++c) put(gWt[l - 1][i][c]);
        }
    }
    __syncthreads();
    __shared__ float s_gb0[W0];
    if (tid < NTACC) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) s += red[w * NTACC + tid];
        int idx = tid;
        if (idx == 0) pp[P] = s;
        else if (idx < 1 + S) pp[m * W0 + T::w_off(NLA) + (idx - 1)] = s;                       // output weights
        else if (idx < 1 + S + W0) { pp[m * W0 + T::b_off(0) + (idx - 1 - S)] = s; s_gb0[idx - 1 - S] = s; }
        else {
            int k = idx - (1 + S + W0);
            const int per = MW + MW * MW;
            const int l = 1 + k / per;
            k %= per;
            if (k < MW) {
                if (k < T::width(l)) pp[m * W0 + T::b_off(l) + k] = s;
            } else {
                k -= MW;
                const int i = k / MW, c = k % MW;
                if (i < T::in_w(l) && c < T::width(l)) pp[m * W0 + T::w_off(l) + c * T::in_w(l) + i] = s;
            }
        }
    }
    __syncthreads();
    // first-layer weight gradient: accumulator row j lives in lane (j % 16) of warp j / 16 (M = 64 layout);
    // S_jc = sum of the three pieces * 2^33 / 4^p(j), then (S_jc - mu_j * gb0_c) / sd_j
    if (lane < 16) {
        const uint32_t j = warp * 16 + lane;
        if (j < m) {
            const float unscale = pow2f(33 - 2 * (int)((j & 7u) >> 1));
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float s = (it > 0 ? (sacc[c] + (sacc[W0 + c] + sacc[2 * W0 + c])) : 0.f) * unscale;
                pp[c * m + j] = __fdiv_rn(s - mu[j] * s_gb0[c], sd[j]);
            }
        }
    }
}

template <int H, int S, int D>
int launch_one_tc(K1Args& a, uint32_t nlist, cudaStream_t st) {
    using C = TcShape<H, S, D>;
    auto kern = k1_tc<H, S, D>;
    static bool configured = false;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        configured = true;
    }
    dim3 grid(a.nchunk, nlist);
    kern<<<grid, 128, C::SMEM, st>>>(a);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

// Tensor-core K1 for a homogeneous launch: tanh, every listed branch of the same architecture with at
// most 64 markers and 3 * W0 <= 16, and a tensor-core store present.  Otherwise *launched stays false.
inline int launch_k1_tc(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                        cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net) {
    *launched = false;
    if (a.act != BANN_TANH || !a.store_tc) return 0;
    const BranchDesc& d0 = descs[single_branch >= 0 ? single_branch : 0];
    uint32_t max_m = d0.m;
    if (single_branch < 0) {
        for (const BranchDesc& d : descs) {
            if (d.nl != d0.nl) return 0;
            for (uint32_t l = 0; l < d.nl; ++l)
                if (d.widths[l] != d0.widths[l]) return 0;
            max_m = std::max(max_m, d.m);
        }
    }
    if (max_m > (uint32_t)kTcMaxMarkers) return 0;
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    const uint32_t nst = a.nst;
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)num_sms * 4 + nlist - 1) / nlist);
    uint32_t nchunk = std::min<uint32_t>(want, std::max<uint32_t>(1, nst));
    uint32_t spc = (nst + nchunk - 1) / nchunk;
    nchunk = (nst + spc - 1) / spc;
#define BANN_TRY_TC(HH, SS, DD)                                                                           \
    if (!*launched && H == HH && S == SS && D == DD) {                                                    \
        a.nchunk = nchunk;                                                                                \
        a.st_per_chunk = spc;                                                                             \
        if (part_io) {                                                                                    \
            if (nchunk == 1) *part_io = bann_net_gsum(net);                                               \
            else {                                                                                        \
                float* p = bann_net_partials(net, (size_t)nlist * nchunk * bann_net_pstride(net));        \
                if (!p) return -2;                                                                        \
                *part_io = p;                                                                             \
            }                                                                                             \
            a.part = *part_io;                                                                            \
        }                                                                                                 \
        *nchunk_io = nchunk;                                                                              \
        int rc = launch_one_tc<HH, SS, DD>(a, nlist, st);                                                 \
        if (rc) return rc;                                                                                \
        *launched = true;                                                                                 \
    }
    BANN_TRY_TC(5, 5, 1)
    BANN_TRY_TC(2, 2, 1)
    BANN_TRY_TC(4, 3, 1)
    BANN_TRY_TC(4, 3, 2)
    BANN_TRY_TC(5, 3, 2)
    BANN_TRY_TC(2, 2, 0)
#undef BANN_TRY_TC
    return 0;
}

}  // namespace bann
