// k1_tc with a dedicated fifth warp.  Same operands, store, tail arithmetic and numerics as k1_tc.cuh; what changes is WHO issues.
// In k1_tc the four compute warps issue the tcgen05.mma themselves (forward row half 0 / 1, backward row half 0 / 1).  The ncu
// source page of round 2 (profiles/r2_k1_tc_ncu_summary.md) shows what that costs: the two backward-issuing warps spend ~21 % of
// their iteration in the elected lane's UTCHMMA queue stalls, the barrier wait before them and the reconvergence behind them, and
// the two forward-issuing warps then wait ~14 % of theirs for the late warps' "expanded" arrivals -- the CTA moves at the pace of
// its slowest warp.  Here warp 4 does nothing but wait on the two 128-arrival barriers, issue the 8 + 16 MMAs of a super-tile and
// request the packed words; the compute warps never leave the FP32 tail, and the backward contraction accumulates in ONE
// tensor-memory accumulator (one issuer, one fixed order).
// Registers: 160 threads x 3 CTAs per SM leave 128 per thread (136 by division, 128 by the allocation granularity).  setmaxnreg cannot help (it dead-locks with a lone fifth warp,
// tests/probe/probe_setmaxnreg.cu); instead the cross-row sums are kept as ONE float per value (TcTail<..., SACC = true>: 41
// registers instead of 82 for [5,5,1]), which brings the compute warps under the limit without spills.
#pragma once
#include "k1_tc.cuh"

namespace bann {

constexpr int kTc5Threads = 160;

// DEFER: where the cross-row sums of a super-tile (layers >= 1, output layer, rss: ~85 FMA-pipe instructions that nothing waits
// for) are taken.  false: inside part 1, as k1_tc.  true: one super-tile later, placed by hand between the first layer's
// activations of the NEXT super-tile, where the warp otherwise only waits for the MUFU pipe (ex2 / rcp, 8 clk per warp
// instruction); the terms they need (activations, rho of the layers >= 1, error: `Mid`) stay in registers until then.
// Measured on cfg3s (100k x 1000 branches x 50 markers, one B200, profiles/r2_k1_tc5_ncu_summary.md): k1_tc 1.641 ms, five warps 1.446,
// five warps + deferred sums 1.387.  Tried and rejected on the way: the sums behind the expansion of the next super-tile (1.395)
// or behind the delta pieces (1.429); the expansion moved up between the units of layer 1 or between the steps of the rho chain
// (1.61 / 1.65: the earlier the wait for the previous backward contraction sits, the less skew between the four compute warps
// of a CTA is tolerated); the loop unrolled twice to save the register copies of the deferred terms (spills, 2.16).
template <int H, int S, int D, int ACT, bool LEAN, int NCT, bool DEFER = true>
__global__ void __launch_bounds__(kTc5Threads, 3) k1_tc5(K1Args a) {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    using TT = TcTail<H, S, D, ACT, true>;
    constexpr int W0 = T::W0, W0P = T::W0P, NN = C::NN;
    constexpr float cA = TT::cA;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // provably warp-uniform
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    const BranchDesc& d = a.descs[b];
    const uint32_t m = d.m, NC = NCT ? (uint32_t)NCT : d.nc, NKS = (NC + 1) >> 1, NCB = a.ncb;
    // ---- shared memory carve-up (as k1_tc)
    uint8_t* sA = smraw + ((128u - (umma::smem_u32(smraw) & 127u)) & 127u);
    const uint32_t sa_bytes = NCB * kTcChunkStride;
    uint8_t* sW = sA + (size_t)2 * sa_bytes;
    uint32_t* sG = reinterpret_cast<uint32_t*>(sW + C::SW);
    uint8_t* sD = reinterpret_cast<uint8_t*>(sG) + (size_t)NCB * 512;
    float* wp = reinterpret_cast<float*>(__builtin_assume_aligned(sD + C::SD, 16));
    float* b0p = wp + ((T::n_tail() + 3) & ~3);
    float* red = b0p + ((T::n_tail() + 3) & ~3) + 2 * W0P;
    // [0] forward MMAs of a super-tile done, [1] backward MMAs done (one tcgen05.commit each); [2] super-tile expanded,
    // [3] delta pieces written (128 arrivals each); [4] packed words landed (bulk copy transaction bytes)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + C::NRED);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 6);

    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;

    // ---- prologue.  At 8 GPUs a CTA covers only 49 super-tiles, so what happens before the first one counts (measured on the
    // 8-GPU shard shape, bench.py --workload cfg3r8: ~10 % of K1): every global load of the staging is issued up front (one
    // memory round trip instead of four dependent ones), the first packed words are requested before the staging, the mean fold
    // of the first-layer bias is a warp reduction per unit instead of one thread walking the markers, and only what is read
    // before it is written gets zeroed (the weight pieces of the k-chunks >= nc; the operand images and the delta pieces are
    // rewritten for every super-tile, rows of the backward accumulator beyond m are never read).
    for (uint32_t k = tid; k < (uint32_t)(C::SW / 16); k += kTc5Threads) reinterpret_cast<uint4*>(sW)[k] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 1);
        umma::mbar_init(&mbar[1], 1);
        umma::mbar_init(&mbar[2], 128);
        umma::mbar_init(&mbar[3], 128);
        umma::mbar_init(&mbar[4], 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, C::TMEM_COLS);
    pdl_launch_dependents();
    pdl_wait();
    if (a.states && a.states[b].status != ST_RUNNING) {
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (warp == 0) umma::tmem_dealloc(*tmem_slot, C::TMEM_COLS);
        return;
    }
    const uint32_t t_begin = chunk * a.st_per_chunk;
    const uint32_t t_end = min(a.nst, t_begin + a.st_per_chunk);
    const uint32_t nit = t_end > t_begin ? t_end - t_begin : 0;
    const uint32_t* gwords = a.store_tc + (d.tc_off >> 2);
    // all global loads of the staging, before anything waits
    const float* th_tail = th + m * W0;
    const uint32_t mw = m * W0;
    constexpr uint32_t NQ = (kTcMaxMarkers * 5 + kTc5Threads - 1) / kTc5Threads;       // W0 <= 5
    float g_th[NQ], g_sd[NQ], g_mu[NQ];
#pragma unroll
    for (uint32_t q = 0; q < NQ; ++q) {
        const uint32_t k = tid + q * kTc5Threads;
        const bool in = k < mw;
        const uint32_t j = in ? k / W0 : 0, c = in ? k % W0 : 0;
        g_th[q] = in ? th[c * m + j] : 0.f;
        g_sd[q] = in ? sd[j] : 1.f;
        g_mu[q] = in ? mu[j] : 0.f;
    }
    const float g_tail = tid < (uint32_t)T::n_tail() ? th_tail[tid] : 0.f;
    const float g_b0 = (lane == 0 && warp < (uint32_t)W0) ? th_tail[T::b_off(0) + warp] : 0.f;
    __syncthreads();                       // barrier initialisation visible to the issuing warp
    if (warp == 4 && nit > 0) {            // the first packed words: their HBM round trip runs under the staging
        if (umma::elect_one()) umma::bulk_load(sG, gwords + (size_t)t_begin * NC * 128, NC * 512u, &mbar[4]);
        __syncwarp();
    }
    {   // tail parameters; what feeds an activated layer >= 1 carries the activation's pre-scale (TcTail::stage_tail)
        auto put_tail = [&](uint32_t k, float w) {
            const bool scaled = TT::NLA > 1 && ((int)k < T::w_off(TT::NLA) || (int)k >= T::b_off(TT::NLA > 1 ? 1 : 0));
            wp[k] = scaled ? w * cA : w;
        };
        if (tid < (uint32_t)T::n_tail()) put_tail(tid, g_tail);
        for (uint32_t k = tid + kTc5Threads; k < (uint32_t)T::n_tail(); k += kTc5Threads) put_tail(k, th_tail[k]);
    }
    float* prod = reinterpret_cast<float*>(sD);                // [W0][64] products mu_j * W'_jc, transient
#pragma unroll
    for (uint32_t q = 0; q < NQ; ++q) {
        const uint32_t k = tid + q * kTc5Threads;
        if (k < mw) {
            const uint32_t j = k / W0, c = k % W0;
            const float w = __fdiv_rn(g_th[q], g_sd[q]);       // bed.rs:354 folded into the first layer
            prod[c * kTcMaxMarkers + j] = g_mu[q] * w;
            const float v = w * pow2f(100 - 2 * (int)((j & 7u) >> 1));
            const float p0 = bf16_round(v), r1 = v - p0, p1 = bf16_round(r1), p2 = bf16_round(r1 - p1);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sW + (j >> 3) * (NN * 16) + (j & 7u) * 2);
            dst[(0 * W0 + c) * 8] = __float2bfloat16_rn(p0);
            dst[(1 * W0 + c) * 8] = __float2bfloat16_rn(p1);
            dst[(2 * W0 + c) * 8] = __float2bfloat16_rn(p2);
        }
    }
    if (tid < (uint32_t)W0P) b0p[tid] = 0.f;
    __syncthreads();
    if (warp < (uint32_t)W0) {             // b0' = (b0 - sum_j mu_j W'_jc) * cA: lanes over the markers, fixed shuffle tree
        float sacc0 = (lane < m ? prod[warp * kTcMaxMarkers + lane] : 0.f) + (lane + 32 < m ? prod[warp * kTcMaxMarkers + lane + 32] : 0.f);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sacc0 += __shfl_xor_sync(0xffffffffu, sacc0, o);
        if (lane == 0) b0p[warp] = (g_b0 - sacc0) * cA;
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t sA_u = umma::smem_u32(sA), sD_u = umma::smem_u32(sD), sW_u = umma::smem_u32(sW);
    constexpr uint32_t idesc_f = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 0, 0, 128, NN);
    constexpr uint32_t idesc_b = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 1, 1, 64, NN);

    const bool bwd = LEAN || !a.fwd_only;
    const bool epilogue = bwd && a.part;

    if (warp == 4) {
        // ================================================================ issuing warp
        const uint64_t dW_f = umma::make_desc(sW_u, NN * 16, 128);
        const uint64_t dA_b = umma::make_desc(sA_u, 128, kTcChunkStride), dD_b = umma::make_desc(sD_u, 128, kTcChunkStride);
        auto issue_bwd = [&](uint32_t e) {      // backward contraction of super-tile e: 16 K steps of 16 rows into ONE accumulator
            umma::mbar_wait(&mbar[3], e & 1u);  // its delta pieces are written (128 arrivals, each behind a proxy fence)
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint64_t base = dA_b + (((e & 1u) * sa_bytes) >> 4);
#pragma unroll
                for (uint32_t ks = 0; ks < kTcRows / 16; ++ks)
                    umma::mma_f16(tmem + 2 * NN, base + ks * 16u, dD_b + ks * 16u, idesc_b, (e | ks) != 0);
                umma::commit(&mbar[1]);
            }
            __syncwarp();
        };
        for (uint32_t e = 0; e < nit; ++e) {      // (the words of the first super-tile were requested in the prologue)
            umma::mbar_wait(&mbar[2], e & 1u);          // super-tile e expanded by all 128 compute threads (and z0 of e - 1 read)
            umma::fence_after_sync();
            if (umma::elect_one()) {
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h)
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks)
                        if (ks < NKS)
                            umma::mma_f16_ts(tmem + h * NN, tmem + 4 * NN + h * 32 + ks * 8, dW_f + ((ks * 2u * (NN * 16)) >> 4), idesc_f, ks > 0);
                umma::commit(&mbar[0]);
                // every compute thread has read the staged words of super-tile e: request those of e + 1
                if (e + 1 < nit) umma::bulk_load(sG, gwords + (size_t)(t_begin + e + 1) * NC * 128, NC * 512u, &mbar[4]);
            }
            __syncwarp();
            if (bwd && e >= 1) issue_bwd(e - 1);
        }
        if (bwd && nit > 0) issue_bwd(nit - 1);
        // leave together with the compute warps (their barriers count all 160 threads)
        umma::fence_before_sync();
        __syncthreads();
        if (epilogue) { __syncthreads(); __syncthreads(); }
        return;
    }
    // ==================================================================== compute warps
    const uint32_t tlane = tmem + ((warp * 32u) << 16);
    const uint32_t tA = tlane + 4 * NN;      // forward A operand in tensor memory (as k1_tc)
    for (uint32_t c = 0; c < 64; c += 4) umma::tmem_st4(tA + c, 0u, 0u, 0u, 0u);
    umma::tmem_st_wait();

    const f2 zero2 = dup2(0.f);
    typename TT::Acc A;
    A.clear();
    const size_t eoff = a.out_per_entry ? (size_t)li * a.n : 0;
    const size_t toff = (a.target_mode == TGT_PER_ENTRY) ? (size_t)li * a.n : 0;
    const float* tsrc = (!LEAN && a.target_mode == TGT_RESID_PLUS_PRED) ? a.resid : (a.tgt ? a.tgt + toff : nullptr);

    auto expand = [&](uint32_t buf) {
        uint8_t* rowA = sA + buf * sa_bytes + tid * 16;
        uint32_t xw[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) xw[i] = ((uint32_t)i < NC) ? sG[i * 128 + tid] : 0u;
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if ((uint32_t)i < NC) {
                const uint32_t x = xw[i], y = x >> 8;
                const uint4 oa = make_uint4(x & 0x00030003u, x & 0x000C000Cu, x & 0x00300030u, x & 0x00C000C0u);
                const uint4 ob = make_uint4(y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride) = oa;
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride + 128 * 16) = ob;
                umma::tmem_st4(tA + 4 * i, oa.x, oa.y, oa.z, oa.w);
                umma::tmem_st4(tA + 32 + 4 * i, ob.x, ob.y, ob.z, ob.w);
            }
        umma::tmem_st_wait();
        umma::fence_before_sync();
    };
    auto load_targets = [&](uint32_t st) -> f2 {
        const uint32_t rA = st * kTcRows + tid, rB = rA + 128;
        if (!tsrc || st >= t_end) return zero2;
        return mk2(rA < a.n ? __ldg(tsrc + rA) : 0.f, rB < a.n ? __ldg(tsrc + rB) : 0.f);
    };
    f2 tg_next = load_targets(t_begin);
    // DEFER: the previous super-tile's terms.  One `Mid` serves both super-tiles: the fields of the layers >= 1 and the error are
    // read by the deferred sums BEFORE this super-tile's forward pass overwrites them; only the first layer's activations
    // overlap (written while the sums still run) and keep a copy, a0p.  All zero at the start: the first sums add nothing.
    typename TT::Mid M;
    f2 a0p[TT::MW];
    if constexpr (DEFER) {
#pragma unroll
        for (int i = 0; i < TT::MW; ++i) {
            a0p[i] = zero2;
#pragma unroll
            for (int l = 0; l < TT::NLA; ++l) M.act[l][i] = zero2;
#pragma unroll
            for (int l = 0; l < TT::NL1; ++l) M.sgl[l][i] = zero2;
        }
        M.e = zero2;
    }
    if (nit > 0) {
        umma::mbar_wait(&mbar[4], 0);
        expand(0);
        if (!bwd) umma::fence_async_smem();
        umma::mbar_arrive(&mbar[2]);          // phase 0: super-tile 0 expanded
    }
    for (uint32_t it = 0; it < nit; ++it) {
        const uint32_t st = t_begin + it, buf = it & 1u;
        const uint32_t rowA_g = st * kTcRows + tid, rowB_g = rowA_g + 128;
        const bool vA = rowA_g < a.n, vB = rowB_g < a.n;
        f2 tg = tg_next;
        tg_next = load_targets(st + 1);
        // two slices of the deferred sums BEFORE the wait for this super-tile's z0: work instead of spinning when the forward
        // contraction is not done yet (cfg3: 12.472 -> 12.408 ms per launch)
        constexpr int KPRE = DEFER ? (TT::NSLICE < 2 ? TT::NSLICE : 2) : 0;
        if constexpr (KPRE > 0) {
            if (bwd) {
#pragma unroll
                for (int k = 0; k < KPRE; ++k) TT::accumulate_slice(M, a0p, A, k);
            }
        }
        umma::mbar_wait(&mbar[0], it & 1u);
        umma::fence_after_sync();
        float accA[16], accB[16];
        umma::tmem_ld16x2(tlane, tlane + NN, accA, accB);
        umma::fence_before_sync();      // these reads precede the next forward MMA (ordered by the "expanded" barrier below)

        f2 yh, sg0[W0], ef0;
        if constexpr (DEFER) {
            TT::part1x(accA, accB, wp, b0p, tg, !LEAN && a.target_mode == TGT_RESID_PLUS_PRED, mk2(vA ? 1.f : 0.f, vB ? 1.f : 0.f), bwd, M,
                       yh, sg0, ef0, [&](int c) { if (bwd) TT::accumulate_slice(M, a0p, A, KPRE + c); },
                       [&]() {
                           if (bwd) {
#pragma unroll
                               for (int k = KPRE + W0; k < TT::NSLICE; ++k) TT::accumulate_slice(M, a0p, A, k);
                           }
                       });
        } else {
            TT::part1(accA, accB, wp, b0p, tg, !LEAN && a.target_mode == TGT_RESID_PLUS_PRED, mk2(vA ? 1.f : 0.f, vB ? 1.f : 0.f), bwd, A,
                      yh, sg0, ef0);
        }
        if (!LEAN) {
            auto put = [&](float* dst, uint32_t row, float v, int accumulate) {
                if (!dst || row >= a.n) return;
                float* p = dst + eoff + row;
                if (accumulate > 0) *p += v;
                else if (accumulate < 0) *p -= v;
                else *p = v;
            };
            if (a.target_mode == TGT_RESID_PLUS_PRED) {
                put(a.tgt_out, rowA_g, lo2(tg), 0); put(a.tgt_out, rowB_g, hi2(tg), 0);
                put(a.prev_out, rowA_g, lo2(yh), 0); put(a.prev_out, rowB_g, hi2(yh), 0);   // net.rs:279
            }
            put(a.yhat_out, rowA_g, lo2(yh), a.yhat_accumulate);
            put(a.yhat_out, rowB_g, hi2(yh), a.yhat_accumulate);
        }
        // the backward contraction of the previous super-tile read buffer buf^1 and the delta buffer
        if (bwd && it > 0) umma::mbar_wait(&mbar[1], (it - 1) & 1u);
        if (it + 1 < nit) {
            umma::mbar_wait(&mbar[4], (it + 1) & 1u);
            expand(buf ^ 1u);
        }
        if (!bwd) umma::fence_async_smem();
        umma::mbar_arrive(&mbar[2]);          // phase it + 1
        if (!bwd) continue;
        {
            f2 v[W0];
            TT::delta0(sg0, ef0, A, v);
            TT::store_pieces(v, sD + tid * 16);
        }
        umma::fence_async_smem();
        umma::mbar_arrive(&mbar[3]);          // phase it
        if constexpr (DEFER) {
#pragma unroll
            for (int i = 0; i < TT::MW; ++i) a0p[i] = M.act[0][i];
        }
    }
    if constexpr (DEFER) {                     // the last super-tile's sums
        if (bwd) {
#pragma unroll
            for (int k = 0; k < TT::NSLICE; ++k) TT::accumulate_slice(M, a0p, A, k);
        }
    }
    const bool has_bwd = epilogue && nit > 0;
    float sacc[16];
    if (has_bwd) {
        umma::mbar_wait(&mbar[1], (nit - 1) & 1u);
        umma::fence_after_sync();
        umma::tmem_ld16(tlane + 2 * NN, sacc);
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, C::TMEM_COLS);
    if (!epilogue) return;

    float* pp = a.part + ((size_t)li * a.nchunk + chunk) * a.pstride;
    __shared__ float s_gb0[W0];
    TT::template reduce_and_store<true>(A, red, warp, lane, tid, pp, m, d.P, s_gb0);   // recursive halving: ~41 shuffles instead of 205
    if (lane < 16) {
        const uint32_t j = warp * 16 + lane;      // M = 64 accumulator: row j in lane j % 16 of warp j / 16
        if (j < m) {
            const float unscale = pow2f(33 - 2 * (int)((j & 7u) >> 1));
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float s = (nit > 0 ? (sacc[c] + (sacc[W0 + c] + sacc[2 * W0 + c])) : 0.f) * unscale;
                pp[c * m + j] = __fdiv_rn(s - mu[j] * s_gb0[c], sd[j]);
            }
        }
    }
}

}  // namespace bann
