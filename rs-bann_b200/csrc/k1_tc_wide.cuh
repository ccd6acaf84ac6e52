// Tensor-core K1 for branches with 65 .. 512 markers (first-layer width <= 5): the k1_tc scheme, K-blocked.
//
// A branch is cut into blocks of 64 markers (8 chunks of 8).  Per 256-row super-tile the CTA streams the blocks TWICE:
//   forward pass : packed words -> TENSOR MEMORY (tcgen05.st of the AND-expanded words: the A operand of the TS-form MMA, no
//                  shared-memory traffic) -> MMA fwd (M128 N16 K16 x 4, two row halves) accumulating z0 over all blocks;
//   tail         : exactly as in k1_tc (packed f32x2, one row pair per thread), delta_0 pieces -> shared memory;
//   backward pass: packed words -> one of two shared-memory operand buffers -> MMA bwd (M64 N16 K16 x 16) into the block's own
//                  16 accumulator columns, which keep accumulating over ALL super-tiles of the CTA.
// The packed words of a block (8 chunks x 128 row pairs x 4 B = 4 KB, contiguous in the tensor-core store) arrive by bulk
// async copies into a 4-slot ring, requested 4 blocks ahead; the second pass re-reads them (from L2).  Every stage is
// decoupled by mbarriers; warps 0-3 expand / run the tail, warp 4 requests the loads and issues every MMA.
// Shared memory: 2 x 32 KB operands + 16 KB ring + 8 KB delta pieces + 16 KB weight pieces -> 2 CTAs per SM;
// tensor memory: 32 + 16 x ceil(m/64) accumulator columns + 64 forward-operand columns <= 224 (256 allocated).
#pragma once
#include "k1_tc.cuh"

namespace bann {

constexpr int kTcwMaxMarkers = 512;
constexpr int kTcwBlockChunks = 8;                       // chunks (of 8 markers) per block = one M = 64 backward tile
constexpr uint32_t kTcwBlockBytes = kTcwBlockChunks * kTcChunkStride;   // 32 KB: one expanded block
constexpr uint32_t kTcwRing = 4;                         // packed-word ring slots (4 KB each)

template <int H, int S, int D>
struct TcwShape {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    static constexpr int NN = C::NN;
    static constexpr int TMEM_COLS = 256;
    static constexpr size_t SW = (size_t)(kTcwMaxMarkers / 8) * NN * 16;       // weight pieces of all 64 k-chunks
    static constexpr size_t SMEM = 2 * (size_t)kTcwBlockBytes + kTcwRing * (size_t)kTcwBlockChunks * 512 + C::SD + SW + C::MISC + 128 + 128;
};

constexpr int kTcwThreads = 160;   // warps 0-3: expansion, tail, epilogue; warp 4: ring loads + MMA issue (measured on cfg2: +50 %)

template <int H, int S, int D, int ACT, bool LEAN>
__global__ void __launch_bounds__(kTcwThreads, 2) k1_tcw(K1Args a) {
    using T = TailShape<H, S, D>;
    using C = TcShape<H, S, D>;
    using CW = TcwShape<H, S, D>;
    using TT = TcTail<H, S, D, ACT>;
    constexpr int W0 = T::W0, W0P = T::W0P, NN = C::NN;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    const BranchDesc& d = a.descs[b];        // descriptors and branch lists are written once, at net creation
    const uint32_t m = d.m, NC = d.nc, NKB = (NC + kTcwBlockChunks - 1) / kTcwBlockChunks;
    const uint32_t issuer = 4;
    // ---- shared memory carve-up
    uint8_t* sA = smraw + ((128u - (umma::smem_u32(smraw) & 127u)) & 127u);       // 2 operand buffers of 8 chunks
    uint32_t* sG = reinterpret_cast<uint32_t*>(sA + 2 * kTcwBlockBytes);           // ring: [slot][chunk][128] packed words
    uint8_t* sD = reinterpret_cast<uint8_t*>(sG) + kTcwRing * kTcwBlockChunks * 512;
    uint8_t* sW = sD + C::SD;                                                     // [k-chunk][n] x 16 B, all blocks
    float* wp = reinterpret_cast<float*>(sW + CW::SW);              // tail parameters, plain floats (FFMA2 broadcasts a scalar register)
    float* b0p = wp + 2 * ((T::n_tail() + 3) & ~3);                  // (layout size as before)
    float* red = b0p + 2 * W0P;
    // mbarriers: [0..1] backward MMAs that read operand buffer 0/1 done; [2..3] buffer 0/1 expanded (128 arrivals);
    //            [4..7] ring slot 0..3 landed; [8] forward MMAs of a block done; [9] forward block expanded (128 arrivals)
    uint64_t* mbar = reinterpret_cast<uint64_t*>(red + C::NRED);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 10);

    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;

    // ---- one-time setup
    {
        const uint32_t nz = (uint32_t)((sW + CW::SW - sA) / 16);
        for (uint32_t k = tid; k < nz; k += kTcwThreads) reinterpret_cast<uint4*>(sA)[k] = make_uint4(0, 0, 0, 0);
    }
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 1); umma::mbar_init(&mbar[1], 1);
        umma::mbar_init(&mbar[2], 128); umma::mbar_init(&mbar[3], 128);
        for (int k = 0; k < (int)kTcwRing; ++k) umma::mbar_init(&mbar[4 + k], 1);
        umma::mbar_init(&mbar[8], 1); umma::mbar_init(&mbar[9], 128);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, CW::TMEM_COLS);
    // programmatic dependent launch (common.cuh): the shared / tensor memory prologue above ran under the previous kernel
    pdl_launch_dependents();
    pdl_wait();
    if (a.states && a.states[b].status != ST_RUNNING) {      // early-rejected / finished branch: release the tensor memory and leave
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (warp == 0) umma::tmem_dealloc(*tmem_slot, CW::TMEM_COLS);
        return;
    }
    TT::stage_tail(th + m * W0, wp, tid, kTcwThreads);
    __syncthreads();
    // first-layer weights: W' = W0 / sd, its three bf16 pieces (scaled as in k1_tc), and the mean-folded bias
    float bacc[W0];
#pragma unroll
    for (int c = 0; c < W0; ++c) bacc[c] = 0.f;
    for (uint32_t j = tid; j < m && tid < 128; j += 128) {
#pragma unroll
        for (int c = 0; c < W0; ++c) {
            const float w = __fdiv_rn(th[c * m + j], sd[j]);
            bacc[c] = fmaf(mu[j], w, bacc[c]);
            const float v = w * pow2f(100 - 2 * (int)((j & 7u) >> 1));
            const float p0 = bf16_round(v), r1 = v - p0, p1 = bf16_round(r1), p2 = bf16_round(r1 - p1);
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(sW + (j >> 3) * (NN * 16) + (j & 7u) * 2);
            dst[(0 * W0 + c) * 8] = __float2bfloat16_rn(p0);
            dst[(1 * W0 + c) * 8] = __float2bfloat16_rn(p1);
            dst[(2 * W0 + c) * 8] = __float2bfloat16_rn(p2);
        }
    }
    // fixed-order block sum of the bias fold (lanes -> warps), result b0' = b0 - sum_j mu_j W'_j
#pragma unroll
    for (int c = 0; c < W0; ++c) {
        float v = bacc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0 && warp < 4) red[warp * W0 + c] = v;
    }
    __syncthreads();
    if (tid < W0P) {
        float acc = 0.f;
        if (tid < W0) acc = (th[m * W0 + T::b_off(0) + tid] - (red[tid] + red[W0 + tid] + red[2 * W0 + tid] + red[3 * W0 + tid])) * TT::cA;
        b0p[tid] = acc;
    }
    umma::fence_async_smem();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);
    const uint32_t sA_u = umma::smem_u32(sA), sD_u = umma::smem_u32(sD), sW_u = umma::smem_u32(sW);
    constexpr uint32_t idesc_f = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 0, 0, 128, NN);
    constexpr uint32_t idesc_b = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 1, 1, 64, NN);
    const uint64_t dW_f = umma::make_desc(sW_u, NN * 16, 128);
    const uint64_t dA_b = umma::make_desc(sA_u, 128, kTcChunkStride), dD_b = umma::make_desc(sD_u, 128, kTcChunkStride);

    // ---- persistent per-thread accumulators (as in k1_tc)
    const f2 zero2 = dup2(0.f);
    typename TT::Acc A;
    A.clear();

    const size_t eoff = a.out_per_entry ? (size_t)li * a.n : 0;
    const size_t toff = (a.target_mode == TGT_PER_ENTRY) ? (size_t)li * a.n : 0;
    const uint32_t t_begin = chunk * a.st_per_chunk;
    const uint32_t t_end = min(a.nst, t_begin + a.st_per_chunk);
    const uint32_t nit = t_end > t_begin ? t_end - t_begin : 0;
    const float* tsrc = (!LEAN && a.target_mode == TGT_RESID_PLUS_PRED) ? a.resid : (a.tgt ? a.tgt + toff : nullptr);
    const bool bwd = LEAN || !a.fwd_only;
    const uint32_t npass = bwd ? 2u : 1u;                       // block passes per super-tile
    const uint32_t bpt = npass * NKB;                           // blocks per super-tile in the stream
    const uint32_t nblk = nit * bpt;                            // length of the block stream of this CTA
    const uint32_t* gwords = a.store_tc + (d.tc_off >> 2);

    // The block stream: position q = (super-tile it, pass, marker block kb); buffer q % 2, ring slot q % 4.  Positions are
    // advanced incrementally (no integer divisions in the loop).
    struct Pos { uint32_t it, pass, kb; };
    auto advance = [&](Pos& p) {
        if (++p.kb == NKB) { p.kb = 0; if (++p.pass == npass) { p.pass = 0; ++p.it; } }
    };
    auto chunks_of = [&](uint32_t kb) { return min((uint32_t)kTcwBlockChunks, NC - kb * kTcwBlockChunks); };
    auto issue_load = [&](uint32_t q, const Pos& p) {   // whole issuer warp enters; words of stream block q -> ring slot q % 4
        if (umma::elect_one())
            umma::bulk_load(sG + (q % kTcwRing) * (kTcwBlockChunks * 128),
                            gwords + ((size_t)(t_begin + p.it) * NC + p.kb * kTcwBlockChunks) * 128, chunks_of(p.kb) * 512u,
                            &mbar[4 + (q % kTcwRing)]);
        __syncwarp();
    };
    // Forward operand in TENSOR MEMORY (as in k1_tc / k1_tcx): one buffer of 2 x 32 columns behind the accumulators (32 + 16 x 8
    // columns), written by tcgen05.st -- the forward pass costs no shared-memory traffic at all (the kernel was shared-memory
    // bandwidth bound: every block was written and read twice); single-buffered, the second resident CTA covers the MMA latency.
    // Backward blocks go through the two shared-memory operand buffers as before.  qf / qb count forward / backward blocks.
    constexpr uint32_t tA = 160;
    auto issue_fwd = [&](uint32_t qf, uint32_t kb) {          // whole issuer warp enters
        umma::mbar_wait(&mbar[9], qf & 1u);
        umma::fence_after_sync();
        if (umma::elect_one()) {
            const uint32_t nks = (chunks_of(kb) + 1) >> 1;
            const uint64_t wbase = dW_f + ((kb * kTcwBlockChunks * (NN * 16)) >> 4);
#pragma unroll
            for (uint32_t h = 0; h < 2; ++h)
#pragma unroll
                for (uint32_t ks = 0; ks < 4; ++ks)
                    if (ks < nks)
                        umma::mma_f16_ts(tmem + h * NN, tmem + tA + h * 32u + ks * 8u, wbase + ((ks * 2u * (NN * 16)) >> 4), idesc_f,
                                         (kb | ks) != 0);
            umma::commit(&mbar[8]);
        }
        __syncwarp();
    };
    auto issue_bwd = [&](uint32_t qb, uint32_t it, uint32_t kb) {   // whole issuer warp enters
        const uint32_t buf = qb & 1u;
        umma::mbar_wait(&mbar[2 + buf], (qb >> 1) & 1u);
        umma::fence_after_sync();
        if (umma::elect_one()) {
            const uint64_t base = dA_b + ((buf * kTcwBlockBytes) >> 4);
#pragma unroll
            for (uint32_t ks = 0; ks < kTcRows / 16; ++ks)
                umma::mma_f16(tmem + (2 + kb) * NN, base + ks * 16u, dD_b + ks * 16u, idesc_b, (it | ks) != 0);
            umma::commit(&mbar[buf]);
        }
        __syncwarp();
    };
    Pos pos{0, 0, 0};           // position of stream block q
    Pos lpos{0, 0, 0};          // position of the next block whose words get requested (issuer warp)
    uint32_t lq = 0;
    // forward block: packed words of ring slot q % 4 -> tensor memory (waits: the words, the previous forward block's MMAs)
    auto expand_fwd = [&](uint32_t q, uint32_t qf, uint32_t kb) {
        const uint32_t nch = chunks_of(kb);
        umma::mbar_wait(&mbar[4 + (q % kTcwRing)], (q / kTcwRing) & 1u);
        const uint32_t* src = sG + (q % kTcwRing) * (kTcwBlockChunks * 128) + tid;
        uint32_t x[kTcwBlockChunks];
#pragma unroll
        for (int i = 0; i < kTcwBlockChunks; ++i) x[i] = (uint32_t)i < nch ? src[i * 128] : 0u;   // all loads first
        if (qf >= 1) umma::mbar_wait(&mbar[8], (qf - 1) & 1u);
        umma::fence_after_sync();
        const uint32_t ta = tlane + tA;
#pragma unroll
        for (int i = 0; i < kTcwBlockChunks; ++i) {
            const uint32_t y = x[i] >> 8;
            umma::tmem_st4(ta + 4 * i, x[i] & 0x00030003u, x[i] & 0x000C000Cu, x[i] & 0x00300030u, x[i] & 0x00C000C0u);
            umma::tmem_st4(ta + 32 + 4 * i, y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
        }
        umma::tmem_st_wait();
        umma::fence_before_sync();
        umma::mbar_arrive(&mbar[9]);
    };
    // backward block: packed words -> operand buffer qb % 2 in shared memory (waits: the words, the MMAs that last read the buffer)
    auto expand_bwd = [&](uint32_t q, uint32_t qb, uint32_t kb) {
        const uint32_t buf = qb & 1u, nch = chunks_of(kb);
        umma::mbar_wait(&mbar[4 + (q % kTcwRing)], (q / kTcwRing) & 1u);
        const uint32_t* src = sG + (q % kTcwRing) * (kTcwBlockChunks * 128) + tid;
        uint32_t x[kTcwBlockChunks];
#pragma unroll
        for (int i = 0; i < kTcwBlockChunks; ++i) x[i] = (uint32_t)i < nch ? src[i * 128] : 0u;   // all loads first
        if (qb >= 2) umma::mbar_wait(&mbar[buf], ((qb >> 1) - 1) & 1u);
        uint8_t* rowA = sA + buf * kTcwBlockBytes + tid * 16;
#pragma unroll
        for (int i = 0; i < kTcwBlockChunks; ++i)
            if ((uint32_t)i < nch) {
                const uint32_t y = x[i] >> 8;
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride) =
                    make_uint4(x[i] & 0x00030003u, x[i] & 0x000C000Cu, x[i] & 0x00300030u, x[i] & 0x00C000C0u);
                *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride + 128 * 16) =
                    make_uint4(y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
            }
        umma::fence_async_smem();
        umma::mbar_arrive(&mbar[2 + buf]);
    };

    auto load_targets = [&](uint32_t st) -> f2 {
        const uint32_t rA = st * kTcRows + tid, rB = rA + 128;
        if (!tsrc || st >= t_end) return zero2;
        return mk2(rA < a.n ? __ldg(tsrc + rA) : 0.f, rB < a.n ? __ldg(tsrc + rB) : 0.f);
    };
    f2 tg_next = load_targets(t_begin);
    if (warp == issuer)
        for (; lq < kTcwRing && lq < nblk; ++lq) { issue_load(lq, lpos); advance(lpos); }

    if (warp == issuer) {
        // ---- issuing warp: the whole block stream in order -- wait "expanded" (128 arrivals: for the first backward block this also
        // publishes the delta pieces), issue the MMAs, recycle the ring slot every thread has consumed
        uint32_t qf = 0, qb = 0;
        for (uint32_t qq = 0; qq < nblk; ++qq) {
            if (pos.pass == 0) issue_fwd(qf++, pos.kb);
            else issue_bwd(qb++, pos.it, pos.kb);
            if (lq < nblk) { issue_load(lq, lpos); ++lq; advance(lpos); }
            advance(pos);
        }
        if (qf >= 1) umma::mbar_wait(&mbar[8], (qf - 1) & 1u);
        if (qb >= 1) umma::mbar_wait(&mbar[(qb - 1) & 1u], ((qb - 1) >> 1) & 1u);
        if (qb >= 2) umma::mbar_wait(&mbar[(qb - 2) & 1u], ((qb - 2) >> 1) & 1u);
        umma::fence_before_sync();
        __syncthreads();                       // the compute warps' exits below: one barrier without an epilogue, three with it
        if (bwd && a.part) { __syncthreads(); __syncthreads(); }
        return;
    }
    uint32_t q = 0, qf = 0, qb = 0;
    for (uint32_t it = 0; it < nit; ++it) {
        const uint32_t st = t_begin + it;
        const uint32_t rowA_g = st * kTcRows + tid, rowB_g = rowA_g + 128;
        const bool vA = rowA_g < a.n, vB = rowB_g < a.n;
        f2 tg = tg_next;
        tg_next = load_targets(st + 1);
        // ---- forward pass over the marker blocks
        for (uint32_t kb = 0; kb < NKB; ++kb, ++q, ++qf) expand_fwd(q, qf, kb);
        // z0 is complete when the MMAs of the last forward block are (commits complete in issue order)
        umma::mbar_wait(&mbar[8], (qf - 1) & 1u);
        umma::fence_after_sync();
        float accA[16], accB[16];
        umma::tmem_ld16x2(tlane, tlane + NN, accA, accB);
        umma::fence_before_sync();

        // ---- tail (TcTail, k1_tc.cuh): remaining layers, error, rho_l, cross-row sums of the layers >= 1
        f2 yh, sg0[W0], ef0;
        TT::part1(accA, accB, wp, b0p, tg, !LEAN && a.target_mode == TGT_RESID_PLUS_PRED, mk2(vA ? 1.f : 0.f, vB ? 1.f : 0.f), bwd, A,
                  yh, sg0, ef0);
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
        if (!bwd) continue;

        // delta_0 pieces.  The delta buffer was last read by the previous super-tile's backward MMAs: wait for its last two
        // commits (both operand buffers), long complete after a forward pass and a tail.
        if (qb >= 1) umma::mbar_wait(&mbar[(qb - 1) & 1u], ((qb - 1) >> 1) & 1u);
        if (qb >= 2) umma::mbar_wait(&mbar[(qb - 2) & 1u], ((qb - 2) >> 1) & 1u);
        {
            f2 v[W0];
            TT::delta0(sg0, ef0, A, v);
            TT::store_pieces(v, sD + tid * 16);
        }
        // ---- backward pass over the marker blocks (the first arrival below also publishes the delta pieces)
        for (uint32_t kb = 0; kb < NKB; ++kb, ++q, ++qb) expand_bwd(q, qb, kb);
    }
    // ---- drain: all MMAs done (the forward ones were waited for above; the last two backward blocks cover both operand buffers)
    if (qb >= 1) umma::mbar_wait(&mbar[(qb - 1) & 1u], ((qb - 1) >> 1) & 1u);
    if (qb >= 2) umma::mbar_wait(&mbar[(qb - 2) & 1u], ((qb - 2) >> 1) & 1u);
    umma::fence_after_sync();
    const bool has_bwd = bwd && a.part && nit > 0;
    if (!bwd || !a.part) {
        umma::fence_before_sync();
        __syncthreads();
        if (warp == 0) umma::tmem_dealloc(tmem, CW::TMEM_COLS);
        return;
    }

    // ---- CTA epilogue: cross-row sums of the layers >= 1 (fixed order), then the first-layer gradient per marker block
    float* pp = a.part + ((size_t)li * a.nchunk + chunk) * a.pstride;
    __shared__ float s_gb0[W0];
    TT::reduce_and_store(A, red, warp, lane, tid, pp, m, d.P, s_gb0);
    for (uint32_t kb = 0; kb < NKB; ++kb) {
        float sacc[16];
        if (has_bwd) umma::tmem_ld16(tlane + (2 + kb) * NN, sacc);
        if (lane < 16) {
            const uint32_t j = kb * 64 + warp * 16 + lane;      // M = 64 accumulator: row r in lane r % 16 of warp r / 16
            if (j < m) {
                const float unscale = pow2f(33 - 2 * (int)((j & 7u) >> 1));
#pragma unroll
                for (int c = 0; c < W0; ++c) {
                    const float s = (has_bwd ? (sacc[c] + (sacc[W0 + c] + sacc[2 * W0 + c])) : 0.f) * unscale;
                    pp[c * m + j] = __fdiv_rn(s - mu[j] * s_gb0[c], sd[j]);
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, CW::TMEM_COLS);
}

template <int H, int S, int D, int ACT>
int launch_one_tcw_act(K1Args& a, uint32_t nlist, cudaStream_t st) {
    using CW = TcwShape<H, S, D>;
    static bool configured = false;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(k1_tcw<H, S, D, ACT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW::SMEM));
        BANN_CUDA(cudaFuncSetAttribute(k1_tcw<H, S, D, ACT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CW::SMEM));
        configured = true;
    }
    dim3 grid(a.nchunk, nlist);
    const bool lean = !a.fwd_only && !a.yhat_out && a.target_mode != TGT_RESID_PLUS_PRED && a.tgt;
    if (lean) BANN_CUDA(launch_pdl(k1_tcw<H, S, D, ACT, true>, grid, dim3(kTcwThreads), CW::SMEM, st, a));
    else BANN_CUDA(launch_pdl(k1_tcw<H, S, D, ACT, false>, grid, dim3(kTcwThreads), CW::SMEM, st, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
template <int H, int S, int D>
int launch_one_tcw(K1Args& a, uint32_t nlist, cudaStream_t st) {
    switch (a.act) {
        case BANN_TANH: return launch_one_tcw_act<H, S, D, BANN_TANH>(a, nlist, st);
        case BANN_RELU: return launch_one_tcw_act<H, S, D, BANN_RELU>(a, nlist, st);
        case BANN_LEAKY_RELU: return launch_one_tcw_act<H, S, D, BANN_LEAKY_RELU>(a, nlist, st);
        case BANN_SILU: return launch_one_tcw_act<H, S, D, BANN_SILU>(a, nlist, st);
        default: return launch_one_tcw_act<H, S, D, BANN_IDENTITY>(a, nlist, st);
    }
}

// K-blocked tensor-core K1: homogeneous architecture, any of the five activations, every listed branch with 65..512 markers (so that each has at
// least two marker blocks), 3 * W0 <= 16, tensor-core store present.
#ifdef BANN_K1_TCW_IMPL
int launch_k1_tcw(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                         cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net) {
    *launched = false;
    if (!a.store_tc) return 0;
    const BranchDesc& d0 = descs[single_branch >= 0 ? single_branch : 0];
    uint32_t max_m = d0.m, min_m = d0.m;
    if (single_branch < 0) {
        for (const BranchDesc& d : descs) {
            if (d.nl != d0.nl) return 0;
            for (uint32_t l = 0; l < d.nl; ++l)
                if (d.widths[l] != d0.widths[l]) return 0;
            max_m = std::max(max_m, d.m);
            min_m = std::min(min_m, d.m);
        }
    }
    if (max_m > (uint32_t)kTcwMaxMarkers || min_m <= (uint32_t)kTcMaxMarkers) return 0;
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    const uint32_t nst = a.nst;
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)num_sms * 2 + nlist - 1) / nlist);
    uint32_t nchunk = std::min<uint32_t>(want, std::max<uint32_t>(1, nst));
    uint32_t spc = (nst + nchunk - 1) / nchunk;
    nchunk = (nst + spc - 1) / spc;
#define BANN_TRY_TCW(HH, SS, DD)                                                                          \
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
        int rc = launch_one_tcw<HH, SS, DD>(a, nlist, st);                                                \
        if (rc) return rc;                                                                                \
        *launched = true;                                                                                 \
    }
    BANN_TRY_TCW(5, 5, 1)
    BANN_TRY_TCW(2, 2, 1)
    BANN_TRY_TCW(4, 3, 1)
    BANN_TRY_TCW(3, 3, 1)
    BANN_TRY_TCW(4, 4, 1)
    BANN_TRY_TCW(5, 5, 2)
#undef BANN_TRY_TCW
    return 0;
}
#else
int launch_k1_tcw(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                         cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net);
#endif

}  // namespace bann
