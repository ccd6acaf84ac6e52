// K1 for WIDE branches (BASELINE.json configs[3]: 1000 markers, widths [16,16,16,1]): first-layer width up to 16, any number of
// markers the tensor-core store holds, hidden / summary widths up to 16, tanh.
//
// With 16 first-layer units the split operands are 48 columns wide and a 1000-marker branch has 48 000 backward
// accumulators: they do not fit next to the forward accumulators and a per-row tail in one CTA's tensor memory / registers, so
// the fused kernel of k1_tc / k1_tcw is cut into three passes over a (branch, row chunk):
//   KA  k_tcx_fwd  : a_0 = tanh(X W' + b')          tcgen05, all marker blocks streamed once, accumulators 2 x 48 columns
//   KT  k_tcx_tail : layers >= 1, error, deltas, cross-row sums of the layers >= 1, delta_0 as bf16 pieces   (FP32)
//   KB  k_tcx_bwd  : S = X^T delta_0                 tcgen05 (M = 128: two marker blocks per MMA), one CTA per 512-marker slab
// Same operands as k1_tc: genotypes expanded to bf16 subnormals by one AND per two elements, W' and delta_0 as three bf16 pieces
// whose sum is the f32 value (exact products, f32 accumulation).  The price of the cut is traffic: the packed genotypes are read
// twice and a_0 (64 B per row) / the delta pieces (96 B per row) travel through HBM / L2 once.
#pragma once
#include "k1_tc_wide.cuh"

namespace bann {

constexpr int kTcxMaxMarkers = 2048;       // tensor-core store limit (genotypes.cu)
constexpr int kTcxSlabPairs = 4;           // pairs of marker blocks (128 markers each) per KB CTA: 4 x NN accumulator columns

template <int W0>
struct TcxShape {
    static constexpr int NN = ((3 * W0 + 15) / 16) * 16;                 // accumulator columns: 3 pieces x W0 units, padded
    static constexpr int NQ = NN / 8;                                    // 16-byte n-chunks of the delta operand
    static constexpr uint32_t WBLK = kTcwBlockChunks * NN * 16;          // bytes of W' pieces per marker block
    static constexpr uint32_t SLOT = kTcwBlockChunks * 512 + WBLK;       // ring slot: packed words + W' pieces of one block
    static constexpr uint32_t DP_ST = NQ * kTcChunkStride;               // delta pieces of one super-tile
    static constexpr uint32_t RING_A = 8;                                // KA ring slots (its operand buffers are in tensor memory)
    static constexpr size_t SMEM_A = RING_A * (size_t)SLOT + 256 + 128;
    static constexpr size_t SMEM_B = 2 * (size_t)kTcwBlockBytes + 2 * (size_t)(2 * kTcwBlockChunks) * 512 + DP_ST + 512 + 128;
    static constexpr int TMEM_A = 256;                                   // 2 x NN accumulators + 2 x 2 x 32 operand columns <= 224
    static constexpr int TMEM_B = 256;                                   // kTcxSlabPairs x NN <= 192
};

namespace umma {
__device__ __forceinline__ void expect_tx(uint64_t* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
}  // namespace umma

struct TcxArgs {
    K1Args k;
    uint8_t* wp;          // [entry][wp_stride] W' pieces, block-major [NKB][8 k-chunks][NN][8] bf16, then b0' (W0 floats)
    size_t wp_stride;     // bytes per entry
    float* a0;            // [entry][a0_stride] first-layer activations, row-major [row][W0]
    size_t a0_stride;     // floats per entry
    uint8_t* dp;          // [entry][dp_stride] delta_0 pieces, per super-tile [NQ][256 rows] x 16 B
    size_t dp_stride;     // bytes per entry
    uint32_t nkb_max;     // marker blocks of the widest listed branch
};

// ---- prep: W' = W0 / sd as three bf16 pieces in the forward B-operand layout, b0' = b0 - sum_j mu_j W'_j
template <int W0>
__global__ void __launch_bounds__(128) k_tcx_prep(TcxArgs a) {
    using X = TcxShape<W0>;
    constexpr int NN = X::NN;
    __shared__ float red[4 * W0];
    const uint32_t li = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t b = a.k.list ? a.k.list[li] : li;
    if (a.k.states && a.k.states[b].status != ST_RUNNING) return;
    const BranchDesc& d = a.k.descs[b];
    const uint32_t m = d.m, NKB = (d.nc + kTcwBlockChunks - 1) / kTcwBlockChunks;
    const float* th = a.k.theta + d.param_off;
    const float* mu = a.k.mu + d.col_off;
    const float* sd = a.k.sd + d.col_off;
    uint8_t* out = a.wp + (size_t)li * a.wp_stride;
    float bacc[W0];
#pragma unroll
    for (int c = 0; c < W0; ++c) bacc[c] = 0.f;
    for (uint32_t j = tid; j < NKB * 64; j += 128) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(out + (size_t)(j >> 3) * (NN * 16) + (j & 7u) * 2);
#pragma unroll
        for (int n = 3 * W0; n < NN; ++n) dst[n * 8] = __float2bfloat16_rn(0.f);
#pragma unroll
        for (int c = 0; c < W0; ++c) {
            float p0 = 0.f, p1 = 0.f, p2 = 0.f;
            if (j < m) {
                const float w = __fdiv_rn(th[c * m + j], sd[j]);          // bed.rs:354 folded into the first layer
                bacc[c] = fmaf(mu[j], w, bacc[c]);
                const float v = w * pow2f(100 - 2 * (int)((j & 7u) >> 1));
                p0 = bf16_round(v);
                const float r1 = v - p0;
                p1 = bf16_round(r1);
                p2 = bf16_round(r1 - p1);
            }
            dst[(0 * W0 + c) * 8] = __float2bfloat16_rn(p0);
            dst[(1 * W0 + c) * 8] = __float2bfloat16_rn(p1);
            dst[(2 * W0 + c) * 8] = __float2bfloat16_rn(p2);
        }
    }
#pragma unroll
    for (int c = 0; c < W0; ++c) {
        float v = bacc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp * W0 + c] = v;
    }
    __syncthreads();
    if (tid < W0) {
        float* b0p = reinterpret_cast<float*>(out + (size_t)a.nkb_max * X::WBLK);
        b0p[tid] = th[d.b_off[0] + tid] - (red[tid] + red[W0 + tid] + red[2 * W0 + tid] + red[3 * W0 + tid]);
    }
}

// expand one block of packed words (ring slot) into an operand buffer: [chunk][row] x 16 B, one AND per two elements
__device__ __forceinline__ void tcx_expand(const uint32_t* src, uint8_t* rowA, uint32_t nch) {
    uint32_t x[kTcwBlockChunks];
#pragma unroll
    for (int i = 0; i < kTcwBlockChunks; ++i) x[i] = (uint32_t)i < nch ? src[i * 128] : 0u;
#pragma unroll
    for (int i = 0; i < kTcwBlockChunks; ++i) {
        const uint32_t y = x[i] >> 8;    // chunks beyond nch are written as zeros: the M = 64 operand always reads 8 chunks
        *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride) =
            make_uint4(x[i] & 0x00030003u, x[i] & 0x000C000Cu, x[i] & 0x00300030u, x[i] & 0x00C000C0u);
        *reinterpret_cast<uint4*>(rowA + i * kTcChunkStride + 128 * 16) =
            make_uint4(y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
    }
}

// ------------------------------------------------------------------ KA: a_0 = tanh(X W' + b')
// The forward A operand lives in TENSOR MEMORY (lanes = rows, 32-bit columns = marker pairs, as in k1_tc): the expansion is a
// tcgen05.st per 8 markers and row, no shared-memory operand image, no proxy fence.  Shared memory only holds the ring of
// (packed words | W' pieces) per marker block, 8 slots deep; a slot is reloaded once the MMAs that read its W' pieces are done.
// Warps 0-3 expand (one row pair per thread) and run the epilogue; warp 4 only requests the ring loads and issues the MMAs, so that
// the ~40 clk per tcgen05.mma of the issuing lane never sits on the expansion's critical path.
// ACT: the value KT needs per first-layer unit goes to HBM -- the activation, from which the rectifiers', tanh's and the identity's
// derivatives follow; for SiLU (derivative needs the sigmoid as well) the PRE-activation, and KT evaluates the activation itself.
template <int W0, int ACT>
__global__ void __launch_bounds__(160, 2) k_tcx_fwd(TcxArgs a) {
    using X = TcxShape<W0>;
    constexpr int NN = X::NN;
    constexpr uint32_t RING = X::RING_A;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.k.list ? a.k.list[li] : li;
    const BranchDesc& d = a.k.descs[b];
    const uint32_t NC = d.nc, NKB = (NC + kTcwBlockChunks - 1) / kTcwBlockChunks;
    uint8_t* sR = smraw + ((128u - (umma::smem_u32(smraw) & 127u)) & 127u);   // ring: [slot][words 4 KB | W' pieces]
    float* b0s = reinterpret_cast<float*>(sR + RING * X::SLOT);             // [W0]
    // mbarriers: [0..1] MMAs that read operand buffer 0/1 (and their ring slot) done; [2..3] buffer expanded (128 arrivals);
    //            [4..4+RING) ring slot landed
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b0s + 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 4 + RING);
    const uint8_t* wp_g = a.wp + (size_t)li * a.wp_stride;
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 1); umma::mbar_init(&mbar[1], 1);
        umma::mbar_init(&mbar[2], 128); umma::mbar_init(&mbar[3], 128);
        for (int k = 0; k < (int)RING; ++k) umma::mbar_init(&mbar[4 + k], 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, X::TMEM_A);
    pdl_launch_dependents();       // programmatic dependent launch (common.cuh): the prologue above ran under k_tcx_prep
    pdl_wait();
    if (a.k.states && a.k.states[b].status != ST_RUNNING) {
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (warp == 0) umma::tmem_dealloc(*tmem_slot, X::TMEM_A);
        return;
    }
    if (tid < W0) b0s[tid] = reinterpret_cast<const float*>(wp_g + (size_t)a.nkb_max * X::WBLK)[tid];
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);
    // tensor memory: accumulators [0, 2 NN); operand buffer p, row half h at columns 2 NN + 64 p + 32 h (4 columns per chunk)
    const uint32_t tA = 2 * NN;
    const uint32_t sR_u = umma::smem_u32(sR);
    constexpr uint32_t idesc_f = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 0, 0, 128, NN);

    const uint32_t t_begin = chunk * a.k.st_per_chunk;
    const uint32_t t_end = min(a.k.nst, t_begin + a.k.st_per_chunk);
    const uint32_t nit = t_end > t_begin ? t_end - t_begin : 0;
    const uint32_t nblk = nit * NKB;
    const uint32_t* gwords = a.k.store_tc + (d.tc_off >> 2);
    auto chunks_of = [&](uint32_t kb) { return min((uint32_t)kTcwBlockChunks, NC - kb * kTcwBlockChunks); };
    auto issue_load = [&](uint32_t q, uint32_t it, uint32_t kb) {   // whole issuer warp enters
        if (umma::elect_one()) {
            uint8_t* slot = sR + (q % RING) * X::SLOT;
            const uint32_t wb = chunks_of(kb) * 512u;
            umma::expect_tx(&mbar[4 + (q % RING)], wb + X::WBLK);
            umma::bulk_copy(slot, gwords + ((size_t)(t_begin + it) * NC + kb * kTcwBlockChunks) * 128, wb, &mbar[4 + (q % RING)]);
            umma::bulk_copy(slot + kTcwBlockChunks * 512, wp_g + (size_t)kb * X::WBLK, X::WBLK, &mbar[4 + (q % RING)]);
        }
        __syncwarp();
    };
    // stream position of the next load (issuer warp): block q + RING - 2 is requested while block q is expanded
    uint32_t l_it = 0, l_kb = 0, lq = 0;
    auto advance_l = [&]() { if (++l_kb == NKB) { l_kb = 0; ++l_it; } ++lq; };
    float* a0_g = a.a0 + (size_t)li * a.a0_stride;
    if (warp == 4) {                 // ---- issuer warp
        for (; lq < RING - 2 && lq < nblk;) { issue_load(lq, l_it, l_kb); advance_l(); }
        uint32_t kb = 0;
        for (uint32_t q = 0; q < nblk; ++q) {
            const uint32_t buf = q & 1u, nch = chunks_of(kb);
            if (q >= 2) umma::mbar_wait(&mbar[buf], ((q >> 1) - 1) & 1u);     // MMAs of block q - 2 done: their ring slot is free
            if (lq < nblk) { issue_load(lq, l_it, l_kb); advance_l(); }
            umma::mbar_wait(&mbar[2 + buf], (q >> 1) & 1u);                   // block q expanded by all 128 threads
            umma::fence_after_sync();
            if (umma::elect_one()) {
                const uint32_t nks = (nch + 1) >> 1;
                const uint64_t wbase = umma::make_desc(sR_u + (q % RING) * X::SLOT + kTcwBlockChunks * 512, NN * 16, 128);
#pragma unroll
                for (uint32_t h = 0; h < 2; ++h)
#pragma unroll
                    for (uint32_t ks = 0; ks < 4; ++ks)
                        if (ks < nks)
                            umma::mma_f16_ts(tmem + h * NN, tmem + tA + buf * 64u + h * 32u + ks * 8u,
                                             wbase + ((ks * 2u * (NN * 16)) >> 4), idesc_f, (kb | ks) != 0);
                umma::commit(&mbar[buf]);
            }
            __syncwarp();
            if (++kb == NKB) kb = 0;
        }
        if (nblk >= 1) umma::mbar_wait(&mbar[(nblk - 1) & 1u], ((nblk - 1) >> 1) & 1u);
        if (nblk >= 2) umma::mbar_wait(&mbar[(nblk - 2) & 1u], ((nblk - 2) >> 1) & 1u);
        umma::fence_before_sync();
        __syncthreads();
        return;
    }

    uint32_t q = 0;
    for (uint32_t it = 0; it < nit; ++it) {
        const uint32_t st = t_begin + it;
        for (uint32_t kb = 0; kb < NKB; ++kb, ++q) {
            const uint32_t buf = q & 1u, nch = chunks_of(kb);
            if (q >= 2) umma::mbar_wait(&mbar[buf], ((q >> 1) - 1) & 1u);     // MMAs of block q - 2 done: operand buffer free
            umma::mbar_wait(&mbar[4 + (q % RING)], (q / RING) & 1u);
            {
                umma::fence_after_sync();
                const uint32_t* src = reinterpret_cast<const uint32_t*>(sR + (q % RING) * X::SLOT) + tid;
                uint32_t x[kTcwBlockChunks];
#pragma unroll
                for (int i = 0; i < kTcwBlockChunks; ++i) x[i] = (uint32_t)i < nch ? src[i * 128] : 0u;
                const uint32_t ta = tlane + tA + buf * 64u;
#pragma unroll
                for (int i = 0; i < kTcwBlockChunks; ++i) {
                    const uint32_t y = x[i] >> 8;
                    umma::tmem_st4(ta + 4 * i, x[i] & 0x00030003u, x[i] & 0x000C000Cu, x[i] & 0x00300030u, x[i] & 0x00C000C0u);
                    umma::tmem_st4(ta + 32 + 4 * i, y & 0x00030003u, y & 0x000C000Cu, y & 0x00300030u, y & 0x00C000C0u);
                }
                umma::tmem_st_wait();
                umma::fence_before_sync();
            }
            umma::mbar_arrive(&mbar[2 + buf]);
        }
        // z0 complete when the MMAs of the last block are (commits complete in issue order)
        umma::mbar_wait(&mbar[(q - 1) & 1u], ((q - 1) >> 1) & 1u);
        umma::fence_after_sync();
        const uint32_t rowA_g = st * kTcRows + tid, rowB_g = rowA_g + 128;
#pragma unroll
        for (uint32_t h = 0; h < 2; ++h) {
            const uint32_t tb = tlane + h * NN;
            float v[NN];                 // n = piece * W0 + c
#pragma unroll
            for (int g = 0; g < NN / 16; ++g) umma::tmem_ld16(tb + 16 * g, v + 16 * g);
            const uint32_t row = h ? rowB_g : rowA_g;
            float out[W0];
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float z = v[c] + (v[W0 + c] + v[2 * W0 + c]);
                const float z0 = fmaf(z, 8589934592.f /* 2^33 */, b0s[c]);
                if constexpr (ACT == BANN_TANH) {
                    float e, r;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z0 * 2.8853900817779268f));
                    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
                    out[c] = fmaf(r, -2.f, 1.f);
                } else if constexpr (ACT == BANN_SILU) {
                    out[c] = z0;
                } else {
                    float unused;
                    out[c] = act_s<ACT>(z0, unused);
                }
            }
            if (row < a.k.n) {
                float4* dst = reinterpret_cast<float4*>(a0_g + (size_t)row * W0);
#pragma unroll
                for (int c = 0; c < W0; c += 4) dst[c >> 2] = make_float4(out[c], out[c + 1], out[c + 2], out[c + 3]);
            }
        }
        umma::fence_before_sync();     // the accumulator reads precede the next super-tile's first MMA (ordered by the arrivals on mbar[2 + buf])
    }
    if (nblk >= 2) umma::mbar_wait(&mbar[(nblk - 2) & 1u], ((nblk - 2) >> 1) & 1u);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, X::TMEM_A);
}

// ------------------------------------------------------------------ KT: everything after the first layer, FP32
// thread t owns rows t and 128 + t of a 256-row super-tile (packed f32x2, as in k1_tc); the cross-row sums of the layers >= 1
// (up to 16 x 16 per layer) go through a shared-memory staged register-tiled product: 8 row groups x 16 blocks of 4 x 4 entries.
// derivative of the activation at the pre-activation, from the activation (and aux); tanh keeps its two-instruction form
template <int ACT>
__device__ __forceinline__ f2 dact2(f2 a, f2 aux) {
    if constexpr (ACT == BANN_TANH) return dtanh2(a);
    else if constexpr (ACT == BANN_IDENTITY) return dup2(1.f);
    else return mul2(neg_dact2<ACT>(a, aux), dup2(-1.f));
}

template <int H, int S, int D, int ACT>
__global__ void __launch_bounds__(128, 2) k_tcx_tail(TcxArgs a) {
    using T = TailShape<H, S, D>;
    constexpr int NLA = T::NLA, W0 = T::W0, MW = 16, NQ = TcxShape<W0>::NQ, NN = TcxShape<W0>::NN;
    static_assert(H <= 16 && S <= 16 && (H % 4 == 0) && (S % 4 == 0), "widths must be multiples of 4, at most 16");
    extern __shared__ __align__(16) uint8_t smraw[];
    const K1Args& k = a.k;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = k.list ? k.list[li] : li;
    const BranchDesc& d = k.descs[b];
    const uint32_t m = d.m, P = d.P;
    float* wp = reinterpret_cast<float*>(smraw);                       // tail parameters [n_tail]
    float* As = wp + ((T::n_tail() + 3) & ~3);                         // [256][16] staged inputs of a layer
    float* Ds = As + 256 * 16;                                         // [256][16] staged deltas
    float* Es = Ds + 256 * 16;                                         // [256] errors
    float* red = Es + 256;                                             // [8 row groups][...] final reduction scratch
    const float* th = k.theta + d.param_off;
    // programmatic dependent launch: the parameters were written before k_tcx_prep was launched (an ordinary launch), so staging
    // them may run under KA; its activations are read after the wait
    for (uint32_t i = tid; i < (uint32_t)T::n_tail(); i += 128) wp[i] = th[m * W0 + i];
    pdl_launch_dependents();
    pdl_wait();
    if (k.states && k.states[b].status != ST_RUNNING) return;
    __syncthreads();

    const f2 zero2 = dup2(0.f);
    const size_t eoff = k.out_per_entry ? (size_t)li * k.n : 0;
    const size_t toff = (k.target_mode == TGT_PER_ENTRY) ? (size_t)li * k.n : 0;
    const uint32_t t_begin = chunk * k.st_per_chunk;
    const uint32_t t_end = min(k.nst, t_begin + k.st_per_chunk);
    const float* tsrc = (k.target_mode == TGT_RESID_PLUS_PRED) ? k.resid : (k.tgt ? k.tgt + toff : nullptr);
    const bool bwd = !k.fwd_only;
    const float* a0_g = a.a0 + (size_t)li * a.a0_stride;
    uint8_t* dp_g = a.dp + (size_t)li * a.dp_stride;
    // register-tiled cross-row products: row group rg (rows r * 8 + rg), block (ib, cb) of 4 x 4 entries
    const uint32_t rg = tid >> 4, blk = tid & 15, ib = blk >> 2, cb = blk & 3;
    auto swz = [](uint32_t row, uint32_t col) { return row * 16 + ((((col >> 2) ^ (row >> 1)) & 3u) << 2) + (col & 3u); };   // staged element (row, col)
    f2 gacc[NLA > 1 ? NLA - 1 : 1][4][2];        // 4 x 4 block, packed over column pairs
    float gbacc[NLA], gwo = 0.f;
    f2 rss = zero2;
#pragma unroll
    for (int l = 0; l < (NLA > 1 ? NLA - 1 : 1); ++l)
#pragma unroll
        for (int i = 0; i < 4; ++i) { gacc[l][i][0] = zero2; gacc[l][i][1] = zero2; }
#pragma unroll
    for (int l = 0; l < NLA; ++l) gbacc[l] = 0.f;

    for (uint32_t st = t_begin; st < t_end; ++st) {
        const uint32_t rowA_g = st * kTcRows + tid, rowB_g = rowA_g + 128;
        const bool vA = rowA_g < k.n, vB = rowB_g < k.n;
        f2 act[NLA][MW], aux[NLA][MW];
        {
            const float4* pa = reinterpret_cast<const float4*>(a0_g + (size_t)rowA_g * W0);
            const float4* pb = reinterpret_cast<const float4*>(a0_g + (size_t)rowB_g * W0);
#pragma unroll
            for (int c = 0; c < W0; c += 4) {
                const float4 xa = vA ? pa[c >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 xb = vB ? pb[c >> 2] : make_float4(0.f, 0.f, 0.f, 0.f);
                act[0][c] = mk2(xa.x, xb.x); act[0][c + 1] = mk2(xa.y, xb.y);
                act[0][c + 2] = mk2(xa.z, xb.z); act[0][c + 3] = mk2(xa.w, xb.w);
            }
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                if constexpr (ACT == BANN_SILU) act[0][c] = act2<ACT>(act[0][c], aux[0][c]);      // KA stored the pre-activation
                else aux[0][c] = act[0][c];
            }
        }
        f2 tg = zero2;
        if (tsrc) tg = mk2(vA ? __ldg(tsrc + rowA_g) : 0.f, vB ? __ldg(tsrc + rowB_g) : 0.f);
#pragma unroll
        for (int l = 1; l < NLA; ++l) {
#pragma unroll
            for (int c = 0; c < MW; ++c) {
                if (c < T::width(l)) {
                    f2 zz = dup2(wp[T::b_off(l) + c]);
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) zz = fma2(act[l - 1][i], dup2(wp[T::w_off(l) + c * T::in_w(l) + i]), zz);
                    act[l][c] = act2<ACT>(mul2(zz, dup2(act_prescale<ACT>())), aux[l][c]);
                }
            }
        }
        f2 yh = zero2;
#pragma unroll
        for (int i = 0; i < S; ++i) yh = fma2(act[NLA - 1][i], dup2(wp[T::w_off(NLA) + i]), yh);
        if (k.target_mode == TGT_RESID_PLUS_PRED) tg = add2(tg, yh);                      // net.rs:280
        const f2 e = mul2(fma2(tg, dup2(-1.f), yh), mk2(vA ? 1.f : 0.f, vB ? 1.f : 0.f));    // branch_sampler.rs:821
        {
            auto put = [&](float* dst, uint32_t row, float v, int accumulate) {
                if (!dst || row >= k.n) return;
                float* p = dst + eoff + row;
                if (accumulate > 0) *p += v;
                else if (accumulate < 0) *p -= v;
                else *p = v;
            };
            if (k.target_mode == TGT_RESID_PLUS_PRED) {
                put(k.tgt_out, rowA_g, lo2(tg), 0); put(k.tgt_out, rowB_g, hi2(tg), 0);
                put(k.prev_out, rowA_g, lo2(yh), 0); put(k.prev_out, rowB_g, hi2(yh), 0);   // net.rs:279
            }
            put(k.yhat_out, rowA_g, lo2(yh), k.yhat_accumulate);
            put(k.yhat_out, rowB_g, hi2(yh), k.yhat_accumulate);
        }
        if (!bwd) continue;

        rss = fma2(e, e, rss);
        // stage (input, delta) of a layer for the whole super-tile, then every thread accumulates its 4 x 4 block over 32 rows
        auto stage = [&](const f2* in, int nin, const f2* dl, int nout) {
            __syncthreads();                                   // the previous product has been read
            // rows are 64 B apart: the 16-byte group g of row r sits at group g ^ ((r >> 1) & 3), which makes the 128-bit
            // stores of eight consecutive rows (one quarter-warp) hit eight different bank groups (rows t and 128 + t swizzle alike)
            float* ra = As + tid * 16; float* rb = As + (128 + tid) * 16;
            float* da = Ds + tid * 16; float* db = Ds + (128 + tid) * 16;
            const int sw = (int)((tid >> 1) & 3u);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const int o = ((i >> 2) ^ sw) << 2;
                if (i < nin) {
                    *reinterpret_cast<float4*>(ra + o) = make_float4(lo2(in[i]), lo2(in[i + 1]), lo2(in[i + 2]), lo2(in[i + 3]));
                    *reinterpret_cast<float4*>(rb + o) = make_float4(hi2(in[i]), hi2(in[i + 1]), hi2(in[i + 2]), hi2(in[i + 3]));
                }
                if (i < nout) {
                    *reinterpret_cast<float4*>(da + o) = make_float4(lo2(dl[i]), lo2(dl[i + 1]), lo2(dl[i + 2]), lo2(dl[i + 3]));
                    *reinterpret_cast<float4*>(db + o) = make_float4(hi2(dl[i]), hi2(dl[i + 1]), hi2(dl[i + 2]), hi2(dl[i + 3]));
                }
            }
            __syncthreads();
        };
        // output layer: gWo_i = sum_rows a_last[i] e ; delta of the summary layer
        f2 delta[MW];
#pragma unroll
        for (int i = 0; i < S; ++i) delta[i] = mul2(dact2<ACT>(act[NLA - 1][i], aux[NLA - 1][i]), mul2(e, dup2(wp[T::w_off(NLA) + i])));
        __syncthreads();
        Es[tid] = lo2(e); Es[128 + tid] = hi2(e);
#pragma unroll
        for (int l = NLA - 1; l >= 1; --l) {
            stage(act[l - 1], T::in_w(l), delta, T::width(l));
            if ((int)(4 * ib) < T::in_w(l) && (int)(4 * cb) < T::width(l)) {
#pragma unroll 4
                for (uint32_t r = 0; r < 32; ++r) {
                    const uint32_t row = r * 8 + rg, sw = (row >> 1) & 3u;
                    const float4 x = *reinterpret_cast<const float4*>(As + row * 16 + 4 * (ib ^ sw));
                    const float4 y = *reinterpret_cast<const float4*>(Ds + row * 16 + 4 * (cb ^ sw));
                    const float xs[4] = {x.x, x.y, x.z, x.w};
                    const f2 y01 = mk2(y.x, y.y), y23 = mk2(y.z, y.w);      // packed over the delta columns: 8 FFMA2 per row
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        gacc[l - 1][i][0] = fma2(y01, dup2(xs[i]), gacc[l - 1][i][0]);
                        gacc[l - 1][i][1] = fma2(y23, dup2(xs[i]), gacc[l - 1][i][1]);
                    }
                }
            }
            if (blk < (uint32_t)T::width(l)) {                  // bias of layer l: column sums of the staged deltas
                float s = gbacc[l];
                for (uint32_t r = 0; r < 32; ++r) s += Ds[swz(r * 8 + rg, blk)];
                gbacc[l] = s;
            }
            f2 nd[MW];
#pragma unroll
            for (int i = 0; i < MW; ++i) nd[i] = zero2;
#pragma unroll
            for (int c = 0; c < MW; ++c)
                if (c < T::width(l)) {
#pragma unroll
                    for (int i = 0; i < MW; ++i)
                        if (i < T::in_w(l)) nd[i] = fma2(delta[c], dup2(wp[T::w_off(l) + c * T::in_w(l) + i]), nd[i]);
                }
#pragma unroll
            for (int i = 0; i < MW; ++i)
                if (i < T::in_w(l)) delta[i] = mul2(dact2<ACT>(act[l - 1][i], aux[l - 1][i]), nd[i]);
        }
        // first layer: bias sums from the staged delta_0; output weights from the staged a_last and the errors
        stage(act[NLA - 1], S, delta, W0);
        if (blk < (uint32_t)W0) {
            float s = gbacc[0];
            for (uint32_t r = 0; r < 32; ++r) s += Ds[swz(r * 8 + rg, blk)];
            gbacc[0] = s;
        }
        if (blk < (uint32_t)S) {
            float s = gwo;
            for (uint32_t r = 0; r < 32; ++r) s = fmaf(As[swz(r * 8 + rg, blk)], Es[r * 8 + rg], s);
            gwo = s;
        }
        // delta_0 -> three bf16 pieces per unit by truncation (exact), n = piece * W0 + unit, in KB's operand layout
        {
            uint8_t* dst = dp_g + (size_t)st * TcxShape<W0>::DP_ST + tid * 16;
            uint32_t pa[NN], pb[NN];
#pragma unroll
            for (int n = 0; n < NN; ++n) { pa[n] = 0u; pb[n] = 0u; }
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                f2 v = mul2(delta[c], dup2(1.2676506002282294e30f));   // 2^100, exact
#pragma unroll
                for (int piece = 0; piece < 3; ++piece) {
                    const uint32_t ua = __float_as_uint(lo2(v)) & 0xFFFF0000u, ub = __float_as_uint(hi2(v)) & 0xFFFF0000u;
                    pa[piece * W0 + c] = ua; pb[piece * W0 + c] = ub;
                    if (piece < 2) v = add2(v, mk2(-__uint_as_float(ua), -__uint_as_float(ub)));
                }
            }
#pragma unroll
            for (int qq = 0; qq < NQ; ++qq) {
                uint4 wa, wb;
                wa.x = __byte_perm(pa[8 * qq], pa[8 * qq + 1], 0x7632); wa.y = __byte_perm(pa[8 * qq + 2], pa[8 * qq + 3], 0x7632);
                wa.z = __byte_perm(pa[8 * qq + 4], pa[8 * qq + 5], 0x7632); wa.w = __byte_perm(pa[8 * qq + 6], pa[8 * qq + 7], 0x7632);
                wb.x = __byte_perm(pb[8 * qq], pb[8 * qq + 1], 0x7632); wb.y = __byte_perm(pb[8 * qq + 2], pb[8 * qq + 3], 0x7632);
                wb.z = __byte_perm(pb[8 * qq + 4], pb[8 * qq + 5], 0x7632); wb.w = __byte_perm(pb[8 * qq + 6], pb[8 * qq + 7], 0x7632);
                *reinterpret_cast<uint4*>(dst + qq * kTcChunkStride) = wa;
                *reinterpret_cast<uint4*>(dst + qq * kTcChunkStride + 128 * 16) = wb;
            }
        }
    }
    if (!bwd || !k.part) return;

    // ---- CTA epilogue: sum the 8 row groups in ascending order (fixed order), write the partials of this (entry, chunk)
    float* pp = k.part + ((size_t)li * k.nchunk + chunk) * k.pstride;
    __syncthreads();
    {   // rss: lanes -> warps
        float v = lo2(rss) + hi2(rss);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) Es[warp] = v;
    }
    __syncthreads();
    if (tid == 0) pp[P] = (Es[0] + Es[1]) + (Es[2] + Es[3]);
    // every (rg, blk) thread publishes its sums; thread blk of row group 0 adds the 8 groups
    constexpr int PER = 16 * (NLA > 1 ? NLA - 1 : 1) + NLA + 1;       // values per thread
    {
        float* mine = red + (size_t)tid * PER;
        int idx = 0;
#pragma unroll
        for (int l = 0; l < (NLA > 1 ? NLA - 1 : 1); ++l)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) mine[idx++] = (c & 1) ? hi2(gacc[l][i][c >> 1]) : lo2(gacc[l][i][c >> 1]);
#pragma unroll
        for (int l = 0; l < NLA; ++l) mine[idx++] = gbacc[l];
        mine[idx++] = gwo;
    }
    __syncthreads();
    if (tid < 16) {
        const uint32_t bi = tid >> 2, bc = tid & 3;
        int idx = 0;
        auto total = [&](int at) {
            float s = 0.f;
#pragma unroll
            for (int g = 0; g < 8; ++g) s += red[(size_t)(g * 16 + tid) * PER + at];
            return s;
        };
#pragma unroll
        for (int l = 1; l < NLA; ++l)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float s = total(idx++);
                    const int ii = 4 * bi + i, cc = 4 * bc + c;
                    if (ii < T::in_w(l) && cc < T::width(l)) pp[m * W0 + T::w_off(l) + cc * T::in_w(l) + ii] = s;
                }
        if (NLA == 1) idx += 16;
#pragma unroll
        for (int l = 0; l < NLA; ++l) {
            const float s = total(idx++);
            if ((int)tid < T::width(l)) pp[m * W0 + T::b_off(l) + tid] = s;
        }
        const float s = total(idx++);
        if ((int)tid < S) pp[m * W0 + T::w_off(NLA) + tid] = s;
    }
}

// ------------------------------------------------------------------ KB: S = X^T delta_0 for one slab of 512 markers
// The unit is a PAIR of marker blocks (128 markers): one M = 128 MMA per 16 rows covers both (at M = 64 the tensor core runs at
// half rate), its accumulator fills all 128 lanes x NN columns.  One 64 KB operand image per CTA (the second resident CTA covers
// the MMA latency), a 2-slot ring of packed words, the delta pieces of the current super-tile.  Warps 0-3 expand and run the
// epilogue, warp 4 requests the loads and issues the MMAs (as in KA).
template <int W0>
__global__ void __launch_bounds__(160, 2) k_tcx_bwd(TcxArgs a) {
    using X = TcxShape<W0>;
    constexpr int NN = X::NN;
    constexpr uint32_t PAIR_CH = 2 * kTcwBlockChunks;                        // 16 chunks of 8 markers
    extern __shared__ __align__(16) uint8_t smraw[];
    const K1Args& k = a.k;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t li = blockIdx.y, chunk = blockIdx.x, slab = blockIdx.z;
    const uint32_t b = k.list ? k.list[li] : li;
    const BranchDesc& d = k.descs[b];
    const uint32_t m = d.m, NC = d.nc, NP = (NC + PAIR_CH - 1) / PAIR_CH;     // pairs of the branch
    const uint32_t p0 = slab * kTcxSlabPairs;
    if (p0 >= NP) { pdl_launch_dependents(); pdl_wait(); return; }
    const uint32_t np = min((uint32_t)kTcxSlabPairs, NP - p0);
    uint8_t* sA = smraw + ((128u - (umma::smem_u32(smraw) & 127u)) & 127u);   // operand image: 16 chunks x 256 rows x 16 B
    uint32_t* sG = reinterpret_cast<uint32_t*>(sA + 2 * kTcwBlockBytes);     // ring [2 slots][16 chunks][128] packed words
    uint8_t* sD = reinterpret_cast<uint8_t*>(sG) + 2 * PAIR_CH * 512;        // delta pieces of the current super-tile
    float* gb0 = reinterpret_cast<float*>(sD + X::DP_ST);                    // [W0]
    // mbarriers: [0] MMAs that read the operand image done; [2] image expanded (128 arrivals); [4..5] ring slot landed;
    //            [8] delta pieces landed
    uint64_t* mbar = reinterpret_cast<uint64_t*>(gb0 + 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 10);
    if (tid == 0) {
        umma::mbar_init(&mbar[0], 1);
        umma::mbar_init(&mbar[2], 128);
        umma::mbar_init(&mbar[4], 1); umma::mbar_init(&mbar[5], 1);
        umma::mbar_init(&mbar[8], 1);
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, X::TMEM_B);
    pdl_launch_dependents();       // programmatic dependent launch: the prologue above ran under KT
    pdl_wait();
    if (k.states && k.states[b].status != ST_RUNNING) {
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (warp == 0) umma::tmem_dealloc(*tmem_slot, X::TMEM_B);
        return;
    }
    float* pp = k.part + ((size_t)li * k.nchunk + chunk) * k.pstride;
    if (tid < W0) gb0[tid] = pp[d.b_off[0] + tid];          // written by KT for the same (entry, chunk)
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((warp * 32u) << 16);
    const uint32_t sA_u = umma::smem_u32(sA), sD_u = umma::smem_u32(sD);
    constexpr uint32_t idesc_b = umma::make_idesc(umma::FMT_BF16, umma::FMT_BF16, 1, 1, 128, NN);
    const uint64_t dA_b = umma::make_desc(sA_u, 128, kTcChunkStride), dD_b = umma::make_desc(sD_u, 128, kTcChunkStride);

    const uint32_t t_begin = chunk * k.st_per_chunk;
    const uint32_t t_end = min(k.nst, t_begin + k.st_per_chunk);
    const uint32_t nit = t_end > t_begin ? t_end - t_begin : 0;
    const uint32_t nq = nit * np;                                            // length of the pair stream of this CTA
    const uint32_t* gwords = k.store_tc + (d.tc_off >> 2);
    const uint8_t* dp_g = a.dp + (size_t)li * a.dp_stride;
    auto chunks_of = [&](uint32_t p) { return min(PAIR_CH, NC - p * PAIR_CH); };     // chunks of pair p
    auto issue_load = [&](uint32_t q, uint32_t it, uint32_t pl) {   // whole issuer warp enters
        if (umma::elect_one())
            umma::bulk_load(sG + (q & 1u) * (PAIR_CH * 128), gwords + ((size_t)(t_begin + it) * NC + (p0 + pl) * PAIR_CH) * 128,
                            chunks_of(p0 + pl) * 512u, &mbar[4 + (q & 1u)]);
        __syncwarp();
    };
    auto issue_dp = [&](uint32_t it) {
        if (umma::elect_one()) umma::bulk_load(sD, dp_g + (size_t)(t_begin + it) * X::DP_ST, X::DP_ST, &mbar[8]);
        __syncwarp();
    };
    if (warp == 4) {                 // ---- issuer warp
        uint32_t l_it = 0, l_p = 0, lq = 0;
        auto advance_l = [&]() { if (++l_p == np) { l_p = 0; ++l_it; } ++lq; };
        for (; lq < 2 && lq < nq;) { issue_load(lq, l_it, l_p); advance_l(); }
        if (nit > 0) issue_dp(0);
        uint32_t q = 0;
        for (uint32_t it = 0; it < nit; ++it) {
            if (it > 0) {
                // the delta buffer is read by the MMAs of the previous super-tile (they complete in order): reload after the last one
                umma::mbar_wait(&mbar[0], (q - 1) & 1u);
                issue_dp(it);
            }
            for (uint32_t pl = 0; pl < np; ++pl, ++q) {
                umma::mbar_wait(&mbar[2], q & 1u);                            // pair q expanded by all 128 threads
                if (lq < nq) { issue_load(lq, l_it, l_p); advance_l(); }      // every thread has consumed ring slot q % 2
                if (pl == 0) umma::mbar_wait(&mbar[8], it & 1u);              // delta pieces of this super-tile
                umma::fence_after_sync();
                if (umma::elect_one()) {
#pragma unroll
                    for (uint32_t ks = 0; ks < kTcRows / 16; ++ks)
                        umma::mma_f16(tmem + pl * NN, dA_b + ks * 16u, dD_b + ks * 16u, idesc_b, (it | ks) != 0);
                    umma::commit(&mbar[0]);
                }
                __syncwarp();
            }
        }
        if (nq >= 1) umma::mbar_wait(&mbar[0], (nq - 1) & 1u);
        umma::fence_before_sync();
        __syncthreads();
        return;
    }
    for (uint32_t q = 0; q < nq; ++q) {
        const uint32_t pl = q % np;
        const uint32_t nch = chunks_of(p0 + pl);
        umma::mbar_wait(&mbar[4 + (q & 1u)], (q >> 1) & 1u);
        const uint32_t* src = sG + (q & 1u) * (PAIR_CH * 128) + tid;
        if (q >= 1) umma::mbar_wait(&mbar[0], (q - 1) & 1u);                  // MMAs of pair q - 1 done: operand image free
        tcx_expand(src, sA + tid * 16, min(nch, (uint32_t)kTcwBlockChunks));
        tcx_expand(src + kTcwBlockChunks * 128, sA + kTcwBlockBytes + tid * 16, nch > kTcwBlockChunks ? nch - kTcwBlockChunks : 0u);
        umma::fence_async_smem();
        umma::mbar_arrive(&mbar[2]);
    }
    if (nq >= 1) umma::mbar_wait(&mbar[0], (nq - 1) & 1u);
    umma::fence_after_sync();
    // first-layer weight gradient of the slab: accumulator row r of pair pl lives in lane r (M = 128 layout)
    const float* mu = k.mu + d.col_off;
    const float* sd = k.sd + d.col_off;
    for (uint32_t pl = 0; pl < np; ++pl) {
        float v[NN];
        if (nit > 0) {
#pragma unroll
            for (int g = 0; g < NN / 16; ++g) umma::tmem_ld16(tlane + pl * NN + 16 * g, v + 16 * g);
        }
        const uint32_t j = (p0 + pl) * 128 + tid;
        if (j < m) {
            const float unscale = pow2f(33 - 2 * (int)((j & 7u) >> 1));
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float s = (nit > 0 ? (v[c] + (v[W0 + c] + v[2 * W0 + c])) : 0.f) * unscale;
                pp[c * m + j] = __fdiv_rn(s - mu[j] * gb0[c], sd[j]);
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, X::TMEM_B);
}

float* bann_net_tcx_buffer(bann_net* net, int which, size_t bytes);   // net.cu: grows the three work buffers on demand

template <int H, int S, int D, int ACT>
int launch_one_tcx_act(TcxArgs& a, uint32_t nlist, uint32_t nslab, bool bwd, cudaStream_t st) {
    using T = TailShape<H, S, D>;
    using X = TcxShape<T::W0>;
    constexpr int NLA = T::NLA;
    constexpr int PER = 16 * (NLA > 1 ? NLA - 1 : 1) + NLA + 1;
    const size_t smem_t = ((size_t)((T::n_tail() + 3) & ~3) + 2 * 256 * 16 + 256 + (size_t)128 * PER) * sizeof(float) + 16;
    static bool configured = false;
    if (!configured) {
        BANN_CUDA(cudaFuncSetAttribute(k_tcx_fwd<T::W0, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X::SMEM_A));
        BANN_CUDA(cudaFuncSetAttribute(k_tcx_bwd<T::W0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)X::SMEM_B));
        BANN_CUDA(cudaFuncSetAttribute(k_tcx_tail<H, S, D, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
        configured = true;
    }
    k_tcx_prep<T::W0><<<nlist, 128, 0, st>>>(a);
    BANN_LAUNCHED();
    dim3 grid(a.k.nchunk, nlist);
    BANN_CUDA(launch_pdl(k_tcx_fwd<T::W0, ACT>, grid, dim3(160), X::SMEM_A, st, a));
    BANN_LAUNCHED();
    BANN_CUDA(launch_pdl(k_tcx_tail<H, S, D, ACT>, grid, dim3(128), smem_t, st, a));
    BANN_LAUNCHED();
    if (bwd) {
        dim3 gridb(a.k.nchunk, nlist, nslab);
        BANN_CUDA(launch_pdl(k_tcx_bwd<T::W0>, gridb, dim3(160), X::SMEM_B, st, a));
        BANN_LAUNCHED();
    }
    BANN_CUDA(cudaGetLastError());
    return 0;
}

template <int H, int S, int D>
int launch_one_tcx(TcxArgs& a, uint32_t nlist, uint32_t nslab, bool bwd, cudaStream_t st) {
    switch (a.k.act) {
        case BANN_TANH: return launch_one_tcx_act<H, S, D, BANN_TANH>(a, nlist, nslab, bwd, st);
        case BANN_RELU: return launch_one_tcx_act<H, S, D, BANN_RELU>(a, nlist, nslab, bwd, st);
        case BANN_LEAKY_RELU: return launch_one_tcx_act<H, S, D, BANN_LEAKY_RELU>(a, nlist, nslab, bwd, st);
        case BANN_SILU: return launch_one_tcx_act<H, S, D, BANN_SILU>(a, nlist, nslab, bwd, st);
        default: return launch_one_tcx_act<H, S, D, BANN_IDENTITY>(a, nlist, nslab, bwd, st);
    }
}

// Wide-branch tensor-core K1: homogeneous architecture, tensor-core store present, widths in the instantiated set.
#ifdef BANN_K1_TCX_IMPL
int launch_k1_tcx(const std::vector<BranchDesc>& descs, int single_branch, K1Args& k, uint32_t nlist, int num_sms,
                         cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net) {
    *launched = false;
    if (!k.store_tc) return 0;
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
    if (max_m > (uint32_t)kTcxMaxMarkers) return 0;
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    const uint32_t nst = k.nst;
    const uint32_t nkb_max = ((max_m + 7) / 8 + kTcwBlockChunks - 1) / kTcwBlockChunks;
    const uint32_t nslab = ((nkb_max + 1) / 2 + kTcxSlabPairs - 1) / kTcxSlabPairs;
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)num_sms * 2 + nlist - 1) / nlist);
    uint32_t nchunk = std::min<uint32_t>(want, std::max<uint32_t>(1, nst));
    uint32_t spc = (nst + nchunk - 1) / nchunk;
    nchunk = (nst + spc - 1) / spc;
#define BANN_TRY_TCX(HH, SS, DD)                                                                          \
    if (!*launched && H == HH && S == SS && D == DD) {                                                    \
        using XX = TcxShape<TailShape<HH, SS, DD>::W0>;                                                   \
        TcxArgs a;                                                                                        \
        k.nchunk = nchunk;                                                                                \
        k.st_per_chunk = spc;                                                                             \
        if (part_io) {                                                                                    \
            if (nchunk == 1) *part_io = bann_net_gsum(net);                                               \
            else {                                                                                        \
                float* p = bann_net_partials(net, (size_t)nlist * nchunk * bann_net_pstride(net));        \
                if (!p) return -2;                                                                        \
                *part_io = p;                                                                             \
            }                                                                                             \
            k.part = *part_io;                                                                            \
        }                                                                                                 \
        *nchunk_io = nchunk;                                                                              \
        a.nkb_max = nkb_max;                                                                              \
        a.wp_stride = ((size_t)nkb_max * XX::WBLK + 64 + 127) & ~(size_t)127;                             \
        a.a0_stride = (size_t)nst * kTcRows * TailShape<HH, SS, DD>::W0;                                  \
        a.dp_stride = (size_t)nst * XX::DP_ST;                                                            \
        a.wp = reinterpret_cast<uint8_t*>(bann_net_tcx_buffer(net, 0, (size_t)nlist * a.wp_stride));      \
        a.a0 = bann_net_tcx_buffer(net, 1, (size_t)nlist * a.a0_stride * sizeof(float));                  \
        a.dp = reinterpret_cast<uint8_t*>(bann_net_tcx_buffer(net, 2, k.fwd_only ? 16 : (size_t)nlist * a.dp_stride)); \
        if (!a.wp || !a.a0 || !a.dp) return -2;                                                           \
        a.k = k;                                                                                          \
        int rc = launch_one_tcx<HH, SS, DD>(a, nlist, nslab, !k.fwd_only, st);                            \
        if (rc) return rc;                                                                                \
        *launched = true;                                                                                 \
    }
    BANN_TRY_TCX(16, 16, 2)
    BANN_TRY_TCX(16, 16, 1)
    BANN_TRY_TCX(12, 12, 2)
    BANN_TRY_TCX(12, 12, 1)
    BANN_TRY_TCX(8, 8, 2)
    BANN_TRY_TCX(8, 8, 1)
    BANN_TRY_TCX(8, 4, 1)
    BANN_TRY_TCX(4, 4, 1)
#undef BANN_TRY_TCX
    return 0;
}
#else
int launch_k1_tcx(const std::vector<BranchDesc>& descs, int single_branch, K1Args& k, uint32_t nlist, int num_sms,
                         cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net);
#endif

}  // namespace bann
