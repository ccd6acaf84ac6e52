// Tuned K1 for narrow branches: fused forward + backward, FP32 FFMA, one warp per 128-row tile.
//
// Template <H, S, D, NP>: D hidden layers of width H, summary width S (compile time, fully
// unrolled), up to 32*NP markers per branch.  Layout of the work inside a warp:
//   forward  : lane = row-quad q (4 individuals = one packed byte per marker); the first-layer
//              pre-activations of its 4 rows live in registers; markers stream from shared memory
//              four at a time (one 32-bit word), W' = W0/sd rows are broadcast float4 loads.
//   tail     : the lane runs the remaining tiny layers + backward deltas for its 4 rows; the
//              cross-row sums of every layer >= 1 (gW_l, gb_l, rss) accumulate in per-lane
//              registers over ALL tiles of the CTA and are reduced once at the end.
//   backward : lane = marker (pass p: marker 32p + lane); delta_0 of the tile is broadcast from
//              shared memory; sum_i g_ij delta_0[i,c] accumulates in registers over all tiles.
// The packed tile is read from HBM once and used for both directions.  Standardisation is folded
// into the first layer (W' = W/sd, b' = b - sum mu W') and unfolded on the gradient.
// Partial sums leave the CTA in a fixed order (deterministic, no float atomics).
#pragma once
#include "kernels.cuh"

struct bann_net;
float* bann_net_partials(bann_net* net, size_t need);
float* bann_net_gsum(bann_net* net);
uint32_t bann_net_pstride(bann_net* net);

namespace bann {

template <int H, int S, int D>
struct TailShape {
    static constexpr int NLA = D + 1;                 // activated layers (hidden..., summary)
    static constexpr int W0 = D > 0 ? H : S;          // width of the layer fed by the genotypes
    static constexpr int MW = H > S ? H : S;
    static constexpr int W0P = (W0 + 3) & ~3;         // padded to float4
    __host__ __device__ static constexpr int width(int l) { return l < D ? H : S; }   // l < NLA
    __host__ __device__ static constexpr int in_w(int l) { return width(l - 1); }     // 1 <= l <= NLA (NLA: output)
    // offsets inside the "tail" parameter block = theta[m*W0 .. P): weights of layers 1..NLA, then all biases
    __host__ __device__ static constexpr int w_off(int l) {   // 1 <= l <= NLA
        int o = 0;
        for (int k = 1; k < l; ++k) o += in_w(k) * width(k);
        return o;
    }
    __host__ __device__ static constexpr int n_tail_w() { return w_off(NLA) + width(NLA - 1); }  // + output weights (S x 1)
    __host__ __device__ static constexpr int b_off(int l) {   // 0 <= l < NLA
        int o = n_tail_w();
        for (int k = 0; k < l; ++k) o += width(k);
        return o;
    }
    __host__ __device__ static constexpr int n_tail() { return b_off(NLA - 1) + width(NLA - 1); }
};

template <int H, int S, int D, int NP, int NW>
__global__ void __launch_bounds__(NW * 32) k1_small(K1Args a) {
    using T = TailShape<H, S, D>;
    constexpr int NLA = T::NLA, W0 = T::W0, W0P = T::W0P, MW = T::MW;
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    if (a.states && a.states[b].status != ST_RUNNING) return;
    const BranchDesc& d = a.descs[b];
    const uint32_t m = d.m, mp = d.m_pad4, wpr = mp >> 2;     // words per row-quad in global memory
    const uint32_t tsw = wpr | 1u;                            // odd word stride in shared memory
    // ---- shared memory carve-up
    float* Wp = reinterpret_cast<float*>(smraw);              // [mp][W0P]   W' = W0 / sd (zero rows beyond m)
    float* b0p = Wp + (size_t)mp * W0P;                       // [W0P]
    float* sp = b0p + W0P;                                    // tail parameters [n_tail]
    float* red = sp + ((T::n_tail() + 3) & ~3);               // cross-warp reduction scratch [NW][...]
    constexpr int NTACC = 1 + S + W0 + (NLA > 1 ? (NLA - 1) * (MW * MW + MW) : 0);
    float* wbase = red + NW * (NTACC > NP * W0 ? NTACC : NP * W0) + 4;
    const uint32_t per_warp_words = ((32 * tsw + 3) & ~3u) + 128 * W0P;
    uint32_t* tilew = reinterpret_cast<uint32_t*>(wbase) + (size_t)warp * per_warp_words;
    float* dbuf = reinterpret_cast<float*>(tilew + ((32 * tsw + 3) & ~3u));   // [128][W0P] delta_0
    const uint8_t* tileb = reinterpret_cast<const uint8_t*>(tilew);

    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;
    // ---- stage parameters
    for (uint32_t k = tid; k < mp * W0P; k += NW * 32) {
        const uint32_t j = k / W0P, c = k % W0P;
        Wp[k] = (j < m && c < W0) ? __fdiv_rn(th[c * m + j], sd[j]) : 0.f;
    }
    for (uint32_t k = tid; k < (uint32_t)T::n_tail(); k += NW * 32) sp[k] = th[m * W0 + k];
    __syncthreads();
    if (tid < W0P) {
        float acc = 0.f;
        if (tid < W0) {
            for (uint32_t j = 0; j < m; ++j) acc = fmaf(mu[j], Wp[j * W0P + tid], acc);
            acc = sp[T::b_off(0) + tid] - acc;
        }
        b0p[tid] = acc;
    }
    __syncthreads();

    // ---- persistent per-lane accumulators
    float acc0[NP][W0];                  // backward first layer: lane = marker
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int c = 0; c < W0; ++c) acc0[p][c] = 0.f;
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
    const uint32_t t_begin = chunk * a.tiles_per_chunk;
    const uint32_t t_end = min(a.ntiles, t_begin + a.tiles_per_chunk);
    for (uint32_t t = t_begin + warp; t < t_end; t += NW) {
        // ---- load the packed tile (coalesced 16-byte reads, re-strided to an odd word stride)
        {
            const uint4* src = reinterpret_cast<const uint4*>(a.store + d.tile_off + (size_t)t * (kTileQuads * mp));
            const uint32_t nvec = (kTileQuads * wpr) >> 2;    // 32*wpr words is a multiple of 4
            __syncwarp();
            for (uint32_t v = lane; v < nvec; v += 32) {
                const uint4 x = __ldg(src + v);
                const uint32_t w0i = v * 4;
                const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t wi = w0i + u;
                    tilew[(wi / wpr) * tsw + (wi % wpr)] = xs[u];
                }
            }
            __syncwarp();
        }
        // ---- forward, first layer: lane = row-quad
        float z[4][W0];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < W0; ++c) z[r][c] = b0p[c];
        const uint32_t* myrow = tilew + lane * tsw;
        for (uint32_t jw = 0; jw < wpr; ++jw) {
            const uint32_t word = myrow[jw];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float w[W0P];
                const float4* wr = reinterpret_cast<const float4*>(Wp + (size_t)(4 * jw + k) * W0P);
#pragma unroll
                for (int v = 0; v < W0P / 4; ++v) {
                    const float4 f = wr[v];
                    w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float gf = (float)((word >> (8 * k + 2 * r)) & 3u);
#pragma unroll
                    for (int c = 0; c < W0; ++c) z[r][c] = fmaf(gf, w[c], z[r][c]);
                }
            }
        }
        // ---- tail: remaining layers, error, backward deltas, for the lane's 4 rows
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const uint32_t row = t * kTileRows + lane * 4 + r;
            const bool valid = row < a.n;
            float act[NLA][MW], dh[NLA][MW];
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                const float h = act_h(a.act, z[r][c]);
                act[0][c] = h;
                dh[0][c] = act_dh(a.act, z[r][c], h);
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
                        const float h = act_h(a.act, zz);
                        act[l][c] = h;
                        dh[l][c] = act_dh(a.act, zz, h);
                    }
                }
            }
            float yh = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i) yh = fmaf(act[NLA - 1][i], sp[T::w_off(NLA) + i], yh);
            float tg = 0.f;
            if (valid) {
                if (a.target_mode == TGT_RESID_PLUS_PRED) {
                    tg = a.resid[row] + yh;
                    if (a.tgt_out) a.tgt_out[eoff + row] = tg;
                    if (a.prev_out) a.prev_out[eoff + row] = yh;
                } else if (a.tgt) {
                    tg = a.tgt[(a.target_mode == TGT_PER_ENTRY ? (size_t)li * a.n : 0) + row];
                }
                if (a.yhat_out) {
                    if (a.yhat_accumulate > 0) a.yhat_out[eoff + row] += yh;
                    else if (a.yhat_accumulate < 0) a.yhat_out[eoff + row] -= yh;
                    else a.yhat_out[eoff + row] = yh;
                }
            }
            if (a.fwd_only) continue;
            const float e = valid ? yh - tg : 0.f;
            rss = fmaf(e, e, rss);
            float delta[MW];
#pragma unroll
            for (int i = 0; i < S; ++i) {
                gWo[i] = fmaf(act[NLA - 1][i], e, gWo[i]);
                delta[i] = dh[NLA - 1][i] * (e * sp[T::w_off(NLA) + i]);
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
                    if (i < T::in_w(l)) delta[i] = dh[l - 1][i] * nd[i];
            }
            float* drow = dbuf + (lane * 4 + r) * W0P;
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                gb0[c] += delta[c];
                drow[c] = delta[c];
            }
        }
        if (a.fwd_only) continue;
        __syncwarp();
        // ---- backward, first layer: lane = marker
        for (uint32_t q = 0; q < kTileQuads; ++q) {
            float dl[4][W0P];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4* dr = reinterpret_cast<const float4*>(dbuf + (q * 4 + r) * W0P);
#pragma unroll
                for (int v = 0; v < W0P / 4; ++v) {
                    const float4 f = dr[v];
                    dl[r][4 * v] = f.x; dl[r][4 * v + 1] = f.y; dl[r][4 * v + 2] = f.z; dl[r][4 * v + 3] = f.w;
                }
            }
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (32u * p < m) {
                    const uint32_t j = 32 * p + lane;
                    const uint32_t byte = (j < mp) ? tileb[q * tsw * 4 + j] : 0u;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float gf = (float)((byte >> (2 * r)) & 3u);
#pragma unroll
                        for (int c = 0; c < W0; ++c) acc0[p][c] = fmaf(gf, dl[r][c], acc0[p][c]);
                    }
                }
            }
        }
    }
    if (a.fwd_only || !a.part) return;

    // ---- CTA epilogue: fixed-order reduction over lanes and warps, unfold the standardisation
    float* pp = a.part + ((size_t)li * a.nchunk + chunk) * a.pstride;
    const uint32_t P = d.P;
    // (1) per-lane tail accumulators -> warp sums (xor tree) -> red[warp][...]
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
        for (int w = 0; w < NW; ++w) s += red[w * NTACC + tid];
        // scatter to param_vec order
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
    // (2) first-layer weight gradient: sum over warps, then (S_jc - mu_j * gb0_c) / sd_j
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (32u * p < m) {
            __syncthreads();
#pragma unroll
            for (int c = 0; c < W0; ++c) red[(warp * W0 + c) * 32 + lane] = acc0[p][c];
            __syncthreads();
            for (uint32_t k = tid; k < 32 * W0; k += NW * 32) {
                const uint32_t c = k / 32, ln = k % 32, j = 32 * p + ln;
                if (j < m) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) s += red[(w * W0 + c) * 32 + ln];
                    pp[c * m + j] = __fdiv_rn(s - mu[j] * s_gb0[c], sd[j]);
                }
            }
        }
    }
}

template <int H, int S, int D, int NP, int NW>
size_t k1_small_smem(uint32_t mp) {
    using T = TailShape<H, S, D>;
    constexpr int NTACC = 1 + S + T::W0 + (T::NLA > 1 ? (T::NLA - 1) * (T::MW * T::MW + T::MW) : 0);
    const uint32_t tsw = (mp >> 2) | 1u;
    size_t fl = (size_t)mp * T::W0P + T::W0P + ((T::n_tail() + 3) & ~3);
    size_t redn = (size_t)NW * (NTACC > 32 * T::W0 ? NTACC : 32 * T::W0) + 4;
    size_t perw = ((32 * tsw + 3) & ~3u) + 128 * T::W0P;
    return (fl + redn + (size_t)NW * perw) * 4 + 16;
}

struct SmallKey { int H, S, D; };

template <int H, int S, int D, int NP, int NW>
int launch_one_small(K1Args& a, uint32_t nlist, uint32_t mp, cudaStream_t st) {
    size_t smem = k1_small_smem<H, S, D, NP, NW>(mp);
    auto kern = k1_small<H, S, D, NP, NW>;
    static size_t configured = 0;
    if (smem > configured) {
        BANN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(a.nchunk, nlist);
    kern<<<grid, NW * 32, smem, st>>>(a);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

// picks an instantiation for a homogeneous launch (all listed branches share the architecture and
// fit the marker bound); otherwise leaves *launched = false and the generic kernel runs.
inline int launch_k1_small(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                           cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net) {
    *launched = false;
    // architecture of the launch: single branch, or all branches (must be homogeneous)
    const BranchDesc& d0 = descs[single_branch >= 0 ? single_branch : 0];
    uint32_t max_m = d0.m, max_mp = d0.m_pad4;
    if (single_branch < 0) {
        for (const BranchDesc& d : descs) {
            if (d.nl != d0.nl) return 0;
            for (uint32_t l = 0; l < d.nl; ++l)
                if (d.widths[l] != d0.widths[l]) return 0;
            max_m = std::max(max_m, d.m);
            max_mp = std::max(max_mp, d.m_pad4);
        }
    }
    const int D = (int)d0.nl - 2;
    const int S = (int)d0.widths[d0.nl - 2];
    const int H = D > 0 ? (int)d0.widths[0] : S;
    for (int l = 0; l < D; ++l)
        if ((int)d0.widths[l] != H) return 0;
    // the small kernel wants fewer, fatter chunks: every CTA keeps NW warps busy on its own tiles
    constexpr int NW = 8;
    uint32_t ntiles = a.ntiles;
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)num_sms * 2 + nlist - 1) / nlist);
    uint32_t nchunk = std::min<uint32_t>(want, std::max<uint32_t>(1, ntiles / NW));
    uint32_t tpc = (ntiles + nchunk - 1) / nchunk;
    nchunk = (ntiles + tpc - 1) / tpc;
#define BANN_TRY(HH, SS, DD, NPP)                                                                         \
    if (!*launched && H == HH && S == SS && D == DD && max_m <= 32u * NPP &&                              \
        k1_small_smem<HH, SS, DD, NPP, NW>(max_mp) <= 200 * 1024) {                                       \
        a.nchunk = nchunk;                                                                                \
        a.tiles_per_chunk = tpc;                                                                          \
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
        int rc = launch_one_small<HH, SS, DD, NPP, NW>(a, nlist, max_mp, st);                             \
        if (rc) return rc;                                                                                \
        *launched = true;                                                                                 \
    }
    BANN_TRY(5, 5, 1, 2)
    BANN_TRY(5, 5, 1, 4)
    BANN_TRY(5, 5, 1, 16)
    BANN_TRY(2, 2, 1, 4)
    BANN_TRY(2, 2, 1, 16)
    BANN_TRY(4, 3, 1, 2)
    BANN_TRY(4, 3, 2, 2)
    BANN_TRY(5, 3, 2, 4)
    BANN_TRY(4, 2, 0, 2)
#undef BANN_TRY
    return 0;
}

}  // namespace bann
