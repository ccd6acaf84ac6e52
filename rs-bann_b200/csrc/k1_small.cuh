// Tuned K1 for narrow branches: fused forward + backward, FP32 FFMA, one warp per 128-row tile.
//
// Template <H, S, D, NP>: D hidden layers of width H, summary width S (compile time, fully
// unrolled, tanh), up to 32*NP markers per branch.  Work layout inside a warp:
//   load     : the tile (32 row-quads x 4*wpr bytes, wpr odd by construction of the store) is
//              copied HBM -> shared memory with 16-byte cp.async, double buffered, so the next
//              tile streams in while the current one is being computed.
//   forward  : lane = row-quad q (4 individuals = one packed byte per marker); the first-layer
//              pre-activations of its 4 rows live in registers; markers stream from shared memory
//              four at a time (one 32-bit word, conflict free because wpr is odd); W' rows are
//              broadcast float4 loads.  Decode is LOP3 + I2FP: float(word & (3 << s)) = g * 2^s,
//              the power of two is folded into the staged weights / undone exactly afterwards.
//   tail     : the lane runs the remaining tiny layers + backward deltas for its 4 rows; the
//              cross-row sums of every layer >= 1 (gW_l, gb_l, rss) accumulate in per-lane
//              registers over ALL tiles of the CTA and are reduced once at the end.
//   backward : lane = marker (pass p: marker 32p + lane); delta_0 of the tile is staged
//              transposed ([unit][row], conflict-free float4 stores, broadcast float4 loads);
//              sum_i g_ij delta_0[i,c] accumulates in registers over all tiles.
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

// tanh with ~3 ulp error, branch free: odd Taylor polynomial below 0.25, 1 - 2/(e^{2|x|}+1) above.
__device__ __forceinline__ float fast_tanh(float x) {
    const float ax = fabsf(x);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
    const float big = copysignf(fmaf(-2.f, r, 1.f), x);
    const float x2 = x * x;
    float p = fmaf(x2, 0.021869488536155203f, -0.05396825396825397f);
    p = fmaf(x2, p, 0.13333333333333333f);
    p = fmaf(x2, p, -0.3333333333333333f);
    const float small = fmaf(p * x2, x, x);
    return ax < 0.25f ? small : big;
}

// activation_functions.rs:23-45 on one value; `aux` = what the derivative needs besides the activation (SiLU: the sigmoid)
template <int ACT>
__device__ __forceinline__ float act_s(float z, float& aux) {
    if constexpr (ACT == BANN_TANH) { aux = 0.f; return fast_tanh(z); }
    else if constexpr (ACT == BANN_RELU) { aux = 0.f; return fmaxf(z, 0.f); }
    else if constexpr (ACT == BANN_LEAKY_RELU) { aux = 0.f; return fmaxf(z, 0.01f * z); }   // x > 0: x, x < 0: 0.01 x, 0 at 0
    else if constexpr (ACT == BANN_SILU) {
        float e, r;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.f));
        aux = r;
        return z * r;
    } else { aux = 0.f; return z; }
}
// derivative at the pre-activation, from the activation (a > 0 <=> x > 0 for the rectifiers)
template <int ACT>
__device__ __forceinline__ float dact_s(float a, float aux) {
    if constexpr (ACT == BANN_TANH) return 1.f - a * a;
    else if constexpr (ACT == BANN_RELU) return a > 0.f ? 1.f : 0.f;
    else if constexpr (ACT == BANN_LEAKY_RELU) return a > 0.f ? 1.f : (a < 0.f ? 0.01f : 0.f);
    else if constexpr (ACT == BANN_SILU) return a + aux * (1.f - a);
    else return 1.f;
}

// packed FP32 FMA (Blackwell FFMA2): {d0,d1} += {a0,a1} * {b0,b1} in one issue slot
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
    unsigned long long ra, rb, rc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(d0), "f"(d1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(rc));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

template <int H, int S, int D>
struct TailShape {
    static constexpr int NLA = D + 1;                 // activated layers (hidden..., summary)
    static constexpr int W0 = D > 0 ? H : S;          // width of the layer fed by the genotypes
    static constexpr int MW = H > S ? H : S;
    static constexpr int W0P = (W0 + 3) & ~3;         // padded to float4
    static constexpr int W2P = (2 * W0 + 3) & ~3;     // staged first-layer row: every weight duplicated {w,w} for FFMA2
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
    static constexpr int NTACC = 1 + S + W0 + (NLA > 1 ? (NLA - 1) * (MW * MW + MW) : 0);
};

template <int H, int S, int D, int NP, int NW, int ACT>
__global__ void __launch_bounds__(NW * 32, (NP <= 4 ? 2 : 1)) k1_small(K1Args a, int nstage) {
    using T = TailShape<H, S, D>;
    constexpr int NLA = T::NLA, W0 = T::W0, W0P = T::W0P, W2P = T::W2P, MW = T::MW, NTACC = T::NTACC;
    constexpr bool PK = NP <= 4;   // packed backward accumulators (2 per marker-unit) only when registers allow
    extern __shared__ __align__(16) uint8_t smraw[];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    if (a.states && a.states[b].status != ST_RUNNING) return;
    const BranchDesc& d = a.descs[b];
    const uint32_t m = d.m, mp = d.m_pad4, wpr = mp >> 2;     // words per row-quad (odd)
    const uint32_t tile_words = kTileQuads * wpr;             // multiple of 4
    // ---- shared memory carve-up
    float* Wp = reinterpret_cast<float*>(smraw);              // [mp][W2P]   {w,w} pairs, w = W0/sd * 2^-(8*(j%4)); zero rows beyond m
    float* b0p = Wp + (size_t)mp * W2P;                       // [W0P]
    float* sp = b0p + W0P;                                    // tail parameters [n_tail]
    float* red = sp + ((T::n_tail() + 3) & ~3);               // cross-warp reduction scratch
    float* wbase = red + NW * (NTACC > 32 * W0 ? NTACC : 32 * W0) + 4;
    const uint32_t per_warp_words = (uint32_t)nstage * tile_words + W0 * 128;
    uint32_t* wtile = reinterpret_cast<uint32_t*>(wbase) + (size_t)warp * per_warp_words;
    float* dT = reinterpret_cast<float*>(wtile + (size_t)nstage * tile_words);   // [W0][128] delta_0 (transposed)

    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;
    // ---- stage parameters
    for (uint32_t k = tid; k < mp * W2P; k += NW * 32) {
        const uint32_t j = k / W2P, c = (k % W2P) >> 1;
        float v = 0.f;
        if (j < m && c < W0) v = __fdiv_rn(th[c * m + j], sd[j]) * exp2f(-8.f * (float)(j & 3u));   // exact scaling
        Wp[k] = v;
    }
    for (uint32_t k = tid; k < (uint32_t)T::n_tail(); k += NW * 32) sp[k] = th[m * W0 + k];
    __syncthreads();
    if (tid < W0P) {
        float acc = 0.f;
        if (tid < W0) {
            for (uint32_t j = 0; j < m; ++j) acc = fmaf(mu[j], Wp[j * W2P + 2 * tid] * exp2f(8.f * (float)(j & 3u)), acc);
            acc = sp[T::b_off(0) + tid] - acc;
        }
        b0p[tid] = acc;
    }
    __syncthreads();

    // ---- persistent per-lane accumulators
    float acc0[NP][W0][PK ? 2 : 1];      // backward first layer: lane = marker (packed: even / odd rows)
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int c = 0; c < W0; ++c)
#pragma unroll
            for (int e = 0; e < (PK ? 2 : 1); ++e) acc0[p][c][e] = 0.f;
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
    const bool vec_ok = (a.n & 3u) == 0;
    const uint32_t t_begin = chunk * a.tiles_per_chunk;
    const uint32_t t_end = min(a.ntiles, t_begin + a.tiles_per_chunk);
    const uint8_t* gbase = a.store + d.tile_off;
    const uint32_t nvec = tile_words >> 2;

    auto issue_load = [&](uint32_t t, uint32_t stage) {
        const uint4* src = reinterpret_cast<const uint4*>(gbase + (size_t)t * tile_words * 4);
        uint4* dst = reinterpret_cast<uint4*>(wtile + (size_t)stage * tile_words);
        for (uint32_t v = lane; v < nvec; v += 32) cp_async16(dst + v, src + v);
        cp_async_commit();
    };

    uint32_t t = t_begin + warp;
    uint32_t stage = 0;
    if (t < t_end) issue_load(t, 0);
    for (; t < t_end; t += NW) {
        // ---- prefetch the next tile, wait for the current one
        const uint32_t tn = t + NW;
        if (nstage == 2) {
            if (tn < t_end) { issue_load(tn, stage ^ 1); cp_async_wait<1>(); }
            else cp_async_wait<0>();
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
        const uint32_t* tilew = wtile + (size_t)stage * tile_words;
        const uint8_t* tileb = reinterpret_cast<const uint8_t*>(tilew);
        // targets of the lane's 4 rows (issued early, consumed in the tail)
        const uint32_t row0 = t * kTileRows + lane * 4;
        float tg4[4] = {0.f, 0.f, 0.f, 0.f};
        {
            const float* src = (a.target_mode == TGT_RESID_PLUS_PRED) ? a.resid : (a.tgt ? a.tgt + toff : nullptr);
            if (src) {
                if (vec_ok && row0 + 3 < a.n) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(src + row0));
                    tg4[0] = v.x; tg4[1] = v.y; tg4[2] = v.z; tg4[3] = v.w;
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (row0 + r < a.n) tg4[r] = src[row0 + r];
                }
            }
        }
        // ---- forward, first layer: lane = row-quad; z[r][c] accumulates 4^r * (x W')
        float z[4][W0];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < W0; ++c) z[r][c] = 0.f;
        const uint32_t* myrow = tilew + lane * wpr;
        for (uint32_t jw = 0; jw < wpr; ++jw) {
            const uint32_t word = myrow[jw];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float w[W2P];
                const float4* wr = reinterpret_cast<const float4*>(Wp + (size_t)(4 * jw + k) * W2P);
#pragma unroll
                for (int v = 0; v < W2P / 4; ++v) {
                    const float4 f = wr[v];
                    w[4 * v] = f.x; w[4 * v + 1] = f.y; w[4 * v + 2] = f.z; w[4 * v + 3] = f.w;
                }
                float gf[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) gf[r] = (float)(word & (3u << (8 * k + 2 * r)));   // g * 2^(8k+2r), exact
#pragma unroll
                for (int c = 0; c < W0; ++c) {
                    ffma2(z[0][c], z[1][c], gf[0], gf[1], w[2 * c], w[2 * c + 1]);
                    ffma2(z[2][c], z[3][c], gf[2], gf[3], w[2 * c], w[2 * c + 1]);
                }
            }
        }
        // ---- tail: remaining layers, error, backward deltas, for the lane's 4 rows
        float yh4[4], dl0[4][W0];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool valid = row0 + r < a.n;
            const float unscale = 1.f / (float)(1 << (2 * r));
            float act[NLA][MW], aux[NLA][MW];
#pragma unroll
            for (int c = 0; c < W0; ++c) act[0][c] = act_s<ACT>(fmaf(z[r][c], unscale, b0p[c]), aux[0][c]);
#pragma unroll
            for (int l = 1; l < NLA; ++l) {
#pragma unroll
                for (int c = 0; c < MW; ++c) {
                    if (c < T::width(l)) {
                        float zz = sp[T::b_off(l) + c];
#pragma unroll
                        for (int i = 0; i < MW; ++i)
                            if (i < T::in_w(l)) zz = fmaf(act[l - 1][i], sp[T::w_off(l) + c * T::in_w(l) + i], zz);
                        act[l][c] = act_s<ACT>(zz, aux[l][c]);
                    }
                }
            }
            float yh = 0.f;
#pragma unroll
            for (int i = 0; i < S; ++i) yh = fmaf(act[NLA - 1][i], sp[T::w_off(NLA) + i], yh);
            yh4[r] = yh;
            float tg = tg4[r];
            if (a.target_mode == TGT_RESID_PLUS_PRED) { tg = tg + yh; tg4[r] = tg; }   // net.rs:280
            const float e = valid ? yh - tg : 0.f;                                      // branch_sampler.rs:821
            rss = fmaf(e, e, rss);
            float delta[MW];
#pragma unroll
            for (int i = 0; i < S; ++i) {
                gWo[i] = fmaf(act[NLA - 1][i], e, gWo[i]);
                delta[i] = dact_s<ACT>(act[NLA - 1][i], aux[NLA - 1][i]) * (e * sp[T::w_off(NLA) + i]);
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
                    if (i < T::in_w(l)) delta[i] = dact_s<ACT>(act[l - 1][i], aux[l - 1][i]) * nd[i];
            }
#pragma unroll
            for (int c = 0; c < W0; ++c) {
                gb0[c] += delta[c];
                dl0[r][c] = delta[c] * unscale;      // backward decodes g * 4^r from the byte
            }
        }
        // ---- per-row outputs (vector stores when the row block is complete)
        {
            const bool full = vec_ok && row0 + 3 < a.n;
            auto put4 = [&](float* dst, const float v[4], int accumulate) {
                if (!dst) return;
                float* p = dst + eoff + row0;
                if (full && accumulate == 0) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
                else {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (row0 + r < a.n) {
                            if (accumulate > 0) p[r] += v[r];
                            else if (accumulate < 0) p[r] -= v[r];
                            else p[r] = v[r];
                        }
                }
            };
            if (a.target_mode == TGT_RESID_PLUS_PRED) {
                put4(a.tgt_out, tg4, 0);
                put4(a.prev_out, yh4, 0);      // net.rs:279
            }
            put4(a.yhat_out, yh4, a.yhat_accumulate);
        }
        if (!a.fwd_only) {
            // ---- stage delta_0 transposed: dT[c][row], one conflict-free float4 per unit
#pragma unroll
            for (int c = 0; c < W0; ++c)
                *reinterpret_cast<float4*>(dT + c * 128 + lane * 4) = make_float4(dl0[0][c], dl0[1][c], dl0[2][c], dl0[3][c]);
            __syncwarp();
            // ---- backward, first layer: lane = marker
            for (uint32_t q = 0; q < kTileQuads; ++q) {
                float dl[W0][4];
#pragma unroll
                for (int c = 0; c < W0; ++c) {
                    const float4 f = *reinterpret_cast<const float4*>(dT + c * 128 + q * 4);
                    dl[c][0] = f.x; dl[c][1] = f.y; dl[c][2] = f.z; dl[c][3] = f.w;
                }
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    if (p + 1 < NP || 32u * p < m) {   // only the last pass can be empty (dispatch picks the smallest NP)
                        // lanes beyond the row read neighbouring bytes of the CTA's own shared memory; their sums are dropped
                        const uint32_t byte = tileb[q * mp + 32 * p + lane];
                        float gf[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) gf[r] = (float)(byte & (3u << (2 * r)));      // g * 4^r
#pragma unroll
                        for (int c = 0; c < W0; ++c) {
                            if constexpr (PK) {
                                ffma2(acc0[p][c][0], acc0[p][c][1], gf[0], gf[1], dl[c][0], dl[c][1]);
                                ffma2(acc0[p][c][0], acc0[p][c][1], gf[2], gf[3], dl[c][2], dl[c][3]);
                            } else {
#pragma unroll
                                for (int r = 0; r < 4; ++r) acc0[p][c][0] = fmaf(gf[r], dl[c][r], acc0[p][c][0]);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
        stage ^= (nstage == 2) ? 1u : 0u;
        if (nstage == 1 && tn < t_end) issue_load(tn, 0);
    }
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
        for (int w = 0; w < NW; ++w) s += red[w * NTACC + tid];
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
    // first-layer weight gradient: sum over warps, then (S_jc - mu_j * gb0_c) / sd_j
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (32u * p < m) {
            __syncthreads();
#pragma unroll
            for (int c = 0; c < W0; ++c) red[(warp * W0 + c) * 32 + lane] = PK ? acc0[p][c][0] + acc0[p][c][PK ? 1 : 0] : acc0[p][c][0];
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
size_t k1_small_smem(uint32_t mp, int nstage) {
    using T = TailShape<H, S, D>;
    size_t fl = (size_t)mp * T::W2P + T::W0P + ((T::n_tail() + 3) & ~3);
    size_t redn = (size_t)NW * (T::NTACC > 32 * T::W0 ? T::NTACC : 32 * T::W0) + 4;
    size_t perw = (size_t)nstage * kTileQuads * (mp >> 2) + (size_t)T::W0 * 128;
    return (fl + redn + (size_t)NW * perw) * 4 + 16;
}

template <int H, int S, int D, int NP, int NW, int ACT>
int launch_one_small_act(K1Args& a, uint32_t nlist, uint32_t mp, cudaStream_t st) {
    int nstage = 2;
    size_t smem = k1_small_smem<H, S, D, NP, NW>(mp, 2);
    if (smem > 100 * 1024) { nstage = 1; smem = k1_small_smem<H, S, D, NP, NW>(mp, 1); }
    auto kern = k1_small<H, S, D, NP, NW, ACT>;
    static size_t configured = 0;
    if (smem > configured) {
        BANN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(a.nchunk, nlist);
    kern<<<grid, NW * 32, smem, st>>>(a, nstage);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

template <int H, int S, int D, int NP, int NW>
int launch_one_small(K1Args& a, uint32_t nlist, uint32_t mp, cudaStream_t st) {
    switch (a.act) {
        case BANN_TANH: return launch_one_small_act<H, S, D, NP, NW, BANN_TANH>(a, nlist, mp, st);
        case BANN_RELU: return launch_one_small_act<H, S, D, NP, NW, BANN_RELU>(a, nlist, mp, st);
        case BANN_LEAKY_RELU: return launch_one_small_act<H, S, D, NP, NW, BANN_LEAKY_RELU>(a, nlist, mp, st);
        case BANN_SILU: return launch_one_small_act<H, S, D, NP, NW, BANN_SILU>(a, nlist, mp, st);
        default: return launch_one_small_act<H, S, D, NP, NW, BANN_IDENTITY>(a, nlist, mp, st);
    }
}

// picks an instantiation for a homogeneous launch (all listed branches share the architecture and
// fit the marker bound); otherwise leaves *launched = false and the generic kernel runs.
#ifdef BANN_K1_SMALL_IMPL
int launch_k1_small(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                           cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net) {
    *launched = false;
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
    // fewer, fatter chunks than the generic kernel: every CTA keeps NW warps busy on its own tiles
    constexpr int NW = 8;
    uint32_t ntiles = a.ntiles;
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)num_sms * 2 + nlist - 1) / nlist);
    uint32_t nchunk = std::min<uint32_t>(want, std::max<uint32_t>(1, ntiles / NW));
    uint32_t tpc = (ntiles + nchunk - 1) / nchunk;
    nchunk = (ntiles + tpc - 1) / tpc;
#define BANN_TRY(HH, SS, DD, NPP)                                                                         \
    if (!*launched && H == HH && S == SS && D == DD && max_m <= 32u * NPP &&                              \
        k1_small_smem<HH, SS, DD, NPP, NW>(max_mp, 1) <= 200 * 1024) {                                    \
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
    BANN_TRY(2, 2, 0, 2)
#undef BANN_TRY
    return 0;
}
#else
int launch_k1_small(const std::vector<BranchDesc>& descs, int single_branch, K1Args& a, uint32_t nlist, int num_sms,
                           cudaStream_t st, bool* launched, uint32_t* nchunk_io, float** part_io, bann_net* net);
#endif

}  // namespace bann
