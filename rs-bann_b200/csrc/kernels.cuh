// Device kernels of the HMC hot path (sm_100a).
//
//  K1  k1_generic / k1_small<...>  fused branch forward + backward over packed genotypes:
//      replaces forward_feed + backpropagate (branch_sampler.rs:743-782,813-875), the
//      host decode + upload + standardise of bed.rs:325-355 and the N-vector passes of
//      net.rs:279-280.  One genotype read serves forward and backward.
//  KR  k_reduce_partials           fixed-order reduction of the per-CTA partials.
//  K2  k2_step                     gradient assembly under the prior (a8), momentum half steps
//      and position step (a9), Hamiltonian (a7, a11), early-reject and U-turn checks.
//  k_hmc_init / k_accept           step sizes (a10), momenta, accept/reject (a11).
#pragma once

#include "comm.cuh"
#include "common.cuh"

namespace bann {

// ------------------------------------------------------------------ activations
// activation_functions.rs:23-45.  dhdx is evaluated on the pre-activation.
__device__ __forceinline__ float act_h(int act, float x) {
    switch (act) {
        case BANN_TANH: return tanhf(x);
        case BANN_RELU: return x > 0.f ? x : 0.f;
        case BANN_LEAKY_RELU: return x > 0.f ? x : (x < 0.f ? 0.01f * x : 0.f);
        case BANN_SILU: return x * (1.f / (1.f + expf(-x)));
        default: return x;
    }
}
__device__ __forceinline__ float act_dh(int act, float x, float hx) {
    switch (act) {
        case BANN_TANH: return 1.f - hx * hx;
        case BANN_RELU: return x > 0.f ? 1.f : 0.f;
        case BANN_LEAKY_RELU: return x > 0.f ? 1.f : (x < 0.f ? 0.01f : 0.f);
        case BANN_SILU: {
            float sg = 1.f / (1.f + expf(-x));
            return hx + sg * (1.f - hx);
        }
        default: return 1.f;
    }
}

// deterministic block-wide sum (fixed tree), result valid in every thread
template <int NT>
__device__ __forceinline__ float block_sum(float v, float* scratch /* >= NT/32 floats */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) scratch[w] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) t += scratch[i];
    return t;
}

// ------------------------------------------------------------------ K1
enum : int { TGT_SHARED = 0, TGT_PER_ENTRY = 1, TGT_RESID_PLUS_PRED = 2 };

struct K1Args {
    const uint8_t* store;
    const BranchDesc* descs;
    const float* theta;        // arena
    const float* mu;           // gathered per-branch column means
    const float* sd;
    const uint32_t* list;      // branch ids of this launch (NULL: entry i == branch i)
    const BranchState* states; // skip entries whose status != RUNNING (NULL: run all)
    uint32_t n;                // local rows
    uint32_t ntiles;
    uint32_t tiles_per_chunk;
    uint32_t nchunk;
    int target_mode;
    const float* tgt;          // TGT_SHARED: [n]; TGT_PER_ENTRY: [entry][n]
    const float* resid;        // TGT_RESID_PLUS_PRED: target = resid + own prediction
    float* tgt_out;            //   ... written here (shared [n] or per entry)
    float* prev_out;           //   ... own prediction written here (may be NULL)
    int out_per_entry;         // tgt_out / prev_out / yhat_out indexed per entry
    float* yhat_out;           // optional prediction output
    int yhat_accumulate;       // +1: yhat_out[i] += yhat (Net::predict); -1: -= (initialize_stats); 0: =
    int fwd_only;
    float* part;               // [entry][chunk][pstride]; d_rss in param_vec order, rss at [P]
    uint32_t pstride;
    int act;
    // tensor-core path (k1_tc.cuh)
    const uint32_t* store_tc;  // NULL: no tensor-core store
    uint32_t nst;              // 256-row super-tiles
    uint32_t st_per_chunk;
    uint32_t ncb;              // 8-marker chunks per expanded-genotype buffer (max over the listed branches)
    uint32_t nc_uniform;       // chunks per row when every listed branch has the same count, else 0
    int tc_variant;            // k1_tc launches: BANN_TC_FOUR_WARPS / FIVE_WARPS / FIVE_WARPS_PLAIN (bann.h); set to 101 by the launcher when k1_tc5 ran
};

__device__ __forceinline__ void locate_param(const BranchDesc& d, uint32_t k, int& layer, uint32_t& row,
                                             uint32_t& col, bool& is_bias) {
    // param_vec order: weights of all layers (column-major), then biases
    const int nl = (int)d.nl;
    if (k >= d.b_off[0] && nl > 1) {
        is_bias = true;
        int l = 0;
        while (l + 1 < nl - 1 && k >= d.b_off[l + 1]) ++l;
        layer = l; col = k - d.b_off[l]; row = 0;
        return;
    }
    is_bias = false;
    int l = 0;
    while (l + 1 < nl && k >= d.w_off[l + 1]) ++l;
    layer = l;
    uint32_t r = k - d.w_off[l];
    col = r / d.in_dim[l];
    row = r % d.in_dim[l];
}

// Generic (any depth / widths / activation) fused forward+backward.  128 threads, one row per
// thread in the forward phase, one parameter per thread in the accumulation phase.  Slow but
// shape-agnostic; k1_small<> below is the tuned path for narrow branches.
#ifdef BANN_NET_TU   // non-template kernel: defined once, in net.cu's translation unit
__global__ void __launch_bounds__(128) k1_generic(K1Args a) {
    extern __shared__ float smf[];
    const uint32_t tid = threadIdx.x;
    const uint32_t li = blockIdx.y, chunk = blockIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    if (a.states && a.states[b].status != ST_RUNNING) return;
    const BranchDesc& d = a.descs[b];
    const uint32_t P = d.P, m = d.m, mp = d.m_pad4, nl = d.nl, w0 = d.widths[0];
    const uint32_t SW = d.sumw | 1u;
    float* sp = smf;
    float* as_ = sp + ((P + 3) & ~3u);
    float* ds_ = as_ + 128 * SW;
    float* es_ = ds_ + 128 * SW;
    float* red = es_ + 128;      // 8
    float* b0p = red + 8;        // w0
    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;

    for (uint32_t k = tid; k < P; k += 128) sp[k] = th[k];
    __syncthreads();
    // fold the standardisation (bed.rs:354) into the first layer: W' = W / sd, b' = b - sum mu W'
    for (uint32_t k = tid; k < m * w0; k += 128) sp[d.w_off[0] + k] = __fdiv_rn(sp[d.w_off[0] + k], sd[k % m]);
    __syncthreads();
    for (uint32_t c = tid; c < w0; c += 128) {
        float acc = 0.f;
        for (uint32_t j = 0; j < m; ++j) acc = fmaf(mu[j], sp[d.w_off[0] + c * m + j], acc);
        b0p[c] = sp[d.b_off[0] + c] - acc;
    }
    float* pp = a.part ? a.part + ((size_t)li * a.nchunk + chunk) * a.pstride : nullptr;
    if (pp)
        for (uint32_t k = tid; k <= P; k += 128) pp[k] = 0.f;
    __syncthreads();

    const size_t eoff = a.out_per_entry ? (size_t)li * a.n : 0;
    const uint32_t t_begin = chunk * a.tiles_per_chunk;
    const uint32_t t_end = min(a.ntiles, t_begin + a.tiles_per_chunk);
    for (uint32_t t = t_begin; t < t_end; ++t) {
        const uint8_t* tile = a.store + d.tile_off + (size_t)t * (kTileQuads * mp);
        const uint32_t r = tid, q = r >> 2, sh = 2 * (r & 3);
        const uint32_t row = t * kTileRows + r;
        const bool valid = row < a.n;
        float* ar = as_ + r * SW;
        float* dr = ds_ + r * SW;
        // ---- forward, first layer (packed genotypes)
        for (uint32_t c0 = 0; c0 < w0; c0 += 8) {
            float z[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) z[k] = (c0 + k < w0) ? b0p[c0 + k] : 0.f;
            for (uint32_t j = 0; j < m; ++j) {
                const uint32_t g = (tile[q * mp + j] >> sh) & 3u;
                if (g) {
                    const float gf = (float)g;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (c0 + k < w0) z[k] = fmaf(gf, sp[d.w_off[0] + (c0 + k) * m + j], z[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c0 + k < w0) {
                    float h = act_h(a.act, z[k]);
                    ar[d.a_off[0] + c0 + k] = h;
                    dr[d.a_off[0] + c0 + k] = act_dh(a.act, z[k], h);
                }
        }
        // ---- forward, remaining activated layers
        for (uint32_t l = 1; l + 1 < nl; ++l) {
            const uint32_t in = d.in_dim[l], out = d.widths[l];
            for (uint32_t c = 0; c < out; ++c) {
                float z = sp[d.b_off[l] + c];
                for (uint32_t i = 0; i < in; ++i) z = fmaf(ar[d.a_off[l - 1] + i], sp[d.w_off[l] + c * in + i], z);
                float h = act_h(a.act, z);
                ar[d.a_off[l] + c] = h;
                dr[d.a_off[l] + c] = act_dh(a.act, z, h);
            }
        }
        // ---- output neuron (no bias, no activation)
        const uint32_t L = nl - 2, sL = d.widths[L];
        float yh = 0.f;
        for (uint32_t i = 0; i < sL; ++i) yh = fmaf(ar[d.a_off[L] + i], sp[d.w_off[nl - 1] + i], yh);
        float tg = 0.f;
        if (valid) {
            if (a.target_mode == TGT_RESID_PLUS_PRED) {
                tg = a.resid[row] + yh;                         // net.rs:280
                if (a.tgt_out) a.tgt_out[eoff + row] = tg;
                if (a.prev_out) a.prev_out[eoff + row] = yh;    // net.rs:279
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
        const float e = valid ? yh - tg : 0.f;                  // branch_sampler.rs:821
        es_[r] = e;
        // ---- backward deltas (row local)
        for (uint32_t i = 0; i < sL; ++i) dr[d.a_off[L] + i] *= e * sp[d.w_off[nl - 1] + i];
        for (uint32_t l = L; l >= 1; --l) {
            const uint32_t in = d.in_dim[l], out = d.widths[l];
            for (uint32_t i = 0; i < in; ++i) {
                float err = 0.f;
                for (uint32_t c = 0; c < out; ++c) err = fmaf(dr[d.a_off[l] + c], sp[d.w_off[l] + c * in + i], err);
                dr[d.a_off[l - 1] + i] *= err;
            }
        }
        __syncthreads();
        // ---- accumulate d_rss: one parameter per thread, rows in fixed order
        for (uint32_t k = tid; k < P; k += 128) {
            int l; uint32_t i, c; bool isb;
            locate_param(d, k, l, i, c, isb);
            float s = 0.f;
            if (isb) {
                for (uint32_t rr = 0; rr < 128; ++rr) s += ds_[rr * SW + d.a_off[l] + c];
            } else if (l == (int)nl - 1) {
                for (uint32_t rr = 0; rr < 128; ++rr) s = fmaf(as_[rr * SW + d.a_off[L] + i], es_[rr], s);
            } else if (l >= 1) {
                for (uint32_t rr = 0; rr < 128; ++rr)
                    s = fmaf(as_[rr * SW + d.a_off[l - 1] + i], ds_[rr * SW + d.a_off[l] + c], s);
            } else {
                for (uint32_t qq = 0; qq < kTileQuads; ++qq) {
                    const uint32_t byte = tile[qq * mp + i];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint32_t g = (byte >> (2 * u)) & 3u;
                        if (g) s = fmaf((float)g, ds_[(qq * 4 + u) * SW + d.a_off[0] + c], s);
                    }
                }
            }
            pp[k] += s;
        }
        float e2 = block_sum<128>(e * e, red);
        if (tid == 0) pp[P] += e2;
        __syncthreads();
    }
    if (a.fwd_only || !pp) return;
    __syncthreads();
    // undo the folding for the first-layer weight gradient:
    // d/dW[j,c] = (sum_i g_ij delta_ic - mu_j * sum_i delta_ic) / sd_j
    for (uint32_t k = tid; k < m * w0; k += 128) {
        const uint32_t j = k % m, c = k / m;
        pp[d.w_off[0] + k] = __fdiv_rn(pp[d.w_off[0] + k] - mu[j] * pp[d.b_off[0] + c], sd[j]);
    }
}
#endif

inline size_t k1_generic_smem(const BranchDesc& d) {
    size_t SW = d.sumw | 1u;
    return (((size_t)d.P + 3) & ~3ull) * 4 + 2 * 128 * SW * 4 + 128 * 4 + 8 * 4 + (size_t)d.widths[0] * 4 + 16;
}

// ------------------------------------------------------------------ KR
// gsum[entry][k] = sum over chunks (ascending) of part[entry][chunk][k], f64 accumulation.
// One block = 32 consecutive entries k x 8 warps; warp w sums the chunks c = w, w + 8, ... (two interleaved f64 running
// sums, coalesced 128-byte loads), then the eight warp sums are added in ascending order.  Fixed order => deterministic;
// enough blocks and loads in flight that a single-branch launch with hundreds of chunks reduces in a few microseconds
// (a thread-per-entry loop over 391 chunks took 60 us: one L2 round trip per chunk).
#ifdef BANN_NET_TU   // non-template kernel: defined once, in net.cu's translation unit
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ part, float* __restrict__ gsum, uint32_t nchunk,
                                                         uint32_t pstride, const uint32_t* __restrict__ list,
                                                         const BranchDesc* __restrict__ descs,
                                                         const BranchState* __restrict__ states, XrComm xc) {
    __shared__ double sm[8][32];
    pdl_wait();                    // the partials come from the kernel launched just before
    pdl_launch_dependents();
    const uint32_t li = blockIdx.y;
    const uint32_t b = list ? list[li] : li;
    // single rank: nothing to do for an early-rejected / finished branch.  Sharded rows: the host has handed out an epoch for
    // this launch, so the exchange below must run on every rank whatever the status (comm.cuh; the sums are not used then).
    if (states && states[b].status != ST_RUNNING && xc.world <= 1) return;
    const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t k = blockIdx.x * 32 + lane;
    const bool live = k <= descs[b].P;
    double s0 = 0.0, s1 = 0.0;
    if (live) {
        const float* p = part + (size_t)li * nchunk * pstride + k;
        uint32_t c = w;
        for (; c + 8 < nchunk; c += 16) {
            const float v0 = p[(size_t)c * pstride], v1 = p[(size_t)(c + 8) * pstride];
            s0 += (double)v0;
            s1 += (double)v1;
        }
        if (c < nchunk) s0 += (double)p[(size_t)c * pstride];
    }
    sm[w][lane] = s0 + s1;
    __syncthreads();
    if (w == 0 && live) {
        double s = sm[0][lane];
#pragma unroll
        for (int i = 1; i < 8; ++i) s += sm[i][lane];
        // row-sharded sequential schedule: the sum over ranks happens right here (comm.cuh), rank order
        gsum[(size_t)li * pstride + k] = xr_sum(xc, li * pstride + k, (float)s);
    }
}
#endif

// ------------------------------------------------------------------ prior helpers
__device__ __forceinline__ float param_prior_precision(const BranchDesc& d, const float* prec, int model, int layer,
                                                       uint32_t row, bool is_bias) {
    if (is_bias) return prec[d.bp_off[layer]];
    if (model == BANN_STD_NORMAL) return 1.f;
    const bool ard = (model == BANN_RIDGE_ARD || model == BANN_LASSO_ARD) && layer < (int)d.nl - 1;
    return prec[d.wp_off[layer] + (ard ? row : 0)];
}

struct K2Args {
    const BranchDesc* descs;
    const uint32_t* list;
    BranchState* states;
    float* theta;
    float* theta0;
    float* mom;
    float* grad;
    const float* eps;
    const float* prec;
    const float* gsum;     // [entry][pstride]
    uint32_t pstride;
    int model;
    float max_h_err;
    int mode_init;         // 1: first evaluation (H_init), no momentum update before the check
    int is_last;           // 1: do not start the next leapfrog step
    float* traj_params;    // optional [L][P] (single-entry launches only)
    float* traj_ldg;
    float* traj_h;         // [L+1]
    const float* num_ldg;  // --num-grad (single-entry launches): the numerical log-density gradient replaces the analytical one
};

// One block per branch.  Reference sequence per leapfrog step (branch_sampler.rs:1239-1284):
//   p += e/2 g; theta += e p; g = grad(theta); p += e/2 g; H check; U-turn
// This kernel runs [g = grad; p += e/2 g; H check; U-turn; p += e/2 g; theta += e p], i.e. the
// same sequence cut after the position update, so that K1 is the only pass over the data.
#ifdef BANN_NET_TU   // non-template kernel: defined once, in net.cu's translation unit
__global__ void __launch_bounds__(256) k2_step(K2Args a) {
    __shared__ float red[8];
    __shared__ int s_status;
    __shared__ BranchDesc s_desc;
    const uint32_t li = blockIdx.x, tid = threadIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    pdl_launch_dependents();       // the next K1 may run its shared-memory / tensor-memory prologue under this kernel
    pdl_wait();                    // gsum comes from the kernel launched just before
    BranchState& st = a.states[b];
    if (st.status != ST_RUNNING) return;
    // the descriptor in shared memory: locate_param walks its offset tables for every parameter, and from global memory every
    // step of that walk was a dependent L2 round trip on the critical path of a leapfrog step
    static_assert(sizeof(BranchDesc) % 4 == 0 && sizeof(BranchDesc) / 4 <= 256, "descriptor copy assumes <= 256 words");
    if (tid < sizeof(BranchDesc) / 4) reinterpret_cast<uint32_t*>(&s_desc)[tid] = reinterpret_cast<const uint32_t*>(&a.descs[b])[tid];
    __syncthreads();
    const BranchDesc& d = s_desc;
    const uint32_t P = d.P;
    float* th = a.theta + d.param_off;
    const float* th0 = a.theta0 + d.param_off;
    float* p = a.mom + d.param_off;
    float* gr = a.grad + d.param_off;
    const float* ep = a.eps + d.param_off;
    const float* pr = a.prec + d.prec_off;
    const float* gs = a.gsum + (size_t)li * a.pstride;
    const float lam_e = pr[d.ep_off];
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);

    float kin = 0.f, prior = 0.f, uturn = 0.f;
    for (uint32_t k = tid; k < P; k += 256) {
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        const float w = th[k];
        float g;
        if (isb) {
            g = -(lam_e * gs[k]);                                  // branch_sampler.rs:322-331
            if (a.model == BANN_STD_NORMAL) prior -= 0.5f * w * w; // std_normal_branch.rs:153-156 (Q5)
        } else {
            const float lam = param_prior_precision(d, pr, a.model, l, row, false);
            if (a.model == BANN_STD_NORMAL) {
                g = -(lam_e * gs[k] + w);                          // std_normal_branch.rs:160-169
                prior -= 0.5f * w * w;
            } else if (lasso) {
                const float sg = (w > 0.f) ? 1.f : (w < 0.f ? -1.f : 0.f);  // af_helpers.rs:53-58
                g = -(lam_e * gs[k] + lam * sg);                   // lasso_base.rs:175-185, lasso_ard.rs:196-218
                prior -= lam * fabsf(w);
            } else {
                g = -(lam_e * gs[k] + lam * w);                    // ridge_base.rs:175-184, ridge_ard.rs:196-219
                prior -= 0.5f * lam * w * w;
            }
        }
        if (a.num_ldg) g = a.num_ldg[k];                           // branch_sampler.rs:1232-1247 (mcmc_cfg.num_grad)
        gr[k] = g;
        float pk = p[k];
        if (!a.mode_init) {
            pk = pk + (0.5f * ep[k]) * g;                          // momentum.rs:129-136
            p[k] = pk;
        }
        kin = fmaf(pk, pk, kin);
        uturn = fmaf(w - th0[k], pk, uturn);
    }
    kin = block_sum<256>(kin, red);
    prior = block_sum<256>(prior, red);
    uturn = block_sum<256>(uturn, red);
    if (tid == 0) {
        const float rss = gs[P];
        const float wrt_e = -1.0f * lam_e * (rss / 2.0f);          // branch_sampler.rs:100-102
        const float ld = prior + wrt_e;
        const float negh = ld - 0.5f * kin;                        // branch_sampler.rs:878-883
        st.rss = rss;
        st.log_density = ld;
        st.neg_h_cur = negh;
        if (a.mode_init) {
            st.neg_h_init = negh;
            if (a.traj_h) a.traj_h[0] = negh;
        } else {
            const int step = st.steps_done;
            st.steps_done = step + 1;
            if (a.traj_h) a.traj_h[step + 1] = negh;
            if (fabsf(negh - st.neg_h_init) > a.max_h_err) st.status = ST_REJECTED_EARLY;   // :1264-1279
            else if (st.u_turn_step < 0 && uturn < 0.f) st.u_turn_step = step;                // :1281-1284
        }
        s_status = st.status;
    }
    __syncthreads();
    const int status = s_status;
    if (!a.mode_init && a.traj_params) {
        const int step = st.steps_done - 1;
        for (uint32_t k = tid; k < P; k += 256) {
            a.traj_params[(size_t)step * P + k] = th[k];
            a.traj_ldg[(size_t)step * P + k] = gr[k];
        }
    }
    if (status == ST_REJECTED_EARLY) {
        for (uint32_t k = tid; k < P; k += 256) th[k] = th0[k];   // :1277
        return;
    }
    if (!a.is_last) {
        for (uint32_t k = tid; k < P; k += 256) {
            const float e = ep[k];
            const float pk = p[k] + (0.5f * e) * gr[k];
            p[k] = pk;
            th[k] = th[k] + e * pk;                               // params.rs:728-738
        }
    }
}
#endif

// ------------------------------------------------------------------ HMC init / accept
struct InitArgs {
    const BranchDesc* descs;
    const uint32_t* list;
    BranchState* states;
    const float* theta;
    float* theta0;
    float* mom;
    float* eps;
    const float* prec;
    int model;
    int step_mode;
    float factor;
    float L;
    const float* inj_momenta;        // single-entry launches: [P]; inj_arena: the parameter arena layout (entry at param_off)
    const float* inj_step_uniforms;  // the same
    int inj_arena;
    uint64_t seed;
    uint64_t stream_base;            // Philox stream = stream_base + branch
    int keep_momenta;                // momenta already in place (bann_leapfrog_host)
};

__device__ __forceinline__ void philox_normal_pair(Philox& ph, float& a, float& b) {
    uint32_t r[4];
    ph.next(r);
    const float u1 = u01_open(r[0]), u2 = u01_half_open(r[1]);
    const float rad = sqrtf(-2.f * logf(u1));
    float s, c;
    sincospif(2.f * u2, &s, &c);
    a = rad * c;
    b = rad * s;
}

__device__ __forceinline__ float step_size_for(const InitArgs& a, const BranchDesc& d, const float* pr, int l, uint32_t row, bool isb,
                               float u) {
    const float PI = 3.14159265358979323846f;  // std::f32::consts::PI
    const float f = a.factor, L = a.L;
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    switch (a.step_mode) {
        case BANN_STEP_UNIFORM: return f;                                             // branch_sampler.rs:706-732
        case BANN_STEP_RANDOM: return __fmul_rn(u, __fmul_rn(powf((float)d.P, -0.25f), f));  // :654-681
        case BANN_STEP_STD_SCALED: {                                                   // ridge_base.rs:52-80
            if (isb) return __fmul_rn(f, __fdiv_rn(1.f, __fsqrt_rn(pr[d.bp_off[l]])));
            return __fmul_rn(f, __fsqrt_rn(__fdiv_rn(1.f, pr[d.wp_off[l]])));
        }
        default: break;
    }
    // Izmailov
    if (isb) {
        const float lam = pr[d.bp_off[l]];
        const float den = __fmul_rn(__fmul_rn(2.f, __fsqrt_rn(lam)), L);
        if (a.model == BANN_STD_NORMAL) return __fdiv_rn(PI, den);                     // std_normal_branch.rs:95-103
        if (a.model == BANN_LASSO_ARD) return __fdiv_rn(__fmul_rn(f, PI), den);        // lasso_ard.rs:105-112
        return __fmul_rn(f, __fdiv_rn(PI, den));                                       // ridge_base.rs:96-106 etc.
    }
    const bool ard_layer = (a.model == BANN_RIDGE_ARD || a.model == BANN_LASSO_ARD) && l < (int)d.nl - 1;
    const float lam = pr[d.wp_off[l] + (ard_layer ? row : 0)];
    if (a.model == BANN_STD_NORMAL) return __fdiv_rn(PI, __fmul_rn(__fmul_rn(2.f, __fsqrt_rn(lam)), L));
    if (lasso) {
        const float den = __fmul_rn(__fmul_rn(4.f, lam), L);
        if (ard_layer) return __fmul_rn(f, __fdiv_rn(1.f, den));                       // lasso_ard.rs:83-94
        return __fdiv_rn(f, den);                                                      // lasso_base.rs:89-96
    }
    const float den = __fmul_rn(__fmul_rn(2.f, __fsqrt_rn(lam)), L);
    if (ard_layer) return __fmul_rn(f, __fdiv_rn(PI, den));                            // ridge_ard.rs:76-87
    return __fdiv_rn(__fmul_rn(f, PI), den);                                           // ridge_base.rs:87-94
}

#ifdef BANN_NET_TU   // non-template kernel: defined once, in net.cu's translation unit
__global__ void __launch_bounds__(256) k_hmc_init(InitArgs a) {
    const uint32_t li = blockIdx.x, tid = threadIdx.x;
    const uint32_t b = a.list ? a.list[li] : li;
    const BranchDesc& d = a.descs[b];
    const uint32_t P = d.P;
    const float* th = a.theta + d.param_off;
    float* th0 = a.theta0 + d.param_off;
    float* p = a.mom + d.param_off;
    float* ep = a.eps + d.param_off;
    const float* pr = a.prec + d.prec_off;
    const float* inj_mom = a.inj_momenta ? a.inj_momenta + (a.inj_arena ? d.param_off : 0) : nullptr;
    const float* inj_su = a.inj_step_uniforms ? a.inj_step_uniforms + (a.inj_arena ? d.param_off : 0) : nullptr;
    // momenta: pairs (2k, 2k+1) from one Philox block each -> independent of the thread count
    for (uint32_t k2 = tid; 2 * k2 < P; k2 += 256) {
        const uint32_t k = 2 * k2;
        if (a.keep_momenta) break;
        if (inj_mom) {
            p[k] = inj_mom[k];
            if (k + 1 < P) p[k + 1] = inj_mom[k + 1];
        } else {
            Philox ph(a.seed, a.stream_base + b, (uint64_t)k2);
            float x, y;
            philox_normal_pair(ph, x, y);
            p[k] = x;
            if (k + 1 < P) p[k + 1] = y;
        }
    }
    for (uint32_t k = tid; k < P; k += 256) {
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        float u = 0.f;
        if (a.step_mode == BANN_STEP_RANDOM) {
            if (inj_su) u = inj_su[k];
            else {
                Philox ph(a.seed ^ 0x5bd1e995u, a.stream_base + b, (uint64_t)k);
                uint32_t r[4];
                ph.next(r);
                u = u01_half_open(r[0]);
            }
        }
        ep[k] = step_size_for(a, d, pr, l, row, isb, u);
        th0[k] = th[k];
    }
    if (tid == 0) {
        BranchState& st = a.states[b];
        st.status = ST_RUNNING;
        st.steps_done = 0;
        st.u_turn_step = -1;
        st.neg_h_init = st.neg_h_cur = st.log_density = st.rss = st.log_acc = 0.f;
    }
}
#endif

// accept_or_reject_hmc_state (branch_sampler.rs:928-962) on the Hamiltonian of the last step
// (the reference recomputes the forward pass; the value is the same, Q13).
#ifdef BANN_NET_TU   // non-template kernel: defined once, in net.cu's translation unit
__global__ void __launch_bounds__(256) k_accept(const BranchDesc* descs, const uint32_t* list, BranchState* states,
                                                float* theta, const float* theta0, const float* inj_u, uint64_t seed,
                                                uint64_t stream_base) {
    __shared__ int s_status;
    const uint32_t li = blockIdx.x, tid = threadIdx.x;
    const uint32_t b = list ? list[li] : li;
    BranchState& st = states[b];
    if (tid == 0) {
        if (st.status == ST_RUNNING) {
            const float log_acc = st.neg_h_cur - st.neg_h_init;
            const float prob = (log_acc >= 0.f) ? 1.f : expf(log_acc);
            float u;
            if (inj_u) u = inj_u[li];
            else {
                Philox ph(seed ^ 0xa511e9b3u, stream_base + b, 0);
                uint32_t r[4];
                ph.next(r);
                u = u01_half_open(r[0]);
            }
            st.log_acc = log_acc;
            st.status = (u < prob) ? ST_ACCEPTED : ST_REJECTED;   // :546-548 (NaN -> rejected)
        }
        s_status = st.status;
    }
    __syncthreads();
    if (s_status == ST_REJECTED) {
        const BranchDesc& d = descs[b];
        for (uint32_t k = tid; k < d.P; k += 256) theta[d.param_off + k] = theta0[d.param_off + k];  // :1293-1296
    }
}
#endif

}  // namespace bann
