// Chain-level device kernels: Gibbs precision draws (a12), residual / output-bias bookkeeping
// and log posterior density (a13).  All tiny; they exist so that a branch visit of
// Net::train (net.rs:258-334) needs no host synchronisation.
#pragma once

#include "kernels.cuh"

namespace bann {

// ------------------------------------------------------------------ gamma variates
__device__ __forceinline__ float philox_u(Philox& ph) {
    uint32_t r[4];
    ph.next(r);
    return u01_open(r[0]);
}

// Marsaglia-Tsang; shape < 1 boosted.  (rand_distr::Gamma uses the same method; its stream is
// third-party and not reproduced -- parity runs inject the variates.)
__device__ float philox_std_gamma(Philox& ph, float shape) {
    float boost = 1.f;
    if (shape < 1.f) {
        boost = powf(philox_u(ph), 1.f / shape);
        shape += 1.f;
    }
    const float dd = shape - 1.f / 3.f;
    const float c = 1.f / sqrtf(9.f * dd);
    for (int it = 0; it < 1000; ++it) {
        float x, y;
        philox_normal_pair(ph, x, y);
        float v = 1.f + c * x;
        if (v <= 0.f) continue;
        v = v * v * v;
        const float u = philox_u(ph);
        const float x2 = x * x;
        if (u < 1.f - 0.0331f * x2 * x2) return boost * dd * v;
        if (logf(u) < 0.5f * x2 + dd * (1.f - v + logf(v))) return boost * dd * v;
    }
    return boost * dd;
}

struct GibbsArgs {
    const BranchDesc* descs;
    const uint32_t* list;       // group visits: one block per listed branch; NULL: the single branch `b`
    uint32_t b;
    const float* theta;
    float* prec;
    NetGlobals* G;              // read only here: a group's members all see the globals frozen at group start
    float* ow_others;           // [B] out: global output-weight stat minus the branch's own (branch_struct.rs:27)
    float* own_old;             // [B] out: the branch's own output-weight stat before the transition
    Hyper6 hyper;
    int model;
    float n_total;
    int fixed_param_precisions;
    int do_draws;               // 0: only cfg.update_global_params + ow_others (from_cfg)
    const float* inj;           // injected standard-gamma variates (consumption order) or NULL; entry li at inj + li * inj_stride
    uint32_t n_inj;
    uint32_t inj_stride;
    uint64_t seed;
    uint64_t stream_base;       // Philox stream = stream_base + branch
};

__device__ __forceinline__ float gamma_variate(const GibbsArgs& a, uint32_t idx, float shape) {
    if (a.inj) return idx < a.n_inj ? a.inj[(size_t)blockIdx.x * a.inj_stride + idx] : 1.f;
    const uint32_t b = a.list ? a.list[blockIdx.x] : a.b;
    Philox ph(a.seed ^ 0x9e3779b97f4a7c15ull, a.stream_base + b, (uint64_t)idx * 4096ull);
    return philox_std_gamma(ph, shape);
}
// gibbs_steps.rs:76-94,115-129
__device__ __forceinline__ float ridge_post(const GibbsArgs& a, uint32_t idx, float k, float s, float stat, float n) {
    const float shape = k + n / 2.f;
    const float scale = 2.f * s / (2.f + s * stat);
    return gamma_variate(a, idx, shape) * scale;
}
// gibbs_steps.rs:25-57
__device__ __forceinline__ float lasso_post(const GibbsArgs& a, uint32_t idx, float k, float s, float stat, float n) {
    const float shape = k + n;
    const float scale = s / (1.f + s * stat);
    return gamma_variate(a, idx, shape) * scale;
}

// One block.  cfg.update_global_params (branch_cfg.rs:59-63), from_cfg's subtraction of the own
// output-weight statistic, sample_error_precision (branch_sampler.rs:190-202),
// sample_prior_precisions (per prior) and sample_output_weight_precisions (:178-188).
__global__ void __launch_bounds__(256) k_gibbs(GibbsArgs a) {
    __shared__ float red[8];
    const uint32_t tid = threadIdx.x;
    const uint32_t bix = a.list ? a.list[blockIdx.x] : a.b;
    const BranchDesc& d = a.descs[bix];
    const float* th = a.theta + d.param_off;
    float* pr = a.prec + d.prec_off;
    const int nl = (int)d.nl, last = nl - 1;
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    const bool ard = (a.model == BANN_RIDGE_ARD || a.model == BANN_LASSO_ARD);

    // own output-weight statistic (summary_stat_fn; StdNormal's device variant is a sum of squares)
    float own = 0.f;
    for (uint32_t i = tid; i < d.in_dim[last]; i += 256) {
        const float w = th[d.w_off[last] + i];
        own += lasso ? fabsf(w) : w * w;
    }
    own = block_sum<256>(own, red);
    const float others = a.G->ow_reg_sum - own;
    __syncthreads();
    if (tid == 0) {
        a.ow_others[bix] = others;
        if (a.own_old) a.own_old[bix] = own;
        pr[d.ep_off] = a.G->error_precision;
        pr[d.wp_off[last]] = a.G->output_layer_precision;
    }
    if (!a.do_draws) return;
    uint32_t idx = 0;
    if (tid == 0) {
        float k, s;
        layer_prior(a.hyper, last, nl, k, s);   // Q9: output-layer hyperparameters
        pr[d.ep_off] = ridge_post(a, idx, k, s, a.G->resid_ss, a.n_total);
    }
    idx += 1;
    if (a.fixed_param_precisions) return;
    if (a.model == BANN_STD_NORMAL) {
        if (tid == 0) pr[d.wp_off[last]] = 1.0f;   // std_normal_branch.rs:178-189
        return;
    }
    for (int l = 0; l < last; ++l) {
        float k, s;
        layer_prior(a.hyper, l, nl, k, s);
        const uint32_t in = d.in_dim[l], out = d.widths[l];
        const float* W = th + d.w_off[l];
        if (ard) {   // ridge_ard.rs:271-292, lasso_ard.rs:268-289: one group per input row
            for (uint32_t r = tid; r < in; r += 256) {
                float stat = 0.f;
                for (uint32_t c = 0; c < out; ++c) {
                    const float w = W[c * in + r];
                    stat += lasso ? fabsf(w) : w * w;
                }
                pr[d.wp_off[l] + r] = lasso ? lasso_post(a, idx + r, k, s, stat, (float)out)
                                            : ridge_post(a, idx + r, k, s, stat, (float)out);
            }
            idx += in;
        } else {     // ridge_base.rs:235-245, lasso_base.rs:235-245: whole layer
            float stat = 0.f;
            for (uint32_t i = tid; i < in * out; i += 256) {
                const float w = W[i];
                stat += lasso ? fabsf(w) : w * w;
            }
            stat = block_sum<256>(stat, red);
            if (tid == 0)
                pr[d.wp_off[l]] = lasso ? lasso_post(a, idx, k, s, stat, (float)(in * out))
                                        : ridge_post(a, idx, k, s, stat, (float)(in * out));
            idx += 1;
        }
        float bs = 0.f;   // biases are always ridge
        for (uint32_t c = tid; c < out; c += 256) {
            const float v = th[d.b_off[l] + c];
            bs += v * v;
        }
        bs = block_sum<256>(bs, red);
        if (tid == 0) pr[d.bp_off[l]] = ridge_post(a, idx, k, s, bs, (float)out);
        idx += 1;
    }
    if (tid == 0) {
        float k, s;
        layer_prior(a.hyper, last, nl, k, s);
        const float stat = others + own;
        pr[d.wp_off[last]] = lasso ? lasso_post(a, idx, k, s, stat, a.G->ow_num_params)
                                   : ridge_post(a, idx, k, s, stat, a.G->ow_num_params);
    }
}

// ------------------------------------------------------------------ residual bookkeeping
// r = t - (accepted ? y_new : y_prev)   (net.rs:295,299), block partials of sum r^2 and
// sum (r + bias_old).
__global__ void __launch_bounds__(256) k_resid_after_hmc(float* __restrict__ r, const float* __restrict__ t,
                                                         const float* __restrict__ ynew,
                                                         const float* __restrict__ yprev, uint32_t n,
                                                         const BranchState* st, const NetGlobals* G,
                                                         float* __restrict__ part /* [2*gridDim.x] */) {
    __shared__ float red[8];
    const bool acc = st->status == ST_ACCEPTED;
    const float bias = G->output_bias;
    float ss = 0.f, sb = 0.f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float v = t[i] - (acc ? ynew[i] : yprev[i]);
        r[i] = v;
        ss = fmaf(v, v, ss);
        sb += v + bias;   // net.rs:321
    }
    ss = block_sum<256>(ss, red);
    sb = block_sum<256>(sb, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = ss;
        part[2 * blockIdx.x + 1] = sb;
    }
}

// r = (r + bias_old) - bias_new (net.rs:321,332), partials of sum r^2 / sum r of the result
__global__ void __launch_bounds__(256) k_resid_apply_bias(float* __restrict__ r, uint32_t n, const float* bias_old_new,
                                                          float* __restrict__ part) {
    __shared__ float red[8];
    const float bo = bias_old_new[0], bn = bias_old_new[1];
    float ss = 0.f, s = 0.f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float v = (r[i] + bo) - bn;
        r[i] = v;
        ss = fmaf(v, v, ss);
        s += v;
    }
    ss = block_sum<256>(ss, red);
    s = block_sum<256>(s, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = ss;
        part[2 * blockIdx.x + 1] = s;
    }
}

// plain statistics of the residual (used by initialize_stats and after set_targets)
__global__ void __launch_bounds__(256) k_resid_stats(const float* __restrict__ r, uint32_t n, float* __restrict__ part) {
    __shared__ float red[8];
    float ss = 0.f, s = 0.f;
    for (uint32_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const float v = r[i];
        ss = fmaf(v, v, ss);
        s += v;
    }
    ss = block_sum<256>(ss, red);
    s = block_sum<256>(s, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = ss;
        part[2 * blockIdx.x + 1] = s;
    }
}

__global__ void k_resid_reduce(const float* __restrict__ part, uint32_t nblk, NetGlobals* G, XrComm xc) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double ss = 0.0, s = 0.0;
        for (uint32_t i = 0; i < nblk; ++i) {
            ss += part[2 * i];
            s += part[2 * i + 1];
        }
        ss = xr_sum(xc, 0, ss);
        s = xr_sum(xc, 1, s);
        G->resid_ss = (float)ss;
        G->resid_sum = (float)s;
    }
}

__global__ void k_init_residual(float* __restrict__ r, const float* __restrict__ y, uint32_t n, const NetGlobals* G) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = y[i] - G->output_bias;   // net.rs:160
}

// ------------------------------------------------------------------ LPD of one branch
// log_density_joint_components_curr_internal_state (branch_sampler.rs:307-318) for the four
// ridge/lasso priors; StdNormal is unimplemented in the reference (Q6/H10) -- extension:
// local = -1/2 sum theta^2, output term 0.
__device__ void branch_lpd_terms(const BranchDesc& d, const float* th, const float* pr, int model, const Hyper6& hy,
                                 float ow_others, float ow_num, float* red, float& out_w, float& local) {
    const uint32_t tid = threadIdx.x;
    const int nl = (int)d.nl, last = nl - 1;
    const bool lasso = (model == BANN_LASSO_BASE || model == BANN_LASSO_ARD);
    const bool ard = (model == BANN_RIDGE_ARD || model == BANN_LASSO_ARD);
    if (model == BANN_STD_NORMAL) {
        float s = 0.f;
        for (uint32_t k = tid; k < d.P; k += 256) s += th[k] * th[k];
        s = block_sum<256>(s, red);
        local = -0.5f * s;
        out_w = 0.f;
        return;
    }
    float ld_b = 0.f, ld_w = 0.f;
    for (int l = 0; l < last; ++l) {
        float shape, scale;
        layer_prior(hy, l, nl, shape, scale);
        const uint32_t in = d.in_dim[l], out = d.widths[l];
        // biases: branch_sampler.rs:260-279
        float bs = 0.f;
        for (uint32_t c = tid; c < out; c += 256) bs += th[d.b_off[l] + c] * th[d.b_off[l] + c];
        bs = block_sum<256>(bs, red);
        const float lb = pr[d.bp_off[l]];
        ld_b -= lb * (bs / 2.f + 1.f / scale);
        ld_b += (shape + ((float)out - 2.f) / 2.f) * logf(lb);
        const float* W = th + d.w_off[l];
        if (ard) {   // ridge_ard.rs:119-148, lasso_ard.rs:123-151
            float t1 = 0.f, t2 = 0.f;
            for (uint32_t r = tid; r < in; r += 256) {
                float stat = 0.f;
                for (uint32_t c = 0; c < out; ++c) {
                    const float w = W[c * in + r];
                    stat += lasso ? fabsf(w) : w * w;
                }
                const float lam = pr[d.wp_off[l] + r];
                if (lasso) {
                    t1 += (stat + 1.f / scale) * lam;
                    t2 += (shape + (float)out - 1.f) * logf(lam);
                } else {
                    t1 += (stat / 2.f + 1.f / scale) * lam;
                    t2 += (shape + ((float)out - 2.f) / 2.f) * logf(lam);
                }
            }
            t1 = block_sum<256>(t1, red);
            t2 = block_sum<256>(t2, red);
            ld_w = ld_w - t1 + t2;
        } else {     // ridge_base.rs:117-136, lasso_base.rs:119-138
            float stat = 0.f;
            for (uint32_t i = tid; i < in * out; i += 256) stat += lasso ? fabsf(W[i]) : W[i] * W[i];
            stat = block_sum<256>(stat, red);
            const float lam = pr[d.wp_off[l]];
            const float nvar = (float)(in * out);
            if (lasso) {
                ld_w -= (stat + 1.f / scale) * lam;
                ld_w += (shape + nvar - 1.f) * logf(lam);
            } else {
                ld_w -= (stat / 2.f + 1.f / scale) * lam;
                ld_w += (shape + (nvar - 2.f) / 2.f) * logf(lam);
            }
        }
    }
    local = ld_b + ld_w;
    // output weights: ridge_base.rs:138-157, lasso_base.rs:140-158 (same for ARD)
    float shape, scale;
    layer_prior(hy, last, nl, shape, scale);
    float own = 0.f;
    for (uint32_t i = tid; i < d.in_dim[last]; i += 256) {
        const float w = th[d.w_off[last] + i];
        own += lasso ? fabsf(w) : w * w;
    }
    own = block_sum<256>(own, red);
    const float g = own + ow_others;
    const float lam = pr[d.wp_off[last]];
    if (lasso) out_w = -(g + 1.f / scale) * lam + (shape + ow_num - 1.f) * logf(lam);
    else out_w = -((0.5f * g) + 1.f / scale) * lam + (shape + (ow_num - 2.f) / 2.f) * logf(lam);
}

struct FinishArgs {
    const BranchDesc* descs;
    uint32_t b;
    const float* theta;
    const float* prec;
    NetGlobals* G;
    const BranchState* st;     // NULL in initialize_stats (always update the LPD)
    const float* ow_others;
    float* lpd_local;          // [B]
    Hyper6 hyper;
    int model;
    float n_total;
    const float* part;         // residual partials [2*nblk]: sum r^2, sum(r + bias_old)
    uint32_t nblk;
    float* bias_old_new;       // out [2]
    int update_bias;           // 1 in a train visit, 0 in initialize_stats
    int* error_flag;
    XrComm xc;                 // row-sharded runs: the residual sums are summed over ranks here
};

// One block: update_lpd_from_branch (net.rs:173-185) when accepted, to_cfg + global params
// (branch_sampler.rs:155-171, params.rs:41-56), training counters, ML output bias (net.rs:43-45).
__global__ void __launch_bounds__(256) k_visit_finish(FinishArgs a) {
    __shared__ float red[8];
    const uint32_t tid = threadIdx.x;
    const BranchDesc& d = a.descs[a.b];
    const float* th = a.theta + d.param_off;
    const float* pr = a.prec + d.prec_off;
    const int last = (int)d.nl - 1;
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    const int status = a.st ? a.st->status : ST_ACCEPTED;
    double ss = 0.0, sb = 0.0;   // every thread computes the same fixed-order sums
    for (uint32_t i = 0; i < a.nblk; ++i) {
        ss += a.part[2 * i];
        sb += a.part[2 * i + 1];
    }
    if (a.xc.world > 1) {        // one thread exchanges, everybody reads the rank-ordered totals
        __shared__ double tot[2];
        if (tid == 0) {
            tot[0] = xr_sum(a.xc, 0, ss);
            tot[1] = xr_sum(a.xc, 1, sb);
        }
        __syncthreads();
        ss = tot[0];
        sb = tot[1];
    }
    const float others = *a.ow_others;
    if (status == ST_ACCEPTED) {
        float out_w, local;
        branch_lpd_terms(d, th, pr, a.model, a.hyper, others, a.G->ow_num_params, red, out_w, local);
        if (tid == 0) {
            a.lpd_local[a.b] = local;
            a.G->lpd_out_w = out_w;
            float k, s;
            layer_prior(a.hyper, last, (int)d.nl, k, s);
            const float le = pr[d.ep_off];
            // log_posterior_density.rs:49-60
            a.G->lpd_rss = logf(le) * (k + (a.n_total - 2.f) / 2.f) - le * ((float)ss / 2.f + 1.f / s);
        }
    }
    float own = 0.f;
    for (uint32_t i = tid; i < d.in_dim[last]; i += 256) {
        const float w = th[d.w_off[last] + i];
        own += lasso ? fabsf(w) : w * w;
    }
    own = block_sum<256>(own, red);
    if (tid == 0) {
        NetGlobals& G = *a.G;
        if (a.st) {
            G.num_samples += 1;                                    // train_stats.rs:48-56
            if (status == ST_ACCEPTED) G.num_accepted += 1;
            if (status == ST_REJECTED_EARLY) G.num_early_rejected += 1;
            G.visit_counter += 1;
            G.error_precision = pr[d.ep_off];                      // params.rs:41-56
            G.output_layer_precision = pr[d.wp_off[last]];
            const float reg = others + own;
            if (reg < 0.f || isnan(reg)) atomicExch(a.error_flag, 1);   // params.rs:49-54
            G.ow_reg_sum = reg;
        }
        if (a.update_bias) {
            a.bias_old_new[0] = G.output_bias;
            G.output_bias = (float)sb / a.n_total;                 // net.rs:43-45
            a.bias_old_new[1] = G.output_bias;
        }
        G.resid_ss = (float)ss;
    }
}


// ------------------------------------------------------------------ block-Jacobi group visits (bann_visit_group, bann_sweep with group_size > 1)
// Every member of a group runs the inner loop of Net::train (net.rs:258-334) against the residual and the global parameters
// FROZEN at group start; residual, globals, LPD terms, counters and the output bias are updated once, after all members
// finished (the checker's visit_group in tests/ is held to the same specification; a group of one member is visit_branch).

// r = r0 - sum over the accepted members, in list order, of (y_new - (t - r0))   (t = r0 + y_prev, net.rs:280,295);
// block partials of sum r^2 and sum (r + bias_old) as k_resid_after_hmc
__global__ void __launch_bounds__(256) k_resid_group(float* __restrict__ r, const float* __restrict__ T, const float* __restrict__ ynew,
                                                     uint32_t n, const uint32_t* __restrict__ list, uint32_t nlist,
                                                     const BranchState* __restrict__ states, const NetGlobals* G,
                                                     float* __restrict__ part /* [2*gridDim.x] */) {
    __shared__ float red[8];
    __shared__ uint8_t acc_flag[4096];
    const float bias = G->output_bias;
    float ss = 0.f, sb = 0.f;
    for (uint32_t base = blockIdx.x * 256; base < n; base += gridDim.x * 256) {   // uniform trip count per block: barriers inside
        const uint32_t i = base + threadIdx.x;
        const float r0 = i < n ? r[i] : 0.f;
        float v = r0;
        for (uint32_t l0 = 0; l0 < nlist; l0 += 4096) {
            const uint32_t cnt = min(4096u, nlist - l0);
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < cnt; k += 256) acc_flag[k] = states[list[l0 + k]].status == ST_ACCEPTED ? 1 : 0;
            __syncthreads();
            if (i < n)
                for (uint32_t k = 0; k < cnt; ++k)
                    if (acc_flag[k]) {
                        const size_t o = (size_t)(l0 + k) * n + i;
                        v = v - (ynew[o] - (T[o] - r0));
                    }
        }
        if (i < n) {
            r[i] = v;
            ss = fmaf(v, v, ss);
            sb += v + bias;
        }
    }
    ss = block_sum<256>(ss, red);
    sb = block_sum<256>(sb, red);
    if (threadIdx.x == 0) {
        part[2 * blockIdx.x] = ss;
        part[2 * blockIdx.x + 1] = sb;
    }
}

struct GroupArgs {
    const BranchDesc* descs;
    const uint32_t* list;
    uint32_t nlist;
    const float* theta;
    const float* prec;
    NetGlobals* G;
    const BranchState* states;
    const float* own_old;      // [B] by branch (k_gibbs)
    float* own_new;            // [B] by branch
    float* lpd_local;          // [B]
    Hyper6 hyper;
    int model;
    float n_total;
    const float* part;         // residual partials [2*nblk]
    uint32_t nblk;
    float* bias_old_new;       // out [2]
    int* error_flag;
    XrComm xc;
};

// one block per member: own output-weight statistic after the transition; LPD local term of an accepted member
__global__ void __launch_bounds__(256) k_group_stats(GroupArgs a) {
    __shared__ float red[8];
    const uint32_t tid = threadIdx.x, b = a.list[blockIdx.x];
    const BranchDesc& d = a.descs[b];
    const float* th = a.theta + d.param_off;
    const float* pr = a.prec + d.prec_off;
    const int last = (int)d.nl - 1;
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    float own = 0.f;
    for (uint32_t i = tid; i < d.in_dim[last]; i += 256) {
        const float w = th[d.w_off[last] + i];
        own += lasso ? fabsf(w) : w * w;
    }
    own = block_sum<256>(own, red);
    if (tid == 0) a.own_new[b] = own;
    if (a.states[b].status == ST_ACCEPTED) {
        float out_w, local;
        branch_lpd_terms(d, th, pr, a.model, a.hyper, 0.f, a.G->ow_num_params, red, out_w, local);
        if (tid == 0) a.lpd_local[b] = local;
    }
}

// one block: the members' bookkeeping in list order (what k_visit_finish does per visit), the group's residual sums,
// LPD terms of the last accepted member, ML output bias
__global__ void __launch_bounds__(256) k_group_finish(GroupArgs a) {
    const uint32_t tid = threadIdx.x;
    double ss = 0.0, sb = 0.0;
    for (uint32_t i = 0; i < a.nblk; ++i) {
        ss += a.part[2 * i];
        sb += a.part[2 * i + 1];
    }
    if (a.xc.world > 1) {
        __shared__ double tot[2];
        if (tid == 0) {
            tot[0] = xr_sum(a.xc, 0, ss);
            tot[1] = xr_sum(a.xc, 1, sb);
        }
        __syncthreads();
        ss = tot[0];
        sb = tot[1];
    }
    if (tid != 0) return;
    NetGlobals& G = *a.G;
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    float reg = G.ow_reg_sum;
    int last_acc = -1;
    float g_last = 0.f;
    for (uint32_t li = 0; li < a.nlist; ++li) {
        const uint32_t b = a.list[li];
        const BranchDesc& d = a.descs[b];
        const float* pr = a.prec + d.prec_off;
        const int status = a.states[b].status;
        const float others = reg - a.own_old[b];                // from_cfg (branch_struct.rs:27) against the running global
        reg = others + a.own_new[b];                            // to_cfg (branch_sampler.rs:155-171)
        if (reg < 0.f || isnan(reg)) atomicExch(a.error_flag, 1);   // params.rs:49-54
        G.num_samples += 1;                                     // train_stats.rs:48-56
        if (status == ST_ACCEPTED) { G.num_accepted += 1; last_acc = (int)li; g_last = a.own_new[b] + others; }
        if (status == ST_REJECTED_EARLY) G.num_early_rejected += 1;
        G.visit_counter += 1;
        G.error_precision = pr[d.ep_off];                       // params.rs:41-56
        G.output_layer_precision = pr[d.wp_off[(int)d.nl - 1]];
    }
    G.ow_reg_sum = reg;
    if (last_acc >= 0) {                                        // update_lpd_from_branch (net.rs:173-185) of the last accepted member
        const uint32_t b = a.list[last_acc];
        const BranchDesc& d = a.descs[b];
        const float* pr = a.prec + d.prec_off;
        const int last = (int)d.nl - 1;
        float shape, scale;
        layer_prior(a.hyper, last, (int)d.nl, shape, scale);
        if (a.model == BANN_STD_NORMAL) G.lpd_out_w = 0.f;
        else {
            const float lam = pr[d.wp_off[last]];
            if (lasso) G.lpd_out_w = -(g_last + 1.f / scale) * lam + (shape + G.ow_num_params - 1.f) * logf(lam);
            else G.lpd_out_w = -((0.5f * g_last) + 1.f / scale) * lam + (shape + (G.ow_num_params - 2.f) / 2.f) * logf(lam);
        }
        const float le = pr[d.ep_off];
        G.lpd_rss = logf(le) * (shape + (a.n_total - 2.f) / 2.f) - le * ((float)ss / 2.f + 1.f / scale);   // log_posterior_density.rs:49-60
    }
    a.bias_old_new[0] = G.output_bias;
    G.output_bias = (float)sb / a.n_total;                      // net.rs:43-45
    a.bias_old_new[1] = G.output_bias;
    G.resid_ss = (float)ss;
}

}  // namespace bann
