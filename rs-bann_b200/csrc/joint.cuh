// Flag-gated sampler modes of BranchSampler: joint HMC over parameters AND precisions
// (hmc_step_joint, branch_sampler.rs:1070-1178), joint gradient ascent (gradient_descent_joint, :1019-1066) and the
// helpers of the line-search gradient ascent (gradient_descent, :964-1017).  One branch at a time (the reference's
// sequential schedule): K1 provides the raw sums [gW | gb | rss] of the forward+backward pass, the kernels here turn them
// into the joint gradient / joint log density and advance the state, exactly like k2_step does for the plain sampler.
#pragma once
#include "chain.cuh"
#include "kernels.cuh"

namespace bann {

enum : int { JM_HMC_INIT = 0, JM_HMC_STEP = 1, JM_GD_STEP = 2, JM_EVAL = 3 };

struct JointArgs {
    const BranchDesc* descs;
    uint32_t b;
    BranchState* st;          // &states[b]
    float* theta;             // parameter arenas (base pointers)
    float* theta0;
    float* mom;
    float* grad;
    float* eps;
    float* prec;              // precision arena (base pointer)
    float* prec0;             // per-visit workspace, [nprec] each
    float* pmom;
    float* pgrad;
    float* peps;
    const float* gsum;        // [P + 1] raw sums of the K1 pass at the current parameters
    const float* ow_others;   // output-weight statistic of all OTHER branches (branch_struct.rs:27)
    const NetGlobals* G;
    Hyper6 hyper;
    int model;
    float n_total;            // y_train.elements()
    float max_h_err;
    int mode;                 // JM_*
    int is_last;
    float gd_step;            // gradient_descent_joint: hmc_step_size_factor
    float* traj_params;       // optional [L][P]
    float* traj_prec;         // optional [L][Q]
    float* traj_ldg;          // optional [L][P + Q] (BranchLogDensityGradientJoint::param_vec order, gradient.rs:66-97)
    float* traj_h;            // optional [L + 1]
    float* out;               // JM_EVAL: [0] rss, [1] log_density_joint, [2] log_density
    float* kin_out;           // scratch [1]: kinetic energy of the last step (for the accept kernel)
};

// log_density_gradient_joint (branch_sampler.rs:406-422) and log_density_joint (:292-305) / log_density (:72-78) at the
// current state.  Writes grad[P] and pgrad[Q]; returns the joint and the plain log density (valid in every thread).
__device__ void joint_eval(const JointArgs& a, const BranchDesc& d, float* red, float& ld_joint, float& ld_plain) {
    const uint32_t tid = threadIdx.x;
    const float* th = a.theta + d.param_off;
    const float* pr = a.prec + d.prec_off;
    float* gr = a.grad + d.param_off;
    const float* gs = a.gsum;
    const int nl = (int)d.nl, last = nl - 1;
    const bool lasso = (a.model == BANN_LASSO_BASE || a.model == BANN_LASSO_ARD);
    const bool ard = (a.model == BANN_RIDGE_ARD || a.model == BANN_LASSO_ARD);
    const float lam_e = pr[d.ep_off];
    const float rss = gs[d.P];

    // parameters: weights under the prior (per-prior log_density_gradient_wrt_weights), l2-regularised biases (:334-345)
    for (uint32_t k = tid; k < d.P; k += 256) {
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        const float w = th[k];
        float g;
        if (isb) g = -1.0f * pr[d.bp_off[l]] * w - lam_e * gs[k];
        else {
            const float lam = param_prior_precision(d, pr, a.model, l, row, false);
            if (lasso) g = -(lam_e * gs[k] + lam * ((w > 0.f) ? 1.f : (w < 0.f ? -1.f : 0.f)));
            else g = -(lam_e * gs[k] + lam * w);
        }
        gr[k] = g;
    }
    float ldw_joint = 0.f, ldw_plain = 0.f, ldb_joint = 0.f;
    for (int l = 0; l < last; ++l) {
        float shape, scale;
        layer_prior(a.hyper, l, nl, shape, scale);
        const uint32_t in = d.in_dim[l], out = d.widths[l];
        const float* W = th + d.w_off[l];
        // biases: density :260-279, precision gradient :348-367
        float bs = 0.f;
        for (uint32_t c = tid; c < out; c += 256) bs += th[d.b_off[l] + c] * th[d.b_off[l] + c];
        bs = block_sum<256>(bs, red);
        const float lb = pr[d.bp_off[l]];
        ldb_joint -= lb * (bs / 2.f + 1.f / scale);
        ldb_joint += (shape + ((float)out - 2.f) / 2.f) * logf(lb);
        if (tid == 0) a.pgrad[d.bp_off[l]] = (2.f * shape + ((float)out - 2.f)) / (2.f * lb) - 1.f / scale - bs / 2.f;
        if (ard) {   // ridge_ard.rs:119-148,171-194,221-236; lasso_ard.rs:123-151,173-194,220-234
            float t1 = 0.f, t2 = 0.f, t3 = 0.f;
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
                    t3 += stat * lam;
                    a.pgrad[d.wp_off[l] + r] = (shape + (float)in - 1.f) / lam - 1.f / scale - stat;   // Q8: #rows
                } else {
                    t1 += (stat / 2.f + 1.f / scale) * lam;
                    t2 += (shape + ((float)out - 2.f) / 2.f) * logf(lam);
                    t3 += 0.5f * stat * lam;
                    a.pgrad[d.wp_off[l] + r] = (2.f * shape + (float)in - 2.f) / (2.f * lam) - 1.f / scale - stat / 2.f;
                }
            }
            t1 = block_sum<256>(t1, red);
            t2 = block_sum<256>(t2, red);
            t3 = block_sum<256>(t3, red);
            ldw_joint = ldw_joint - t1 + t2;
            ldw_plain -= t3;
        } else {     // ridge_base.rs:117-136,159-173,186-200; lasso_base.rs:119-138,160-173,187-201
            float stat = 0.f;
            for (uint32_t i = tid; i < in * out; i += 256) stat += lasso ? fabsf(W[i]) : W[i] * W[i];
            stat = block_sum<256>(stat, red);
            const float lam = pr[d.wp_off[l]];
            const float nvar = (float)(in * out);
            if (lasso) {
                ldw_joint -= (stat + 1.f / scale) * lam;
                ldw_joint += (shape + nvar - 1.f) * logf(lam);
                ldw_plain -= stat * lam;
                if (tid == 0) a.pgrad[d.wp_off[l]] = (shape + nvar - 1.f) / lam - 1.f / scale - stat;
            } else {
                ldw_joint -= (stat / 2.f + 1.f / scale) * lam;
                ldw_joint += (shape + (nvar - 2.f) / 2.f) * logf(lam);
                ldw_plain -= (stat / 2.f) * lam;
                if (tid == 0) a.pgrad[d.wp_off[l]] = (2.f * shape + nvar - 2.f) / (2.f * lam) - 1.f / scale - stat / 2.f;
            }
        }
    }
    // output weights: shared precision, statistic = own + others (ridge_base.rs:138-157, lasso_base.rs:140-158, same for ARD)
    float shape, scale;
    layer_prior(a.hyper, last, nl, shape, scale);
    float own = 0.f;
    for (uint32_t i = tid; i < d.in_dim[last]; i += 256) {
        const float w = th[d.w_off[last] + i];
        own += lasso ? fabsf(w) : w * w;
    }
    own = block_sum<256>(own, red);
    const float gstat = own + *a.ow_others;
    const float lam_o = pr[d.wp_off[last]];
    const float ow_num = a.G->ow_num_params;
    float out_w;
    if (lasso) {
        out_w = -(gstat + 1.f / scale) * lam_o + (shape + ow_num - 1.f) * logf(lam_o);
        ldw_plain -= own * lam_o;
        if (tid == 0) a.pgrad[d.wp_off[last]] = (shape + ow_num - 1.f) / lam_o - 1.f / scale - gstat;
    } else {
        out_w = -((0.5f * gstat) + 1.f / scale) * lam_o + (shape + (ow_num - 2.f) / 2.f) * logf(lam_o);
        ldw_plain -= 0.5f * own * lam_o;
        if (tid == 0) a.pgrad[d.wp_off[last]] = (2.f * shape + ow_num - 2.f) / (2.f * lam_o) - 1.f / scale - gstat / 2.f;
    }
    // error precision: density :240-257, gradient :369-378 (output-layer hyperparameters, last_rss of this pass)
    const float wrt_e_joint = (shape + (a.n_total - 2.f) / 2.f) * logf(lam_e) - lam_e * (rss / 2.f + 1.f / scale);
    if (tid == 0) a.pgrad[d.ep_off] = (2.f * shape + a.n_total - 2.f) / (2.f * lam_e) - 1.f / scale - rss / 2.f;
    ld_joint = ((ldw_joint + out_w) + ldb_joint) + wrt_e_joint;          // :292-305
    ld_plain = ldw_plain + (-1.0f * lam_e * (rss / 2.0f));              // :72-78 (bias term 0)
    __syncthreads();                                                    // grad / pgrad visible to the whole block
}

// One block, one branch.  HMC: [eval; p += e/2 g; H check; p += e/2 g; q += e p] -- the reference's leapfrog
// (:1126-1131) cut after the position update so that K1 stays the only pass over the data (as k2_step).
// Gradient ascent: [eval; q += s g].
__global__ void __launch_bounds__(256) k2_joint(JointArgs a) {
    __shared__ float red[8];
    __shared__ int s_status;
    const uint32_t tid = threadIdx.x;
    const BranchDesc& d = a.descs[a.b];
    BranchState& st = *a.st;
    if (a.mode != JM_EVAL && st.status != ST_RUNNING) return;
    const uint32_t P = d.P, Q = d.nprec;
    float* th = a.theta + d.param_off;
    float* pr = a.prec + d.prec_off;
    float* p = a.mom + d.param_off;
    float* gr = a.grad + d.param_off;
    const float* ep = a.eps + d.param_off;
    float ld_joint, ld_plain;
    joint_eval(a, d, red, ld_joint, ld_plain);
    const float rss = a.gsum[P];
    if (a.mode == JM_EVAL) {
        if (tid == 0 && a.out) { a.out[0] = rss; a.out[1] = ld_joint; a.out[2] = ld_plain; }
        return;
    }
    if (a.mode == JM_GD_STEP) {
        if (!a.is_last) {
            for (uint32_t k = tid; k < P; k += 256) th[k] = __fadd_rn(th[k], __fmul_rn(a.gd_step, gr[k]));        // params.rs:740-749
            for (uint32_t q = tid; q < Q; q += 256) pr[q] = __fadd_rn(pr[q], __fmul_rn(a.gd_step, a.pgrad[q]));   // params.rs:357-367
            if (tid == 0) st.steps_done += 1;
            return;
        }
        // :1053-1065: Rejected (state restored) when the error precision ended <= 0
        const bool bad = pr[d.ep_off] <= 0.0f;
        __syncthreads();
        if (bad) {
            for (uint32_t k = tid; k < P; k += 256) th[k] = a.theta0[d.param_off + k];
            for (uint32_t q = tid; q < Q; q += 256) pr[q] = a.prec0[q];
        }
        if (tid == 0) {
            st.rss = rss;
            st.log_density = ld_joint;
            st.neg_h_cur = ld_joint;
            st.status = bad ? ST_REJECTED : ST_ACCEPTED;
        }
        return;
    }
    // ---- HMC
    float kin = 0.f;
    for (uint32_t k = tid; k < P; k += 256) {
        float pk = p[k];
        if (a.mode == JM_HMC_STEP) {
            pk = pk + (0.5f * ep[k]) * gr[k];            // momentum.rs:32-58
            p[k] = pk;
        }
        kin = fmaf(pk, pk, kin);
    }
    for (uint32_t q = tid; q < Q; q += 256) {
        float pk = a.pmom[q];
        if (a.mode == JM_HMC_STEP) {
            pk = pk + (a.peps[q] * 0.5f) * a.pgrad[q];
            a.pmom[q] = pk;
        }
        kin = fmaf(pk, pk, kin);
    }
    kin = 0.5f * block_sum<256>(kin, red);               // momentum.rs:83-104
    if (tid == 0) {
        const float negh = ld_joint - kin;               // neg_hamiltonian_joint, :886-903
        st.rss = rss;
        st.log_density = ld_plain;                       // what accept_or_reject_hmc_state evaluates (:928-962)
        st.neg_h_cur = negh;
        *a.kin_out = kin;
        if (a.mode == JM_HMC_INIT) {
            st.neg_h_init = negh;
            if (a.traj_h) a.traj_h[0] = negh;
        } else {
            const int step = st.steps_done;
            st.steps_done = step + 1;
            if (a.traj_h) a.traj_h[step + 1] = negh;
            if (fabsf(negh - st.neg_h_init) > a.max_h_err) st.status = ST_REJECTED_EARLY;   // :1146-1162
        }
        s_status = st.status;
    }
    __syncthreads();
    const int status = s_status;
    if (a.mode == JM_HMC_STEP && a.traj_params) {
        const int step = st.steps_done - 1;
        for (uint32_t k = tid; k < P; k += 256) {
            a.traj_params[(size_t)step * P + k] = th[k];
            a.traj_ldg[(size_t)step * (P + Q) + k] = gr[k];
        }
        for (uint32_t q = tid; q < Q; q += 256) {
            a.traj_prec[(size_t)step * Q + q] = pr[q];
            a.traj_ldg[(size_t)step * (P + Q) + P + q] = a.pgrad[q];
        }
    }
    if (status == ST_REJECTED_EARLY) {
        for (uint32_t k = tid; k < P; k += 256) th[k] = a.theta0[d.param_off + k];
        for (uint32_t q = tid; q < Q; q += 256) pr[q] = a.prec0[q];
        return;
    }
    if (!a.is_last) {
        for (uint32_t k = tid; k < P; k += 256) {
            const float e = ep[k];
            const float pk = p[k] + (0.5f * e) * gr[k];
            p[k] = pk;
            th[k] = th[k] + e * pk;                      // params.rs:728-738
        }
        for (uint32_t q = tid; q < Q; q += 256) {
            const float e = a.peps[q];
            const float pk = a.pmom[q] + (e * 0.5f) * a.pgrad[q];
            a.pmom[q] = pk;
            pr[q] = pr[q] + e * pk;                      // params.rs:344-355
        }
    }
}

struct JointInitArgs {
    const BranchDesc* descs;
    uint32_t b;
    BranchState* st;
    const float* theta;
    float* theta0;
    float* mom;
    float* eps;
    const float* prec;
    float* prec0;
    float* pmom;
    float* peps;
    float factor;
    int hmc;                         // 1: momenta + random step sizes; 0: gradient ascent (only the saved state)
    const float* inj_momenta;        // [P + Q] or NULL
    const float* inj_step_uniforms;  // [P + Q] or NULL
    uint64_t seed;
    uint64_t stream;
};

// sample_joint_momentum (:611-641) + random_step_sizes with the joint factor (P + Q)^(-1/4) f (:654-704) + saved state
__global__ void __launch_bounds__(256) k_joint_init(JointInitArgs a) {
    const uint32_t tid = threadIdx.x;
    const BranchDesc& d = a.descs[a.b];
    const uint32_t P = d.P, Q = d.nprec, T = P + Q;
    const float* th = a.theta + d.param_off;
    const float* pr = a.prec + d.prec_off;
    for (uint32_t k = tid; k < P; k += 256) a.theta0[d.param_off + k] = th[k];
    for (uint32_t q = tid; q < Q; q += 256) a.prec0[q] = pr[q];
    if (a.hmc) {
        const float prop = __fmul_rn(powf((float)P + (float)Q, -0.25f), a.factor);
        for (uint32_t k2 = tid; 2 * k2 < T; k2 += 256) {
            float v[2];
            if (a.inj_momenta) {
                v[0] = a.inj_momenta[2 * k2];
                v[1] = (2 * k2 + 1 < T) ? a.inj_momenta[2 * k2 + 1] : 0.f;
            } else {
                Philox ph(a.seed, a.stream, (uint64_t)k2);
                philox_normal_pair(ph, v[0], v[1]);
            }
            for (int i = 0; i < 2; ++i) {
                const uint32_t k = 2 * k2 + i;
                if (k < P) a.mom[d.param_off + k] = v[i];
                else if (k < T) a.pmom[k - P] = v[i];
            }
        }
        for (uint32_t k = tid; k < T; k += 256) {
            float u;
            if (a.inj_step_uniforms) u = a.inj_step_uniforms[k];
            else {
                Philox ph(a.seed ^ 0x5bd1e995u, a.stream, (uint64_t)k);
                uint32_t r[4];
                ph.next(r);
                u = u01_half_open(r[0]);
            }
            const float e = __fmul_rn(u, prop);
            if (k < P) a.eps[d.param_off + k] = e;
            else a.peps[k - P] = e;
        }
    }
    if (tid == 0) {
        BranchState& st = *a.st;
        st.status = ST_RUNNING;
        st.steps_done = 0;
        st.u_turn_step = -1;
        st.neg_h_init = st.neg_h_cur = st.log_density = st.rss = st.log_acc = 0.f;
    }
}

// accept_or_reject_hmc_state (:928-962) as hmc_step_joint calls it: NON-joint log density of the final state minus the
// joint kinetic energy, against the joint initial Hamiltonian; rejected -> parameters and precisions restored (:1166-1171)
__global__ void __launch_bounds__(256) k_joint_accept(const BranchDesc* descs, uint32_t b, BranchState* stp, float* theta,
                                                      const float* theta0, float* prec, const float* prec0, const float* kin,
                                                      const float* inj_u, uint64_t seed, uint64_t stream) {
    __shared__ int s_status;
    const uint32_t tid = threadIdx.x;
    const BranchDesc& d = descs[b];
    BranchState& st = *stp;
    if (tid == 0) {
        if (st.status == ST_RUNNING) {
            const float h_final = st.log_density - *kin;
            const float log_acc = h_final - st.neg_h_init;
            const float prob = (log_acc >= 0.f) ? 1.f : expf(log_acc);
            float u;
            if (inj_u) u = inj_u[0];
            else {
                Philox ph(seed ^ 0xa511e9b3u, stream, 0);
                uint32_t r[4];
                ph.next(r);
                u = u01_half_open(r[0]);
            }
            st.neg_h_cur = h_final;
            st.log_acc = log_acc;
            st.status = (u < prob) ? ST_ACCEPTED : ST_REJECTED;   // NaN -> rejected
        }
        s_status = st.status;
    }
    __syncthreads();
    if (s_status == ST_REJECTED) {
        for (uint32_t k = tid; k < d.P; k += 256) theta[d.param_off + k] = theta0[d.param_off + k];
        for (uint32_t q = tid; q < d.nprec; q += 256) prec[d.prec_off + q] = prec0[q];
    }
}

// ---- helpers of gradient_descent (:964-1017)
// theta = base + s * g  (probe_gradient_step / descend_gradient, params.rs:740-749; separate multiply and add as ArrayFire)
__global__ void __launch_bounds__(256) k_gd_axpy(const BranchDesc* descs, uint32_t b, float* theta, const float* base,
                                                 const float* grad, float s) {
    const BranchDesc& d = descs[b];
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < d.P; k += gridDim.x * 256)
        theta[d.param_off + k] = __fadd_rn(base[d.param_off + k], __fmul_rn(s, grad[d.param_off + k]));
}
__global__ void __launch_bounds__(256) k_gd_copy(const BranchDesc* descs, uint32_t b, float* dst, const float* src) {
    const BranchDesc& d = descs[b];
    for (uint32_t k = blockIdx.x * 256 + threadIdx.x; k < d.P; k += gridDim.x * 256) dst[d.param_off + k] = src[d.param_off + k];
}
// own output-weight statistic subtracted from the global one WITHOUT touching the branch's precisions
// (from_cfg, branch_struct.rs:27) -- for the stand-alone entry points; a train visit uses k_gibbs for this
__global__ void __launch_bounds__(256) k_ow_others(const BranchDesc* descs, uint32_t b, const float* theta, const NetGlobals* G,
                                                   int model, float* ow_others) {
    __shared__ float red[8];
    const BranchDesc& d = descs[b];
    const int last = (int)d.nl - 1;
    const bool lasso = (model == BANN_LASSO_BASE || model == BANN_LASSO_ARD);
    float own = 0.f;
    for (uint32_t i = threadIdx.x; i < d.in_dim[last]; i += 256) {
        const float w = theta[d.param_off + d.w_off[last] + i];
        own += lasso ? fabsf(w) : w * w;
    }
    own = block_sum<256>(own, red);
    if (threadIdx.x == 0) *ow_others = G->ow_reg_sum - own;
}
// end of gradient_descent: always Accepted, log_density(params, precisions, rss) of the final state (:991-1002)
__global__ void k_gd_finish(BranchState* st, const float* gsum, uint32_t P, const float* log_density, int steps) {
    st->status = ST_ACCEPTED;
    st->steps_done = steps;
    st->u_turn_step = -1;
    st->rss = gsum[P];
    st->log_density = *log_density;
    st->neg_h_init = st->neg_h_cur = *log_density;
    st->log_acc = 0.f;
}

}  // namespace bann
