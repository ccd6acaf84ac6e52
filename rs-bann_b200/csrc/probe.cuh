// Per-row diagnostics of a branch: Net::activations (net/net.rs:509-518 -> forward_feed, branch_sampler.rs:743-782) and
// BranchSampler::effect_sizes / Net::population_effect_sizes (branch_sampler.rs:784-811, net/net.rs:529-543).
// Off the sampler's hot path (the reference runs them from separate subcommands on saved models): one thread per row,
// any depth / widths / activation, same folded standardisation as K1.
#pragma once
#include "kernels.cuh"

namespace bann {

struct ProbeArgs {
    const uint8_t* store;
    const BranchDesc* descs;
    uint32_t b;
    const float* theta;
    const float* mu;
    const float* sd;
    uint32_t n, ntiles;
    int act;
    float* acts_out;   // optional: [a_0 (n x w_0) | a_1 | ... | yhat (n x 1)], each column-major
    float* es_out;     // optional: n x m column-major, yhat * d yhat / d x (the reference seeds the back-propagation with yhat)
    float* dsum_part;  // optional: [ntiles][w_0] per-tile column sums of the first-layer deltas (for the population mean)
};

__global__ void __launch_bounds__(128) k_branch_probe(ProbeArgs a) {
    extern __shared__ float smf[];
    const uint32_t tid = threadIdx.x;
    const BranchDesc& d = a.descs[a.b];
    const uint32_t P = d.P, m = d.m, mp = d.m_pad4, nl = d.nl, w0 = d.widths[0];
    const uint32_t SW = d.sumw | 1u;
    float* sp = smf;
    float* as_ = sp + ((P + 3) & ~3u);
    float* ds_ = as_ + 128 * SW;
    float* b0p = ds_ + 128 * SW + 128 + 8;
    const float* th = a.theta + d.param_off;
    const float* mu = a.mu + d.col_off;
    const float* sd = a.sd + d.col_off;
    for (uint32_t k = tid; k < P; k += 128) sp[k] = th[k];
    __syncthreads();
    for (uint32_t k = tid; k < m * w0; k += 128) sp[d.w_off[0] + k] = __fdiv_rn(sp[d.w_off[0] + k], sd[k % m]);   // bed.rs:354 folded
    __syncthreads();
    for (uint32_t c = tid; c < w0; c += 128) {
        float acc = 0.f;
        for (uint32_t j = 0; j < m; ++j) acc = fmaf(mu[j], sp[d.w_off[0] + c * m + j], acc);
        b0p[c] = sp[d.b_off[0] + c] - acc;
    }
    __syncthreads();
    const uint32_t L = nl - 2, sL = d.widths[L];
    for (uint32_t t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
        const uint8_t* tile = a.store + d.tile_off + (size_t)t * (kTileQuads * mp);
        const uint32_t r = tid, q = r >> 2, sh = 2 * (r & 3);
        const uint32_t row = t * kTileRows + r;
        const bool valid = row < a.n;
        float* ar = as_ + r * SW;
        float* dr = ds_ + r * SW;
        for (uint32_t c = 0; c < w0; ++c) {
            float z = b0p[c];
            for (uint32_t j = 0; j < m; ++j) {
                const uint32_t g = (tile[q * mp + j] >> sh) & 3u;
                if (g) z = fmaf((float)g, sp[d.w_off[0] + c * m + j], z);
            }
            const float h = act_h(a.act, z);
            ar[d.a_off[0] + c] = h;
            dr[d.a_off[0] + c] = act_dh(a.act, z, h);
        }
        for (uint32_t l = 1; l + 1 < nl; ++l) {
            const uint32_t in = d.in_dim[l], out = d.widths[l];
            for (uint32_t c = 0; c < out; ++c) {
                float z = sp[d.b_off[l] + c];
                for (uint32_t i = 0; i < in; ++i) z = fmaf(ar[d.a_off[l - 1] + i], sp[d.w_off[l] + c * in + i], z);
                const float h = act_h(a.act, z);
                ar[d.a_off[l] + c] = h;
                dr[d.a_off[l] + c] = act_dh(a.act, z, h);
            }
        }
        float yh = 0.f;
        for (uint32_t i = 0; i < sL; ++i) yh = fmaf(ar[d.a_off[L] + i], sp[d.w_off[nl - 1] + i], yh);
        if (a.acts_out && valid) {
            for (uint32_t k = 0; k < d.sumw; ++k) a.acts_out[(size_t)k * a.n + row] = ar[k];   // a_off[l] + c == column index
            a.acts_out[(size_t)d.sumw * a.n + row] = yh;
        }
        if (a.es_out || a.dsum_part) {
            // branch_sampler.rs:791-809: error = yhat W_last^T; per layer delta = h'(z) * error, error = delta W_l^T
            const float seed = valid ? yh : 0.f;
            for (uint32_t i = 0; i < sL; ++i) dr[d.a_off[L] + i] *= seed * sp[d.w_off[nl - 1] + i];
            for (uint32_t l = L; l >= 1; --l) {
                const uint32_t in = d.in_dim[l], out = d.widths[l];
                for (uint32_t i = 0; i < in; ++i) {
                    float err = 0.f;
                    for (uint32_t c = 0; c < out; ++c) err = fmaf(dr[d.a_off[l] + c], sp[d.w_off[l] + c * in + i], err);
                    dr[d.a_off[l - 1] + i] *= err;
                }
            }
            if (a.es_out && valid) {
                for (uint32_t j = 0; j < m; ++j) {       // w.r.t. the STANDARDISED input: the unfolded first-layer weights
                    float err = 0.f;
                    for (uint32_t c = 0; c < w0; ++c) err = fmaf(dr[d.a_off[0] + c], th[c * m + j], err);
                    a.es_out[(size_t)j * a.n + row] = err;
                }
            }
            __syncthreads();
            if (a.dsum_part)
                for (uint32_t c = tid; c < w0; c += 128) {
                    float s = 0.f;
                    for (uint32_t rr = 0; rr < 128; ++rr) s += ds_[rr * SW + d.a_off[0] + c];   // fixed order
                    a.dsum_part[(size_t)t * w0 + c] = s;
                }
            __syncthreads();
        }
    }
}

// population_effect_sizes (net.rs:529-543): column means of the effect sizes.  Linear in the first-layer deltas:
// sum_rows es[row, j] = sum_c W0[j, c] * sum_rows delta0[row, c]
__global__ void __launch_bounds__(128) k_population_effects(const BranchDesc* descs, uint32_t b, const float* theta,
                                                            const float* dsum_part, uint32_t ntiles, float n_total,
                                                            float* out /* m */) {
    extern __shared__ float sd_[];   // [w0]
    const BranchDesc& d = descs[b];
    const uint32_t w0 = d.widths[0], m = d.m;
    for (uint32_t c = threadIdx.x; c < w0; c += 128) {
        double s = 0.0;
        for (uint32_t t = 0; t < ntiles; ++t) s += dsum_part[(size_t)t * w0 + c];
        sd_[c] = (float)s;
    }
    __syncthreads();
    const float* W0 = theta + d.param_off;
    for (uint32_t j = blockIdx.x * 128 + threadIdx.x; j < m; j += gridDim.x * 128) {
        float s = 0.f;
        for (uint32_t c = 0; c < w0; ++c) s = fmaf(W0[c * m + j], sd_[c], s);
        out[j] = s / n_total;
    }
}

}  // namespace bann
