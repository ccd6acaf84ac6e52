/*
 * CPU restatement (C + OpenMP) of the reference's op sequence for one branch -- TEST
 * INFRASTRUCTURE and CPU-baseline timing only (see oracle/__init__.py).  Never linked or loaded
 * by the product (rs-bann_b200/).
 *
 * Follows, in the reference repository:
 *   io/bed.rs:325-355            host LUT decode of the branch's columns to dense f32 + standardise
 *   branch_sampler.rs:743-782    forward_feed
 *   branch_sampler.rs:813-875    backpropagate (uses e, not 2e)
 *   branch_sampler.rs:1239-1284  leapfrog: half step, full step, gradient, half step,
 *                                neg_hamiltonian (a SECOND forward pass)
 *   momentum.rs:129-158, params.rs:728-738
 * Prior handling is reduced to a per-parameter precision vector `lam` (0 for biases) and a
 * lasso flag, which covers the five prior variants' non-joint density / gradient.
 * The reference runs these as ArrayFire (BLAS + elementwise) calls; this port threads over rows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXL 8
#define MAXW 64

static const float CODE_TO_VALUE[4] = {2.f, 0.f, 1.f, 0.f}; /* io/bed_lookup_tables.rs:4 */

int bann_cpu_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* pin the OpenMP team size explicitly (torchrun exports OMP_NUM_THREADS=1 to its workers); returns the size in effect */
int bann_cpu_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

void bann_cpu_decode_std(const uint8_t* payload, uint64_t n, const uint64_t* cols, uint32_t m, const float* means,
                         const float* stds, float* X) {
    uint64_t bpc = (n + 3) / 4;
#pragma omp parallel for schedule(static)
    for (int64_t j = 0; j < (int64_t)m; ++j) {
        const uint8_t* col = payload + cols[j] * bpc;
        float mu = means[cols[j]], sd = stds[cols[j]];
        float* x = X + (uint64_t)j * n;
        for (uint64_t i = 0; i < n; ++i) x[i] = (CODE_TO_VALUE[(col[i >> 2] >> (2 * (i & 3))) & 3] - mu) / sd;
    }
}

static inline float act_h(int act, float x) {
    switch (act) {
        case 0: return tanhf(x);
        case 1: return x > 0.f ? x : 0.f;
        case 2: return x > 0.f ? x : (x < 0.f ? 0.01f * x : 0.f);
        case 3: return x * (1.f / (1.f + expf(-x)));
        default: return x;
    }
}
static inline float act_dh(int act, float x, float hx) {
    switch (act) {
        case 0: return 1.f - hx * hx;
        case 1: return x > 0.f ? 1.f : 0.f;
        case 2: return x > 0.f ? 1.f : (x < 0.f ? 0.01f : 0.f);
        case 3: { float sg = 1.f / (1.f + expf(-x)); return hx + sg * (1.f - hx); }
        default: return 1.f;
    }
}

typedef struct { uint32_t nl, m, P; uint32_t w[MAXL], in[MAXL], woff[MAXL], boff[MAXL]; } shape_t;

static void make_shape(shape_t* s, uint32_t m, const uint32_t* widths, uint32_t nl) {
    s->nl = nl; s->m = m;
    uint32_t prev = m, off = 0;
    for (uint32_t l = 0; l < nl; ++l) { s->w[l] = widths[l]; s->in[l] = prev; s->woff[l] = off; off += prev * widths[l]; prev = widths[l]; }
    for (uint32_t l = 0; l + 1 < nl; ++l) { s->boff[l] = off; off += widths[l]; }
    s->P = off;
}

#define BLK 256

/* first-layer pre-activations for a block of rows: Z[c][ii] = b0[c] + sum_j X[i0+ii, j] W0[j, c]
 * (column-major X: the inner loop runs over contiguous rows and vectorises) */
static inline void first_layer_block(const shape_t* s, const float* th, const float* X, uint64_t n, uint64_t i0,
                                     uint32_t nb, float Z[MAXW][BLK]) {
    for (uint32_t c = 0; c < s->w[0]; ++c) {
        float b = th[s->boff[0] + c];
        for (uint32_t ii = 0; ii < nb; ++ii) Z[c][ii] = b;
    }
    for (uint32_t j = 0; j < s->m; ++j) {
        const float* xj = X + (uint64_t)j * n + i0;
        for (uint32_t c = 0; c < s->w[0]; ++c) {
            float w = th[s->woff[0] + (uint64_t)c * s->m + j];
            float* z = Z[c];
            for (uint32_t ii = 0; ii < nb; ++ii) z[ii] += xj[ii] * w;
        }
    }
}

/* remaining layers for one row given its first-layer pre-activations; returns yhat */
static inline float tail_row(const shape_t* s, int act, const float* th, const float* z0, float a[MAXL][MAXW],
                             float dh[MAXL][MAXW]) {
    const uint32_t nl = s->nl;
    for (uint32_t c = 0; c < s->w[0]; ++c) {
        float h = act_h(act, z0[c]);
        a[0][c] = h; dh[0][c] = act_dh(act, z0[c], h);
    }
    for (uint32_t l = 1; l + 1 < nl; ++l)
        for (uint32_t c = 0; c < s->w[l]; ++c) {
            float z = th[s->boff[l] + c];
            const float* W = th + s->woff[l] + c * s->in[l];
            for (uint32_t k = 0; k < s->in[l]; ++k) z += a[l - 1][k] * W[k];
            float h = act_h(act, z);
            a[l][c] = h; dh[l][c] = act_dh(act, z, h);
        }
    float yh = 0.f;
    const float* Wo = th + s->woff[nl - 1];
    for (uint32_t k = 0; k < s->in[nl - 1]; ++k) yh += a[nl - 2][k] * Wo[k];
    return yh;
}

float bann_cpu_rss(const float* X, const float* y, uint64_t n, uint32_t m, const uint32_t* widths, uint32_t nl, int act,
                   const float* theta, float* yhat) {
    shape_t s; make_shape(&s, m, widths, nl);
    double rss = 0.0;
    int64_t nblk = (int64_t)((n + BLK - 1) / BLK);
#pragma omp parallel for schedule(static) reduction(+ : rss)
    for (int64_t bi = 0; bi < nblk; ++bi) {
        float Z[MAXW][BLK];
        uint64_t i0 = (uint64_t)bi * BLK;
        uint32_t nb = (uint32_t)((n - i0 < BLK) ? (n - i0) : BLK);
        first_layer_block(&s, theta, X, n, i0, nb, Z);
        for (uint32_t ii = 0; ii < nb; ++ii) {
            float a[MAXL][MAXW], dh[MAXL][MAXW], z0[MAXW];
            for (uint32_t c = 0; c < s.w[0]; ++c) z0[c] = Z[c][ii];
            float yh = tail_row(&s, act, theta, z0, a, dh);
            if (yhat) yhat[i0 + ii] = yh;
            float e = yh - y[i0 + ii];
            rss += (double)(e * e);
        }
    }
    return (float)rss;
}

/* backpropagate: d_rss (param_vec order) and rss */
float bann_cpu_backprop(const float* X, const float* y, uint64_t n, uint32_t m, const uint32_t* widths, uint32_t nl,
                        int act, const float* theta, float* d_rss) {
    shape_t s; make_shape(&s, m, widths, nl);
    const uint32_t P = s.P;
    int nt = bann_cpu_num_threads();
    double* acc = (double*)calloc((size_t)nt * (P + 1), sizeof(double));
    int64_t nblk = (int64_t)((n + BLK - 1) / BLK);
#pragma omp parallel
    {
#ifdef _OPENMP
        int tidx = omp_get_thread_num();
#else
        int tidx = 0;
#endif
        double* g = acc + (size_t)tidx * (P + 1);
        float Z[MAXW][BLK];   /* pre-activations, then delta_0 */
#pragma omp for schedule(static)
        for (int64_t bi = 0; bi < nblk; ++bi) {
            uint64_t i0 = (uint64_t)bi * BLK;
            uint32_t nb = (uint32_t)((n - i0 < BLK) ? (n - i0) : BLK);
            first_layer_block(&s, theta, X, n, i0, nb, Z);
            for (uint32_t ii = 0; ii < nb; ++ii) {
                float a[MAXL][MAXW], dh[MAXL][MAXW], delta[MAXW], nd[MAXW], z0[MAXW];
                for (uint32_t c = 0; c < s.w[0]; ++c) z0[c] = Z[c][ii];
                float yh = tail_row(&s, act, theta, z0, a, dh);
                float e = yh - y[i0 + ii];
                g[P] += (double)(e * e);
                const uint32_t L = nl - 2;
                const float* Wo = theta + s.woff[nl - 1];
                for (uint32_t k = 0; k < s.w[L]; ++k) { g[s.woff[nl - 1] + k] += a[L][k] * e; delta[k] = dh[L][k] * (e * Wo[k]); }
                for (uint32_t l = L; l >= 1; --l) {
                    for (uint32_t k = 0; k < s.in[l]; ++k) nd[k] = 0.f;
                    for (uint32_t c = 0; c < s.w[l]; ++c) {
                        g[s.boff[l] + c] += delta[c];
                        const float* W = theta + s.woff[l] + c * s.in[l];
                        for (uint32_t k = 0; k < s.in[l]; ++k) { g[s.woff[l] + c * s.in[l] + k] += a[l - 1][k] * delta[c]; nd[k] += delta[c] * W[k]; }
                    }
                    for (uint32_t k = 0; k < s.in[l]; ++k) delta[k] = dh[l - 1][k] * nd[k];
                }
                for (uint32_t c = 0; c < s.w[0]; ++c) { g[s.boff[0] + c] += delta[c]; Z[c][ii] = delta[c]; }
            }
            /* gW0[j, c] += sum_ii X[i0+ii, j] * delta_0[ii, c]  (contiguous over rows) */
            for (uint32_t j = 0; j < m; ++j) {
                const float* xj = X + (uint64_t)j * n + i0;
                for (uint32_t c = 0; c < s.w[0]; ++c) {
                    const float* dz = Z[c];
                    float t = 0.f;
#pragma omp simd reduction(+ : t)
                    for (uint32_t ii = 0; ii < nb; ++ii) t += xj[ii] * dz[ii];
                    g[s.woff[0] + (uint64_t)c * m + j] += t;
                }
            }
        }
    }
    float rss = 0.f;
    for (uint32_t k = 0; k <= P; ++k) {
        double t = 0.0;
        for (int q = 0; q < nt; ++q) t += acc[(size_t)q * (P + 1) + k];
        if (k < P) d_rss[k] = (float)t;
        else rss = (float)t;
    }
    free(acc);
    return rss;
}

static float log_density(uint32_t P, const float* th, const float* lam, int lasso, int stdn_bias_l2, uint32_t nbias,
                         float lam_e, float rss) {
    double prior = 0.0;
    for (uint32_t k = 0; k < P; ++k) {
        if (k >= P - nbias) { if (stdn_bias_l2) prior -= 0.5 * th[k] * th[k]; continue; }
        prior -= lasso ? lam[k] * fabsf(th[k]) : 0.5f * lam[k] * th[k] * th[k];
    }
    return (float)prior + (-1.0f * lam_e * (rss / 2.0f));
}

/* L leapfrog steps in the reference's op sequence; theta/mom updated in place; returns -H */
float bann_cpu_leapfrog(const float* X, const float* y, uint64_t n, uint32_t m, const uint32_t* widths, uint32_t nl,
                        int act, int lasso, int stdn, float* theta, float* mom, const float* eps, const float* lam,
                        float lam_e, uint32_t L, float* grad_scratch) {
    shape_t s; make_shape(&s, m, widths, nl);
    const uint32_t P = s.P;
    uint32_t nbias = 0;
    for (uint32_t l = 0; l + 1 < nl; ++l) nbias += s.w[l];
    float* g = grad_scratch;
    float negh = 0.f;
    bann_cpu_backprop(X, y, n, m, widths, nl, act, theta, g);
    for (uint32_t k = 0; k < P; ++k) {
        float sg = theta[k] > 0.f ? 1.f : (theta[k] < 0.f ? -1.f : 0.f);
        g[k] = (k >= P - nbias) ? -(lam_e * g[k]) : -(lam_e * g[k] + lam[k] * (lasso ? sg : theta[k]));
    }
    for (uint32_t step = 0; step < L; ++step) {
        for (uint32_t k = 0; k < P; ++k) { mom[k] += 0.5f * eps[k] * g[k]; theta[k] += eps[k] * mom[k]; }
        bann_cpu_backprop(X, y, n, m, widths, nl, act, theta, g);
        double kin = 0.0;
        for (uint32_t k = 0; k < P; ++k) {
            float sg = theta[k] > 0.f ? 1.f : (theta[k] < 0.f ? -1.f : 0.f);
            g[k] = (k >= P - nbias) ? -(lam_e * g[k]) : -(lam_e * g[k] + lam[k] * (lasso ? sg : theta[k]));
            mom[k] += 0.5f * eps[k] * g[k];
            kin += (double)mom[k] * mom[k];
        }
        float rss = bann_cpu_rss(X, y, n, m, widths, nl, act, theta, NULL); /* Q13: second forward pass */
        negh = log_density(P, theta, lam, lasso, stdn, nbias, lam_e, rss) - 0.5f * (float)kin;
    }
    return negh;
}
