// Host orchestration behind the C ABI: device-resident Net state, branch visits, grouped
// full-network leapfrog.  Mirrors Net<B>::train / predict / gradient (net/net.rs) and
// BranchSampler::hmc_step (net/branch/branch_sampler.rs:1192-1299); see include/bann.h.
#define BANN_NET_TU 1     // this translation unit owns the non-template kernels of kernels.cuh
#include <algorithm>
#include <cmath>

#include "chain.cuh"
#include "joint.cuh"
#include "probe.cuh"
#include "k1_small.cuh"
#include "k1_tc.cuh"
#include "k1_tc_wide.cuh"
#include "k1_tcx.cuh"
#include "k1_tcp.cuh"
#include "store.cuh"

using // which <= 64-marker tensor-core kernel gradient / leapfrog launches use unless bann_net_select_k1_tc_variant says otherwise
#ifndef BANN_TC_DEFAULT_VARIANT
#define BANN_TC_DEFAULT_VARIANT BANN_TC_FIVE_WARPS
#endif

namespace bann;

struct bann_net {
    bann_ctx* ctx = nullptr;
    bann_genotypes* gen = nullptr;
    int model = 0, act = 0;
    Hyper6 hyper;
    uint64_t B = 0;
    uint32_t n = 0;
    float n_total = 0.f;
    std::vector<BranchDesc> descs;
    BranchDesc* d_descs = nullptr;
    uint64_t total_params = 0, total_prec = 0;
    uint32_t maxP = 0, pstride = 0;
    size_t max_generic_smem = 0;
    float *d_theta = nullptr, *d_theta0 = nullptr, *d_mom = nullptr, *d_grad = nullptr, *d_eps = nullptr;
    float* d_prec = nullptr;
    BranchState* d_states = nullptr;
    NetGlobals* d_G = nullptr;
    float *d_y = nullptr, *d_r = nullptr, *d_t = nullptr, *d_prev = nullptr, *d_ynew = nullptr;
    float* d_part = nullptr;
    size_t part_cap = 0;
    float* d_gsum = nullptr;
    size_t gsum_cap = 0;
    float* d_rpart = nullptr;
    uint32_t rblk = 0;
    float* d_ow_others = nullptr;   // [B] global output-weight stat minus the branch's own, as of the branch's last Gibbs / from_cfg
    float *d_own_old = nullptr, *d_own_new = nullptr;   // [B] own output-weight stat before / after a transition (group visits)
    uint32_t* d_order = nullptr;    // branch order of the current sweep (group visits read sub-ranges of it)
    size_t order_cap = 0;
    float *d_Tg = nullptr, *d_Yg = nullptr;   // group visits: per-member targets / final predictions [G][n]
    size_t tg_cap = 0, yg_cap = 0;
    float* d_inj_grp = nullptr;     // injected randomness of a group visit: momenta + step uniforms (arena layout), u[B], gammas[B][stride]
    uint32_t inj_grp_stride = 0;
    float* d_bias2 = nullptr;
    float* d_lpd_local = nullptr;
    int* d_errflag = nullptr;
    uint32_t* d_list_all = nullptr;
    float* d_inj = nullptr;   // scratch for injected randomness: momenta[maxP] uniforms[maxP] u[1] gammas[...]
    size_t inj_gamma_cap = 0;
    float* d_T = nullptr;     // grouped per-branch targets [B][n]
    int grouped_per_branch = 0;
    uint64_t grouped_seed = 0;
    float* d_traj = nullptr;
    float* d_numgrad = nullptr;   // numerical_ldg: [maxP] gradient + [1] base log density
    size_t traj_cap = 0;
    float* d_scratchB = nullptr;  // [3*B] gather buffer
    uint64_t visit_seq = 0;
    int k1_mode = BANN_K1_AUTO;   // which K1 kernel launch_k1 may pick (bann_net_select_k1)
    int tc_variant = BANN_TC_DEFAULT_VARIANT;   // k1_tc (four warps) or k1_tc5 (dedicated issuing warp) (bann_net_select_k1_tc_variant)
    const char* last_k1 = "none"; // kernel family the last fused forward+backward launch used (bann_net_last_k1_kernel)
    int hmc_path = BANN_HMC_AUTO; // per-branch transitions: persistent cooperative kernel where eligible / launch per step (bann_net_select_hmc_path)
    uint2* d_tcp_words = nullptr;        // tagged exchange words of the persistent kernel: [4][kTcpMaxGrid][pstride] partials, [4][kTcpSumCopies][pstride] sums
    uint32_t tcp_tag = 0;                // every tag in d_tcp_words (and in the ranks' tables behind the bulk-exchange region) is <= tcp_tag
    uint32_t tcp_launches = 0;           // persistent launches so far (its parity selects the slot pair)
    // zero-padded architectures: widths no tensor-core kernel is instantiated for run through the next larger instantiated
    // architecture on a padded copy of the parameters (padded units have zero weights and biases: activation 0, delta 0)
    bool pad_built = false, pad_active = false;
    std::vector<BranchDesc> pad_descs;   // per branch: the padded description, or the real one when there is no target
    std::vector<uint8_t> pad_ok;
    BranchDesc* d_pad_descs = nullptr;
    float* d_pad_theta = nullptr;
    uint32_t pad_pstride = 0;
    float* d_pad_gsum = nullptr; size_t pad_gsum_cap = 0;
    float* d_pad_part = nullptr; size_t pad_part_cap = 0;
    uint64_t padded_launches = 0;
    uint64_t persistent_launches = 0;
    float *h_pin_a = nullptr, *h_pin_b = nullptr;  // pinned staging for bann_net_gradient (callers with pageable buffers)
    float *d_dense_in = nullptr, *d_dense_out = nullptr;   // dense host-facing layouts of the parameters / gradients + rss
    uint64_t sum_params = 0;
    // joint HMC / gradient-ascent modes (joint.cuh), one branch at a time: [prec0 | pmom | pgrad | peps] x maxQ,
    // injected momenta / step uniforms (maxP + maxQ each), accept uniform + kinetic energy + 3 outputs
    uint32_t maxQ = 0;
    float* d_jws = nullptr;
    float* d_tcx[3] = {nullptr, nullptr, nullptr};   // k1_tcx.cuh work buffers
    // bulk peer-memory exchange of the grouped schedule (comm.cuh: XgComm), connected by bann_net_comm_connect
    uint8_t* xg_region[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool xg_ipc[8] = {false, false, false, false, false, false, false, false};
    bool xg_connected = false;
    uint32_t xg_epoch = 0;
    uint64_t xg_cap = 0;            // floats per in / out buffer
    unsigned int* d_xg_counter = nullptr;
    size_t tcx_cap[3] = {0, 0, 0};
};

static int ensure_cap(float** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    BANN_CUDA(cudaMalloc(p, need * sizeof(float)));
    *cap = need;
    return 0;
}

namespace bann {
int launch_hmc_persistent(const BranchDesc& d0, int act, TcpArgs& a, int num_sms, cudaStream_t st, bool* launched);
}

// ------------------------------------------------------------------ K1 launch
struct K1Launch {
    const uint32_t* list = nullptr;  // device
    uint32_t nlist = 0;
    const BranchState* states = nullptr;
    int target_mode = TGT_SHARED;
    const float* tgt = nullptr;
    const float* resid = nullptr;
    float* tgt_out = nullptr;
    float* prev_out = nullptr;
    int out_per_entry = 0;
    float* yhat_out = nullptr;
    int yhat_accumulate = 0;
    int fwd_only = 0;
    // alternative store (predict on test data)
    const bann_genotypes* store = nullptr;
    const BranchDesc* descs_dev = nullptr;
    int single_branch = -1;  // host index of the only branch in the list (for kernel selection), -1: all
    bool xr = false;         // sum the reduced [gW | gb | rss] over ranks inside KR (sequential schedule, world > 1)
    bool xg = false;         // sum them over ranks with the bulk exchange after KR (launches over many branches, world > 1)
};

static XrComm xr_none() {
    XrComm c;
    memset(&c, 0, sizeof(c));
    c.world = 1;
    return c;
}
// sequential-exact entry points on sharded rows need the peer-memory exchange
static int need_comm(bann_net* net) {
    if (net->ctx->world > 1 && !net->ctx->xr_connected)
        BANN_FAIL("rows are sharded over ranks: call bann_ctx_comm_handle / bann_ctx_comm_connect first");
    return 0;
}
static bool sharded(const bann_net* net) { return net->ctx->world > 1; }

// ---- bulk exchange (comm.cuh): descriptor of the NEXT exchange, all-reduce(sum) in place, all-gather of 1/world slices
// [flags | in[2][cap] | out[2][cap]] floats, then the rank-level sums table of the persistent HMC kernel: [4][8][pstride] tagged words
static size_t xg_table_offset(uint64_t cap) { return 256 + (size_t)4 * cap * sizeof(float); }
static size_t xg_region_bytes(uint64_t cap, uint32_t pstride) { return xg_table_offset(cap) + (size_t)4 * 8 * pstride * sizeof(uint2); }
static XgComm xg_next(bann_net* net) {
    XgComm c;
    memset(&c, 0, sizeof(c));
    for (int r = 0; r < net->ctx->world; ++r) c.region[r] = net->xg_region[r];
    c.rank = (uint32_t)net->ctx->rank;
    c.world = (uint32_t)net->ctx->world;
    c.epoch = ++net->xg_epoch;
    c.cap = net->xg_cap;
    c.error_flag = net->d_errflag;
    c.counter = net->d_xg_counter;
    return c;
}
static uint32_t xg_grid(const bann_net* net, uint64_t n4) {
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)net->ctx->num_sms * 2, (n4 + 255) / 256));
}
// buf[0 .. count) <- sum over ranks, identical bits on every rank; count is rounded up to a multiple of 4 (buf must hold that)
static int xg_allreduce(bann_net* net, float* buf, uint64_t count) {
    if (!net->xg_connected) BANN_FAIL("rows are sharded over ranks: call bann_net_comm_handle / bann_net_comm_connect first");
    const uint64_t c4 = (count + 3) & ~3ull, n4 = c4 / 4, per4 = (n4 + net->ctx->world - 1) / net->ctx->world;
    if (c4 > net->xg_cap) BANN_FAIL("bulk exchange: more values than the exchange region holds");
    cudaStream_t st = net->ctx->stream;
    XgComm c = xg_next(net);
    k_xg_publish<<<xg_grid(net, n4), 256, 0, st>>>(c, buf, c4);
    BANN_LAUNCHED();
    k_xg_reduce_scatter<<<xg_grid(net, per4), 256, 0, st>>>(c, buf, c4, per4);
    BANN_LAUNCHED();
    k_xg_all_gather<<<xg_grid(net, n4), 256, 0, st>>>(c, buf, c4, per4);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
// every rank holds its slice [rank * per4, (rank + 1) * per4) (float4 units) of buf: fill in the others' slices
static int xg_allgather_slices(bann_net* net, float* buf, uint64_t count) {
    if (!net->xg_connected) BANN_FAIL("rows are sharded over ranks: call bann_net_comm_handle / bann_net_comm_connect first");
    const uint64_t c4 = (count + 3) & ~3ull, n4 = c4 / 4, per4 = (n4 + net->ctx->world - 1) / net->ctx->world;
    if (c4 > net->xg_cap) BANN_FAIL("bulk exchange: more values than the exchange region holds");
    cudaStream_t st = net->ctx->stream;
    XgComm c = xg_next(net);
    k_xg_publish_slice<<<xg_grid(net, per4), 256, 0, st>>>(c, buf, c4, per4);
    BANN_LAUNCHED();
    k_xg_all_gather<<<xg_grid(net, n4), 256, 0, st>>>(c, buf, c4, per4);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
// this rank's share [lo, hi) of a host-facing vector of `count` floats (the slices of xg_allgather_slices)
static void xg_slice(const bann_net* net, uint64_t count, uint64_t* lo, uint64_t* hi) {
    const uint64_t c4 = (count + 3) & ~3ull, n4 = c4 / 4, per4 = (n4 + net->ctx->world - 1) / net->ctx->world;
    *lo = std::min<uint64_t>(count, 4 * per4 * net->ctx->rank);
    *hi = std::min<uint64_t>(count, 4 * per4 * (net->ctx->rank + 1));
}

// ------------------------------------------------------------------ zero-padded architectures
// index of real parameter k of `dr` inside the padded layout `dp` (same depth, widths >= the real ones)
__device__ __forceinline__ uint32_t padded_index(const BranchDesc& dr, const BranchDesc& dp, uint32_t k) {
    int l; uint32_t row, col; bool isb;
    locate_param(dr, k, l, row, col, isb);
    return isb ? dp.b_off[l] + col : dp.w_off[l] + col * dp.in_dim[l] + row;
}
__global__ void __launch_bounds__(256) k_pad_params(const BranchDesc* __restrict__ real, const BranchDesc* __restrict__ pad,
                                                    const uint32_t* __restrict__ list, const float* __restrict__ theta,
                                                    float* __restrict__ theta_pad) {
    const uint32_t b = list ? list[blockIdx.x] : blockIdx.x;
    const BranchDesc& dr = real[b];
    const BranchDesc& dp = pad[b];
    for (uint32_t k = threadIdx.x; k < dr.P; k += 256) theta_pad[dp.param_off + padded_index(dr, dp, k)] = theta[dr.param_off + k];
}
__global__ void __launch_bounds__(256) k_unpad_sums(const BranchDesc* __restrict__ real, const BranchDesc* __restrict__ pad,
                                                    const uint32_t* __restrict__ list, const float* __restrict__ gsum_pad,
                                                    uint32_t pstride_pad, float* __restrict__ gsum, uint32_t pstride) {
    const uint32_t li = blockIdx.x, b = list ? list[li] : li;
    const BranchDesc& dr = real[b];
    const BranchDesc& dp = pad[b];
    const float* src = gsum_pad + (size_t)li * pstride_pad;
    float* dst = gsum + (size_t)li * pstride;
    for (uint32_t k = threadIdx.x; k < dr.P; k += 256) dst[k] = src[padded_index(dr, dp, k)];
    if (threadIdx.x == 0) dst[dr.P] = src[dp.P];     // rss
}

// the padded description of every branch: hidden / summary widths raised to the next instantiated pair, same depth
static int build_padded(bann_net* net) {
    if (net->pad_built) return 0;
    net->pad_descs = net->descs;
    net->pad_ok.assign(net->B, 0);
    uint64_t poff = 0;
    uint32_t maxP = 0;
    for (uint64_t b = 0; b < net->B; ++b) {
        BranchDesc& d = net->pad_descs[b];
        const int D = (int)d.nl - 2;
        const uint32_t S = d.widths[d.nl - 2];
        uint32_t H = D > 0 ? d.widths[0] : S;
        bool homogeneous = true;
        for (int l = 0; l < D; ++l) homogeneous = homogeneous && d.widths[l] == H;
        const uint32_t mx = std::max(H, S);
        uint32_t target = 0;
        if (homogeneous && D >= 0 && D <= 2) {
            if (mx <= 5 && d.m <= 512 && (D > 0 || d.m <= 64)) target = 5;          // k1_tc / k1_tcw: (5,5,0|1|2)
            else if (D >= 1 && mx <= 8) target = 8;                                  // k1_tcx: (8,8,1|2), (12,12,1|2), (16,16,1|2)
            else if (D >= 1 && mx <= 12) target = 12;
            else if (D >= 1 && mx <= 16) target = 16;
        }
        if (target) {
            net->pad_ok[b] = 1;
            for (uint32_t l = 0; l + 1 < d.nl; ++l) d.widths[l] = target;
            uint32_t prev = d.m, off = 0, aoff = 0;
            for (uint32_t l = 0; l < d.nl; ++l) {
                d.in_dim[l] = prev;
                d.w_off[l] = off;
                off += prev * d.widths[l];
                if (l + 1 < d.nl) { d.a_off[l] = aoff; aoff += d.widths[l]; }
                prev = d.widths[l];
            }
            d.sumw = aoff;
            for (uint32_t l = 0; l + 1 < d.nl; ++l) { d.b_off[l] = off; off += d.widths[l]; }
            d.P = off;
        }
        d.param_off = poff;
        poff += (d.P + 3) & ~3u;
        maxP = std::max(maxP, d.P);
    }
    net->pad_pstride = (maxP + 1 + 3) & ~3u;
    BANN_CUDA(cudaMalloc(&net->d_pad_descs, net->B * sizeof(BranchDesc)));
    BANN_CUDA(cudaMemcpy(net->d_pad_descs, net->pad_descs.data(), net->B * sizeof(BranchDesc), cudaMemcpyHostToDevice));
    BANN_CUDA(cudaMalloc(&net->d_pad_theta, (poff + 4) * sizeof(float)));
    BANN_CUDA(cudaMemset(net->d_pad_theta, 0, (poff + 4) * sizeof(float)));          // the padded slots stay zero for good
    BANN_CHECK(ensure_cap(&net->d_pad_gsum, &net->pad_gsum_cap, (size_t)net->B * net->pad_pstride));
    net->pad_built = true;
    return 0;
}

static int launch_k1(bann_net* net, const K1Launch& L, bool reduce) {
    const bann_genotypes* g = L.store ? L.store : net->gen;
    cudaStream_t st = net->ctx->stream;
    uint32_t ntiles = g->ntiles;
    // chunking: enough CTAs to fill the machine, never more chunks than tiles
    uint32_t want = (uint32_t)std::max<uint64_t>(1, ((uint64_t)net->ctx->num_sms * 4 + L.nlist - 1) / L.nlist);
    uint32_t nchunk = std::min(want, ntiles);
    uint32_t tpc = (ntiles + nchunk - 1) / nchunk;
    nchunk = (ntiles + tpc - 1) / tpc;
    if (L.fwd_only) {  // no partials: one CTA per few tiles
        tpc = std::max<uint32_t>(1, tpc);
    }
    float* part = nullptr;
    if (!L.fwd_only) {
        BANN_CHECK(ensure_cap(&net->d_gsum, &net->gsum_cap, (size_t)L.nlist * net->pstride));
        if (nchunk == 1) part = net->d_gsum;
        else {
            BANN_CHECK(ensure_cap(&net->d_part, &net->part_cap, (size_t)L.nlist * nchunk * net->pstride));
            part = net->d_part;
        }
    }
    K1Args a;
    a.store = g->d_store;
    a.descs = L.descs_dev ? L.descs_dev : net->d_descs;
    a.theta = net->d_theta;
    a.mu = g->d_mu;
    a.sd = g->d_sd;
    a.list = L.list;
    a.states = L.states;
    a.n = (uint32_t)g->n;
    a.ntiles = ntiles;
    a.tiles_per_chunk = tpc;
    a.nchunk = nchunk;
    a.target_mode = L.target_mode;
    a.tgt = L.tgt;
    a.resid = L.resid;
    a.tgt_out = L.tgt_out;
    a.prev_out = L.prev_out;
    a.out_per_entry = L.out_per_entry;
    a.yhat_out = L.yhat_out;
    a.yhat_accumulate = L.yhat_accumulate;
    a.fwd_only = L.fwd_only;
    a.part = part;
    a.pstride = net->pstride;
    a.act = net->act;
    a.store_tc = g->d_store_tc;
    a.nst = g->nst;
    a.st_per_chunk = 0;
    a.ncb = 8;
    a.nc_uniform = 0;
    a.tc_variant = net->tc_variant;

    bool launched = false;
    if (net->k1_mode == BANN_K1_AUTO || net->k1_mode == BANN_K1_TENSOR) {
        int r = launch_k1_tc(net->descs, L.single_branch, a, L.nlist, net->ctx->num_sms, st, &launched,
                             &nchunk, L.fwd_only ? nullptr : &part, net);
        if (r != 0) return r;
        if (launched)
            net->last_k1 = a.tc_variant == 101 ? "k1_tc5<H,S,D,ACT> (tcgen05 + tensor memory, dedicated issuing warp, <= 64 markers per branch)"
                                               : "k1_tc<H,S,D,ACT> (tcgen05 + tensor memory, <= 64 markers per branch)";
        if (!launched) {   // 65..512 markers per branch: the K-blocked variant
            r = launch_k1_tcw(net->descs, L.single_branch, a, L.nlist, net->ctx->num_sms, st, &launched, &nchunk,
                              L.fwd_only ? nullptr : &part, net);
            if (r != 0) return r;
            if (launched) net->last_k1 = "k1_tcw<H,S,D,ACT> (tcgen05, K-blocked, 65..512 markers per branch)";
        }
        if (!launched) {   // wide first layers (up to 16 units) / more markers: the three-pass variant
            r = launch_k1_tcx(net->descs, L.single_branch, a, L.nlist, net->ctx->num_sms, st, &launched, &nchunk,
                              L.fwd_only ? nullptr : &part, net);
            if (r != 0) return r;
            if (launched) net->last_k1 = "k1_tcx: k_tcx_fwd + k_tcx_tail + k_tcx_bwd (tcgen05, three passes, first-layer width <= 16)";
        }
        // widths outside the instantiated set: the next larger instantiated architecture on a zero-padded copy of the parameters
        if (!launched && !L.descs_dev && !L.store) {
            BANN_CHECK(build_padded(net));
            bool all_ok = true;
            if (L.single_branch >= 0) all_ok = net->pad_ok[L.single_branch] != 0;
            else for (uint64_t b = 0; b < net->B && all_ok; ++b) all_ok = net->pad_ok[b] != 0;
            if (all_ok) {
                k_pad_params<<<L.nlist, 256, 0, st>>>(net->d_descs, net->d_pad_descs, L.list, net->d_theta, net->d_pad_theta);
                BANN_LAUNCHED();
                BANN_CHECK(ensure_cap(&net->d_pad_gsum, &net->pad_gsum_cap, (size_t)L.nlist * net->pad_pstride));
                K1Args ap = a;
                ap.descs = net->d_pad_descs;
                ap.theta = net->d_pad_theta;
                ap.pstride = net->pad_pstride;
                float* ppart = nullptr;
                uint32_t pchunk = nchunk;
                net->pad_active = true;        // the kernels' buffer hooks (partials, sums, stride) now answer for the padded layout
                int r = launch_k1_tc(net->pad_descs, L.single_branch, ap, L.nlist, net->ctx->num_sms, st, &launched, &pchunk,
                                     L.fwd_only ? nullptr : &ppart, net);
                if (r == 0 && !launched)
                    r = launch_k1_tcw(net->pad_descs, L.single_branch, ap, L.nlist, net->ctx->num_sms, st, &launched, &pchunk,
                                      L.fwd_only ? nullptr : &ppart, net);
                if (r == 0 && !launched)
                    r = launch_k1_tcx(net->pad_descs, L.single_branch, ap, L.nlist, net->ctx->num_sms, st, &launched, &pchunk,
                                      L.fwd_only ? nullptr : &ppart, net);
                net->pad_active = false;
                if (r != 0) return r;
                if (launched) {
                    net->padded_launches += 1;
                    net->last_k1 = "tensor-core kernel of the next larger instantiated architecture on zero-padded parameters";
                    if (!L.fwd_only) {
                        if (ppart != net->d_pad_gsum) {
                            dim3 grid((net->pad_pstride + 31) / 32, L.nlist);
                            BANN_CUDA(launch_pdl(k_reduce_partials, grid, dim3(256), 0, st, (const float*)ppart, net->d_pad_gsum, pchunk,
                                                 net->pad_pstride, L.list, (const BranchDesc*)net->d_pad_descs, L.states, xr_none()));
                            BANN_LAUNCHED();
                        }
                        k_unpad_sums<<<L.nlist, 256, 0, st>>>(net->d_descs, net->d_pad_descs, L.list, net->d_pad_gsum, net->pad_pstride,
                                                              net->d_gsum, net->pstride);
                        BANN_LAUNCHED();
                        BANN_CUDA(cudaGetLastError());
                        part = net->d_gsum;      // the sums are complete and in place: what follows is the cross-rank exchange only
                        nchunk = 1;
                    }
                }
            }
        }
        if (!launched && net->k1_mode == BANN_K1_TENSOR)
            BANN_FAIL("tensor-core K1 requested but the launch is not eligible (homogeneous architecture, widths up to 16, depth up to 2)");
    }
    if (!launched && !g->d_store)
        BANN_FAIL("the byte-tile store was released (bann_genotypes_release_byte_store) and no tensor-core kernel is eligible for this launch");
    if (!launched && net->k1_mode != BANN_K1_GENERIC) {
        int r = launch_k1_small(net->descs, L.single_branch, a, L.nlist, net->ctx->num_sms, st, &launched,
                                &nchunk, L.fwd_only ? nullptr : &part, net);
        if (r != 0) return r;
        if (launched) net->last_k1 = "k1_small<H,S,D,NP,NW,ACT> (FFMA)";
    }
    if (!launched) {
        net->last_k1 = "k1_generic (shape-agnostic)";
        dim3 grid(nchunk, L.nlist);
        k1_generic<<<grid, 128, net->max_generic_smem, st>>>(a);
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
    }
    if (reduce && !L.fwd_only && (part != net->d_gsum || L.xr)) {   // nchunk == 1 with L.xr: in place, exchange only
        if (L.xr && (size_t)L.nlist * net->pstride > kXrCap) BANN_FAIL("peer-memory exchange: too many values in one launch");
        dim3 grid((net->pstride + 31) / 32, L.nlist);
        BANN_CUDA(launch_pdl(k_reduce_partials, grid, dim3(256), 0, st, (const float*)part, net->d_gsum, nchunk, net->pstride, L.list,
                             a.descs, L.states, L.xr ? xr_next(net->ctx, net->d_errflag) : xr_none()));
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
    }
    if (L.xg && !L.fwd_only) BANN_CHECK(xg_allreduce(net, net->d_gsum, (uint64_t)L.nlist * net->pstride));
    return 0;
}

// hooks used by k1_small.cuh to size its partial buffers
float* bann_net_partials(bann_net* net, size_t need) {
    if (net->pad_active) {
        if (ensure_cap(&net->d_pad_part, &net->pad_part_cap, need) != 0) return nullptr;
        return net->d_pad_part;
    }
    if (ensure_cap(&net->d_part, &net->part_cap, need) != 0) return nullptr;
    return net->d_part;
}
float* bann_net_gsum(bann_net* net) { return net->pad_active ? net->d_pad_gsum : net->d_gsum; }
// work buffers of the three-pass wide kernel (k1_tcx.cuh): 0 = W' pieces, 1 = first-layer activations, 2 = delta pieces
namespace bann {
float* bann_net_tcx_buffer(bann_net* net, int which, size_t bytes) {
    if (ensure_cap(&net->d_tcx[which], &net->tcx_cap[which], (bytes + 3) / 4 + 64) != 0) return nullptr;
    return net->d_tcx[which];
}
}  // namespace bann
uint32_t bann_net_pstride(bann_net* net) { return net->pad_active ? net->pad_pstride : net->pstride; }

// gradient under the prior without touching the HMC state (log_density_gradient, a8)
__global__ void __launch_bounds__(256) k_grad_only(const BranchDesc* descs, const uint32_t* list, const float* theta,
                                                   const float* prec, const float* gsum, uint32_t pstride, int model,
                                                   float* grad) {
    const uint32_t li = blockIdx.x;
    const uint32_t b = list ? list[li] : li;
    const BranchDesc& d = descs[b];
    const float* th = theta + d.param_off;
    const float* pr = prec + d.prec_off;
    const float* gs = gsum + (size_t)li * pstride;
    const float lam_e = pr[d.ep_off];
    const bool lasso = (model == BANN_LASSO_BASE || model == BANN_LASSO_ARD);
    for (uint32_t k = threadIdx.x; k < d.P; k += 256) {
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        const float w = th[k];
        float g;
        if (isb) g = -(lam_e * gs[k]);
        else {
            const float lam = param_prior_precision(d, pr, model, l, row, false);
            if (model == BANN_STD_NORMAL) g = -(lam_e * gs[k] + w);
            else if (lasso) g = -(lam_e * gs[k] + lam * ((w > 0.f) ? 1.f : (w < 0.f ? -1.f : 0.f)));
            else g = -(lam_e * gs[k] + lam * w);
        }
        grad[d.param_off + k] = g;
    }
}

// log_density(params, precisions, rss) (branch_sampler.rs:72-78)
__global__ void __launch_bounds__(256) k_log_density(const BranchDesc* descs, uint32_t b, const float* theta,
                                                     const float* prec, int model, float rss, float* out) {
    __shared__ float red[8];
    const BranchDesc& d = descs[b];
    const float* th = theta + d.param_off;
    const float* pr = prec + d.prec_off;
    const bool lasso = (model == BANN_LASSO_BASE || model == BANN_LASSO_ARD);
    float prior = 0.f;
    for (uint32_t k = threadIdx.x; k < d.P; k += 256) {
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        const float w = th[k];
        if (isb) {
            if (model == BANN_STD_NORMAL) prior -= 0.5f * w * w;
        } else {
            const float lam = param_prior_precision(d, pr, model, l, row, false);
            if (model == BANN_STD_NORMAL) prior -= 0.5f * w * w;
            else if (lasso) prior -= lam * fabsf(w);
            else prior -= 0.5f * lam * w * w;
        }
    }
    prior = block_sum<256>(prior, red);
    if (threadIdx.x == 0) *out = prior + (-1.0f * pr[d.ep_off] * (rss / 2.0f));
}

// ---- numerical_ldg (branch_sampler.rs:480-504, the reference's own debugging aid: "DO NOT run this in production code"):
// forward differences (log_density(theta + delta e_k) - log_density(theta)) / delta with NUMERICAL_DELTA = 0.001 (:30), one
// fused pass per parameter.  The perturbed vector is walked exactly as the reference does (+= delta, evaluate, -= delta).
constexpr float kNumericalDelta = 0.001f;
__global__ void k_numgrad_poke(float* p, float delta) { *p += delta; }
// log density at the current theta from the rss the fused pass just reduced; base == NULL: store it, else the difference quotient
__global__ void __launch_bounds__(256) k_numgrad_eval(const BranchDesc* descs, uint32_t b, const float* theta, const float* prec,
                                                      int model, const float* rss_ptr, const float* base, float delta, float* out) {
    __shared__ float red[8];
    const BranchDesc& d = descs[b];
    const float* th = theta + d.param_off;
    const float* pr = prec + d.prec_off;
    const bool lasso = (model == BANN_LASSO_BASE || model == BANN_LASSO_ARD);
    float prior = 0.f;
    for (uint32_t k = threadIdx.x; k < d.P; k += 256) {
        int l; uint32_t row, col; bool isb;
        locate_param(d, k, l, row, col, isb);
        const float w = th[k];
        if (isb) {
            if (model == BANN_STD_NORMAL) prior -= 0.5f * w * w;
        } else {
            const float lam = param_prior_precision(d, pr, model, l, row, false);
            if (model == BANN_STD_NORMAL) prior -= 0.5f * w * w;
            else if (lasso) prior -= lam * fabsf(w);
            else prior -= 0.5f * lam * w * w;
        }
    }
    prior = block_sum<256>(prior, red);
    if (threadIdx.x == 0) {
        const float ld = prior + (-1.0f * pr[d.ep_off] * (*rss_ptr / 2.0f));
        *out = base ? (ld - *base) / delta : ld;
    }
}

__global__ void k_gather_states(const BranchState* st, uint32_t B, float* h_init, float* h_cur, int* status) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    h_init[b] = st[b].neg_h_init;
    h_cur[b] = st[b].neg_h_cur;
    status[b] = st[b].status;
}

__global__ void k_count_status(const BranchState* st, uint32_t B, unsigned long long* out /*[2]*/) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (st[b].status == ST_ACCEPTED) atomicAdd(&out[0], 1ull);
    if (st[b].status == ST_REJECTED_EARLY) atomicAdd(&out[1], 1ull);
}

__global__ void k_gather_rss(const float* gsum, uint32_t pstride, const BranchDesc* descs, uint32_t B, float* out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) out[b] = gsum[(size_t)b * pstride + descs[b].P];
}

// ------------------------------------------------------------------ HMC driver
struct HmcRun {
    const uint32_t* list;   // device
    uint32_t nlist;
    int single_branch;
    int first_mode;         // target mode of the first evaluation
    const float* tgt;       // shared target (d_t / d_y) or per-entry base
    int later_mode;
    const float* inj_mom = nullptr;
    const float* inj_su = nullptr;
    const float* inj_u = nullptr;
    int inj_arena = 0;      // injected momenta / step uniforms in the parameter arena layout (group visits)
    int out_per_entry = 0;  // tgt_out / prev_out / ynew_out hold one row vector per list entry (group visits)
    bool xg = false;        // sharded rows: all-reduce the sums inside the library (else the caller does, bann_grouped_phase_a / _b)
    uint64_t seed = 0;
    uint64_t stream_base = 0;
    float* traj_params = nullptr;
    float* traj_ldg = nullptr;
    float* traj_h = nullptr;
    float* traj_num_ldg = nullptr;   // [L][P]: numerical_ldg per step (mcmc_cfg.num_grad_traj)
};

static int hmc_init_and_first_eval(bann_net* net, const bann_mcmc_cfg* cfg, const HmcRun& R, int keep_momenta) {
    cudaStream_t st = net->ctx->stream;
    const bool ard = (net->model == BANN_RIDGE_ARD || net->model == BANN_LASSO_ARD);
    if (cfg->hmc_step_size_mode == BANN_STEP_STD_SCALED && ard)
        BANN_FAIL("StdScaled step sizes are not implemented for ARD priors (reference returns empty vectors, ridge_ard.rs:56-68)");
    if (cfg->hmc_step_size_mode < 0 || cfg->hmc_step_size_mode > 3) BANN_FAIL("bad step size mode");
    InitArgs ia;
    ia.descs = net->d_descs;
    ia.list = R.list;
    ia.states = net->d_states;
    ia.theta = net->d_theta;
    ia.theta0 = net->d_theta0;
    ia.mom = net->d_mom;
    ia.eps = net->d_eps;
    ia.prec = net->d_prec;
    ia.model = net->model;
    ia.step_mode = cfg->hmc_step_size_mode;
    ia.factor = cfg->hmc_step_size_factor;
    ia.L = (float)cfg->hmc_integration_length;
    ia.inj_momenta = R.inj_mom;
    ia.inj_step_uniforms = R.inj_su;
    ia.inj_arena = R.inj_arena;
    ia.seed = R.seed;
    ia.stream_base = R.stream_base;
    ia.keep_momenta = keep_momenta;
    k_hmc_init<<<R.nlist, 256, 0, st>>>(ia);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

static K2Args make_k2(bann_net* net, const bann_mcmc_cfg* cfg, const HmcRun& R, int mode_init, int is_last) {
    K2Args a;
    a.descs = net->d_descs;
    a.list = R.list;
    a.states = net->d_states;
    a.theta = net->d_theta;
    a.theta0 = net->d_theta0;
    a.mom = net->d_mom;
    a.grad = net->d_grad;
    a.eps = net->d_eps;
    a.prec = net->d_prec;
    a.gsum = net->d_gsum;
    a.pstride = net->pstride;
    a.model = net->model;
    a.max_h_err = cfg->hmc_max_hamiltonian_error;
    a.mode_init = mode_init;
    a.is_last = is_last;
    a.traj_params = R.traj_params;
    a.traj_ldg = R.traj_ldg;
    a.traj_h = R.traj_h;
    a.num_ldg = nullptr;
    return a;
}

// numerical_ldg of branch b against the device target vector `tgt` at the CURRENT parameters: P + 1 fused passes.
// Result in net->d_numgrad[0 .. P); d_gsum holds the sums of the LAST perturbed evaluation afterwards (callers re-evaluate).
static int numerical_ldg_async(bann_net* net, uint32_t b, const float* tgt) {
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    if (!net->d_numgrad) BANN_CUDA(cudaMalloc(&net->d_numgrad, ((size_t)net->maxP + 4) * sizeof(float)));
    float* base = net->d_numgrad + net->maxP;
    K1Launch k;
    k.list = net->d_list_all + b;
    k.nlist = 1;
    k.single_branch = (int)b;
    k.target_mode = TGT_SHARED;
    k.tgt = tgt;
    k.xr = sharded(net);
    BANN_CHECK(launch_k1(net, k, true));
    k_numgrad_eval<<<1, 256, 0, st>>>(net->d_descs, b, net->d_theta, net->d_prec, net->model, net->d_gsum + d.P, nullptr,
                                      kNumericalDelta, base);
    BANN_LAUNCHED();
    for (uint32_t pix = 0; pix < d.P; ++pix) {
        k_numgrad_poke<<<1, 1, 0, st>>>(net->d_theta + d.param_off + pix, kNumericalDelta);
        BANN_LAUNCHED();
        BANN_CHECK(launch_k1(net, k, true));
        k_numgrad_eval<<<1, 256, 0, st>>>(net->d_descs, b, net->d_theta, net->d_prec, net->model, net->d_gsum + d.P, base,
                                          kNumericalDelta, net->d_numgrad + pix);
        BANN_LAUNCHED();
        k_numgrad_poke<<<1, 1, 0, st>>>(net->d_theta + d.param_off + pix, -kNumericalDelta);
        BANN_LAUNCHED();
    }
    BANN_CUDA(cudaGetLastError());
    return 0;
}

// full HMC transition for the listed branches (sequential-exact when nlist == 1)
static int run_hmc(bann_net* net, const bann_mcmc_cfg* cfg, const HmcRun& R, float* ynew_out, float* tgt_out,
                   float* prev_out, const float* resid) {
    cudaStream_t st = net->ctx->stream;
    BANN_CHECK(hmc_init_and_first_eval(net, cfg, R, 0));
    const uint32_t Lsteps = cfg->hmc_integration_length;
    // ---- one branch, no per-step recording: the whole trajectory in ONE cooperative launch (k1_tcp.cuh) where the branch is
    //      eligible; otherwise (and for A/B tests: bann_net_select_hmc_path) three launches per leapfrog step.  On sharded rows
    //      every rank runs the kernel on its shard and the cross-rank sums travel inside it (eligibility must not depend on the
    //      rank: it is decided on the largest shard)
    const uint64_t max_shard = (net->gen->n_total + net->ctx->world - 1) / net->ctx->world;
    if (R.nlist == 1 && R.single_branch >= 0 && (!sharded(net) || net->xg_connected)
        && (max_shard + kTcRows - 1) / kTcRows <= (uint64_t)net->ctx->num_sms * kTcpMaxTiles && !R.traj_params && !R.traj_h && !cfg->num_grad && !cfg->num_grad_traj
        && net->hmc_path != BANN_HMC_LAUNCHES && net->k1_mode == BANN_K1_AUTO
        && (R.first_mode == TGT_RESID_PLUS_PRED || R.first_mode == TGT_SHARED)) {
        const uint32_t b = (uint32_t)R.single_branch;
        const size_t tcp_words = (size_t)(4 * kTcpMaxGrid + 4 * kTcpSumCopies) * net->pstride;
        if (sharded(net) && net->tcp_tag > 0xffffffffu - (Lsteps + 2u)) BANN_FAIL("persistent HMC kernel on sharded rows: 2^32 evaluations reached, re-create the net");
        if (!net->d_tcp_words || net->tcp_tag > 0xffffffffu - (Lsteps + 2u)) {      // first use / the 32-bit tags would wrap
            if (!net->d_tcp_words) BANN_CUDA(cudaMalloc(&net->d_tcp_words, tcp_words * sizeof(uint2)));
            BANN_CUDA(cudaMemsetAsync(net->d_tcp_words, 0, tcp_words * sizeof(uint2), st));
            net->tcp_tag = 0;
        }
        BANN_CHECK(ensure_cap(&net->d_gsum, &net->gsum_cap, (size_t)net->pstride));
        TcpArgs t;
        memset(&t, 0, sizeof(t));
        t.store_tc = net->gen->d_store_tc;
        t.descs = net->d_descs;
        t.b = b;
        t.mu = net->gen->d_mu;
        t.sd = net->gen->d_sd;
        t.n = (uint32_t)net->gen->n;
        t.nst = net->gen->nst;
        t.L = Lsteps;
        t.state = net->d_states + b;
        t.theta = net->d_theta;
        t.theta0 = net->d_theta0;
        t.mom = net->d_mom;
        t.grad = net->d_grad;
        t.eps = net->d_eps;
        t.prec = net->d_prec;
        t.model = net->model;
        t.max_h_err = cfg->hmc_max_hamiltonian_error;
        if (R.first_mode == TGT_RESID_PLUS_PRED) { t.resid = resid; t.tgt_out = tgt_out; t.prev_out = prev_out; }
        else t.tgt = R.tgt;
        t.ynew_out = ynew_out;
        t.gsum = net->d_gsum;
        t.pstride = net->pstride;
        t.part = net->d_tcp_words;
        t.sums = net->d_tcp_words + (size_t)4 * kTcpMaxGrid * net->pstride;
        t.tag_base = net->tcp_tag;
        t.launch_par = net->tcp_launches & 1u;
        t.world = (uint32_t)net->ctx->world;
        t.rank = (uint32_t)net->ctx->rank;
        if (sharded(net))
            for (int r = 0; r < net->ctx->world; ++r)
                t.rank_sums[r] = reinterpret_cast<uint2*>(net->xg_region[r] + xg_table_offset(net->xg_cap));
        t.error_flag = net->d_errflag;
        bool launched = false;
        BANN_CHECK(launch_hmc_persistent(net->descs[b], net->act, t, net->ctx->num_sms, st, &launched));
        if (launched) {
            net->tcp_tag += Lsteps + 1u;              // (same sequence on every rank: the tags must agree)
            net->tcp_launches += 1;
            net->persistent_launches += 1;
            net->last_k1 = "k_hmc_persistent<H,S,D,ACT> (whole trajectory, operands resident in tensor / shared memory)";
            k_accept<<<R.nlist, 256, 0, st>>>(net->d_descs, R.list, net->d_states, net->d_theta, net->d_theta0, R.inj_u, R.seed,
                                              R.stream_base);
            BANN_LAUNCHED();
            BANN_CUDA(cudaGetLastError());
            return 0;
        }
        if (net->hmc_path == BANN_HMC_PERSISTENT)
            BANN_FAIL("persistent HMC kernel requested but the branch is not eligible (<= 64 markers, an instantiated architecture, <= 4 super-tiles per SM: ~150k rows)");
    }
    K1Launch k;
    k.list = R.list;
    k.nlist = R.nlist;
    k.single_branch = R.single_branch;
    k.states = net->d_states;
    // sequential schedule and small groups (up to kXrCap values per step: 112 branches of [5,5,1] x 50 markers): the latency-critical
    // push exchange inside KR; launches over more branches: the bulk exchange (three kernels) after KR
    k.xr = sharded(net) && R.list != nullptr && (R.nlist == 1 || (R.xg && (size_t)R.nlist * net->pstride <= kXrCap));
    k.xg = sharded(net) && !k.xr && R.xg;
    k.target_mode = R.first_mode;
    k.tgt = R.tgt;
    k.resid = resid;
    k.tgt_out = tgt_out;
    k.prev_out = prev_out;
    k.out_per_entry = R.out_per_entry;
    if (Lsteps == 0) k.yhat_out = ynew_out;
    BANN_CHECK(launch_k1(net, k, true));
    K2Args a = make_k2(net, cfg, R, 1, Lsteps == 0);
    // --num-grad / --num-grad-traj (branch_sampler.rs:1232-1261; single-branch transitions): numerical_ldg at the parameters
    // the fused pass above just evaluated, then that pass again (shared target) so that K2 finds the unperturbed sums
    const bool numgrad = (cfg->num_grad || (cfg->num_grad_traj && R.traj_num_ldg)) && R.nlist == 1 && R.single_branch >= 0;
    if ((cfg->num_grad || cfg->num_grad_traj) && !numgrad && R.nlist != 1)
        BANN_FAIL("numerical gradients are a per-branch debugging aid (group_size 1)");
    const float* ng_tgt = (R.first_mode == TGT_RESID_PLUS_PRED) ? tgt_out : R.tgt;
    auto numgrad_point = [&](uint32_t step /* 0 = initial evaluation */) -> int {
        const uint32_t b = (uint32_t)R.single_branch;
        const BranchDesc& d = net->descs[b];
        BANN_CHECK(numerical_ldg_async(net, b, ng_tgt));
        if (cfg->num_grad_traj && R.traj_num_ldg && step >= 1)
            BANN_CUDA(cudaMemcpyAsync(R.traj_num_ldg + (size_t)(step - 1) * d.P, net->d_numgrad, d.P * sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
        K1Launch k2 = k;                 // the evaluation at the unperturbed parameters again: sums for K2
        k2.target_mode = TGT_SHARED;
        k2.tgt = ng_tgt;
        k2.resid = nullptr; k2.tgt_out = nullptr; k2.prev_out = nullptr; k2.yhat_out = nullptr;
        k2.states = nullptr;
        BANN_CHECK(launch_k1(net, k2, true));
        a.num_ldg = cfg->num_grad ? net->d_numgrad : nullptr;
        return 0;
    };
    if (numgrad && (cfg->num_grad)) BANN_CHECK(numgrad_point(0));
    BANN_CUDA(launch_pdl(k2_step, dim3(R.nlist), dim3(256), 0, st, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    k.target_mode = R.later_mode;
    k.tgt = (R.first_mode == TGT_RESID_PLUS_PRED) ? tgt_out : R.tgt;
    k.resid = nullptr;
    k.tgt_out = nullptr;
    k.prev_out = nullptr;
    for (uint32_t s = 1; s <= Lsteps; ++s) {
        k.yhat_out = (s == Lsteps) ? ynew_out : nullptr;
        BANN_CHECK(launch_k1(net, k, true));
        a.mode_init = 0;
        a.is_last = (s == Lsteps);
        if (numgrad) BANN_CHECK(numgrad_point(s));
        BANN_CUDA(launch_pdl(k2_step, dim3(R.nlist), dim3(256), 0, st, a));
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
    }
    k_accept<<<R.nlist, 256, 0, st>>>(net->d_descs, R.list, net->d_states, net->d_theta, net->d_theta0, R.inj_u, R.seed,
                                      R.stream_base);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ joint HMC / gradient ascent (joint.cuh)
struct JointWs {
    float *prec0, *pmom, *pgrad, *peps, *inj_mom, *inj_su, *tail;   // tail: [0] accept u, [1] kinetic, [2..4] eval out, [5] log density
};
static JointWs joint_ws(bann_net* net) {
    JointWs w;
    const size_t q = net->maxQ, t = (size_t)net->maxP + net->maxQ;
    w.prec0 = net->d_jws;
    w.pmom = w.prec0 + q;
    w.pgrad = w.pmom + q;
    w.peps = w.pgrad + q;
    w.inj_mom = w.peps + q;
    w.inj_su = w.inj_mom + t;
    w.tail = w.inj_su + t;
    return w;
}
static int joint_supported(bann_net* net) {
    if (net->model == BANN_STD_NORMAL)
        BANN_FAIL("joint sampling / joint gradient ascent: the joint density is unimplemented!() for StdNormal in the reference (std_normal_branch.rs)");
    return 0;
}
struct JointRun {
    uint32_t b = 0;
    int first_mode = TGT_SHARED;     // target mode of the first evaluation (TGT_RESID_PLUS_PRED inside a train visit)
    const float* tgt = nullptr;      // shared target
    const float* resid = nullptr;
    float* tgt_out = nullptr;
    float* prev_out = nullptr;
    float* ynew_out = nullptr;
    const float* inj_mom = nullptr;  // device, P + Q
    const float* inj_su = nullptr;   // device, P + Q
    const float* inj_u = nullptr;    // device, 1
    uint64_t seed = 0, stream = 0;
    float *traj_params = nullptr, *traj_prec = nullptr, *traj_ldg = nullptr, *traj_h = nullptr;
    bool ow_from_gibbs = false;      // d_ow_others already set by k_gibbs (train visit)
};
static JointArgs make_joint(bann_net* net, const bann_mcmc_cfg* cfg, const JointRun& R, int mode, int is_last) {
    JointWs w = joint_ws(net);
    JointArgs a;
    a.descs = net->d_descs;
    a.b = R.b;
    a.st = net->d_states + R.b;
    a.theta = net->d_theta;
    a.theta0 = net->d_theta0;
    a.mom = net->d_mom;
    a.grad = net->d_grad;
    a.eps = net->d_eps;
    a.prec = net->d_prec;
    a.prec0 = w.prec0;
    a.pmom = w.pmom;
    a.pgrad = w.pgrad;
    a.peps = w.peps;
    a.gsum = net->d_gsum;
    a.ow_others = net->d_ow_others + R.b;
    a.G = net->d_G;
    a.hyper = net->hyper;
    a.model = net->model;
    a.n_total = net->n_total;
    a.max_h_err = cfg ? cfg->hmc_max_hamiltonian_error : 0.f;
    a.mode = mode;
    a.is_last = is_last;
    a.gd_step = cfg ? cfg->hmc_step_size_factor : 0.f;
    a.traj_params = R.traj_params;
    a.traj_prec = R.traj_prec;
    a.traj_ldg = R.traj_ldg;
    a.traj_h = R.traj_h;
    a.out = w.tail + 2;
    a.kin_out = w.tail + 1;
    return a;
}
static K1Launch joint_k1(bann_net* net, const JointRun& R, bool first) {
    K1Launch k;
    k.list = net->d_list_all + R.b;
    k.nlist = 1;
    k.single_branch = (int)R.b;
    k.states = net->d_states;
    k.xr = sharded(net);
    if (first) {
        k.target_mode = R.first_mode;
        k.tgt = R.tgt;
        k.resid = R.resid;
        k.tgt_out = R.tgt_out;
        k.prev_out = R.prev_out;
    } else {
        k.target_mode = TGT_SHARED;
        k.tgt = (R.first_mode == TGT_RESID_PLUS_PRED) ? R.tgt_out : R.tgt;
    }
    return k;
}
static int joint_prepare(bann_net* net, const JointRun& R, int hmc, const bann_mcmc_cfg* cfg) {
    cudaStream_t st = net->ctx->stream;
    JointWs w = joint_ws(net);
    if (!R.ow_from_gibbs) {
        k_ow_others<<<1, 256, 0, st>>>(net->d_descs, R.b, net->d_theta, net->d_G, net->model, net->d_ow_others + R.b);
        BANN_LAUNCHED();
    }
    JointInitArgs ia;
    ia.descs = net->d_descs;
    ia.b = R.b;
    ia.st = net->d_states + R.b;
    ia.theta = net->d_theta;
    ia.theta0 = net->d_theta0;
    ia.mom = net->d_mom;
    ia.eps = net->d_eps;
    ia.prec = net->d_prec;
    ia.prec0 = w.prec0;
    ia.pmom = w.pmom;
    ia.peps = w.peps;
    ia.factor = cfg->hmc_step_size_factor;
    ia.hmc = hmc;
    ia.inj_momenta = R.inj_mom;
    ia.inj_step_uniforms = R.inj_su;
    ia.seed = R.seed;
    ia.stream = R.stream;
    k_joint_init<<<1, 256, 0, st>>>(ia);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
// hmc_step_joint (branch_sampler.rs:1070-1178): always Random step sizes with the joint factor (:1094-1101)
static int run_hmc_joint(bann_net* net, const bann_mcmc_cfg* cfg, const JointRun& R) {
    cudaStream_t st = net->ctx->stream;
    BANN_CHECK(joint_supported(net));
    BANN_CHECK(joint_prepare(net, R, 1, cfg));
    const uint32_t Ls = cfg->hmc_integration_length;
    K1Launch k = joint_k1(net, R, true);
    if (Ls == 0) k.yhat_out = R.ynew_out;
    BANN_CHECK(launch_k1(net, k, true));
    k2_joint<<<1, 256, 0, st>>>(make_joint(net, cfg, R, JM_HMC_INIT, Ls == 0));
    BANN_LAUNCHED();
    k = joint_k1(net, R, false);
    for (uint32_t s = 1; s <= Ls; ++s) {
        k.yhat_out = (s == Ls) ? R.ynew_out : nullptr;
        BANN_CHECK(launch_k1(net, k, true));
        k2_joint<<<1, 256, 0, st>>>(make_joint(net, cfg, R, JM_HMC_STEP, s == Ls));
        BANN_LAUNCHED();
    }
    JointWs w = joint_ws(net);
    k_joint_accept<<<1, 256, 0, st>>>(net->d_descs, R.b, net->d_states + R.b, net->d_theta, net->d_theta0, net->d_prec, w.prec0,
                                      w.tail + 1, R.inj_u, R.seed, R.stream);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
// gradient_descent_joint (branch_sampler.rs:1019-1066)
static int run_gd_joint(bann_net* net, const bann_mcmc_cfg* cfg, const JointRun& R) {
    cudaStream_t st = net->ctx->stream;
    BANN_CHECK(joint_supported(net));
    BANN_CHECK(joint_prepare(net, R, 0, cfg));
    const uint32_t Ls = cfg->hmc_integration_length;
    for (uint32_t s = 0; s <= Ls; ++s) {
        K1Launch k = joint_k1(net, R, s == 0);
        k.yhat_out = (s == Ls) ? R.ynew_out : nullptr;
        BANN_CHECK(launch_k1(net, k, true));
        k2_joint<<<1, 256, 0, st>>>(make_joint(net, cfg, R, JM_GD_STEP, s == Ls));
        BANN_LAUNCHED();
    }
    BANN_CUDA(cudaGetLastError());
    return 0;
}
// gradient_descent (branch_sampler.rs:964-1017): the line search compares RSS values on the host, so every probe is one
// K1 pass + a 4-byte read-back (a debugging / point-estimate mode in the reference, not the sampler's hot loop)
static int run_gd(bann_net* net, const bann_mcmc_cfg* cfg, const JointRun& R, float* step_sizes_out, uint32_t* num_probes) {
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[R.b];
    const uint32_t Ls = cfg->hmc_integration_length;
    const unsigned gb = (d.P + 255) / 256;
    JointWs w = joint_ws(net);
    BANN_CHECK(joint_prepare(net, R, 0, cfg));   // theta0 = theta (the base of every probe), state RUNNING
    auto eval_grad = [&](bool first, bool want_yhat) -> int {
        K1Launch k = joint_k1(net, R, first);
        k.yhat_out = want_yhat ? R.ynew_out : nullptr;
        BANN_CHECK(launch_k1(net, k, true));
        k_grad_only<<<1, 256, 0, st>>>(net->d_descs, net->d_list_all + R.b, net->d_theta, net->d_prec, net->d_gsum, net->pstride,
                                       net->model, net->d_grad);
        BANN_LAUNCHED();
        return 0;
    };
    uint32_t probes = 0;
    auto probe = [&](float s, float* rss) -> int {   // probe_gradient_step (:1005-1017)
        k_gd_axpy<<<gb, 256, 0, st>>>(net->d_descs, R.b, net->d_theta, net->d_theta0, net->d_grad, s);
        BANN_LAUNCHED();
        K1Launch k = joint_k1(net, R, false);
        BANN_CHECK(launch_k1(net, k, true));
        BANN_CUDA(cudaMemcpyAsync(rss, net->d_gsum + d.P, sizeof(float), cudaMemcpyDeviceToHost, st));
        BANN_CUDA(cudaStreamSynchronize(st));
        ++probes;
        return 0;
    };
    BANN_CHECK(eval_grad(true, Ls == 0));
    for (uint32_t it = 0; it < Ls; ++it) {
        float step = cfg->hmc_step_size_factor, prev, r2, curr;
        BANN_CHECK(probe(step, &prev));
        BANN_CHECK(probe(2.0f * step, &r2));
        const float fac = (r2 < prev) ? 2.0f : 0.5f;
        step *= fac;
        BANN_CHECK(probe(step, &curr));
        while (curr < prev) {
            prev = curr;
            step *= fac;
            BANN_CHECK(probe(step, &curr));
        }
        step /= fac;
        if (step_sizes_out) step_sizes_out[it] = step;
        k_gd_axpy<<<gb, 256, 0, st>>>(net->d_descs, R.b, net->d_theta, net->d_theta0, net->d_grad, step);   // descend_gradient
        BANN_LAUNCHED();
        k_gd_copy<<<gb, 256, 0, st>>>(net->d_descs, R.b, net->d_theta0, net->d_theta);
        BANN_LAUNCHED();
        BANN_CHECK(eval_grad(false, it + 1 == Ls));
    }
    if (num_probes) *num_probes = probes;
    // the last gradient evaluation ran at the final parameters: its rss / prediction are those of :991-1002
    float rss = 0.f;
    BANN_CUDA(cudaMemcpyAsync(&rss, net->d_gsum + d.P, sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    k_log_density<<<1, 256, 0, st>>>(net->d_descs, R.b, net->d_theta, net->d_prec, net->model, rss, w.tail + 5);
    BANN_LAUNCHED();
    k_gd_finish<<<1, 1, 0, st>>>(net->d_states + R.b, net->d_gsum, d.P, w.tail + 5, (int)Ls);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}
// injected randomness of the joint sampler: P + Q momenta / step uniforms, 1 accept uniform
static int stage_inject_joint(bann_net* net, const bann_rng_inject* inj, uint32_t T, JointRun* R) {
    if (!inj) return 0;
    cudaStream_t st = net->ctx->stream;
    JointWs w = joint_ws(net);
    if (inj->momenta) {
        BANN_CUDA(cudaMemcpyAsync(w.inj_mom, inj->momenta, T * sizeof(float), cudaMemcpyHostToDevice, st));
        R->inj_mom = w.inj_mom;
    }
    if (inj->step_uniforms) {
        BANN_CUDA(cudaMemcpyAsync(w.inj_su, inj->step_uniforms, T * sizeof(float), cudaMemcpyHostToDevice, st));
        R->inj_su = w.inj_su;
    }
    if (inj->accept_uniform) {
        BANN_CUDA(cudaMemcpyAsync(w.tail, inj->accept_uniform, sizeof(float), cudaMemcpyHostToDevice, st));
        R->inj_u = w.tail;
    }
    return 0;
}

static int stage_inject(bann_net* net, const bann_rng_inject* inj, uint32_t P, HmcRun* R, const float** d_gam,
                        uint32_t* n_gam) {
    cudaStream_t st = net->ctx->stream;
    if (d_gam) { *d_gam = nullptr; *n_gam = 0; }
    if (!inj) return 0;
    float* base = net->d_inj;
    if (inj->momenta && R) {
        BANN_CUDA(cudaMemcpyAsync(base, inj->momenta, P * sizeof(float), cudaMemcpyHostToDevice, st));
        R->inj_mom = base;
    }
    if (inj->step_uniforms && R) {
        BANN_CUDA(cudaMemcpyAsync(base + net->maxP, inj->step_uniforms, P * sizeof(float), cudaMemcpyHostToDevice, st));
        R->inj_su = base + net->maxP;
    }
    if (inj->accept_uniform && R) {
        BANN_CUDA(cudaMemcpyAsync(base + 2 * (size_t)net->maxP, inj->accept_uniform, sizeof(float), cudaMemcpyHostToDevice, st));
        R->inj_u = base + 2 * (size_t)net->maxP;
    }
    if (inj->std_gammas && d_gam) {
        if (inj->num_std_gammas > net->inj_gamma_cap) BANN_FAIL("too many injected gamma variates");
        float* gp = base + 2 * (size_t)net->maxP + 4;
        BANN_CUDA(cudaMemcpyAsync(gp, inj->std_gammas, inj->num_std_gammas * sizeof(float), cudaMemcpyHostToDevice, st));
        *d_gam = gp;
        *n_gam = inj->num_std_gammas;
    }
    return 0;
}

// Gibbs draws of one branch (list == NULL) or of every member of a group (one block each, globals frozen)
static int launch_gibbs(bann_net* net, uint32_t b, const bann_mcmc_cfg* cfg, int do_draws, const float* d_gam,
                        uint32_t n_gam, uint64_t seed, uint64_t stream_base, const uint32_t* list = nullptr, uint32_t nlist = 1,
                        uint32_t inj_stride = 0) {
    GibbsArgs g;
    g.descs = net->d_descs;
    g.list = list;
    g.b = b;
    g.theta = net->d_theta;
    g.prec = net->d_prec;
    g.G = net->d_G;
    g.ow_others = net->d_ow_others;
    g.own_old = net->d_own_old;
    g.hyper = net->hyper;
    g.model = net->model;
    g.n_total = net->n_total;
    g.fixed_param_precisions = cfg ? cfg->fixed_param_precisions : 0;
    g.do_draws = do_draws;
    g.inj = d_gam;
    g.n_inj = n_gam;
    g.inj_stride = inj_stride;
    g.seed = seed;
    g.stream_base = stream_base;
    k_gibbs<<<list ? nlist : 1, 256, 0, net->ctx->stream>>>(g);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

static int check_error_flag(bann_net* net) {
    int flag = 0;
    BANN_CUDA(cudaMemcpyAsync(&flag, net->d_errflag, sizeof(int), cudaMemcpyDeviceToHost, net->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    if (flag == 2) BANN_FAIL("peer-memory exchange timed out: a rank did not issue the matching call");
    if (flag == 3) BANN_FAIL("persistent HMC kernel: a CTA waited too long for the partial sums of the others");
    if (flag) BANN_FAIL("Invalid output weight summary statistic (negative or NaN), params.rs:49-54");
    return 0;
}

// one iteration of the inner loop of Net::train, fully asynchronous
struct VisitTraj {   // device buffers of one visit's trajectory (trajectory.rs:4-43), all optional
    float *params = nullptr, *prec = nullptr, *ldg = nullptr, *h = nullptr, *num_ldg = nullptr;
};
static int visit_async(bann_net* net, uint32_t b, const bann_mcmc_cfg* cfg, const bann_rng_inject* inj, uint64_t seed,
                       const VisitTraj* vt = nullptr) {
    cudaStream_t st = net->ctx->stream;
    BANN_CHECK(need_comm(net));
    HmcRun R;
    if (vt) { R.traj_params = vt->params; R.traj_ldg = vt->ldg; R.traj_h = vt->h; R.traj_num_ldg = vt->num_ldg; }
    R.list = net->d_list_all + b;
    R.nlist = 1;
    R.single_branch = (int)b;
    R.first_mode = TGT_RESID_PLUS_PRED;
    R.later_mode = TGT_SHARED;
    R.tgt = nullptr;
    R.seed = seed;
    R.stream_base = net->visit_seq * net->B;
    const float* d_gam = nullptr;
    uint32_t n_gam = 0;
    const bool joint = cfg->joint_hmc || cfg->gradient_descent_joint;
    if (cfg->gradient_descent || joint) {   // net.rs:268-290: no Gibbs draws in the joint modes; dispatch order as the reference
        JointRun J;
        J.b = b;
        J.first_mode = TGT_RESID_PLUS_PRED;
        J.resid = net->d_r;
        J.tgt_out = net->d_t;
        J.prev_out = net->d_prev;
        J.ynew_out = net->d_ynew;
        J.seed = seed;
        J.stream = net->visit_seq * net->B + b;
        J.ow_from_gibbs = true;
        if (vt && cfg->joint_hmc && !cfg->gradient_descent && !cfg->gradient_descent_joint) {
            J.traj_params = vt->params; J.traj_prec = vt->prec; J.traj_ldg = vt->ldg; J.traj_h = vt->h;
        }
        if (cfg->gradient_descent) {
            BANN_CHECK(stage_inject(net, inj, net->descs[b].P, nullptr, &d_gam, &n_gam));
            BANN_CHECK(launch_gibbs(net, b, cfg, joint ? 0 : 1, d_gam, n_gam, seed, net->visit_seq * net->B));
            BANN_CHECK(run_gd(net, cfg, J, nullptr, nullptr));
        } else {
            BANN_CHECK(launch_gibbs(net, b, cfg, 0, nullptr, 0, seed, net->visit_seq * net->B));
            if (cfg->gradient_descent_joint) BANN_CHECK(run_gd_joint(net, cfg, J));
            else {
                BANN_CHECK(stage_inject_joint(net, inj, net->descs[b].P + net->descs[b].nprec, &J));
                BANN_CHECK(run_hmc_joint(net, cfg, J));
            }
        }
    } else {
        BANN_CHECK(stage_inject(net, inj, net->descs[b].P, &R, &d_gam, &n_gam));
        BANN_CHECK(launch_gibbs(net, b, cfg, 1, d_gam, n_gam, seed, net->visit_seq * net->B));   // net.rs:261-277
        BANN_CHECK(run_hmc(net, cfg, R, net->d_ynew, net->d_t, net->d_prev, net->d_r));               // net.rs:279-290
    }
    k_resid_after_hmc<<<net->rblk, 256, 0, st>>>(net->d_r, net->d_t, net->d_ynew, net->d_prev, net->n,
                                                 net->d_states + b, net->d_G, net->d_rpart);      // net.rs:292-300
    BANN_LAUNCHED();
    FinishArgs f;
    f.descs = net->d_descs;
    f.b = b;
    f.theta = net->d_theta;
    f.prec = net->d_prec;
    f.G = net->d_G;
    f.st = net->d_states + b;
    f.ow_others = net->d_ow_others + b;
    f.lpd_local = net->d_lpd_local;
    f.hyper = net->hyper;
    f.model = net->model;
    f.n_total = net->n_total;
    f.part = net->d_rpart;
    f.nblk = net->rblk;
    f.bias_old_new = net->d_bias2;
    f.update_bias = 1;
    f.error_flag = net->d_errflag;
    f.xc = sharded(net) ? xr_next(net->ctx, net->d_errflag) : xr_none();
    k_visit_finish<<<1, 256, 0, st>>>(f);                                                          // net.rs:296,303-305,320-330
    BANN_LAUNCHED();
    k_resid_apply_bias<<<net->rblk, 256, 0, st>>>(net->d_r, net->n, net->d_bias2, net->d_rpart);   // net.rs:321,332
    BANN_LAUNCHED();
    k_resid_reduce<<<1, 32, 0, st>>>(net->d_rpart, net->rblk, net->d_G,
                                     sharded(net) ? xr_next(net->ctx, net->d_errflag) : xr_none());
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    net->visit_seq += 1;
    return 0;
}

static int read_hmc_result(bann_net* net, uint32_t b, bann_hmc_result* out) {
    BranchState s;
    BANN_CUDA(cudaMemcpyAsync(&s, net->d_states + b, sizeof(s), cudaMemcpyDeviceToHost, net->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    out->status = s.status;
    out->log_density = s.log_density;
    out->neg_h_init = s.neg_h_init;
    out->neg_h_final = s.neg_h_cur;
    out->steps_done = (uint32_t)s.steps_done;
    out->u_turn_step = s.u_turn_step;
    return 0;
}

static int refresh_resid_stats(bann_net* net) {
    cudaStream_t st = net->ctx->stream;
    k_resid_stats<<<net->rblk, 256, 0, st>>>(net->d_r, net->n, net->d_rpart);
    BANN_LAUNCHED();
    k_resid_reduce<<<1, 32, 0, st>>>(net->d_rpart, net->rblk, net->d_G,
                                     (sharded(net) && net->ctx->xr_connected) ? xr_next(net->ctx, net->d_errflag) : xr_none());
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ block-Jacobi group visit
// `members`: device list of `num` distinct branch ids.  All of them run Gibbs + HMC against the residual and the globals frozen
// at group start (t_b = r + yhat_b, one row vector per member), then residual / globals / LPD / counters / output bias are
// updated once (chain.cuh: k_resid_group, k_group_stats, k_group_finish).  Asynchronous.
struct GroupInject {          // device pointers (or NULL): see bann_visit_group
    const float *mom = nullptr, *su = nullptr, *u = nullptr, *gam = nullptr;
    uint32_t n_gam = 0, gam_stride = 0;
};
static int visit_group_async(bann_net* net, const uint32_t* members, uint32_t num, const bann_mcmc_cfg* cfg, uint64_t seed,
                             const GroupInject* gi) {
    cudaStream_t st = net->ctx->stream;
    if (cfg->joint_hmc || cfg->gradient_descent || cfg->gradient_descent_joint)
        BANN_FAIL("group visits (group_size > 1) run the HMC sampler; the joint / gradient-descent modes are sequential (group_size 1)");
    BANN_CHECK(need_comm(net));
    if (sharded(net) && !net->xg_connected)
        BANN_FAIL("group visits on sharded rows need the bulk exchange: call bann_net_comm_handle / bann_net_comm_connect first");
    BANN_CHECK(ensure_cap(&net->d_Tg, &net->tg_cap, (size_t)num * net->n));
    BANN_CHECK(ensure_cap(&net->d_Yg, &net->yg_cap, (size_t)num * net->n));
    const uint64_t stream_base = net->visit_seq * net->B;
    BANN_CHECK(launch_gibbs(net, 0, cfg, 1, gi ? gi->gam : nullptr, gi ? gi->n_gam : 0, seed, stream_base, members, num,
                            gi ? gi->gam_stride : 0));                                             // net.rs:261-277 per member
    HmcRun R;
    R.list = members;
    R.nlist = num;
    R.single_branch = -1;
    R.first_mode = TGT_RESID_PLUS_PRED;       // t_b = r + yhat_b (net.rs:279-280), written per member
    R.later_mode = TGT_PER_ENTRY;
    R.tgt = nullptr;
    R.out_per_entry = 1;
    R.seed = seed;
    R.stream_base = stream_base;
    R.xg = true;
    if (gi) { R.inj_mom = gi->mom; R.inj_su = gi->su; R.inj_u = gi->u; R.inj_arena = 1; }
    BANN_CHECK(run_hmc(net, cfg, R, net->d_Yg, net->d_Tg, nullptr, net->d_r));                     // net.rs:282-290 per member
    k_resid_group<<<net->rblk, 256, 0, st>>>(net->d_r, net->d_Tg, net->d_Yg, net->n, members, num, net->d_states, net->d_G,
                                             net->d_rpart);                                        // net.rs:292-300 over the group
    BANN_LAUNCHED();
    GroupArgs g;
    g.descs = net->d_descs;
    g.list = members;
    g.nlist = num;
    g.theta = net->d_theta;
    g.prec = net->d_prec;
    g.G = net->d_G;
    g.states = net->d_states;
    g.own_old = net->d_own_old;
    g.own_new = net->d_own_new;
    g.lpd_local = net->d_lpd_local;
    g.hyper = net->hyper;
    g.model = net->model;
    g.n_total = net->n_total;
    g.part = net->d_rpart;
    g.nblk = net->rblk;
    g.bias_old_new = net->d_bias2;
    g.error_flag = net->d_errflag;
    g.xc = sharded(net) ? xr_next(net->ctx, net->d_errflag) : xr_none();
    k_group_stats<<<num, 256, 0, st>>>(g);
    BANN_LAUNCHED();
    k_group_finish<<<1, 256, 0, st>>>(g);                                                          // net.rs:296,303-305,320-330
    BANN_LAUNCHED();
    k_resid_apply_bias<<<net->rblk, 256, 0, st>>>(net->d_r, net->n, net->d_bias2, net->d_rpart);   // net.rs:321,332
    BANN_LAUNCHED();
    k_resid_reduce<<<1, 32, 0, st>>>(net->d_rpart, net->rblk, net->d_G,
                                     sharded(net) ? xr_next(net->ctx, net->d_errflag) : xr_none());
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    net->visit_seq += 1;
    return 0;
}

static int upload_order(bann_net* net, const uint64_t* order, uint64_t num) {
    std::vector<uint32_t> h(num);
    std::vector<uint8_t> seen(net->B, 0);
    for (uint64_t i = 0; i < num; ++i) {
        if (order[i] >= net->B) BANN_FAIL("branch index out of range");
        h[i] = (uint32_t)order[i];
    }
    if (net->order_cap < num) {
        if (net->d_order) cudaFree(net->d_order);
        net->d_order = nullptr;
        net->order_cap = 0;
        BANN_CUDA(cudaMalloc(&net->d_order, num * sizeof(uint32_t)));
        net->order_cap = num;
    }
    // the previous sweep's kernels may still read the list: the copy is ordered behind them on the same stream
    BANN_CUDA(cudaMemcpyAsync(net->d_order, h.data(), num * sizeof(uint32_t), cudaMemcpyHostToDevice, net->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));   // h goes out of scope
    (void)seen;
    return 0;
}

extern "C" {

int bann_net_create(bann_ctx* ctx, bann_genotypes* gen, int model_type, int activation,
                    const bann_branch_layout* layouts, const float hyper[6], bann_net** out) {
    if (!ctx || !gen || !layouts || !hyper || !out) BANN_FAIL("NULL argument");
    if (model_type < 0 || model_type > 4) BANN_FAIL("unknown model type");
    if (activation < 0 || activation > 4) BANN_FAIL("unknown activation function");
    BANN_CUDA(cudaSetDevice(ctx->device));
    bann_net* net = new bann_net();
    net->ctx = ctx;
    net->gen = gen;
    net->model = model_type;
    net->act = activation;
    for (int i = 0; i < 6; ++i) net->hyper.v[i] = hyper[i];
    net->B = gen->num_branches;
    net->n = (uint32_t)gen->n;
    net->n_total = (float)gen->n_total;
    const bool ard = (model_type == BANN_RIDGE_ARD || model_type == BANN_LASSO_ARD);
    net->descs.resize(net->B);
    uint64_t poff = 0, qoff = 0;
    uint32_t max_gam = 4;
    for (uint64_t b = 0; b < net->B; ++b) {
        const bann_branch_layout& L = layouts[b];
        if (L.num_layers < 2 || L.num_layers > BANN_MAX_LAYERS) { delete net; BANN_FAIL("num_layers must be in [2, 8]"); }
        if (L.widths[L.num_layers - 1] != 1) { delete net; BANN_FAIL("output layer width must be 1"); }
        BranchDesc& d = net->descs[b];
        memset(&d, 0, sizeof(d));
        d.m = gen->m_b[b];
        d.m_pad4 = gen->m_pad4[b];
        d.nl = L.num_layers;
        d.tile_off = gen->tile_off[b];
        d.col_off = gen->col_off[b];
        d.tc_off = gen->tc_off[b];
        d.nc = (d.m + 7) / 8;
        uint32_t prev = d.m, off = 0, aoff = 0, gam = 1;
        for (uint32_t l = 0; l < d.nl; ++l) {
            if (L.widths[l] == 0) { delete net; BANN_FAIL("layer width 0"); }
            d.widths[l] = L.widths[l];
            d.in_dim[l] = prev;
            d.w_off[l] = off;
            off += prev * L.widths[l];
            if (l + 1 < d.nl) { d.a_off[l] = aoff; aoff += L.widths[l]; }
            prev = L.widths[l];
        }
        d.sumw = aoff;
        for (uint32_t l = 0; l + 1 < d.nl; ++l) { d.b_off[l] = off; off += d.widths[l]; }
        d.P = off;
        uint32_t q = 0;
        for (uint32_t l = 0; l < d.nl; ++l) {
            d.wp_off[l] = q;
            d.wp_len[l] = (ard && l + 1 < d.nl) ? d.in_dim[l] : 1;
            q += d.wp_len[l];
            if (l + 1 < d.nl) gam += d.wp_len[l] + 1;
        }
        for (uint32_t l = 0; l + 1 < d.nl; ++l) d.bp_off[l] = q++;
        d.ep_off = q++;
        d.nprec = q;
        d.param_off = poff;
        d.prec_off = qoff;
        if (net->sum_params + d.P > 0xFFFFFFFFull) { delete net; BANN_FAIL("more than 2^32 parameters"); }
        d.dense_off = (uint32_t)net->sum_params;
        net->sum_params += d.P;
        poff += (d.P + 3) & ~3u;
        qoff += d.nprec;
        net->maxP = std::max(net->maxP, d.P);
        net->maxQ = std::max(net->maxQ, d.nprec);
        max_gam = std::max(max_gam, gam + 1);
        size_t sm = k1_generic_smem(d);
        net->max_generic_smem = std::max(net->max_generic_smem, sm);
    }
    if (net->max_generic_smem > 227 * 1024) { delete net; BANN_FAIL("branch too large for the generic kernel's shared memory"); }
    net->total_params = poff;
    net->total_prec = qoff;
    net->pstride = (net->maxP + 1 + 3) & ~3u;
    BANN_CHECK(ensure_cap(&net->d_gsum, &net->gsum_cap, (size_t)net->B * net->pstride));   // all-reduce buffer, fixed address
    // per-CTA partials of the largest launch (entries x chunks <= 4 CTAs per SM + entries): sized once so that no
    // allocation (a device-wide synchronisation) ever happens between the launches of a visit
    BANN_CHECK(ensure_cap(&net->d_part, &net->part_cap, ((size_t)ctx->num_sms * 4 + net->B + 8) * net->pstride));
    cudaStream_t st = ctx->stream;
    size_t pb = poff * sizeof(float);
    BANN_CUDA(cudaMalloc(&net->d_descs, net->B * sizeof(BranchDesc)));
    BANN_CUDA(cudaMemcpyAsync(net->d_descs, net->descs.data(), net->B * sizeof(BranchDesc), cudaMemcpyHostToDevice, st));
    BANN_CUDA(cudaMalloc(&net->d_theta, pb));
    BANN_CUDA(cudaMalloc(&net->d_theta0, pb));
    BANN_CUDA(cudaMalloc(&net->d_mom, pb));
    BANN_CUDA(cudaMalloc(&net->d_grad, pb));
    BANN_CUDA(cudaMalloc(&net->d_eps, pb));
    BANN_CUDA(cudaMemsetAsync(net->d_theta, 0, pb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_theta0, 0, pb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_mom, 0, pb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_grad, 0, pb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_eps, 0, pb, st));
    BANN_CUDA(cudaMalloc(&net->d_prec, qoff * sizeof(float)));
    {
        std::vector<float> ones(qoff, 1.0f);
        BANN_CUDA(cudaMemcpyAsync(net->d_prec, ones.data(), qoff * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaStreamSynchronize(st));
    }
    BANN_CUDA(cudaMalloc(&net->d_states, net->B * sizeof(BranchState)));
    BANN_CUDA(cudaMemsetAsync(net->d_states, 0xff, net->B * sizeof(BranchState), st));   // status = -1 (idle)
    BANN_CUDA(cudaMalloc(&net->d_G, sizeof(NetGlobals)));
    NetGlobals G;
    memset(&G, 0, sizeof(G));
    G.error_precision = 2.0f;            // architectures.rs:229-235
    G.output_layer_precision = 0.05f;    // architectures.rs:16
    G.lpd_rss = -INFINITY;               // log_posterior_density.rs:19-25
    G.lpd_out_w = -INFINITY;
    BANN_CUDA(cudaMemcpyAsync(net->d_G, &G, sizeof(G), cudaMemcpyHostToDevice, st));
    if (net->n == 0) { delete net; BANN_FAIL("this rank holds no individuals (empty row shard): use fewer ranks"); }
    size_t nb = (size_t)net->n * sizeof(float);
    BANN_CUDA(cudaMalloc(&net->d_y, nb));
    BANN_CUDA(cudaMalloc(&net->d_r, nb));
    BANN_CUDA(cudaMalloc(&net->d_t, nb));
    BANN_CUDA(cudaMalloc(&net->d_prev, nb));
    BANN_CUDA(cudaMalloc(&net->d_ynew, nb));
    BANN_CUDA(cudaMemsetAsync(net->d_y, 0, nb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_r, 0, nb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_t, 0, nb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_prev, 0, nb, st));
    BANN_CUDA(cudaMemsetAsync(net->d_ynew, 0, nb, st));
    net->rblk = std::min<uint32_t>((net->n + 255) / 256, (uint32_t)ctx->num_sms * 4);
    BANN_CUDA(cudaMalloc(&net->d_rpart, 2 * (size_t)net->rblk * sizeof(float)));
    BANN_CUDA(cudaMalloc(&net->d_ow_others, 3 * net->B * sizeof(float)));
    BANN_CUDA(cudaMemsetAsync(net->d_ow_others, 0, 3 * net->B * sizeof(float), st));
    net->d_own_old = net->d_ow_others + net->B;
    net->d_own_new = net->d_ow_others + 2 * net->B;
    BANN_CUDA(cudaMalloc(&net->d_bias2, 2 * sizeof(float)));
    BANN_CUDA(cudaMalloc(&net->d_lpd_local, net->B * sizeof(float)));
    {
        std::vector<float> ninf(net->B, -INFINITY);
        BANN_CUDA(cudaMemcpyAsync(net->d_lpd_local, ninf.data(), net->B * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaStreamSynchronize(st));
    }
    BANN_CUDA(cudaMalloc(&net->d_errflag, sizeof(int)));
    BANN_CUDA(cudaMemsetAsync(net->d_errflag, 0, sizeof(int), st));
    {
        std::vector<uint32_t> ids(net->B);
        for (uint64_t b = 0; b < net->B; ++b) ids[b] = (uint32_t)b;
        BANN_CUDA(cudaMalloc(&net->d_list_all, net->B * sizeof(uint32_t)));
        BANN_CUDA(cudaMemcpyAsync(net->d_list_all, ids.data(), net->B * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaStreamSynchronize(st));
    }
    net->inj_gamma_cap = max_gam;
    BANN_CUDA(cudaMalloc(&net->d_inj, (2 * (size_t)net->maxP + 4 + max_gam) * sizeof(float)));
    BANN_CUDA(cudaMalloc(&net->d_scratchB, 3 * net->B * sizeof(float) + 16));
    BANN_CUDA(cudaMalloc(&net->d_jws, (4 * (size_t)net->maxQ + 2 * ((size_t)net->maxP + net->maxQ) + 16) * sizeof(float)));
    BANN_CUDA(cudaFuncSetAttribute(k1_generic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)net->max_generic_smem));
    BANN_CUDA(cudaStreamSynchronize(st));
    *out = net;
    return 0;
}

void bann_net_destroy(bann_net* net) {
    if (net) {
        for (int r = 0; r < kXrMaxWorld; ++r) {
            if (!net->xg_region[r]) continue;
            if (r == net->ctx->rank) cudaFree(net->xg_region[r]);
            else if (net->xg_ipc[r]) cudaIpcCloseMemHandle(net->xg_region[r]);
            net->xg_region[r] = nullptr;
        }
        cudaFree(net->d_xg_counter);
    }
    if (!net) return;
    cudaFree(net->d_descs); cudaFree(net->d_theta); cudaFree(net->d_theta0); cudaFree(net->d_mom);
    cudaFree(net->d_grad); cudaFree(net->d_eps); cudaFree(net->d_prec); cudaFree(net->d_states);
    cudaFree(net->d_G); cudaFree(net->d_y); cudaFree(net->d_r); cudaFree(net->d_t); cudaFree(net->d_prev);
    cudaFree(net->d_ynew); cudaFree(net->d_part); cudaFree(net->d_gsum); cudaFree(net->d_rpart);
    cudaFree(net->d_ow_others); cudaFree(net->d_order); cudaFree(net->d_Tg); cudaFree(net->d_Yg); cudaFree(net->d_inj_grp); cudaFree(net->d_bias2); cudaFree(net->d_lpd_local); cudaFree(net->d_errflag);
    cudaFree(net->d_list_all); cudaFree(net->d_inj); cudaFree(net->d_T); cudaFree(net->d_traj); cudaFree(net->d_numgrad);
    cudaFree(net->d_tcp_words); cudaFree(net->d_pad_descs); cudaFree(net->d_pad_theta); cudaFree(net->d_pad_gsum); cudaFree(net->d_pad_part); cudaFree(net->d_scratchB); cudaFree(net->d_jws); cudaFree(net->d_dense_in); cudaFree(net->d_dense_out);
    for (int i = 0; i < 3; ++i) cudaFree(net->d_tcx[i]);
    if (net->h_pin_a) cudaFreeHost(net->h_pin_a);
    if (net->h_pin_b) cudaFreeHost(net->h_pin_b);
    delete net;
}

int bann_net_branch_sizes(bann_net* net, uint64_t b, uint64_t* num_params, uint64_t* num_precisions) {
    if (!net) BANN_FAIL("NULL net");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    if (num_params) *num_params = net->descs[b].P;
    if (num_precisions) *num_precisions = net->descs[b].nprec;
    return 0;
}

int bann_net_set_branch(bann_net* net, uint64_t b, const float* param_vec, const float* precision_vec) {
    if (!net) BANN_FAIL("NULL net");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    const BranchDesc& d = net->descs[b];
    cudaStream_t st = net->ctx->stream;
    if (param_vec) BANN_CUDA(cudaMemcpyAsync(net->d_theta + d.param_off, param_vec, d.P * sizeof(float), cudaMemcpyHostToDevice, st));
    if (precision_vec) BANN_CUDA(cudaMemcpyAsync(net->d_prec + d.prec_off, precision_vec, d.nprec * sizeof(float), cudaMemcpyHostToDevice, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_net_get_branch(bann_net* net, uint64_t b, float* param_vec, float* precision_vec) {
    if (!net) BANN_FAIL("NULL net");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    const BranchDesc& d = net->descs[b];
    cudaStream_t st = net->ctx->stream;
    if (param_vec) BANN_CUDA(cudaMemcpyAsync(param_vec, net->d_theta + d.param_off, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (precision_vec) BANN_CUDA(cudaMemcpyAsync(precision_vec, net->d_prec + d.prec_off, d.nprec * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int copy_all(bann_net* net, float* dev, float* host_out, const float* host_in) {
    // branches are stored with 4-float aligned offsets; the host view is densely concatenated
    cudaStream_t st = net->ctx->stream;
    std::vector<float> tmp(net->total_params);
    if (host_in) {
        size_t k = 0;
        for (uint64_t b = 0; b < net->B; ++b) {
            const BranchDesc& d = net->descs[b];
            memcpy(&tmp[d.param_off], host_in + k, d.P * sizeof(float));
            k += d.P;
        }
        BANN_CUDA(cudaMemcpyAsync(dev, tmp.data(), net->total_params * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaStreamSynchronize(st));
    } else {
        BANN_CUDA(cudaMemcpyAsync(tmp.data(), dev, net->total_params * sizeof(float), cudaMemcpyDeviceToHost, st));
        BANN_CUDA(cudaStreamSynchronize(st));
        size_t k = 0;
        for (uint64_t b = 0; b < net->B; ++b) {
            const BranchDesc& d = net->descs[b];
            memcpy(host_out + k, &tmp[d.param_off], d.P * sizeof(float));
            k += d.P;
        }
    }
    return 0;
}

int bann_net_set_all_params(bann_net* net, const float* param_vecs, const float* precision_vecs) {
    if (!net || !param_vecs) BANN_FAIL("NULL argument");
    BANN_CHECK(copy_all(net, net->d_theta, nullptr, param_vecs));
    if (precision_vecs) {
        BANN_CUDA(cudaMemcpyAsync(net->d_prec, precision_vecs, net->total_prec * sizeof(float), cudaMemcpyHostToDevice, net->ctx->stream));
        BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    }
    return 0;
}

int bann_net_get_all_params(bann_net* net, float* param_vecs, float* precision_vecs) {
    if (!net || !param_vecs) BANN_FAIL("NULL argument");
    BANN_CHECK(copy_all(net, net->d_theta, param_vecs, nullptr));
    if (precision_vecs) {
        BANN_CUDA(cudaMemcpyAsync(precision_vecs, net->d_prec, net->total_prec * sizeof(float), cudaMemcpyDeviceToHost, net->ctx->stream));
        BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    }
    return 0;
}

int bann_net_set_globals(bann_net* net, const float g[5]) {
    if (!net || !g) BANN_FAIL("NULL argument");
    NetGlobals G;
    cudaStream_t st = net->ctx->stream;
    BANN_CUDA(cudaMemcpyAsync(&G, net->d_G, sizeof(G), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    G.error_precision = g[0];
    G.output_layer_precision = g[1];
    G.ow_reg_sum = g[2];
    G.ow_num_params = g[3];
    G.output_bias = g[4];
    BANN_CUDA(cudaMemcpyAsync(net->d_G, &G, sizeof(G), cudaMemcpyHostToDevice, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_net_get_globals(bann_net* net, float g[5]) {
    if (!net || !g) BANN_FAIL("NULL argument");
    NetGlobals G;
    BANN_CUDA(cudaMemcpyAsync(&G, net->d_G, sizeof(G), cudaMemcpyDeviceToHost, net->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    g[0] = G.error_precision; g[1] = G.output_layer_precision; g[2] = G.ow_reg_sum; g[3] = G.ow_num_params;
    g[4] = G.output_bias;
    return 0;
}

int bann_net_set_targets(bann_net* net, const float* y) {
    if (!net || !y) BANN_FAIL("NULL argument");
    cudaStream_t st = net->ctx->stream;
    BANN_CUDA(cudaMemcpyAsync(net->d_y, y, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, st));
    BANN_CUDA(cudaMemcpyAsync(net->d_r, net->d_y, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    BANN_CHECK(refresh_resid_stats(net));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_net_get_residual(bann_net* net, float* r) {
    if (!net || !r) BANN_FAIL("NULL argument");
    BANN_CUDA(cudaMemcpyAsync(r, net->d_r, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToHost, net->ctx->stream));
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    return 0;
}

int bann_net_set_residual(bann_net* net, const float* r) {
    if (!net || !r) BANN_FAIL("NULL argument");
    BANN_CUDA(cudaMemcpyAsync(net->d_r, r, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, net->ctx->stream));
    BANN_CHECK(refresh_resid_stats(net));
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    return 0;
}

int bann_net_init_residual(bann_net* net) {
    if (!net) BANN_FAIL("NULL net");
    BANN_CHECK(need_comm(net));
    cudaStream_t st = net->ctx->stream;
    k_init_residual<<<(net->n + 255) / 256, 256, 0, st>>>(net->d_r, net->d_y, net->n, net->d_G);
    BANN_LAUNCHED();
    for (uint64_t b = 0; b < net->B; ++b) {
        BANN_CHECK(launch_gibbs(net, (uint32_t)b, nullptr, 0, nullptr, 0, 0, 0));   // update_global_params + from_cfg
        K1Launch k;
        k.list = net->d_list_all + b;
        k.nlist = 1;
        k.single_branch = (int)b;
        k.fwd_only = 1;
        k.yhat_out = net->d_r;
        k.yhat_accumulate = -1;                                                      // net.rs:166
        BANN_CHECK(launch_k1(net, k, false));
        k_resid_stats<<<net->rblk, 256, 0, st>>>(net->d_r, net->n, net->d_rpart);
        BANN_LAUNCHED();
        FinishArgs f;
        f.descs = net->d_descs; f.b = (uint32_t)b; f.theta = net->d_theta; f.prec = net->d_prec; f.G = net->d_G;
        f.st = nullptr; f.ow_others = net->d_ow_others + b; f.lpd_local = net->d_lpd_local; f.hyper = net->hyper;
        f.model = net->model; f.n_total = net->n_total; f.part = net->d_rpart; f.nblk = net->rblk;
        f.bias_old_new = net->d_bias2; f.update_bias = 0; f.error_flag = net->d_errflag;
        f.xc = sharded(net) ? xr_next(net->ctx, net->d_errflag) : xr_none();
        k_visit_finish<<<1, 256, 0, st>>>(f);                                         // net.rs:167
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
    }
    BANN_CHECK(refresh_resid_stats(net));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_branch_fwd_bwd(bann_net* net, uint64_t b, const float* target, float* rss, float* ldg, float* d_rss,
                        float* yhat) {
    if (!net) BANN_FAIL("NULL net");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    BANN_CHECK(need_comm(net));
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    const float* tgt = net->d_y;
    if (target) {
        BANN_CUDA(cudaMemcpyAsync(net->d_t, target, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, st));
        tgt = net->d_t;
    }
    K1Launch k;
    k.list = net->d_list_all + b;
    k.nlist = 1;
    k.single_branch = (int)b;
    k.target_mode = TGT_SHARED;
    k.tgt = tgt;
    k.yhat_out = yhat ? net->d_ynew : nullptr;
    k.xr = sharded(net);
    BANN_CHECK(launch_k1(net, k, true));
    if (ldg) {
        k_grad_only<<<1, 256, 0, st>>>(net->d_descs, net->d_list_all + b, net->d_theta, net->d_prec, net->d_gsum,
                                       net->pstride, net->model, net->d_grad);
        BANN_LAUNCHED();
        BANN_CUDA(cudaGetLastError());
        BANN_CUDA(cudaMemcpyAsync(ldg, net->d_grad + d.param_off, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (d_rss) BANN_CUDA(cudaMemcpyAsync(d_rss, net->d_gsum, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (rss) BANN_CUDA(cudaMemcpyAsync(rss, net->d_gsum + d.P, sizeof(float), cudaMemcpyDeviceToHost, st));
    if (yhat) BANN_CUDA(cudaMemcpyAsync(yhat, net->d_ynew, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_branch_log_density(bann_net* net, uint64_t b, float rss, float* out) {
    if (!net || !out) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    cudaStream_t st = net->ctx->stream;
    k_log_density<<<1, 256, 0, st>>>(net->d_descs, (uint32_t)b, net->d_theta, net->d_prec, net->model, rss, net->d_scratchB);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    BANN_CUDA(cudaMemcpyAsync(out, net->d_scratchB, sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_branch_numerical_ldg(bann_net* net, uint64_t b, const float* target, float* out) {
    if (!net || !out) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    BANN_CHECK(need_comm(net));
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    const float* tgt = net->d_y;
    if (target) {
        BANN_CUDA(cudaMemcpyAsync(net->d_t, target, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, st));
        tgt = net->d_t;
    }
    std::vector<float> keep(d.P);       // the reference reloads the original parameter vector at the end (:502)
    BANN_CUDA(cudaMemcpyAsync(keep.data(), net->d_theta + d.param_off, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CHECK(numerical_ldg_async(net, (uint32_t)b, tgt));
    BANN_CUDA(cudaMemcpyAsync(out, net->d_numgrad, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    BANN_CUDA(cudaMemcpyAsync(net->d_theta + d.param_off, keep.data(), d.P * sizeof(float), cudaMemcpyHostToDevice, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_branch_step_sizes(bann_net* net, uint64_t b, const bann_mcmc_cfg* cfg, const float* step_uniforms, float* out) {
    if (!net || !cfg || !out) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    // k_hmc_init also overwrites theta0 / momenta / state of the branch: harmless outside a transition
    HmcRun R;
    R.list = net->d_list_all + b;
    R.nlist = 1;
    R.single_branch = (int)b;
    R.first_mode = TGT_SHARED;
    R.later_mode = TGT_SHARED;
    R.tgt = net->d_y;
    bann_rng_inject inj;
    memset(&inj, 0, sizeof(inj));
    inj.step_uniforms = step_uniforms;
    BANN_CHECK(stage_inject(net, &inj, d.P, &R, nullptr, nullptr));
    BANN_CHECK(hmc_init_and_first_eval(net, cfg, R, 0));
    BANN_CUDA(cudaMemcpyAsync(out, net->d_eps + d.param_off, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaMemsetAsync(net->d_states + b, 0xff, sizeof(BranchState), st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_hmc_step(bann_net* net, uint64_t b, const float* target, const bann_mcmc_cfg* cfg, const bann_rng_inject* inj,
                  bann_hmc_result* out, bann_trajectory* traj, float* yhat_out) {
    if (!net || !cfg) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    BANN_CHECK(need_comm(net));
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    const float* tgt = net->d_y;
    if (target) {
        BANN_CUDA(cudaMemcpyAsync(net->d_t, target, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, st));
        tgt = net->d_t;
    }
    HmcRun R;
    R.list = net->d_list_all + b;
    R.nlist = 1;
    R.single_branch = (int)b;
    R.first_mode = TGT_SHARED;
    R.later_mode = TGT_SHARED;
    R.tgt = tgt;
    R.seed = 0x243f6a8885a308d3ull;
    R.stream_base = net->visit_seq * net->B;
    BANN_CHECK(stage_inject(net, inj, d.P, &R, nullptr, nullptr));
    const uint32_t Ls = cfg->hmc_integration_length;
    if (traj && (traj->params || traj->ldg || traj->hamiltonian)) {
        size_t need = 2 * (size_t)Ls * d.P + Ls + 1;
        BANN_CHECK(ensure_cap(&net->d_traj, &net->traj_cap, need));
        BANN_CUDA(cudaMemsetAsync(net->d_traj, 0, need * sizeof(float), st));
        R.traj_params = net->d_traj;
        R.traj_ldg = net->d_traj + (size_t)Ls * d.P;
        R.traj_h = net->d_traj + 2 * (size_t)Ls * d.P;
    }
    BANN_CHECK(run_hmc(net, cfg, R, net->d_ynew, nullptr, nullptr, nullptr));
    net->visit_seq += 1;
    if (traj && R.traj_params) {
        if (traj->params) BANN_CUDA(cudaMemcpyAsync(traj->params, R.traj_params, (size_t)Ls * d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (traj->ldg) BANN_CUDA(cudaMemcpyAsync(traj->ldg, R.traj_ldg, (size_t)Ls * d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (traj->hamiltonian) BANN_CUDA(cudaMemcpyAsync(traj->hamiltonian, R.traj_h, (Ls + 1) * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (yhat_out) BANN_CUDA(cudaMemcpyAsync(yhat_out, net->d_ynew, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out) BANN_CHECK(read_hmc_result(net, (uint32_t)b, out));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

static int joint_entry_common(bann_net* net, uint64_t b, const float* target, const bann_mcmc_cfg* cfg, JointRun* R) {
    if (!net || !cfg) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    BANN_CHECK(need_comm(net));
    R->b = (uint32_t)b;
    R->tgt = net->d_y;
    if (target) {
        BANN_CUDA(cudaMemcpyAsync(net->d_t, target, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, net->ctx->stream));
        R->tgt = net->d_t;
    }
    R->ynew_out = net->d_ynew;
    R->stream = net->visit_seq * net->B + b;
    return 0;
}

int bann_branch_joint(bann_net* net, uint64_t b, const float* target, float* rss, float* log_density_joint,
                      float* log_density, float* ldg_joint) {
    JointRun R;
    bann_mcmc_cfg cfg;
    memset(&cfg, 0, sizeof(cfg));
    BANN_CHECK(joint_entry_common(net, b, target, &cfg, &R));
    BANN_CHECK(joint_supported(net));
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    JointWs w = joint_ws(net);
    k_ow_others<<<1, 256, 0, st>>>(net->d_descs, R.b, net->d_theta, net->d_G, net->model, net->d_ow_others + R.b);
    BANN_LAUNCHED();
    K1Launch k = joint_k1(net, R, true);
    k.states = nullptr;
    BANN_CHECK(launch_k1(net, k, true));
    k2_joint<<<1, 256, 0, st>>>(make_joint(net, &cfg, R, JM_EVAL, 1));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    float o[3];
    BANN_CUDA(cudaMemcpyAsync(o, w.tail + 2, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (ldg_joint) {
        BANN_CUDA(cudaMemcpyAsync(ldg_joint, net->d_grad + d.param_off, d.P * sizeof(float), cudaMemcpyDeviceToHost, st));
        BANN_CUDA(cudaMemcpyAsync(ldg_joint + d.P, w.pgrad, d.nprec * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    BANN_CUDA(cudaStreamSynchronize(st));
    if (rss) *rss = o[0];
    if (log_density_joint) *log_density_joint = o[1];
    if (log_density) *log_density = o[2];
    return 0;
}

int bann_hmc_step_joint(bann_net* net, uint64_t b, const float* target, const bann_mcmc_cfg* cfg, const bann_rng_inject* inj,
                        bann_hmc_result* out, bann_trajectory_joint* traj, float* yhat_out) {
    JointRun R;
    BANN_CHECK(joint_entry_common(net, b, target, cfg, &R));
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    const size_t P = d.P, Q = d.nprec, Ls = cfg->hmc_integration_length;
    R.seed = 0x243f6a8885a308d3ull;
    BANN_CHECK(stage_inject_joint(net, inj, (uint32_t)(P + Q), &R));
    const bool want_traj = traj && (traj->params || traj->precisions || traj->ldg || traj->hamiltonian);
    if (want_traj) {
        const size_t need = Ls * P + Ls * Q + Ls * (P + Q) + Ls + 1;
        BANN_CHECK(ensure_cap(&net->d_traj, &net->traj_cap, need));
        BANN_CUDA(cudaMemsetAsync(net->d_traj, 0, need * sizeof(float), st));
        R.traj_params = net->d_traj;
        R.traj_prec = R.traj_params + Ls * P;
        R.traj_ldg = R.traj_prec + Ls * Q;
        R.traj_h = R.traj_ldg + Ls * (P + Q);
    }
    BANN_CHECK(run_hmc_joint(net, cfg, R));
    net->visit_seq += 1;
    if (want_traj) {
        if (traj->params) BANN_CUDA(cudaMemcpyAsync(traj->params, R.traj_params, Ls * P * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (traj->precisions) BANN_CUDA(cudaMemcpyAsync(traj->precisions, R.traj_prec, Ls * Q * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (traj->ldg) BANN_CUDA(cudaMemcpyAsync(traj->ldg, R.traj_ldg, Ls * (P + Q) * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (traj->hamiltonian) BANN_CUDA(cudaMemcpyAsync(traj->hamiltonian, R.traj_h, (Ls + 1) * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    if (yhat_out) BANN_CUDA(cudaMemcpyAsync(yhat_out, net->d_ynew, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out) BANN_CHECK(read_hmc_result(net, (uint32_t)b, out));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_gradient_descent(bann_net* net, uint64_t b, const float* target, const bann_mcmc_cfg* cfg, bann_hmc_result* out,
                          float* step_sizes_out, uint32_t* num_probes, float* yhat_out) {
    JointRun R;
    BANN_CHECK(joint_entry_common(net, b, target, cfg, &R));
    cudaStream_t st = net->ctx->stream;
    BANN_CHECK(run_gd(net, cfg, R, step_sizes_out, num_probes));
    net->visit_seq += 1;
    if (yhat_out) BANN_CUDA(cudaMemcpyAsync(yhat_out, net->d_ynew, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out) BANN_CHECK(read_hmc_result(net, (uint32_t)b, out));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_gradient_descent_joint(bann_net* net, uint64_t b, const float* target, const bann_mcmc_cfg* cfg, bann_hmc_result* out,
                                float* yhat_out) {
    JointRun R;
    BANN_CHECK(joint_entry_common(net, b, target, cfg, &R));
    cudaStream_t st = net->ctx->stream;
    BANN_CHECK(run_gd_joint(net, cfg, R));
    net->visit_seq += 1;
    if (yhat_out) BANN_CUDA(cudaMemcpyAsync(yhat_out, net->d_ynew, (size_t)net->n * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out) BANN_CHECK(read_hmc_result(net, (uint32_t)b, out));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_gibbs_branch(bann_net* net, uint64_t b, const bann_mcmc_cfg* cfg, const bann_rng_inject* inj) {
    if (!net) BANN_FAIL("NULL net");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    const float* d_gam = nullptr;
    uint32_t n_gam = 0;
    BANN_CHECK(stage_inject(net, inj, net->descs[b].P, nullptr, &d_gam, &n_gam));
    BANN_CHECK(launch_gibbs(net, (uint32_t)b, cfg, 1, d_gam, n_gam, 0x13198a2e03707344ull, net->visit_seq * net->B));
    net->visit_seq += 1;
    BANN_CUDA(cudaStreamSynchronize(net->ctx->stream));
    return 0;
}

int bann_visit_branch(bann_net* net, uint64_t b, const bann_mcmc_cfg* cfg, const bann_rng_inject* inj,
                      bann_hmc_result* out) {
    if (!net || !cfg) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    BANN_CHECK(visit_async(net, (uint32_t)b, cfg, inj, 0x452821e638d01377ull));
    if (out) BANN_CHECK(read_hmc_result(net, (uint32_t)b, out));
    BANN_CHECK(check_error_flag(net));
    return 0;
}

int bann_visit_branch_traj(bann_net* net, uint64_t b, const bann_mcmc_cfg* cfg, uint64_t seed, bann_hmc_result* out,
                           bann_trajectory_joint* traj) {
    if (!net || !cfg || !traj) BANN_FAIL("NULL argument");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    cudaStream_t st = net->ctx->stream;
    const BranchDesc& d = net->descs[b];
    const bool jt = cfg->joint_hmc && !cfg->gradient_descent && !cfg->gradient_descent_joint;
    const size_t P = d.P, Q = jt ? d.nprec : 0, Ls = cfg->hmc_integration_length;
    const bool ng = cfg->num_grad_traj && traj->num_ldg && !jt && !cfg->gradient_descent && !cfg->gradient_descent_joint;
    const size_t need = Ls * P + Ls * Q + Ls * (P + Q) + Ls + 1 + (ng ? Ls * P : 0);
    BANN_CHECK(ensure_cap(&net->d_traj, &net->traj_cap, need));
    BANN_CUDA(cudaMemsetAsync(net->d_traj, 0, need * sizeof(float), st));
    VisitTraj vt;
    vt.params = net->d_traj;
    vt.prec = vt.params + Ls * P;
    vt.ldg = vt.prec + Ls * Q;
    vt.h = vt.ldg + Ls * (P + Q);
    if (ng) vt.num_ldg = vt.h + Ls + 1;
    BANN_CHECK(visit_async(net, (uint32_t)b, cfg, nullptr, seed, &vt));
    if (traj->params) BANN_CUDA(cudaMemcpyAsync(traj->params, vt.params, Ls * P * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (traj->precisions && Q) BANN_CUDA(cudaMemcpyAsync(traj->precisions, vt.prec, Ls * Q * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (traj->ldg) BANN_CUDA(cudaMemcpyAsync(traj->ldg, vt.ldg, Ls * (P + Q) * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (traj->hamiltonian) BANN_CUDA(cudaMemcpyAsync(traj->hamiltonian, vt.h, (Ls + 1) * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (ng) BANN_CUDA(cudaMemcpyAsync(traj->num_ldg, vt.num_ldg, Ls * P * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out) BANN_CHECK(read_hmc_result(net, (uint32_t)b, out));
    BANN_CHECK(check_error_flag(net));
    return 0;
}

int bann_net_stats(bann_net* net, bann_sweep_stats* out) {
    if (!net || !out) BANN_FAIL("NULL argument");
    cudaStream_t st = net->ctx->stream;
    NetGlobals G;
    std::vector<float> loc(net->B);
    BANN_CUDA(cudaMemcpyAsync(&G, net->d_G, sizeof(G), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaMemcpyAsync(loc.data(), net->d_lpd_local, net->B * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    float acc = 0.f;   // log_posterior_density.rs:62-67: sequential f32 sum of the local terms
    for (uint64_t b = 0; b < net->B; ++b) acc += loc[b];
    out->num_samples = G.num_samples;
    out->num_accepted = G.num_accepted;
    out->num_early_rejected = G.num_early_rejected;
    out->mse_train = G.resid_ss / net->n_total;   // net.rs:604-606
    out->lpd = (G.lpd_rss + G.lpd_out_w) + acc;
    out->output_bias = G.output_bias;
    out->error_precision = G.error_precision;
    out->output_layer_precision = G.output_layer_precision;
    return 0;
}

int bann_net_lpd_terms(bann_net* net, float* wrt_rss_and_error_precision, float* wrt_output_weights_and_precision,
                       float* wrt_local_params) {
    if (!net || !wrt_rss_and_error_precision || !wrt_output_weights_and_precision || !wrt_local_params) BANN_FAIL("NULL argument");
    cudaStream_t st = net->ctx->stream;
    NetGlobals G;
    BANN_CUDA(cudaMemcpyAsync(&G, net->d_G, sizeof(G), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaMemcpyAsync(wrt_local_params, net->d_lpd_local, net->B * sizeof(float), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    *wrt_rss_and_error_precision = G.lpd_rss;
    *wrt_output_weights_and_precision = G.lpd_out_w;
    return 0;
}

int bann_sweep(bann_net* net, const bann_mcmc_cfg* cfg, const uint64_t* branch_order, uint64_t num, uint32_t group_size,
               uint64_t seed, bann_sweep_stats* out) {
    if (!net || !cfg || !branch_order) BANN_FAIL("NULL argument");
    if (group_size == 0) BANN_FAIL("group_size must be >= 1");
    if (group_size == 1) {          // sequential-exact schedule: the reference's Gauss-Seidel order
        for (uint64_t i = 0; i < num; ++i) {
            if (branch_order[i] >= net->B) BANN_FAIL("branch index out of range");
            BANN_CHECK(visit_async(net, (uint32_t)branch_order[i], cfg, nullptr, seed));
        }
    } else {                        // block-Jacobi: consecutive groups of the order advance concurrently
        {
            std::vector<uint8_t> seen(net->B, 0);
            for (uint64_t i = 0; i < num; ++i) {
                if (branch_order[i] >= net->B) BANN_FAIL("branch index out of range");
                // a branch twice in one group would race on its own state; twice in a sweep is fine across groups
                if (i % group_size == 0) std::fill(seen.begin(), seen.end(), 0);
                if (seen[branch_order[i]]) BANN_FAIL("a branch appears twice in one group");
                seen[branch_order[i]] = 1;
            }
        }
        BANN_CHECK(upload_order(net, branch_order, num));
        for (uint64_t i = 0; i < num; i += group_size) {
            const uint32_t cnt = (uint32_t)std::min<uint64_t>(group_size, num - i);
            if (cnt == 1) BANN_CHECK(visit_async(net, (uint32_t)branch_order[i], cfg, nullptr, seed));
            else BANN_CHECK(visit_group_async(net, net->d_order + i, cnt, cfg, seed, nullptr));
        }
    }
    BANN_CHECK(check_error_flag(net));
    if (out) BANN_CHECK(bann_net_stats(net, out));
    return 0;
}

int bann_visit_group(bann_net* net, const uint64_t* members, uint64_t num, const bann_mcmc_cfg* cfg, const bann_rng_inject* inj,
                     uint64_t seed, bann_hmc_result* out) {
    if (!net || !members || !cfg || num == 0) BANN_FAIL("NULL / empty argument");
    cudaStream_t st = net->ctx->stream;
    {
        std::vector<uint8_t> seen(net->B, 0);
        for (uint64_t i = 0; i < num; ++i) {
            if (members[i] >= net->B) BANN_FAIL("branch index out of range");
            if (seen[members[i]]) BANN_FAIL("a branch appears twice in one group");
            seen[members[i]] = 1;
        }
    }
    BANN_CHECK(upload_order(net, members, num));
    GroupInject gi;
    if (inj) {
        // staging: [momenta arena | step-uniform arena | u per entry | gammas per entry x stride]
        uint32_t stride = 0;
        for (uint64_t i = 0; i < num; ++i) stride = std::max(stride, inj[i].num_std_gammas);
        stride = std::max(stride, 1u);
        const size_t need = 2 * (size_t)net->total_params + num + (size_t)num * stride;
        if (net->inj_grp_stride < need) {      // (re)allocate: inj_grp_stride doubles as the capacity in floats
            if (net->d_inj_grp) cudaFree(net->d_inj_grp);
            net->d_inj_grp = nullptr;
            net->inj_grp_stride = 0;
            BANN_CUDA(cudaMalloc(&net->d_inj_grp, need * sizeof(float)));
            net->inj_grp_stride = (uint32_t)need;
        }
        std::vector<float> h(need, 0.f);
        float* mom = h.data();
        float* su = mom + net->total_params;
        float* u = su + net->total_params;
        float* gam = u + num;
        bool any_mom = false, any_su = false, any_u = false, any_gam = false;
        for (uint64_t i = 0; i < num; ++i) {
            const BranchDesc& d = net->descs[members[i]];
            if (inj[i].momenta) { memcpy(mom + d.param_off, inj[i].momenta, d.P * sizeof(float)); any_mom = true; }
            if (inj[i].step_uniforms) { memcpy(su + d.param_off, inj[i].step_uniforms, d.P * sizeof(float)); any_su = true; }
            if (inj[i].accept_uniform) { u[i] = *inj[i].accept_uniform; any_u = true; }
            if (inj[i].std_gammas) { memcpy(gam + i * stride, inj[i].std_gammas, inj[i].num_std_gammas * sizeof(float)); any_gam = true; }
        }
        BANN_CUDA(cudaMemcpyAsync(net->d_inj_grp, h.data(), need * sizeof(float), cudaMemcpyHostToDevice, st));
        BANN_CUDA(cudaStreamSynchronize(st));
        float* base = net->d_inj_grp;
        if (any_mom) gi.mom = base;
        if (any_su) gi.su = base + net->total_params;
        if (any_u) gi.u = base + 2 * (size_t)net->total_params;
        if (any_gam) { gi.gam = base + 2 * (size_t)net->total_params + num; gi.n_gam = stride; gi.gam_stride = stride; }
    }
    if (num == 1 && !inj) BANN_CHECK(visit_async(net, (uint32_t)members[0], cfg, nullptr, seed));
    else BANN_CHECK(visit_group_async(net, net->d_order, (uint32_t)num, cfg, seed, inj ? &gi : nullptr));
    BANN_CHECK(check_error_flag(net));
    if (out)
        for (uint64_t i = 0; i < num; ++i) BANN_CHECK(read_hmc_result(net, (uint32_t)members[i], &out[i]));
    return 0;
}

int bann_predict(bann_net* net, bann_genotypes* test, float* yhat) {
    if (!net || !yhat) BANN_FAIL("NULL argument");
    cudaStream_t st = net->ctx->stream;
    bann_genotypes* g = test ? test : net->gen;
    if (g->num_branches != net->B) BANN_FAIL("test genotypes have a different number of branches");
    std::vector<BranchDesc> descs = net->descs;
    for (uint64_t b = 0; b < net->B; ++b) {
        if (g->m_b[b] != net->descs[b].m) BANN_FAIL("test genotypes: branch marker count mismatch");
        descs[b].tile_off = g->tile_off[b];
        descs[b].col_off = g->col_off[b];
        descs[b].tc_off = g->tc_off[b];
    }
    BranchDesc* d_descs = nullptr;
    float* d_out = nullptr;
    BANN_CUDA(cudaMalloc(&d_descs, net->B * sizeof(BranchDesc)));
    BANN_CUDA(cudaMalloc(&d_out, g->n * sizeof(float)));
    int rc = 0;
    do {
        if (cudaMemcpyAsync(d_descs, descs.data(), net->B * sizeof(BranchDesc), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = -2; break; }
        NetGlobals G;
        if (cudaMemcpyAsync(&G, net->d_G, sizeof(G), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = -2; break; }
        cudaStreamSynchronize(st);
        std::vector<float> init(g->n, G.output_bias);   // net.rs:547-552
        if (cudaMemcpyAsync(d_out, init.data(), g->n * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = -2; break; }
        cudaStreamSynchronize(st);
        for (uint64_t b = 0; b < net->B && rc == 0; ++b) {   // net.rs:554-557, branch order preserved
            K1Launch k;
            k.list = net->d_list_all + b;
            k.nlist = 1;
            k.single_branch = (int)b;
            k.fwd_only = 1;
            k.yhat_out = d_out;
            k.yhat_accumulate = 1;
            k.store = g;
            k.descs_dev = d_descs;
            rc = launch_k1(net, k, false);
        }
        if (rc) break;
        if (cudaMemcpyAsync(yhat, d_out, g->n * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = -2; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -2; break; }
    } while (0);
    cudaFree(d_descs);
    cudaFree(d_out);
    if (rc == -2) BANN_FAIL("CUDA error in bann_predict");
    return rc;
}

// ------------------------------------------------------------------ per-row diagnostics (probe.cuh)
static int run_probe(bann_net* net, uint64_t b, bann_genotypes* other, float* acts_host, float* es_host, float* pop_host) {
    if (!net) BANN_FAIL("NULL net");
    if (b >= net->B) BANN_FAIL("branch index out of range");
    cudaStream_t st = net->ctx->stream;
    bann_genotypes* g = other ? other : net->gen;
    if (g->num_branches != net->B || g->m_b[b] != net->descs[b].m) BANN_FAIL("genotypes do not match the net's grouping");
    if (!g->d_store) BANN_FAIL("the byte-tile store was released (bann_genotypes_release_byte_store): the diagnostic kernels read it");
    if (pop_host && sharded(net)) BANN_FAIL("population effect sizes on sharded rows: sum the per-rank effect sizes on the host");
    BranchDesc d = net->descs[b];
    d.tile_off = g->tile_off[b];
    d.col_off = g->col_off[b];
    const size_t n = g->n, na = n * (d.sumw + 1), ne = n * d.m;
    const uint32_t ntiles = g->ntiles, w0 = d.widths[0];
    const size_t smem = k1_generic_smem(d);
    if (smem > 227 * 1024) BANN_FAIL("branch too large for the diagnostic kernel's shared memory");
    BranchDesc* d_desc = nullptr;
    float *d_acts = nullptr, *d_es = nullptr, *d_dsum = nullptr, *d_pop = nullptr;
    int rc = 0;
    do {
        if (cudaMalloc(&d_desc, sizeof(BranchDesc)) != cudaSuccess) { rc = -2; break; }
        if (acts_host && cudaMalloc(&d_acts, na * sizeof(float)) != cudaSuccess) { rc = -2; break; }
        if (es_host && cudaMalloc(&d_es, ne * sizeof(float)) != cudaSuccess) { rc = -2; break; }
        if (pop_host && (cudaMalloc(&d_dsum, (size_t)ntiles * w0 * sizeof(float)) != cudaSuccess ||
                         cudaMalloc(&d_pop, d.m * sizeof(float)) != cudaSuccess)) { rc = -2; break; }
        if (cudaMemcpyAsync(d_desc, &d, sizeof(d), cudaMemcpyHostToDevice, st) != cudaSuccess) { rc = -2; break; }
        if (cudaFuncSetAttribute(k_branch_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { rc = -2; break; }
        ProbeArgs a;
        a.store = g->d_store;
        a.descs = d_desc;
        a.b = 0;
        a.theta = net->d_theta;
        a.mu = g->d_mu;
        a.sd = g->d_sd;
        a.n = (uint32_t)n;
        a.ntiles = ntiles;
        a.act = net->act;
        a.acts_out = d_acts;
        a.es_out = d_es;
        a.dsum_part = d_dsum;
        k_branch_probe<<<std::min<uint32_t>(ntiles, (uint32_t)net->ctx->num_sms * 2), 128, smem, st>>>(a);
        BANN_LAUNCHED();
        if (pop_host) {
            k_population_effects<<<(d.m + 127) / 128, 128, w0 * sizeof(float), st>>>(d_desc, 0, net->d_theta, d_dsum, ntiles,
                                                                                       (float)g->n_total, d_pop);
            BANN_LAUNCHED();
        }
        if (cudaGetLastError() != cudaSuccess) { rc = -2; break; }
        if (acts_host && cudaMemcpyAsync(acts_host, d_acts, na * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = -2; break; }
        if (es_host && cudaMemcpyAsync(es_host, d_es, ne * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = -2; break; }
        if (pop_host && cudaMemcpyAsync(pop_host, d_pop, d.m * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess) { rc = -2; break; }
        if (cudaStreamSynchronize(st) != cudaSuccess) { rc = -2; break; }
    } while (0);
    cudaFree(d_desc); cudaFree(d_acts); cudaFree(d_es); cudaFree(d_dsum); cudaFree(d_pop);
    if (rc == -2) BANN_FAIL(std::string("CUDA error in the diagnostic kernels: ") + cudaGetErrorString(cudaGetLastError()));
    return rc;
}

int bann_branch_activations(bann_net* net, uint64_t b, bann_genotypes* genotypes_or_null, float* out) {
    if (!out) BANN_FAIL("NULL argument");
    return run_probe(net, b, genotypes_or_null, out, nullptr, nullptr);
}

int bann_branch_effect_sizes(bann_net* net, uint64_t b, bann_genotypes* genotypes_or_null, float* effect_sizes,
                             float* population_effect_sizes) {
    if (!effect_sizes && !population_effect_sizes) BANN_FAIL("NULL argument");
    return run_probe(net, b, genotypes_or_null, nullptr, effect_sizes, population_effect_sizes);
}

// ------------------------------------------------------------------ full-network (grouped) operations
// Net::gradient split in two so that a multi-GPU caller can all-reduce the raw sums in between:
//   begin: H2D params / y, fused fwd+bwd over every branch, chunk reduction into the all-reduce buffer
//   end  : gradient under the prior, D2H of gradients and rss
// dense (host-facing: param vecs back to back) <-> arena (16-byte aligned per branch) on the device, so that the host side
// of Net::gradient is ONE copy per direction straight from / into the caller's buffer
__global__ void __launch_bounds__(128) k_dense_to_arena(const BranchDesc* descs, const float* __restrict__ dense, float* __restrict__ arena) {
    const BranchDesc& d = descs[blockIdx.x];
    for (uint32_t k = threadIdx.x; k < d.P; k += 128) arena[d.param_off + k] = dense[d.dense_off + k];
}
__global__ void __launch_bounds__(128) k_arena_to_dense(const BranchDesc* descs, const float* __restrict__ arena, float* __restrict__ dense) {
    const BranchDesc& d = descs[blockIdx.x];
    for (uint32_t k = threadIdx.x; k < d.P; k += 128) dense[d.dense_off + k] = arena[d.param_off + k];
}
static bool host_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

int bann_net_gradient_begin(bann_net* net, const float* param_vecs, const float* y) {
    if (!net) BANN_FAIL("NULL net");
    cudaStream_t st = net->ctx->stream;
    // Pinned (page-locked / registered) caller buffers are copied by DMA directly; pageable ones go through one pinned
    // staging buffer with a single memcpy, so that the copies stay asynchronous either way.
    if (!net->d_dense_in) {
        BANN_CUDA(cudaMalloc(&net->d_dense_in, (std::max<uint64_t>(net->sum_params, 1) + 4) * sizeof(float)));
        BANN_CUDA(cudaMalloc(&net->d_dense_out, (net->sum_params + net->B + 4) * sizeof(float)));
    }
    if (!net->h_pin_a && ((param_vecs && !host_pinned(param_vecs)) || (y && !host_pinned(y))))
        BANN_CUDA(cudaMallocHost(&net->h_pin_a, (net->sum_params + net->n) * sizeof(float)));
    if (param_vecs) {
        // Sharded rows with the bulk exchange connected: the parameters are replicated, so every rank uploads only its
        // 1 / world slice of the vector and the ranks all-gather on the device over NVLink (host traffic / world).
        uint64_t lo = 0, hi = net->sum_params;
        const bool sliced = sharded(net) && net->xg_connected;
        if (sliced) xg_slice(net, net->sum_params, &lo, &hi);
        const float* src = param_vecs + lo;
        if (!host_pinned(param_vecs)) {
            memcpy(net->h_pin_a + lo, param_vecs + lo, (hi - lo) * sizeof(float));
            src = net->h_pin_a + lo;
        }
        if (hi > lo) BANN_CUDA(cudaMemcpyAsync(net->d_dense_in + lo, src, (hi - lo) * sizeof(float), cudaMemcpyHostToDevice, st));
        if (sliced) BANN_CHECK(xg_allgather_slices(net, net->d_dense_in, net->sum_params));
        k_dense_to_arena<<<(unsigned)net->B, 128, 0, st>>>(net->d_descs, net->d_dense_in, net->d_theta);
        BANN_LAUNCHED();
    }
    const float* tgt = net->d_y;
    if (y) {
        const float* src = y;
        if (!host_pinned(y)) {
            memcpy(net->h_pin_a + net->sum_params, y, (size_t)net->n * sizeof(float));
            src = net->h_pin_a + net->sum_params;
        }
        BANN_CUDA(cudaMemcpyAsync(net->d_t, src, (size_t)net->n * sizeof(float), cudaMemcpyHostToDevice, st));
        tgt = net->d_t;
    }
    K1Launch k;
    k.list = nullptr;
    k.nlist = (uint32_t)net->B;
    k.target_mode = TGT_SHARED;
    k.tgt = tgt;
    k.xg = sharded(net) && net->xg_connected;
    return launch_k1(net, k, true);
}

int bann_net_gradient_end(bann_net* net, float* grads, float* rss) {
    if (!net) BANN_FAIL("NULL net");
    cudaStream_t st = net->ctx->stream;
    if (!net->d_dense_out) BANN_FAIL("bann_net_gradient_end without bann_net_gradient_begin");
    k_grad_only<<<(unsigned)net->B, 256, 0, st>>>(net->d_descs, nullptr, net->d_theta, net->d_prec, net->d_gsum,
                                                  net->pstride, net->model, net->d_grad);
    BANN_LAUNCHED();
    const bool pin_g = !grads || host_pinned(grads), pin_r = !rss || host_pinned(rss);
    if (!net->h_pin_b && !(pin_g && pin_r)) BANN_CUDA(cudaMallocHost(&net->h_pin_b, (net->sum_params + net->B) * sizeof(float)));
    // sharded rows + bulk exchange: every rank holds the full result on the device and writes only its 1 / world slice of
    // [grads | rss] to the host (bann_net_gradient_slice tells the caller which)
    uint64_t lo = 0, hi = net->sum_params + net->B;
    if (sharded(net) && net->xg_connected) xg_slice(net, net->sum_params + net->B, &lo, &hi);
    const uint64_t glo = std::min(lo, net->sum_params), ghi = std::min(hi, net->sum_params);
    const uint64_t rlo = std::max(lo, net->sum_params) - net->sum_params, rhi = std::max(hi, net->sum_params) - net->sum_params;
    if (grads) {
        k_arena_to_dense<<<(unsigned)net->B, 128, 0, st>>>(net->d_descs, net->d_grad, net->d_dense_out);
        BANN_LAUNCHED();
        if (ghi > glo)
            BANN_CUDA(cudaMemcpyAsync((pin_g ? grads : net->h_pin_b) + glo, net->d_dense_out + glo, (ghi - glo) * sizeof(float),
                                      cudaMemcpyDeviceToHost, st));
    }
    if (rss) {
        k_gather_rss<<<((unsigned)net->B + 255) / 256, 256, 0, st>>>(net->d_gsum, net->pstride, net->d_descs, (uint32_t)net->B,
                                                                     net->d_dense_out + net->sum_params);
        BANN_LAUNCHED();
        if (rhi > rlo)
            BANN_CUDA(cudaMemcpyAsync((pin_r ? rss : net->h_pin_b + net->sum_params) + rlo, net->d_dense_out + net->sum_params + rlo,
                                      (rhi - rlo) * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    BANN_CUDA(cudaGetLastError());
    BANN_CUDA(cudaStreamSynchronize(st));
    if (grads && !pin_g && ghi > glo) memcpy(grads + glo, net->h_pin_b + glo, (ghi - glo) * sizeof(float));
    if (rss && !pin_r && rhi > rlo) memcpy(rss + rlo, net->h_pin_b + net->sum_params + rlo, (rhi - rlo) * sizeof(float));
    return 0;
}

int bann_net_gradient_slice(bann_net* net, uint64_t* param_lo, uint64_t* param_hi, uint64_t* out_lo, uint64_t* out_hi) {
    if (!net) BANN_FAIL("NULL net");
    uint64_t a = 0, b = net->sum_params, c = 0, d = net->sum_params + net->B;
    if (sharded(net) && net->xg_connected) {
        xg_slice(net, net->sum_params, &a, &b);
        xg_slice(net, net->sum_params + net->B, &c, &d);
    }
    if (param_lo) *param_lo = a;
    if (param_hi) *param_hi = b;
    if (out_lo) *out_lo = c;
    if (out_hi) *out_hi = d;
    return 0;
}

// ---- bulk exchange region of a net (comm.cuh: XgComm); the host layer all-gathers the handles once after bann_net_create
int bann_net_comm_handle(bann_net* net, uint8_t* out) {
    if (!net || !out) BANN_FAIL("NULL argument");
    if (net->ctx->world > kXrMaxWorld) BANN_FAIL("peer-memory exchange supports at most 8 ranks");
    BANN_CUDA(cudaSetDevice(net->ctx->device));
    uint8_t*& mine = net->xg_region[net->ctx->rank];
    if (!mine) {
        net->xg_cap = ((std::max<uint64_t>((uint64_t)net->B * net->pstride, net->sum_params + net->B) + 3) & ~3ull) + 4;
        BANN_CUDA(cudaMalloc(&mine, xg_region_bytes(net->xg_cap, net->pstride)));
        BANN_CUDA(cudaMemset(mine, 0, xg_region_bytes(net->xg_cap, net->pstride)));   // flags / tags 0 = "nothing yet"
        BANN_CUDA(cudaMalloc(&net->d_xg_counter, sizeof(unsigned int)));
        BANN_CUDA(cudaMemset(net->d_xg_counter, 0, sizeof(unsigned int)));
        BANN_CHECK(ensure_cap(&net->d_gsum, &net->gsum_cap, (size_t)net->B * net->pstride + 4));
        BANN_CUDA(cudaDeviceSynchronize());
    }
    return comm_export(mine, net->ctx->device, out);
}

int bann_net_comm_connect(bann_net* net, const uint8_t* handles) {
    if (!net || !handles) BANN_FAIL("NULL argument");
    if (!net->xg_region[net->ctx->rank]) BANN_FAIL("bann_net_comm_connect before bann_net_comm_handle");
    BANN_CUDA(cudaSetDevice(net->ctx->device));
    for (int r = 0; r < net->ctx->world; ++r) {
        if (r == net->ctx->rank) continue;
        void* p = nullptr;
        bool ipc = false;
        BANN_CHECK(comm_import(net->ctx, handles + (size_t)r * BANN_COMM_HANDLE_BYTES, &p, &ipc));
        net->xg_region[r] = reinterpret_cast<uint8_t*>(p);
        net->xg_ipc[r] = ipc;
    }
    net->xg_connected = true;
    net->xg_epoch = 0;
    return 0;
}

int bann_net_comm_connected(bann_net* net) { return net && net->xg_connected ? 1 : 0; }

int bann_pinned_alloc(uint64_t bytes, void** out) {
    if (!out) BANN_FAIL("NULL argument");
    BANN_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return 0;
}
void bann_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

int bann_net_gradient(bann_net* net, const float* param_vecs, const float* y, float* grads, float* rss) {
    if (!net) BANN_FAIL("NULL net");
    if (net->ctx->world > 1 && !net->xg_connected)
        BANN_FAIL("sharded rows: connect the bulk exchange (bann_net_comm_connect), or call bann_net_gradient_begin, all-reduce, bann_net_gradient_end");
    BANN_CHECK(bann_net_gradient_begin(net, param_vecs, y));
    return bann_net_gradient_end(net, grads, rss);
}

static HmcRun grouped_run(bann_net* net) {
    HmcRun R;
    R.list = nullptr;
    R.nlist = (uint32_t)net->B;
    R.single_branch = -1;
    R.first_mode = net->grouped_per_branch ? TGT_RESID_PLUS_PRED : TGT_SHARED;
    R.later_mode = net->grouped_per_branch ? TGT_PER_ENTRY : TGT_SHARED;
    R.tgt = net->grouped_per_branch ? net->d_T : net->d_y;
    R.seed = net->grouped_seed;
    R.stream_base = net->visit_seq * net->B;
    return R;
}

static int grouped_k1(bann_net* net, int first, bool exchange) {
    HmcRun R = grouped_run(net);
    K1Launch k;
    k.list = nullptr;
    k.nlist = (uint32_t)net->B;
    k.states = net->d_states;
    k.xg = exchange && sharded(net) && net->xg_connected;   // in-library all-reduce of the sums (bann_net_comm_connect)
    if (first && net->grouped_per_branch) {
        k.target_mode = TGT_RESID_PLUS_PRED;
        k.resid = net->d_r;
        k.tgt_out = net->d_T;
        k.out_per_entry = 1;
    } else {
        k.target_mode = R.later_mode;
        k.tgt = R.tgt;
    }
    return launch_k1(net, k, true);
}

int bann_grouped_begin(bann_net* net, const bann_mcmc_cfg* cfg, uint64_t seed, int per_branch_targets) {
    if (!net || !cfg) BANN_FAIL("NULL argument");
    net->grouped_per_branch = per_branch_targets ? 1 : 0;
    net->grouped_seed = seed;
    if (per_branch_targets && !net->d_T) BANN_CUDA(cudaMalloc(&net->d_T, (size_t)net->B * net->n * sizeof(float)));
    HmcRun R = grouped_run(net);
    BANN_CHECK(hmc_init_and_first_eval(net, cfg, R, 0));
    BANN_CHECK(grouped_k1(net, 1, true));
    if (net->ctx->world > 1 && !net->xg_connected) return 0;   // caller all-reduces, then bann_grouped_phase_b(is_init = 1)
    K2Args a = make_k2(net, cfg, R, 1, 0);
    BANN_CUDA(launch_pdl(k2_step, dim3(R.nlist), dim3(256), 0, net->ctx->stream, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

int bann_grouped_phase_a(bann_net* net) {
    if (!net) BANN_FAIL("NULL net");
    return grouped_k1(net, 0, false);   // the caller sums over ranks: bann_grouped_allreduce, or its own collective
}

int bann_grouped_allreduce(bann_net* net) {
    if (!net) BANN_FAIL("NULL net");
    if (!sharded(net)) return 0;
    return xg_allreduce(net, net->d_gsum, (uint64_t)net->B * net->pstride);
}

const char* bann_net_last_k1_kernel(bann_net* net) { return net ? net->last_k1 : "none"; }

int bann_grouped_phase_b(bann_net* net, const bann_mcmc_cfg* cfg, int is_init, int is_last) {
    if (!net || !cfg) BANN_FAIL("NULL argument");
    HmcRun R = grouped_run(net);
    K2Args a = make_k2(net, cfg, R, is_init, is_last);
    BANN_CUDA(launch_pdl(k2_step, dim3(R.nlist), dim3(256), 0, net->ctx->stream, a));
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    return 0;
}

int bann_grouped_leapfrog(bann_net* net, const bann_mcmc_cfg* cfg, uint32_t num_steps, int finalize) {
    if (!net || !cfg) BANN_FAIL("NULL argument");
    if (net->ctx->world > 1 && !net->xg_connected)
        BANN_FAIL("sharded rows: connect the bulk exchange (bann_net_comm_connect) or drive bann_grouped_phase_a / _b with your own all-reduce in between");
    for (uint32_t s = 0; s < num_steps; ++s) {
        BANN_CHECK(grouped_k1(net, 0, true));
        BANN_CHECK(bann_grouped_phase_b(net, cfg, 0, (finalize && s + 1 == num_steps) ? 1 : 0));
    }
    return 0;
}

int bann_grouped_finish(bann_net* net, uint64_t seed, uint64_t* num_accepted, uint64_t* num_early_rejected) {
    if (!net) BANN_FAIL("NULL net");
    cudaStream_t st = net->ctx->stream;
    k_accept<<<(unsigned)net->B, 256, 0, st>>>(net->d_descs, nullptr, net->d_states, net->d_theta, net->d_theta0, nullptr,
                                               seed, net->visit_seq * net->B);
    BANN_LAUNCHED();
    unsigned long long* cnt = (unsigned long long*)net->d_scratchB;
    BANN_CUDA(cudaMemsetAsync(cnt, 0, 2 * sizeof(unsigned long long), st));
    k_count_status<<<((unsigned)net->B + 255) / 256, 256, 0, st>>>(net->d_states, (uint32_t)net->B, cnt);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    unsigned long long h[2];
    BANN_CUDA(cudaMemcpyAsync(h, cnt, sizeof(h), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    if (num_accepted) *num_accepted = h[0];
    if (num_early_rejected) *num_early_rejected = h[1];
    net->visit_seq += 1;
    return 0;
}

int bann_grouped_state(bann_net* net, float* neg_h_init, float* neg_h_cur, int32_t* status) {
    if (!net) BANN_FAIL("NULL net");
    cudaStream_t st = net->ctx->stream;
    float* hi = net->d_scratchB;
    float* hc = hi + net->B;
    int* ss = (int*)(hc + net->B);
    k_gather_states<<<((unsigned)net->B + 255) / 256, 256, 0, st>>>(net->d_states, (uint32_t)net->B, hi, hc, ss);
    BANN_LAUNCHED();
    BANN_CUDA(cudaGetLastError());
    if (neg_h_init) BANN_CUDA(cudaMemcpyAsync(neg_h_init, hi, net->B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (neg_h_cur) BANN_CUDA(cudaMemcpyAsync(neg_h_cur, hc, net->B * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (status) BANN_CUDA(cudaMemcpyAsync(status, ss, net->B * sizeof(int), cudaMemcpyDeviceToHost, st));
    BANN_CUDA(cudaStreamSynchronize(st));
    return 0;
}

int bann_allreduce_buffer(bann_net* net, void** dev_ptr, uint64_t* num_floats) {
    if (!net || !dev_ptr || !num_floats) BANN_FAIL("NULL argument");
    BANN_CHECK(ensure_cap(&net->d_gsum, &net->gsum_cap, (size_t)net->B * net->pstride));
    *dev_ptr = net->d_gsum;
    *num_floats = (uint64_t)net->B * net->pstride;
    return 0;
}

int bann_net_algorithmic_bytes(bann_net* net, uint64_t* bytes) {
    if (!net || !bytes) BANN_FAIL("NULL argument");
    // sum_b m_b*ceil(N/4) + 4*N*B + 12*sum_b P_b   (N = rows this rank holds)
    uint64_t sumP = 0;
    for (uint64_t b = 0; b < net->B; ++b) sumP += net->descs[b].P;
    *bytes = net->gen->packed_bytes + 4ull * net->n * net->B + 12ull * sumP;
    return 0;
}

int bann_net_force_generic(bann_net* net, int on) {
    if (!net) BANN_FAIL("NULL net");
    net->k1_mode = on ? BANN_K1_GENERIC : BANN_K1_AUTO;
    return 0;
}

int bann_net_select_hmc_path(bann_net* net, int which) {
    if (!net) BANN_FAIL("NULL net");
    if (which < BANN_HMC_AUTO || which > BANN_HMC_PERSISTENT) BANN_FAIL("unknown HMC path selector");
    net->hmc_path = which;
    return 0;
}

uint64_t bann_net_persistent_launches(bann_net* net) { return net ? net->persistent_launches : 0; }

int bann_net_select_k1_tc_variant(bann_net* net, int which) {
    if (!net) BANN_FAIL("NULL net");
    if (which < BANN_TC_FOUR_WARPS || which > BANN_TC_FIVE_WARPS_PLAIN) BANN_FAIL("unknown k1_tc variant");
    net->tc_variant = which;
    return 0;
}

int bann_net_select_k1(bann_net* net, int which) {
    if (!net) BANN_FAIL("NULL net");
    if (which < BANN_K1_AUTO || which > BANN_K1_GENERIC) BANN_FAIL("unknown K1 selector");
    net->k1_mode = which;
    return 0;
}

}  // extern "C"
