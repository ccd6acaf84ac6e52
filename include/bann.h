/*
 * bann.h -- C ABI of the B200-native rs-bann hot path (libbann_b200.so).
 *
 * The reference (medical-genomics-group/rs-bann) has no FFI boundary today: the seam is a set
 * of Rust traits over the third-party `arrayfire` crate.  This header places the replacement
 * boundary at exactly those traits.  Each entry point cites the
 * reference interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every call returns int32 status: 0 = ok, < 0 = error; bann_last_error() (thread local)
 *     describes the last error.  The library never aborts the process.
 *   - all pointers are HOST pointers owned by the caller unless the name ends in `_dev`.
 *   - matrices are column-major, weights are in x out, as in the reference (ArrayFire).
 *   - parameter vectors use BranchParams::param_vec order (net/params.rs:700-715): all weights
 *     layer by layer (column-major), then all biases.  Precision vectors use
 *     BranchPrecisions::param_vec order (net/params.rs:272-289): weight precisions per layer,
 *     bias precisions per layer, error precision.
 *   - handles are opaque, not thread safe per handle.
 *   - there is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a).
 */
#ifndef BANN_H_
#define BANN_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bann_ctx bann_ctx;
typedef struct bann_genotypes bann_genotypes;
typedef struct bann_net bann_net;

/* net/model_type.rs:6-13 */
enum { BANN_STD_NORMAL = 0, BANN_RIDGE_BASE = 1, BANN_RIDGE_ARD = 2, BANN_LASSO_BASE = 3, BANN_LASSO_ARD = 4 };
/* net/activation_functions.rs:6-12 */
enum { BANN_TANH = 0, BANN_RELU = 1, BANN_LEAKY_RELU = 2, BANN_SILU = 3, BANN_IDENTITY = 4 };
/* net/mcmc_cfg.rs:265-270 */
enum { BANN_STEP_UNIFORM = 0, BANN_STEP_RANDOM = 1, BANN_STEP_STD_SCALED = 2, BANN_STEP_IZMAILOV = 3 };
/* net/branch/branch_sampler.rs:1310-1314 (HMCStepResult) */
enum { BANN_HMC_REJECTED_EARLY = 0, BANN_HMC_REJECTED = 1, BANN_HMC_ACCEPTED = 2 };

#define BANN_MAX_LAYERS 8

/* One branch's architecture: layer_widths = hidden..., summary, 1
 * (net/branch/branch_cfg_builder.rs:285-297, net/branch/branch_cfg.rs:8-16). */
typedef struct {
    uint32_t num_layers;                  /* d + 2 */
    uint32_t widths[BANN_MAX_LAYERS];     /* widths[num_layers-1] must be 1 */
} bann_branch_layout;

/* net/mcmc_cfg.rs:181-204 -- the fields the hot path reads. */
typedef struct {
    float    hmc_step_size_factor;        /* default 1.0 */
    float    hmc_max_hamiltonian_error;   /* default 10.0 */
    uint32_t hmc_integration_length;      /* default 100 */
    int32_t  hmc_step_size_mode;          /* BANN_STEP_* , default Izmailov */
    int32_t  fixed_param_precisions;      /* skip sample_param_precisions (net.rs:273) */
    /* flag-gated sampler modes of Net::train (net/mcmc_cfg.rs:21-26, net/net.rs:268-290); all 0 = hmc_step */
    int32_t  joint_hmc;                   /* hmc_step_joint: precisions are part of the HMC state, no Gibbs draws */
    int32_t  gradient_descent;            /* gradient_descent: line-search ascent on the log density */
    int32_t  gradient_descent_joint;      /* gradient_descent_joint: fixed-step ascent on parameters and precisions */
    /* the reference's debugging aids (net/mcmc_cfg.rs, branch_sampler.rs:1232-1261; "DO NOT run this in production code"):
     * hmc_step integrates with numerical_ldg instead of the analytical gradient / records numerical_ldg per step in the
     * trajectory.  Per-branch transitions only (group_size 1); P_b + 2 fused passes per leapfrog step. */
    int32_t  num_grad;
    int32_t  num_grad_traj;
} bann_mcmc_cfg;

/* Injected randomness for parity runs.  NULL members fall back to the built-in
 * counter-based Philox4x32-10 stream keyed by (seed, visit counter, branch). */
typedef struct {
    const float* momenta;        /* P_b N(0,1) draws, param_vec order (branch_sampler.rs:594-609) */
    const float* accept_uniform; /* 1 value in [0,1)            (branch_sampler.rs:546-548) */
    const float* step_uniforms;  /* P_b U(0,1) draws for BANN_STEP_RANDOM (branch_sampler.rs:654-681) */
    const float* std_gammas;     /* standard-gamma variates in consumption order of one visit:
                                    error precision; per layer l<last: weight precision(s), bias
                                    precision; output weight precision (net.rs:272-275) */
    uint32_t     num_std_gammas;
} bann_rng_inject;

typedef struct {
    int32_t  status;             /* BANN_HMC_* */
    float    log_density;        /* at the final state (valid when accepted) */
    float    neg_h_init;
    float    neg_h_final;
    uint32_t steps_done;
    int32_t  u_turn_step;        /* first step with (theta-theta0).p < 0, or -1 */
} bann_hmc_result;

/* Optional per-step trajectory (net/branch/trajectory.rs:4-11); arrays sized by the caller:
 * params/ldg: L * P_b floats, hamiltonian: L + 1 floats. */
typedef struct {
    float* params;
    float* ldg;
    float* hamiltonian;
} bann_trajectory;

/* The same for hmc_step_joint (net/branch/trajectory.rs:4-43): params L * P_b, precisions L * Q_b,
 * ldg L * (P_b + Q_b) in BranchLogDensityGradientJoint::param_vec order (net/branch/gradient.rs:66-97:
 * weights, biases, weight precisions, bias precisions, error precision), hamiltonian L + 1. */
typedef struct {
    float* params;
    float* precisions;
    float* ldg;
    float* hamiltonian;
    float* num_ldg;              /* NULL, or L * P_b: numerical_ldg per step when cfg->num_grad_traj (Trajectory::num_ldg) */
} bann_trajectory_joint;

/* net/train_stats.rs:23-32 + net/log_posterior_density.rs:62-67 */
typedef struct {
    uint64_t num_samples;
    uint64_t num_accepted;
    uint64_t num_early_rejected;
    float    mse_train;
    float    lpd;
    float    output_bias;
    float    error_precision;
    float    output_layer_precision;
} bann_sweep_stats;

const char* bann_last_error(void);
/* 1 if a CUDA device is usable, 0 otherwise (no compute is attempted). */
int bann_cuda_available(void);

/* ---- lifetime.  One context per process/GPU.  `stream` may be NULL (own stream) or a
 * cudaStream_t (e.g. torch's current stream).  rank/world describe the row shard this
 * process holds; cross-rank sums are the caller's all-reduce over the buffers exposed by
 * bann_allreduce_buffer() (torch.distributed / NCCL plumbing). */
int  bann_ctx_create(int device, void* stream, int rank, int world, bann_ctx** out);
void bann_ctx_destroy(bann_ctx*);
int  bann_ctx_sync(bann_ctx*);

/* ---- cross-rank sums of the sequential-exact schedule.  The reference is single-process;
 * what these replace are its N-row reductions (`sum_all`, `dot`, `matmul(delta^T, X)` in
 * net/branch/branch_sampler.rs:813-875,905-909 and the residual sums of net/net.rs:43-45,604-606), which
 * with sharded rows become sums over ranks.  Each rank allocates an inbox in its own HBM and exports a
 * handle; the host layer all-gathers the world * BANN_COMM_HANDLE_BYTES bytes (rank order) and every rank
 * maps its peers' inboxes (CUDA IPC over NVLink; plain pointers for ranks of the same process).  From then
 * on bann_visit_branch / bann_sweep / bann_hmc_step / bann_branch_fwd_bwd / bann_net_init_residual sum over
 * ranks INSIDE their reduction kernels (8-byte {value, epoch} stores into peer memory, rank-ordered adds:
 * bit-identical results on every rank), with no collective launch and no host involvement.  All ranks
 * must issue the same sequence of calls. */
#define BANN_COMM_HANDLE_BYTES 128
int  bann_ctx_comm_handle(bann_ctx*, uint8_t* handle_out /* BANN_COMM_HANDLE_BYTES */);
int  bann_ctx_comm_connect(bann_ctx*, const uint8_t* all_handles /* world * BANN_COMM_HANDLE_BYTES */);
int  bann_ctx_comm_connected(bann_ctx*);
/* ---- bulk cross-rank sums of the grouped schedule ("all-reduce(sum) of [gW | gb | rss], G x (P_b + 1) floats when
 * G branches are grouped").  With these connected the library is self-sufficient on sharded rows: the all-reduce between the
 * fused forward+backward launch and the parameter update runs INSIDE the library over NVLink peer memory -- reduce-scatter
 * (rank-ordered sums: identical bits on every rank) + all-gather, 2 (world - 1) / world of the bytes per rank, three small
 * kernels on the caller's stream, no NCCL, no host involvement (rs-bann_b200/csrc/comm.cuh: XgComm).  Used by
 * bann_sweep / bann_visit_group with group_size > 1, bann_grouped_begin / _leapfrog and bann_net_gradient.  Set-up mirrors the
 * context's: every rank exports the handle of its exchange region, the host layer all-gathers them once (the only step that
 * needs a host-side transport, e.g. torch.distributed / MPI all_gather of 128 bytes), every rank maps its peers. */
int  bann_net_comm_handle(bann_net*, uint8_t* handle_out /* BANN_COMM_HANDLE_BYTES */);
int  bann_net_comm_connect(bann_net*, const uint8_t* all_handles /* world * BANN_COMM_HANDLE_BYTES */);
int  bann_net_comm_connected(bann_net*);

/* ---- genotypes: replaces BedVM + MarkerGrouping + GroupedGenotypes::x_group_af
 * (io/bed.rs:123-133,193-245,325-355; group/grouping.rs:7-15; data/genotypes.rs:7-48).
 * bed_payload: PLINK variant-major 2-bit payload WITHOUT the 3-byte signature, m * ceil(n/4)
 * bytes, for the n rows this rank holds.  col_means/col_stds: length m, or NULL to compute
 * them on the device exactly as io/bed.rs:231-238 does (sequential f32, population std).
 * branch_offsets[B+1] / col_ids[branch_offsets[B]]: CSR of each branch's marker columns
 * (arbitrary order, may overlap).  n_total: rows over all ranks (statistics / N in formulas). */
int  bann_genotypes_create(bann_ctx*, const uint8_t* bed_payload, uint64_t n, uint64_t n_total, uint64_t m,
                           const float* col_means, const float* col_stds, uint64_t num_branches,
                           const uint64_t* branch_offsets, const uint64_t* col_ids, bann_genotypes** out);
/* Synthetic store generated on the device, in the spirit of BedVM::random (io/bed.rs:136-188):
 * maf_j ~ U(maf_lo, maf_hi), g_ij ~ Binomial(2, maf_j); counter-based RNG keyed by column and
 * GLOBAL row (row_offset + i), so the data are independent of the row sharding.  With world > 1
 * the column statistics must then be set from all-reduced bann_genotypes_col_counts. */
int  bann_genotypes_random(bann_ctx*, uint64_t n, uint64_t row_offset, uint64_t n_total, uint64_t m, uint64_t seed,
                           float maf_lo, float maf_hi, uint64_t num_branches, const uint64_t* branch_offsets,
                           const uint64_t* col_ids, bann_genotypes** out);
void bann_genotypes_destroy(bann_genotypes*);
int  bann_genotypes_col_stats(bann_genotypes*, float* col_means, float* col_stds);
/* per-column counts of decoded values 0,1,2 over the local rows: out[3*m] (for global stats). */
int  bann_genotypes_col_counts(bann_genotypes*, uint64_t* out);
int  bann_genotypes_set_col_stats(bann_genotypes*, const float* col_means, const float* col_stds);
/* test hook: decode branch b to f32 [n x m_b] column-major, raw 0/1/2 or standardised. */
int  bann_genotypes_decode_branch(bann_genotypes*, uint64_t b, int standardized, float* out);
/* the same, decoded from the tensor-core store (bf16-subnormal layout, every branch <= 512 markers) */
int  bann_genotypes_decode_branch_tc(bann_genotypes*, uint64_t b, int standardized, float* out);
int  bann_genotypes_has_tc_store(bann_genotypes*);
int  bann_genotypes_has_byte_store(bann_genotypes*);
/* Frees the byte-tile store (the copy the FFMA / shape-agnostic kernels, the probe kernels and bann_genotypes_decode_branch
 * read) when the tensor-core store exists; later calls that need it fail with a message.  No reference counterpart: the
 * reference keeps one copy of the genotypes (io/bed.rs:431-446). */
int  bann_genotypes_release_byte_store(bann_genotypes*);

/* ---- model state: replaces Vec<BranchCfg> + from_cfg/to_cfg round trips
 * (net/net.rs:76-85, net/branch/branch_struct.rs:12-29, net/branch/branch_sampler.rs:155-171).
 * hyper = dense(shape,scale), summary(shape,scale), output(shape,scale) (net/params.rs:135-142). */
int  bann_net_create(bann_ctx*, bann_genotypes*, int model_type, int activation,
                     const bann_branch_layout* layouts /* one per branch */, const float hyper[6], bann_net** out);
void bann_net_destroy(bann_net*);
int  bann_net_branch_sizes(bann_net*, uint64_t b, uint64_t* num_params, uint64_t* num_precisions);
int  bann_net_set_branch(bann_net*, uint64_t b, const float* param_vec, const float* precision_vec);
int  bann_net_get_branch(bann_net*, uint64_t b, float* param_vec, float* precision_vec);
/* bulk variants over all branches, concatenated in branch order */
int  bann_net_set_all_params(bann_net*, const float* param_vecs, const float* precision_vecs_or_null);
int  bann_net_get_all_params(bann_net*, float* param_vecs, float* precision_vecs_or_null);
/* GlobalParams + OutputBias (net/params.rs:13-56, net/net.rs:29-36):
 * g[0]=error_precision g[1]=output_layer_precision g[2]=output-weight reg_sum (all branches)
 * g[3]=output-weight num_params g[4]=output bias */
int  bann_net_set_globals(bann_net*, const float g[5]);
int  bann_net_get_globals(bann_net*, float g[5]);
int  bann_net_set_targets(bann_net*, const float* y /* n local rows */);
int  bann_net_get_residual(bann_net*, float* r /* n local rows */);
int  bann_net_set_residual(bann_net*, const float* r /* n local rows */);
/* initialize_stats (net/net.rs:158-171): residual = y - bias - sum_b predict_b, LPD terms. */
int  bann_net_init_residual(bann_net*);

/* ---- hot path */
/* backpropagate + log_density_gradient for branch b (branch_sampler.rs:813-875,380-391).
 * target NULL -> the net's targets y.  Outputs (any may be NULL): rss, log-density gradient in
 * param_vec order, raw d_rss (1/2 dRSS/dtheta, Q4), yhat (n). */
int  bann_branch_fwd_bwd(bann_net*, uint64_t b, const float* target, float* rss, float* ldg, float* d_rss,
                         float* yhat);
/* log_density(params, precisions, rss) (branch_sampler.rs:72-78; std_normal_branch.rs:147-158) */
int  bann_branch_log_density(bann_net*, uint64_t b, float rss, float* out);
/* numerical_ldg (branch_sampler.rs:480-504): forward differences of log_density with NUMERICAL_DELTA = 0.001, P_b + 1 fused
 * passes; out: P_b floats in param_vec order.  target NULL: the net's targets.  The parameters are left unchanged. */
int  bann_branch_numerical_ldg(bann_net*, uint64_t b, const float* target, float* out);
/* per-parameter step sizes of the chosen mode (a10), param_vec order */
int  bann_branch_step_sizes(bann_net*, uint64_t b, const bann_mcmc_cfg*, const float* step_uniforms, float* out);
/* hmc_step(x_b, target, cfg) (branch_sampler.rs:1192-1299).  target NULL -> net targets.
 * yhat_out (n, may be NULL) receives the prediction at the final state. */
int  bann_hmc_step(bann_net*, uint64_t b, const float* target, const bann_mcmc_cfg*, const bann_rng_inject*,
                   bann_hmc_result* out, bann_trajectory* traj, float* yhat_out);
/* ---- flag-gated sampler modes.  Q_b = number of precisions of the branch
 * (bann_net_branch_sizes); the joint state is [param_vec | BranchPrecisions::param_vec (net/params.rs:272-289)].
 * The output-weight statistic of the other branches comes from the globals (bann_net_set_globals g[2], g[3]) minus the
 * branch's own (net/branch/branch_struct.rs:27).  StdNormal fails: its joint density is unimplemented!() in the reference. */
/* log_density_gradient_joint (branch_sampler.rs:406-422) + log_density_joint (:292-305) + log_density (:72-78) of the
 * current state against `target` (NULL -> net targets).  ldg_joint: P_b + Q_b floats.  Any output may be NULL. */
int  bann_branch_joint(bann_net*, uint64_t b, const float* target, float* rss, float* log_density_joint,
                       float* log_density, float* ldg_joint);
/* hmc_step_joint (branch_sampler.rs:1070-1178): Random step sizes with the joint factor (P_b + Q_b)^(-1/4) f whatever the
 * configured mode (:1094-1101); inject->momenta / step_uniforms carry P_b + Q_b values.  The accept step evaluates the
 * NON-joint log density against the joint initial Hamiltonian, as the reference does (:928-962,1164-1172). */
int  bann_hmc_step_joint(bann_net*, uint64_t b, const float* target, const bann_mcmc_cfg*, const bann_rng_inject*,
                         bann_hmc_result* out, bann_trajectory_joint* traj, float* yhat_out);
/* gradient_descent (branch_sampler.rs:964-1017): hmc_integration_length ascent steps, each with the doubling / halving
 * line search on probe RSS values.  Always BANN_HMC_ACCEPTED.  step_sizes_out (L, may be NULL): the step size taken in
 * every iteration; num_probes (may be NULL): probe_gradient_step evaluations. */
int  bann_gradient_descent(bann_net*, uint64_t b, const float* target, const bann_mcmc_cfg*, bann_hmc_result* out,
                           float* step_sizes_out, uint32_t* num_probes, float* yhat_out);
/* gradient_descent_joint (branch_sampler.rs:1019-1066): fixed step hmc_step_size_factor on parameters and precisions;
 * BANN_HMC_REJECTED (state restored) when the error precision ends <= 0.  out->log_density: log_density_joint. */
int  bann_gradient_descent_joint(bann_net*, uint64_t b, const float* target, const bann_mcmc_cfg*, bann_hmc_result* out,
                                 float* yhat_out);
/* sample_error_precision + sample_param_precisions against the current residual
 * (branch_sampler.rs:173-202 and the per-prior sample_prior_precisions). */
int  bann_gibbs_branch(bann_net*, uint64_t b, const bann_mcmc_cfg*, const bann_rng_inject*);
/* one iteration of the inner loop of Net::train (net/net.rs:258-334): globals -> cfg, Gibbs,
 * prev_pred, HMC against residual + prev_pred, residual / LPD / globals / output-bias update. */
int  bann_visit_branch(bann_net*, uint64_t b, const bann_mcmc_cfg*, const bann_rng_inject*, bann_hmc_result* out);
/* the same with the built-in RNG keyed by `seed` and the per-step trajectory of the visit's HMC transition recorded
 * (net/branch/trajectory.rs:4-43, the `traj` file of --trajectories): params L * P_b, ldg L * P_b, hamiltonian L + 1; with
 * cfg->joint_hmc also precisions L * Q_b and ldg L * (P_b + Q_b).  Rows after an early rejection stay zero (out->steps_done
 * tells how many are valid).  The ascent modes record nothing. */
int  bann_visit_branch_traj(bann_net*, uint64_t b, const bann_mcmc_cfg*, uint64_t seed, bann_hmc_result* out,
                            bann_trajectory_joint* traj);
/* a full pass over `branch_order` with the built-in RNG keyed by seed; asynchronous, one sync at the end.
 * group_size 1: the reference's sequential (Gauss-Seidel) order, net/net.rs:258-334 visit by visit.
 * group_size G > 1: block-Jacobi -- consecutive groups of G branches of the order advance concurrently: every member draws its
 * precisions and runs its HMC transition against the residual and the global parameters frozen at group start
 * (t_b = r + yhat_b, one fused forward+backward launch over the whole group per leapfrog step); residual
 * (r -= sum over accepted members of yhat_new - yhat_old), global parameters, LPD terms, counters and the ML output bias are
 * updated once per group.  G = number of branches is the schedule of the full-network leapfrog metric. */
int  bann_sweep(bann_net*, const bann_mcmc_cfg*, const uint64_t* branch_order, uint64_t num, uint32_t group_size,
                uint64_t seed, bann_sweep_stats* out);
/* one group visit of the block-Jacobi schedule (see bann_sweep): `members` = num distinct branches; inj = NULL or an array of
 * num per-member injections (parity runs: the oracle's visit_group replayed draw for draw); out = NULL or num results in
 * member order.  One member without injections is bann_visit_branch. */
int  bann_visit_group(bann_net*, const uint64_t* members, uint64_t num, const bann_mcmc_cfg*, const bann_rng_inject* inj_or_null,
                      uint64_t seed, bann_hmc_result* out_or_null);
/* Net::predict (net/net.rs:545-559) on the training genotypes (NULL) or another store. */
int  bann_predict(bann_net*, bann_genotypes* test_or_null, float* yhat);
int  bann_net_stats(bann_net*, bann_sweep_stats* out);
/* ---- per-row diagnostics of saved models (the `activations` / `population-effect-sizes` subcommands and
 * --effect-sizes; off the sampler's hot path).  genotypes NULL -> the training store.
 * Net::activations (net/net.rs:509-518): forward_feed of branch b (branch_sampler.rs:743-782);
 * out = [a_0 (n x w_0) | a_1 (n x w_1) | ... | yhat (n x 1)], every block column-major, n * (sum w_l + 1) floats. */
int  bann_branch_activations(bann_net*, uint64_t b, bann_genotypes* genotypes_or_null, float* out);
/* BranchSampler::effect_sizes (branch_sampler.rs:784-811): the prediction back-propagated to the standardised input,
 * seeded with yhat itself as the reference does; n x m_b column-major.  population_effect_sizes (net/net.rs:529-543):
 * its column means, m_b floats.  Either output may be NULL. */
int  bann_branch_effect_sizes(bann_net*, uint64_t b, bann_genotypes* genotypes_or_null, float* effect_sizes,
                              float* population_effect_sizes);
/* the three groups of terms of LogPosteriorDensity (net/log_posterior_density.rs:9-16), as serialised in a model file;
 * wrt_local_params: one value per branch */
int  bann_net_lpd_terms(bann_net*, float* wrt_rss_and_error_precision, float* wrt_output_weights_and_precision,
                        float* wrt_local_params);

/* ---- full-network (grouped, all branches concurrently) operations */
/* Net::gradient (net/net.rs:520-527): for every branch log_density_gradient(x_b, y).
 * HOST buffers: params in (sum P_b, may be NULL = keep device state), y in (n, may be NULL),
 * grads out (sum P_b), rss out (B).  This is the host-facing full-network fwd+grad call. */
int  bann_net_gradient(bann_net*, const float* param_vecs, const float* y, float* grads, float* rss);
/* Page-locked host buffers for the calls above: pinned (or cudaHostRegister-ed) caller buffers are read / written by DMA
 * directly, one copy per direction; pageable ones are staged through an internal pinned buffer (one extra memcpy each way). */
int  bann_pinned_alloc(uint64_t bytes, void** out);
void bann_pinned_free(void* p);
/* the same in two halves for row-sharded runs WITHOUT the bulk exchange: begin (H2D, fused fwd+bwd, raw sums into the
 * all-reduce buffer) -- caller all-reduces -- end (gradient under the prior, D2H).
 * With the bulk exchange connected (bann_net_comm_connect) bann_net_gradient itself works on sharded rows, and the host
 * traffic is divided by the number of ranks: the parameters are replicated, so every rank reads only ITS 1 / world slice of
 * param_vecs (the ranks all-gather on the device over NVLink) and writes only its slice of [grads | rss];
 * bann_net_gradient_slice reports the element ranges [param_lo, param_hi) of param_vecs and [out_lo, out_hi) of the
 * concatenation [grads | rss] this rank touches (the whole vectors on a single rank).  y is this rank's rows, as always. */
int  bann_net_gradient_begin(bann_net*, const float* param_vecs, const float* y);
int  bann_net_gradient_end(bann_net*, float* grads, float* rss);
int  bann_net_gradient_slice(bann_net*, uint64_t* param_lo, uint64_t* param_hi, uint64_t* out_lo, uint64_t* out_hi);
/* Grouped leapfrog over ALL branches against per-branch targets (schedule G = B).
 * begin: theta0 <- theta, step sizes, momenta (Philox, seed), targets t_b = y (shared) or
 * residual + own prediction; then each step = B branch-leapfrogs.  Device resident, async. */
int  bann_grouped_begin(bann_net*, const bann_mcmc_cfg*, uint64_t seed, int per_branch_targets);
/* num_steps full-network leapfrog steps; finalize != 0: the last one ends the trajectory (no
 * further position update) so that bann_grouped_finish can accept / reject. */
int  bann_grouped_leapfrog(bann_net*, const bann_mcmc_cfg*, uint32_t num_steps, int finalize);
/* split form for multi-GPU: phase A = fwd/bwd + chunk reduction into the all-reduce buffer,
 * (caller all-reduces), phase B = gradient, momentum/position update, Hamiltonian check. */
int  bann_grouped_phase_a(bann_net*);
int  bann_grouped_phase_b(bann_net*, const bann_mcmc_cfg*, int is_init, int is_last);
int  bann_grouped_finish(bann_net*, uint64_t seed, uint64_t* num_accepted, uint64_t* num_early_rejected);
/* per-branch Hamiltonians / status after the last step (B each, may be NULL) */
int  bann_grouped_state(bann_net*, float* neg_h_init, float* neg_h_cur, int32_t* status);

/* device buffer (float) that must be sum-all-reduced across ranks between phase A and B and
 * after residual updates; NULL/0 when nothing is pending.  Exposed so that the host side
 * (torch.distributed over NCCL) can run the collective on the same stream. */
int  bann_allreduce_buffer(bann_net*, void** dev_ptr, uint64_t* num_floats);
/* the same sum over ranks done by the library itself over NVLink peer memory (needs bann_net_comm_connect; no-op on one rank):
 * bann_grouped_phase_a -- bann_grouped_allreduce -- bann_grouped_phase_b is what bann_grouped_leapfrog runs per step */
int  bann_grouped_allreduce(bann_net*);
/* name of the kernel family the last fused forward+backward launch used (profiling / bench bookkeeping); static string */
const char* bann_net_last_k1_kernel(bann_net*);

/* test hook: route every K1 launch through the shape-agnostic kernel (cross-checks the tuned one) */
int  bann_net_force_generic(bann_net*, int on);
/* test / profiling hook: which fused forward+backward kernel launches may use.  AUTO tries the tensor-core
 * kernels (tcgen05: <= 64 markers per branch; K-blocked up to 512; the three-pass wide variant for first-layer widths up
 * to 16 and up to 2048 markers), then the FFMA kernel, then the shape-agnostic one. */
enum { BANN_K1_AUTO = 0, BANN_K1_TENSOR = 1, BANN_K1_FFMA = 2, BANN_K1_GENERIC = 3 };
int  bann_net_select_k1(bann_net*, int which);
/* test / profiling hook: which variant of the <= 64-marker tensor-core kernel a gradient / leapfrog launch uses.
 * FOUR_WARPS: k1_tc, the four compute warps issue the tcgen05.mma themselves; FIVE_WARPS (the default): k1_tc5, a dedicated
 * issuing warp, the cross-row sums of a super-tile taken under the next super-tile's MUFU phase (architectures with at most
 * one hidden layer; the others keep k1_tc); FIVE_WARPS_PLAIN: k1_tc5 with the sums where k1_tc takes them ([5,5,1] only, A/B).
 * Same arithmetic per row; the cross-row sums are taken in a different (fixed) order, so the variants agree to FP32
 * rounding, not bit for bit. */
enum { BANN_TC_FOUR_WARPS = 0, BANN_TC_FIVE_WARPS = 1, BANN_TC_FIVE_WARPS_PLAIN = 2 };
int  bann_net_select_k1_tc_variant(bann_net*, int which);
/* test / profiling hook: how a per-branch HMC transition (bann_hmc_step, bann_visit_branch, bann_sweep with group_size 1) runs.
 * AUTO: the persistent cooperative kernel (the whole L-step trajectory of BranchSampler::hmc_step, branch_sampler.rs:1239-1284,
 * in ONE launch, the branch's operands resident on chip) where the branch is eligible, else three launches per leapfrog step;
 * LAUNCHES: always the launch-per-step path; PERSISTENT: fail loudly when the branch is not eligible. */
enum { BANN_HMC_AUTO = 0, BANN_HMC_LAUNCHES = 1, BANN_HMC_PERSISTENT = 2 };
int  bann_net_select_hmc_path(bann_net*, int which);
/* how many transitions ran through the persistent kernel so far */
uint64_t bann_net_persistent_launches(bann_net*);

/* counters for bench.py: kernels launched by this library since the last reset */
uint64_t bann_launch_count(int reset);
/* algorithmic bytes of one full-network leapfrog step (SURVEY 8d formula) */
int  bann_net_algorithmic_bytes(bann_net*, uint64_t* bytes);

#ifdef __cplusplus
}
#endif
#endif /* BANN_H_ */
