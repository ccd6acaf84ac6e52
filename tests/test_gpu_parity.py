"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bar (BASELINE.md section 4 / north_star):
  * packed-genotype decode, CSR gather, column statistics: BIT-EXACT;
  * rss, gradients, log-density, Hamiltonian trajectories: within FP32 tolerance of the oracle's
    f64 truth, no worse than `TOL_K` x the oracle's own f32-mimic error plus `TOL_REL` of the scale;
  * accept / early-reject decisions identical given injected momenta and uniforms (near-ties,
    |log alpha - log u| below tolerance, are skipped and counted).
"""
import os

import numpy as np
import pytest

from oracle import bed as obed
from oracle import net as onet
from oracle.branch import ACCEPTED, REJECTED, REJECTED_EARLY, Branch, BranchCfg, Hyper
from oracle.branch import MCMCCfg as OCfg
from oracle.branch import make_cfg

pytestmark = pytest.mark.gpu

TOL_K = 8.0       # multiples of the oracle's own f32-vs-f64 error
TOL_REL = 2e-5    # relative to the scale of the quantity
MODELS = ["std_normal", "ridge_base", "ridge_ard", "lasso_base", "lasso_ard"]
HYPER = Hyper(dense=(3.0, 2.0), summary=(2.5, 1.5), output=(4.0, 5.0))


@pytest.fixture(scope="module")
def rb():
    import rs_bann_b200 as rb
    if not rb.cuda_available():
        pytest.skip("no CUDA device")
    return rb


@pytest.fixture(scope="module")
def ctx(rb):
    c = rb.Context(0)
    yield c
    c.close()


def within(gpu, truth, mimic, scale=None, k=TOL_K, rel=TOL_REL):
    gpu, truth, mimic = (np.asarray(a, dtype=np.float64) for a in (gpu, truth, mimic))
    sc = np.max(np.abs(truth)) if scale is None else scale
    tol = k * np.abs(mimic - truth) + rel * max(sc, 1e-30) + 1e-30
    bad = np.abs(gpu - truth) > tol
    assert not bad.any(), (f"max err {np.max(np.abs(gpu - truth)):.3e} vs tol {np.max(tol):.3e}; "
                           f"{bad.sum()} / {bad.size} out of tolerance; scale {sc:.3e}")


class Problem:
    """Synthetic genotypes + grouping + a net of random branches, mirrored in the oracle and on the GPU."""

    def __init__(self, rb, ctx, model, n, group_sizes, hidden, summary, depth=1, act="tanh", seed=0, overlap=False,
                 hyper=HYPER):
        rng = np.random.default_rng(seed)
        m = sum(group_sizes) if not overlap else sum(group_sizes) - (len(group_sizes) - 1)
        g = obed.random_genotypes(n, m, seed=seed + 1)
        self.payload = obed.pack_columns(g)
        self.n, self.m = n, m
        self.means, self.stds = obed.col_stats(self.payload, n, m)
        groups, start = [], 0
        for sz in group_sizes:
            cols = list(range(start, start + sz))
            if overlap:
                rng.shuffle(cols)          # arbitrary order inside a group (Q15)
                start += sz - 1            # neighbouring groups share one marker
            else:
                start += sz
            groups.append(cols)
        self.groups = groups
        self.model, self.hyper, self.act = model, hyper, act
        self.cfgs = []
        for gi, cols in enumerate(groups):
            cfg = make_cfg(model, len(cols), [hidden] * depth, summary, activation=act, rng=rng)
            # perturb everything away from the degenerate defaults (zero biases, inf precisions)
            cfg.biases = [rng.normal(0, 0.3, size=b.shape).astype(np.float32) for b in cfg.biases]
            cfg.weights = [(w * 1.5).astype(np.float32) for w in cfg.weights]
            cfg.weight_precisions = [rng.uniform(0.5, 3.0, size=p.shape).astype(np.float32) for p in cfg.weight_precisions]
            cfg.bias_precisions = [rng.uniform(0.5, 3.0, size=p.shape).astype(np.float32) for p in cfg.bias_precisions]
            cfg.error_precision = float(np.float32(rng.uniform(0.5, 2.0)))
            self.cfgs.append(cfg)
        self.y = rng.normal(0, 1, size=n).astype(np.float32)
        self.gen = rb.Genotypes(ctx, self.payload, n, m, groups)
        self.net = rb.Net(ctx, self.gen, model, [c.layer_widths for c in self.cfgs],
                          hyper=(*hyper.dense, *hyper.summary, *hyper.output), activation=act)
        for b, c in enumerate(self.cfgs):
            self.net.set_branch(b, c.param_vec(), c.precision_vec())
        self.net.set_targets(self.y)

    def x(self, b, dt):
        return obed.submatrix_standardized(self.payload, self.n, self.groups[b], self.means, self.stds, dt)

    def close(self):
        self.net.close()
        self.gen.close()


# ------------------------------------------------------------------ genotype store (bit-exact)
def test_decode_small_bed_overlapping_groups(rb, ctx, golden_dir):
    payload, n, m = obed.read_bed(os.path.join(golden_dir, "small"))
    groups = obed.read_grouping(os.path.join(golden_dir, "small.gene_grouping"))
    # column 4 is monomorphic (std 0, Q3) and belongs to no group; statistics still computed for it
    gen = rb.Genotypes(ctx, payload, n, m, groups)
    mu, sd = gen.col_stats()
    omu, osd = obed.col_stats(payload, n, m)
    assert np.array_equal(mu, omu) and np.array_equal(sd, osd)                       # io/bed.rs:468-482 bit-exact
    for b, cols in enumerate(groups):
        raw = gen.x_group(b, standardized=False)
        assert np.array_equal(raw, obed.decode_columns(payload, n, cols))
        std = gen.x_group(b, standardized=True)
        assert np.array_equal(std, obed.submatrix_standardized(payload, n, cols, omu, osd))   # bed.rs:485-497
    counts = gen.col_counts()
    dec = obed.decode_columns(payload, n, range(m))
    for v in range(3):
        assert np.array_equal(counts[:, v], (dec == v).sum(axis=0))
    gen.close()


@pytest.mark.parametrize("n", [1, 3, 4, 100, 127, 128, 129, 1000, 1030])
def test_decode_ragged_rows(rb, ctx, n):
    m = 37
    g = obed.random_genotypes(max(n, 8), m, seed=n)[:n] if n >= 8 else np.random.default_rng(n).integers(0, 3, size=(n, m))
    payload = obed.pack_columns(g)
    # PLINK "missing" code 01 must decode to 0 (Q2): inject a few, only inside valid rows
    pl = payload.copy()
    bpc = (n + 3) // 4
    if n >= 4:
        pl[0] = (pl[0] & 0xFC) | 0x01
        pl[bpc * 5] = (pl[bpc * 5] & 0xF3) | 0x04
    groups = [list(range(0, 10)), [36, 0, 17], list(range(10, 37))]
    mu, sd = obed.col_stats(pl, n, m)
    sd_safe = np.where(sd == 0, 1, sd).astype(np.float32)
    gen = rb.Genotypes(ctx, pl, n, m, groups, col_means=mu, col_stds=sd_safe)
    for b, cols in enumerate(groups):
        assert np.array_equal(gen.x_group(b, False), obed.decode_columns(pl, n, cols))
        assert np.array_equal(gen.x_group(b, True), obed.submatrix_standardized(pl, n, cols, mu, sd_safe))
    gen.close()
    if n >= 8 and np.all(sd > 0):
        gen2 = rb.Genotypes(ctx, pl, n, m, groups)        # device-computed statistics, bit-exact
        dmu, dsd = gen2.col_stats()
        assert np.array_equal(dmu, mu) and np.array_equal(dsd, sd)
        gen2.close()


def test_random_bed_fixture(rb, ctx, golden_dir):
    payload, n, m = obed.read_bed(os.path.join(golden_dir, "random"))
    groups = obed.uniform_grouping(4, 5)
    gen = rb.Genotypes(ctx, payload, n, m, groups)
    mu, sd = gen.col_stats()
    omu, osd = obed.col_stats(payload, n, m)
    assert np.array_equal(mu, omu) and np.array_equal(sd, osd)
    for b, cols in enumerate(groups):
        assert np.array_equal(gen.x_group(b, True), obed.submatrix_standardized(payload, n, cols, omu, osd))
    gen.close()


# ------------------------------------------------------------------ forward / backward / gradient
SHAPES = [  # n, group sizes, hidden, summary, depth
    (200, [3], 2, 1, 1),
    (1000, [100, 37], 2, 2, 1),          # config-1-like
    (777, [50, 50, 9], 5, 5, 1),         # config-2/3 widths
    (300, [20, 64], 5, 3, 2),
    (260, [33], 16, 16, 2),              # config-4 widths
    (150, [12, 7], 4, 2, 0),             # no hidden layer: summary reads the markers
]


def oracle_fwd_bwd(P, b, target, dt):
    br = Branch(P.cfgs[b], dt)
    rss, gw, gb = br.backpropagate(P.x(b, dt), np.asarray(target, dtype=dt))
    _, lw, lb = br.log_density_gradient(P.x(b, dt), np.asarray(target, dtype=dt))
    return dict(rss=float(rss), d_rss=Branch.join_vec(gw, gb), ldg=Branch.join_vec(lw, lb),
                yhat=br.predict(P.x(b, dt)), ld=float(br.log_density(rss)))


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("shape", SHAPES)
def test_fwd_bwd_parity(rb, ctx, model, shape):
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, model, n, gs, h, s, depth=d, seed=(sum(map(ord, model)) + n) % 1000, overlap=len(gs) > 1)
    try:
        for k1 in (P.net.K1_GENERIC, P.net.K1_FFMA, P.net.K1_AUTO):   # AUTO = tensor-core kernel where eligible
            P.net.select_k1(k1)
            for b in range(len(gs)):
                tgt = P.y if b % 2 == 0 else (P.y * 0.5 + 0.1).astype(np.float32)
                got = P.net.branch_fwd_bwd(b, target=None if b % 2 == 0 else tgt)
                t64, t32 = oracle_fwd_bwd(P, b, tgt, np.float64), oracle_fwd_bwd(P, b, tgt, np.float32)
                within(got["yhat"], t64["yhat"], t32["yhat"])
                within(got["rss"], t64["rss"], t32["rss"])
                within(got["d_rss"], t64["d_rss"], t32["d_rss"])
                within(got["ldg"], t64["ldg"], t32["ldg"])
                within(P.net.branch_log_density(b, t32["rss"]), Branch(P.cfgs[b], np.float64).log_density(t32["rss"]),
                       Branch(P.cfgs[b], np.float32).log_density(np.float32(t32["rss"])))
    finally:
        P.close()


# shapes the tensor-core kernel (k1_tc.cuh) must take: <= 64 markers per branch, 3 * first width <= 16
TC_SHAPES = [  # n, group sizes, hidden, summary, depth
    (7, [5], 5, 5, 1),
    (255, [1, 8, 9], 5, 5, 1),
    (256, [50, 64], 5, 5, 1),
    (257, [16, 17], 2, 2, 1),
    (1030, [50, 50, 50, 33], 5, 5, 1),     # several super-tiles, ragged tail, overlapping groups
    (700, [40, 24], 4, 3, 1),
    (515, [64, 7], 4, 3, 2),
    (300, [20, 64], 5, 3, 2),
    (150, [12, 7], 4, 2, 0),
    # 65..512 markers: the K-blocked variant (k1_tc_wide.cuh), 2..8 marker blocks, ragged last block
    (300, [100], 5, 5, 1),
    (1030, [500, 65], 5, 5, 1),
    (515, [128, 129], 2, 2, 1),
    (700, [512], 4, 3, 1),
    # further instantiations: [3,3,1] (the CLI test's architecture), [4,4,1], two hidden layers of 5
    (600, [20, 64], 3, 3, 1),
    (300, [33, 8], 4, 4, 1),
    (515, [50, 64], 5, 5, 2),
    (700, [100, 400], 3, 3, 1),
    (300, [65, 200], 4, 4, 1),
    (515, [256, 90], 5, 5, 2),
]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tensor_core_store_decode_bit_exact(rb, ctx, shape):
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, "ridge_ard", n, gs, h, s, depth=d, seed=n, overlap=len(gs) > 1)
    try:
        assert P.gen.has_tc_store()
        for b, cols in enumerate(P.groups):
            assert np.array_equal(P.gen.x_group_tc(b, False), obed.decode_columns(P.payload, n, cols))
            assert np.array_equal(P.gen.x_group_tc(b, True), P.x(b, np.float32))
    finally:
        P.close()


@pytest.mark.parametrize("model", ["ridge_ard", "std_normal", "lasso_base"])
@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tensor_core_fwd_bwd_parity(rb, ctx, model, shape):
    """k1_tc (tcgen05, bf16-subnormal genotypes x 3-piece bf16 weights / deltas) against the oracle and the FFMA kernel."""
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, model, n, gs, h, s, depth=d, seed=(sum(map(ord, model)) + 7 * n) % 1000, overlap=len(gs) > 1)
    try:
        for b in range(len(gs)):
            tgt = P.y if b % 2 == 0 else (P.y * 0.5 + 0.1).astype(np.float32)
            P.net.select_k1(P.net.K1_TENSOR)          # fails loudly if the launch is not eligible
            got = P.net.branch_fwd_bwd(b, target=None if b % 2 == 0 else tgt)
            P.net.select_k1(P.net.K1_FFMA)
            ffma = P.net.branch_fwd_bwd(b, target=None if b % 2 == 0 else tgt)
            t64, t32 = oracle_fwd_bwd(P, b, tgt, np.float64), oracle_fwd_bwd(P, b, tgt, np.float32)
            for key in ("yhat", "rss", "d_rss", "ldg"):
                within(got[key], t64[key], t32[key])
                sc = np.max(np.abs(t64[key]))
                assert np.max(np.abs(np.asarray(got[key], dtype=np.float64) - ffma[key])) <= 4e-5 * sc + 1e-30, key
    finally:
        P.close()


def test_tensor_core_kernel_refuses_ineligible_launch(rb, ctx):
    P = Problem(rb, ctx, "ridge_ard", 300, [40], 5, 5, depth=3, seed=3)     # three hidden layers: no tensor-core kernel, not even zero-padded
    try:
        assert P.gen.has_tc_store()
        P.net.select_k1(P.net.K1_TENSOR)
        with pytest.raises(RuntimeError):
            P.net.branch_fwd_bwd(0)
        P.net.select_k1(P.net.K1_AUTO)
        got = P.net.branch_fwd_bwd(0)                                       # AUTO: the shape-agnostic kernel
        t64, t32 = oracle_fwd_bwd(P, 0, P.y, np.float64), oracle_fwd_bwd(P, 0, P.y, np.float32)
        within(got["ldg"], t64["ldg"], t32["ldg"])
    finally:
        P.close()


@pytest.mark.parametrize("act", ["tanh", "silu"])
@pytest.mark.parametrize("shape", [(400, [30, 11], 3, 2, 1),        # -> k1_tc on (5,5,1)
                                   (300, [20, 64], 4, 2, 2),        # -> k1_tc on (5,5,2)
                                   (350, [12, 7], 1, 1, 1),         # single units
                                   (300, [40, 9], 5, 4, 0),         # no hidden layer: the summary layer reads the markers, (5,5,0)
                                   (515, [100, 70], 4, 2, 1),       # 65..512 markers -> k1_tcw on (5,5,1)
                                   (300, [600, 90], 10, 6, 1),      # -> k1_tcx on (12,12,1)
                                   (300, [600], 5, 5, 1),           # [5,5,1] with more than 512 markers -> k1_tcx on (8,8,1)
                                   (260, [80, 33], 7, 7, 2)])       # -> k1_tcx on (8,8,2)
def test_fwd_bwd_zero_padded_architectures(rb, ctx, act, shape):
    """Widths no tensor-core kernel is instantiated for run through the next larger instantiated architecture on a
    zero-padded copy of the parameters (padded units: zero weights and biases, activation 0, delta 0) -- K1_TENSOR fails
    loudly instead of falling back, so passing means the padded launch really ran."""
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, "ridge_ard", n, gs, h, s, depth=d, act=act, seed=8)
    try:
        P.net.select_k1(P.net.K1_TENSOR)
        for b in range(len(gs)):
            got = P.net.branch_fwd_bwd(b)
            assert "zero-padded" in P.net.last_k1_kernel()
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(got["yhat"], t64["yhat"], t32["yhat"])
            within(got["rss"], t64["rss"], t32["rss"])
            within(got["d_rss"], t64["d_rss"], t32["d_rss"])
            within(got["ldg"], t64["ldg"], t32["ldg"])
    finally:
        P.close()


def test_chain_on_zero_padded_architecture_matches_generic_kernel(rb, ctx):
    """Three sweeps of Net::train visits on a [3,2,1] net: the padded tensor-core path against the shape-agnostic kernel
    (same seeds): identical decisions, parameters within the FP32 reduction-order tolerance."""
    res = {}
    for mode in ("auto", "generic"):
        P = Problem(rb, ctx, "ridge_ard", 600, [30, 11, 25], 3, 2, seed=9)
        try:
            P.net.select_k1(P.net.K1_AUTO if mode == "auto" else P.net.K1_GENERIC)
            onet_ = mirror_net(P)
            P.net.set_globals(2.0, 0.05, onet_.g_ow_reg_sum, onet_.g_ow_num_params, 0.0)
            P.net.init_residual()
            cfg = rb.MCMCCfg(hmc_step_size_factor=0.3, hmc_integration_length=10)
            st = None
            for it in range(3):
                st = P.net.sweep(cfg, np.arange(3), seed=100 + it)
            res[mode] = (P.net.get_all_params()[0], st, P.net.last_k1_kernel())
        finally:
            P.close()
    assert "zero-padded" in res["auto"][2] and "generic" in res["generic"][2]
    assert res["auto"][1]["num_accepted"] == res["generic"][1]["num_accepted"]
    assert res["auto"][1]["num_early_rejected"] == res["generic"][1]["num_early_rejected"]
    assert np.allclose(res["auto"][0], res["generic"][0], rtol=5e-4, atol=5e-5)


def test_release_byte_store(rb, ctx):
    """The byte-tile copy of the genotypes can be given back where the tensor-core store exists: the tensor-core kernels keep
    producing the same numbers, everything that reads the released copy fails with a message (no silent fallback)."""
    P = Problem(rb, ctx, "ridge_ard", 700, [40, 21], 5, 5, seed=4)
    try:
        assert P.gen.has_tc_store() and P.gen.has_byte_store()
        before = [P.net.branch_fwd_bwd(b) for b in range(2)]
        P.gen.release_byte_store()
        assert not P.gen.has_byte_store()
        for b in range(2):
            got = P.net.branch_fwd_bwd(b)
            assert np.array_equal(got["ldg"], before[b]["ldg"]) and np.array_equal(got["yhat"], before[b]["yhat"])
            assert "k1_tc" in P.net.last_k1_kernel()
        assert np.array_equal(P.gen.x_group_tc(0, standardized=False), obed.decode_columns(P.payload, P.n, P.groups[0]))
        with pytest.raises(RuntimeError, match="released"):
            P.gen.x_group(0, standardized=False)
        P.net.select_k1(P.net.K1_GENERIC)
        with pytest.raises(RuntimeError, match="released"):
            P.net.branch_fwd_bwd(0)
        P.net.select_k1(P.net.K1_AUTO)
        with pytest.raises(RuntimeError, match="released"):
            P.net.branch_activations(0)
    finally:
        P.close()


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "silu", "identity"])
@pytest.mark.parametrize("shape", [(500, [40, 21], 5, 5, 1),        # k1_tc
                                   (300, [20, 64], 5, 3, 2),        # k1_tc, two hidden layers (alternating-sign backward)
                                   (515, [128, 129], 5, 5, 1),      # k1_tcw (65..512 markers)
                                   (150, [12, 7], 4, 2, 0),         # summary layer reads the markers
                                   (300, [600, 520], 16, 16, 2),    # k1_tcx (three passes; SiLU sends the pre-activation to the tail pass)
                                   (260, [70, 9], 8, 8, 1)])        # k1_tcx, first-layer width 8
def test_fwd_bwd_activations(rb, ctx, act, shape):
    """activation_functions.rs:23-45 through the shape-agnostic kernel AND the tensor-core kernels (K1_TENSOR fails loudly
    when a launch is not eligible: no silent fall-back to the slow kernel for ReLU / LeakyReLU / SiLU / identity nets)."""
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, "ridge_ard", n, gs, h, s, depth=d, act=act, seed=5)
    try:
        for k1 in (P.net.K1_GENERIC, P.net.K1_TENSOR):
            P.net.select_k1(k1)
            for b in range(len(gs)):
                got = P.net.branch_fwd_bwd(b)
                t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
                within(got["yhat"], t64["yhat"], t32["yhat"])
                within(got["rss"], t64["rss"], t32["rss"])
                within(got["d_rss"], t64["d_rss"], t32["d_rss"])
                within(got["ldg"], t64["ldg"], t32["ldg"])
    finally:
        P.close()


@pytest.mark.parametrize("act", ["tanh", "relu", "leaky_relu", "silu", "identity"])
@pytest.mark.parametrize("shape", [(500, [40, 21], 5, 5, 1), (300, [20, 100], 5, 3, 2), (150, [12, 7], 2, 2, 0)])
def test_fwd_bwd_activations_ffma(rb, ctx, act, shape):
    """The FFMA kernel (`k1_small`, the path of byte-tile-only stores and of K1_FFMA) with every activation; the launch
    must really be k1_small's, not the shape-agnostic kernel's."""
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, "ridge_ard", n, gs, h, s, depth=d, act=act, seed=6)
    try:
        P.net.select_k1(P.net.K1_FFMA)
        for b in range(len(gs)):
            got = P.net.branch_fwd_bwd(b)
            assert "k1_small" in P.net.last_k1_kernel()
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(got["yhat"], t64["yhat"], t32["yhat"])
            within(got["rss"], t64["rss"], t32["rss"])
            within(got["d_rss"], t64["d_rss"], t32["d_rss"])
            within(got["ldg"], t64["ldg"], t32["ldg"])
    finally:
        P.close()


@pytest.mark.parametrize("model", ["ridge_ard", "lasso_base", "std_normal"])
def test_numerical_ldg(rb, ctx, model):
    """numerical_ldg (branch_sampler.rs:480-504): forward differences with delta = 0.001.  In f32 a difference quotient of a
    log density of size |ld| carries ~|ld| * 2^-23 / delta of rounding noise, in the reference as much as here: the bar is the
    f64 oracle within that noise (plus the usual mimic-vs-truth term); the analytical gradient must lie inside the same band
    plus the O(delta) truncation error."""
    P = Problem(rb, ctx, model, 400, [12, 9], 4, 3, seed=17)
    try:
        for b in range(2):
            got = P.net.branch_numerical_ldg(b)
            pv_after, _ = P.net.get_branch(b)
            assert np.array_equal(pv_after, P.cfgs[b].param_vec())                 # parameters reloaded (:502)
            x64 = P.x(b, np.float64)
            br64 = Branch(P.cfgs[b], np.float64)
            ld = abs(float(br64.log_density(br64.rss(x64, P.y.astype(np.float64)))))
            t64 = br64.numerical_ldg(x64, P.y.astype(np.float64))
            t32 = Branch(P.cfgs[b], np.float32).numerical_ldg(P.x(b, np.float32), P.y)
            noise = 8 * ld * 2.0 ** -23 / 1e-3
            assert np.all(np.abs(got - t64) <= 8 * np.abs(t32 - t64) + noise + 2e-5 * np.max(np.abs(t64))), \
                (np.max(np.abs(got - t64)), noise)
            ana = P.net.branch_fwd_bwd(b)["ldg"]
            assert np.max(np.abs(ana - got)) <= noise + 0.02 * np.max(np.abs(ana)) + 0.5
    finally:
        P.close()


def test_num_grad_transition_follows_the_analytical_one(rb, ctx):
    """hmc_step with mcmc_cfg.num_grad (branch_sampler.rs:1232-1247): same momenta, numerical instead of analytical gradients --
    the trajectory stays close to the analytical one (short trajectory, small steps) and the decision is the same."""
    P = Problem(rb, ctx, "ridge_base", 300, [10], 3, 2, seed=23)
    try:
        rng = np.random.default_rng(3)
        mom = rng.standard_normal(P.cfgs[0].num_params).astype(np.float32)
        outs = []
        for ng in (False, True):
            P.net.set_branch(0, P.cfgs[0].param_vec(), P.cfgs[0].precision_vec())
            cfg = rb.MCMCCfg(hmc_step_size_factor=0.05, hmc_integration_length=4, num_grad=ng)
            res = P.net.hmc_step(0, cfg, momenta=mom, u=0.0)
            outs.append((res, P.net.get_branch(0)[0].copy()))
        assert outs[0][0].status == outs[1][0].status == rb.HMC_ACCEPTED
        assert abs(outs[0][0].neg_h_final - outs[1][0].neg_h_final) < 2e-2 * abs(outs[0][0].neg_h_final) + 0.5
        assert np.allclose(outs[0][1], outs[1][1], rtol=0, atol=5e-3)
        assert not np.array_equal(outs[0][1], outs[1][1])      # the numerical gradient really was used
    finally:
        P.close()


def test_fixture_branch_golden_values(rb, ctx):
    """The reference's micro-fixture through the CUDA path: raw X is not
    expressible as packed genotypes unless means=0/stds=1, which is exactly the fixture."""
    X = np.array([1, 0, 0, 2, 1, 1, 2, 0, 0, 2, 0, 1]).reshape(4, 3, order="F")
    payload = obed.pack_columns(X)
    gen = rb.Genotypes(ctx, payload, 4, 3, [[0, 1, 2]], col_means=np.zeros(3), col_stds=np.ones(3))
    W = [np.arange(6, dtype=np.float32).reshape(3, 2, order="F"), np.array([[1.], [2.]]), np.array([[2.]])]
    Bs = [np.array([0., 1.]), np.array([2.])]
    y = np.array([0.0, 2.0, 1.0, 1.5], dtype=np.float32)
    exp = {  # ridge_ard.rs:681-700 / lasso_ard.rs (precision 1.0)
        "ridge_ard": ([-0.0005189283, -1.0005465, -2.0000138, -3.0, -4.0, -5.0, -1.0014552, -2.0017552, -5.4986963],
                      [-0.00053271546, -1.2088213e-9, -0.0017552058]),
        "lasso_ard": ([-0.0005189283, -1.0005465, -1.0000138, -1.0, -1.0, -1.0, -1.0014552, -1.0017552, -4.4986963],
                      [-0.00053271546, -1.2088213e-9, -0.0017552058]),
    }
    for model, (ew, eb) in exp.items():
        cfg = make_cfg(model, 3, [2], 1, weights=W, biases=Bs, precision=1.0)
        net = rb.Net(ctx, gen, model, [cfg.layer_widths])
        net.set_branch(0, cfg.param_vec(), cfg.precision_vec())
        net.set_targets(y)
        got = net.branch_fwd_bwd(0)
        assert abs(got["rss"] - 5.248245) < 3e-6                                     # ridge_ard.rs:537
        assert np.allclose(got["yhat"], [1.9997075, 1.9998183, 1.9998049, 1.9998049], atol=3e-6)   # :493
        e = np.array(ew + eb)
        big = np.abs(e) > 1e-2
        assert np.allclose(got["ldg"][big], e[big], rtol=3e-6)
        assert np.allclose(got["ldg"][~big], e[~big], rtol=5e-3, atol=1e-8)          # Q16: tanh near saturation
        net.close()
    gen.close()


# ------------------------------------------------------------------ step sizes
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode", ["uniform", "random", "std_scaled", "izmailov"])
def test_step_sizes(rb, ctx, model, mode):
    P = Problem(rb, ctx, model, 64, [9], 3, 2, seed=3)
    try:
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.7, hmc_integration_length=13, hmc_step_size_mode=mode)
        ocfg = OCfg(hmc_step_size_factor=0.7, hmc_integration_length=13, hmc_step_size_mode=mode)
        u = np.random.default_rng(1).random(P.cfgs[0].num_params, dtype=np.float32)
        br = Branch(P.cfgs[0], np.float32)
        if mode == "std_scaled" and model.endswith("ard"):
            with pytest.raises(rb.BannError):
                P.net.branch_step_sizes(0, cfg)
            with pytest.raises(IndexError):
                br.step_sizes(ocfg)
            return
        ws, bs = br.step_sizes(ocfg, u)
        exp = Branch.join_vec(ws, bs)
        got = P.net.branch_step_sizes(0, cfg, step_uniforms=u)
        assert np.allclose(got, exp, rtol=3e-7, atol=0), np.max(np.abs(got - exp) / np.abs(exp))
    finally:
        P.close()


def test_izmailov_infinite_bias_precision_gives_zero_step(rb, ctx):
    # Q7: default init has zero biases -> ML bias precision +inf -> bias step size 0
    P = Problem(rb, ctx, "std_normal", 64, [9], 3, 2, seed=3)
    try:
        c = P.cfgs[0]
        c.bias_precisions = [np.array([np.inf], dtype=np.float32) for _ in c.bias_precisions]
        P.net.set_branch(0, c.param_vec(), c.precision_vec())
        got = P.net.branch_step_sizes(0, rb.MCMCCfg())
        nw = sum(w.size for w in c.weights)
        assert np.all(got[nw:] == 0.0) and np.all(got[:nw] > 0)
    finally:
        P.close()


# ------------------------------------------------------------------ HMC transition
def run_oracle_hmc(P, b, target, ocfg, mom, u, su, dt):
    br = Branch(P.cfgs[b], dt)
    res = br.hmc_step(P.x(b, dt), np.asarray(target, dtype=dt), ocfg, mom, u, step_uniforms=su, record=True)
    res["params_after"] = br.param_vec()
    return res


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("mode,factor,L", [("izmailov", 1.0, 12), ("uniform", 0.002, 10), ("random", 0.01, 8),
                                           ("uniform", 0.35, 20)])
def test_hmc_step_trajectory_and_decision(rb, ctx, model, mode, factor, L):
    P = Problem(rb, ctx, model, 600, [30, 11], 5, 5, seed=11)
    near_ties = 0
    try:
        rng = np.random.default_rng(99)
        for b in range(2):
            for trial in range(3):
                cfg = rb.MCMCCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode=mode,
                                 hmc_max_hamiltonian_error=10.0)
                ocfg = OCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode=mode,
                            hmc_max_hamiltonian_error=10.0)
                Pn = P.cfgs[b].num_params
                mom = rng.standard_normal(Pn).astype(np.float32)
                u = float(np.float32(rng.random(dtype=np.float32)))
                su = rng.random(Pn, dtype=np.float32)
                P.net.set_branch(b, P.cfgs[b].param_vec(), P.cfgs[b].precision_vec())
                got = P.net.hmc_step(b, cfg, momenta=mom, u=u, step_uniforms=su, trajectory=True)
                o64 = run_oracle_hmc(P, b, P.y, ocfg, mom, u, su, np.float64)
                o32 = run_oracle_hmc(P, b, P.y, ocfg, mom, u, su, np.float32)
                # Hamiltonian at the start: plain FP32 tolerance
                within(got.neg_h_init, o64["h_init"], o32["h_init"], scale=abs(o64["h_init"]))
                nsteps = min(got.steps_done, o64["steps_done"], o32["steps_done"])
                assert nsteps >= 1
                # trajectories: compare while all three agree on being alive; chaotic growth of
                # rounding differences is bounded by comparing against the f32 mimic's own drift
                for s in range(min(nsteps, 6)):
                    within(got.trajectory["params"][s], o64["traj"]["params"][s], o32["traj"]["params"][s],
                           rel=5e-5)
                    within(got.trajectory["ldg"][s], o64["traj"]["ldg"][s], o32["traj"]["ldg"][s], rel=5e-5)
                    within(got.trajectory["hamiltonian"][s + 1], o64["traj"]["hamiltonian"][s + 1],
                           o32["traj"]["hamiltonian"][s + 1],
                           scale=max(abs(o64["h_init"]), abs(o64["traj"]["hamiltonian"][s + 1])), rel=5e-5)
                # decisions: identical unless the f64 truth itself is within tolerance of a boundary
                margin = 1e-3 * max(1.0, abs(o64["h_init"]) * 1e-3)
                if o64["status"] == REJECTED_EARLY or got.status == rb.HMC_REJECTED_EARLY:
                    hs = np.array(o64["traj"]["hamiltonian"])
                    dist = np.min(np.abs(np.abs(hs - hs[0]) - 10.0))
                    if dist < margin:
                        near_ties += 1
                        continue
                    assert got.status == o64["status"] and got.steps_done == o64["steps_done"]
                    continue
                la = o64["log_acc"]
                if abs(min(la, 0.0) - np.log(max(u, 1e-30))) < margin:
                    near_ties += 1
                    continue
                assert got.status == o64["status"], (got.status, o64["status"], la, u)
                pv, _ = P.net.get_branch(b)
                within(pv, o64["params_after"], o32["params_after"], rel=1e-4)
                if got.status == rb.HMC_ACCEPTED:
                    within(got.y_pred, o64["y_pred"], o32["y_pred"], rel=1e-4)
                    within(got.log_density, o64["log_density"], o32["log_density"], scale=abs(o64["h_init"]), rel=5e-5)
        assert near_ties <= 2
    finally:
        P.close()


def test_hmc_early_rejection_restores_params(rb, ctx):
    P = Problem(rb, ctx, "ridge_base", 400, [25], 5, 5, seed=4)
    try:
        cfg = rb.MCMCCfg(hmc_step_size_factor=5.0, hmc_integration_length=30, hmc_step_size_mode="uniform")
        ocfg = OCfg(hmc_step_size_factor=5.0, hmc_integration_length=30, hmc_step_size_mode="uniform")
        mom = np.random.default_rng(0).standard_normal(P.cfgs[0].num_params).astype(np.float32)
        before = P.cfgs[0].param_vec().copy()
        got = P.net.hmc_step(0, cfg, momenta=mom, u=0.5)
        o = run_oracle_hmc(P, 0, P.y, ocfg, mom, 0.5, None, np.float32)
        assert o["status"] == REJECTED_EARLY and got.status == rb.HMC_REJECTED_EARLY
        assert got.steps_done == o["steps_done"]
        pv, _ = P.net.get_branch(0)
        assert np.array_equal(pv, before.astype(np.float32))
    finally:
        P.close()


# ------------------------------------------------------------------ Gibbs draws
@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("fixed", [False, True])
def test_gibbs_precisions(rb, ctx, model, fixed):
    P = Problem(rb, ctx, model, 300, [12, 8], 4, 3, depth=2, seed=21)
    try:
        b = 1
        rng = np.random.default_rng(5)
        glob = dict(error_precision=1.7, output_layer_precision=0.4, ow_reg_sum=3.25, ow_num_params=6.0)
        # make the global statistic consistent: others + own
        own = float(Branch(P.cfgs[b], np.float32).summary_stat(P.cfgs[b].weights[-1]))
        glob["ow_reg_sum"] = 2.0 + own
        P.net.set_globals(**glob)
        r = rng.normal(0, 1.3, size=P.n).astype(np.float32)
        P.net.set_residual(r)
        cfg_o = P.cfgs[b]
        cfg_o.error_precision = glob["error_precision"]
        cfg_o.weight_precisions[-1] = np.array([glob["output_layer_precision"]], dtype=np.float32)
        cfg_o.ow_reg_sum, cfg_o.ow_num_params = glob["ow_reg_sum"], int(glob["ow_num_params"])
        res = {}
        for dt in (np.float32, np.float64):
            br = Branch(cfg_o, dt)
            shapes = br.gibbs_shapes(P.hyper, P.n, fixed)
            gam = [np.float32(np.random.default_rng(77 + i).standard_gamma(s)) for i, s in enumerate(shapes)]
            it = iter(gam)
            draw = lambda shape: next(it)
            br.sample_error_precision(r.astype(dt), P.hyper, draw)
            if not fixed:
                br.sample_param_precisions(P.hyper, draw)
            res[dt] = br.to_cfg().precision_vec()
        P.net.gibbs_branch(b, rb.MCMCCfg(fixed_param_precisions=fixed), std_gammas=np.array(gam, dtype=np.float32))
        _, qv = P.net.get_branch(b)
        within(qv, res[np.float64], res[np.float32], rel=2e-6)
    finally:
        P.close()


# ------------------------------------------------------------------ chain bookkeeping (Net::train)
def mirror_net(P):
    net = onet.Net(model=P.model, hyper=P.hyper, cfgs=[c.astype(np.float32) for c in P.cfgs],
                   groups=P.groups, output_bias=0.0, g_error_precision=2.0, g_output_layer_precision=0.05,
                   g_ow_reg_sum=0.0, g_ow_num_params=0, lpd_local=np.full(len(P.cfgs), -np.inf, dtype=np.float32))
    reg = np.float32(0)
    for c in net.cfgs:
        reg = np.float32(reg + np.float32(Branch(c, np.float32).summary_stat(c.weights[-1])))
        net.g_ow_num_params += c.layer_widths[-2]
    net.g_ow_reg_sum = float(reg)
    return net


@pytest.mark.parametrize("model", ["ridge_ard", "lasso_base", "std_normal", "ridge_base", "lasso_ard"])
def test_train_visits_match_oracle(rb, ctx, model):
    P = Problem(rb, ctx, model, 500, [20, 15, 9], 4, 3, seed=31)
    try:
        onet_ = mirror_net(P)
        P.net.set_globals(2.0, 0.05, onet_.g_ow_reg_sum, onet_.g_ow_num_params, 0.0)
        ocfg = OCfg(hmc_step_size_factor=0.5, hmc_integration_length=6, hmc_step_size_mode="izmailov")
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.5, hmc_integration_length=6, hmc_step_size_mode="izmailov")
        draws = onet.Draws(seed=123)
        resid_o = onet.initialize_stats(onet_, P.payload, P.n, P.means, P.stds, P.y, np.float32)
        P.net.init_residual()
        assert np.allclose(P.net.residual(), resid_o, rtol=0, atol=2e-5)
        st = P.net.stats()
        assert abs(st["lpd"] - onet.lpd_value(onet_)) < 2e-4 * abs(onet.lpd_value(onet_))
        xs = [P.x(b, np.float32) for b in range(3)]
        flips = 0
        for it in range(3):
            for b in draws.order(3):
                b = int(b)
                resid_o, res_o = onet.visit_branch(onet_, b, xs[b], resid_o, ocfg, draws)
                d = draws.log[-1]
                got = P.net.visit_branch(b, cfg, momenta=d["momenta"], u=d["u"], std_gammas=np.array(d["gammas"], dtype=np.float32))
                if got.status != res_o["status"]:
                    flips += 1      # near-tie in f32: resynchronise the device from the oracle
                    for bb, c in enumerate(onet_.cfgs):
                        P.net.set_branch(bb, c.param_vec(), c.precision_vec())
                    P.net.set_residual(resid_o)
                    P.net.set_globals(onet_.g_error_precision, onet_.g_output_layer_precision, onet_.g_ow_reg_sum,
                                      onet_.g_ow_num_params, onet_.output_bias)
                    continue
                pv, qv = P.net.get_branch(b)
                assert np.allclose(pv, onet_.cfgs[b].param_vec(), rtol=2e-4, atol=2e-5)
                assert np.allclose(qv, onet_.cfgs[b].precision_vec(), rtol=2e-4, atol=1e-6)
                assert np.allclose(P.net.residual(), resid_o, rtol=0, atol=5e-4)
                g = P.net.get_globals()
                assert abs(g["output_bias"] - onet_.output_bias) < 1e-5
                assert abs(g["ow_reg_sum"] - onet_.g_ow_reg_sum) < 1e-4 * max(1.0, onet_.g_ow_reg_sum)
                assert abs(g["error_precision"] - onet_.g_error_precision) < 1e-4 * onet_.g_error_precision
        assert flips <= 1
        st = P.net.stats()
        if flips == 0:
            assert st["num_samples"] == onet_.num_samples and st["num_accepted"] == onet_.num_accepted
            assert st["num_early_rejected"] == onet_.num_early_rejected
            lo = onet.lpd_value(onet_)
            assert abs(st["lpd"] - lo) < 5e-4 * abs(lo), (st["lpd"], lo)
        r = P.net.residual()
        assert abs(st["mse_train"] - float(np.dot(r, r) / P.n)) < 1e-4
    finally:
        P.close()


def test_predict_train_and_test_store(rb, ctx):
    P = Problem(rb, ctx, "ridge_ard", 333, [17, 6, 30], 5, 5, seed=8)
    try:
        onet_ = mirror_net(P)
        onet_.output_bias = 0.37
        P.net.set_globals(2.0, 0.05, 0.0, 0, 0.37)
        exp = onet.predict(onet_, P.payload, P.n, P.means, P.stds, np.float32)
        exp64 = onet.predict(onet_, P.payload, P.n, P.means, P.stds, np.float64)
        within(P.net.predict(), exp64, exp)
        # a separate test store with its own statistics (io/bed.rs:193-245 computes them per file)
        g2 = obed.random_genotypes(91, P.m, seed=1234)
        pl2 = obed.pack_columns(g2)
        mu2, sd2 = obed.col_stats(pl2, 91, P.m)
        test = rb.Genotypes(ctx, pl2, 91, P.m, P.groups)
        exp2 = onet.predict(onet_, pl2, 91, mu2, sd2, np.float32)
        exp2_64 = onet.predict(onet_, pl2, 91, mu2, sd2, np.float64)
        within(P.net.predict(test), exp2_64, exp2)
        test.close()
    finally:
        P.close()


def test_net_gradient_pinned_and_pageable_host_buffers_agree(rb, ctx):
    """bann_net_gradient copies page-locked caller buffers by DMA directly and stages pageable ones: same results,
    heterogeneous branch sizes (dense host layout vs the 16-byte aligned device arena)."""
    P = Problem(rb, ctx, "ridge_ard", 700, [13, 50, 7, 64], 5, 5, seed=23)
    try:
        pv = np.concatenate([c.param_vec() for c in P.cfgs]).astype(np.float32)
        g0, r0 = P.net.gradient(pv, P.y)
        pv_h, y_h = rb.pinned_empty(pv.size), rb.pinned_empty(P.n)
        pv_h[:], y_h[:] = pv, P.y
        out = (rb.pinned_empty(pv.size), rb.pinned_empty(len(P.cfgs)))
        g1, r1 = P.net.gradient(pv_h, y_h, out=out)
        assert g1 is out[0] and np.array_equal(g0, g1) and np.array_equal(r0, r1)
        # the per-branch entry point (per-row outputs: k1_tc) takes the cross-row sums in another fixed order than the gradient
        # launch (k1_tc5): FP32 rounding apart by default, bit-identical when the gradient launch is told to use k1_tc as well
        P.net.select_k1_tc_variant(P.net.TC_FOUR_WARPS)
        g4, r4 = P.net.gradient(pv_h, y_h)
        off = 0
        for b, c in enumerate(P.cfgs):
            one = P.net.branch_fwd_bwd(b)
            assert np.array_equal(one["ldg"], g4[off:off + c.num_params]) and one["rss"] == r4[b]
            sc = np.max(np.abs(one["ldg"]))
            assert np.max(np.abs(one["ldg"] - g1[off:off + c.num_params])) <= 2e-5 * sc and abs(one["rss"] - r1[b]) <= 2e-5 * r1[b]
            off += c.num_params
    finally:
        P.close()


# ------------------------------------------------------------------ full network (grouped schedule)
def test_net_gradient_matches_per_branch(rb, ctx):
    P = Problem(rb, ctx, "ridge_ard", 900, [50, 50, 50, 13, 50], 5, 5, seed=2)
    try:
        grads, rss = P.net.gradient()
        k = 0
        for b in range(5):
            one = P.net.branch_fwd_bwd(b)
            n = P.net.num_branch_params(b)
            # two kernels (gradient launch: k1_tc5, per-branch entry point with per-row outputs: k1_tc), two fixed summation orders
            assert np.max(np.abs(grads[k:k + n] - one["ldg"])) <= 4e-5 * np.max(np.abs(one["ldg"]))
            assert abs(rss[b] - one["rss"]) < 2e-5 * one["rss"]
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(grads[k:k + n], t64["ldg"], t32["ldg"])
            k += n
        # host-provided parameters and targets (the e2e form)
        pv, _ = P.net.get_all_params()
        y2 = (P.y * 2).astype(np.float32)
        g2, r2 = P.net.gradient(pv * 0.5, y2)
        P.net.set_all_params(pv * 0.5)
        one = P.net.branch_fwd_bwd(3, target=y2)
        off = sum(P.net.num_branch_params(b) for b in range(3))
        assert np.max(np.abs(g2[off:off + P.net.num_branch_params(3)] - one["ldg"])) <= 4e-5 * np.max(np.abs(one["ldg"]))
    finally:
        P.close()


@pytest.mark.parametrize("model,sizes", [("lasso_ard", [64, 57, 8, 1]), ("std_normal", [7, 56, 49]),
                                         ("ridge_ard", [100, 200, 65, 333])])     # last: K-blocked kernel, 2..6 blocks
def test_tensor_core_grouped_launch_heterogeneous_branches(rb, ctx, model, sizes):
    """One tensor-core launch over branches with different chunk counts (8-chunk operand buffers when a branch has 64 markers)."""
    P = Problem(rb, ctx, model, 1111, sizes, 5, 5, seed=11)
    try:
        P.net.select_k1(P.net.K1_TENSOR)
        grads, rss = P.net.gradient()
        k = 0
        for b in range(len(sizes)):
            n = P.net.num_branch_params(b)
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(grads[k:k + n], t64["ldg"], t32["ldg"])
            within(rss[b], t64["rss"], t32["rss"])
            k += n
    finally:
        P.close()


def test_grouped_leapfrog_conserves_energy_and_matches_kernels(rb, ctx):
    cfg = rb.MCMCCfg(hmc_step_size_factor=0.05, hmc_integration_length=20, hmc_step_size_mode="izmailov")
    out = {}
    for generic in (3, 2, 0):                            # K1_GENERIC, K1_FFMA, K1_AUTO (tensor-core kernel here)
        # fresh net per kernel: the built-in Philox stream is keyed by (seed, visit counter, branch)
        P = Problem(rb, ctx, "ridge_ard", 1200, [50] * 6, 5, 5, seed=6)
        try:
            P.net.select_k1(generic)
            P.net.grouped_begin(cfg, seed=7, per_branch_targets=False)
            P.net.grouped_leapfrog(cfg, 20, finalize=True)
            hi, hc, st = P.net.grouped_state()
            assert np.all(st == 3)                       # still running: no early rejection
            assert np.all(np.abs(hc - hi) < 0.5), np.abs(hc - hi).max()
            acc, early = P.net.grouped_finish(seed=7)
            assert early == 0 and acc >= 4
            out[generic] = (hi.copy(), hc.copy(), P.net.get_all_params()[0])
        finally:
            P.close()
    for other in (2, 0):
        assert np.allclose(out[3][0], out[other][0], rtol=1e-5)
        assert np.allclose(out[3][1], out[other][1], rtol=1e-4)
        assert np.allclose(out[3][2], out[other][2], rtol=1e-3, atol=1e-4)


def test_synthetic_store_is_independent_of_row_sharding(rb, ctx):
    n, m, groups = 1000, 60, obed.uniform_grouping(3, 20)
    full = rb.Genotypes.random(ctx, n, m, groups, seed=5)
    lo = rb.Genotypes.random(ctx, 512, m, groups, seed=5, row_offset=0, n_total=n)
    hi = rb.Genotypes.random(ctx, n - 512, m, groups, seed=5, row_offset=512, n_total=n)
    for b in range(3):
        x = full.x_group(b, standardized=False)
        assert set(np.unique(x)) <= {0.0, 1.0, 2.0}
        assert np.array_equal(x[:512], lo.x_group(b, False)) and np.array_equal(x[512:], hi.x_group(b, False))
    cnt = lo.col_counts() + hi.col_counts()
    assert np.array_equal(cnt, full.col_counts())
    mu, sd = rb.stats_from_counts(cnt, n)
    fmu, fsd = full.col_stats()          # sequential f32 statistics of the full store
    assert np.array_equal(mu, fmu) and np.allclose(sd, fsd, rtol=2e-5)
    # allele frequencies look like U(0.01, 0.5) draws
    assert 0.05 < fmu.mean() / 2 < 0.45
    for g in (full, lo, hi):
        g.close()


def test_grouped_per_branch_targets_equal_residual_plus_prediction(rb, ctx):
    P = Problem(rb, ctx, "lasso_ard", 640, [30, 30, 30], 5, 5, seed=16)
    try:
        P.net.set_globals(2.0, 0.05, 1.0, 15, 0.0)
        P.net.init_residual()
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.02, hmc_integration_length=5)
        P.net.grouped_begin(cfg, seed=1, per_branch_targets=True)
        hi, hc, st = P.net.grouped_state()
        # rss against t_b = r + yhat_b is |r|^2 for every branch (net.rs:279-280), so the initial
        # log densities differ only by the prior term
        r = P.net.residual().astype(np.float64)
        for b in range(3):
            one = P.net.branch_fwd_bwd(b, target=(P.net.residual() + P.net.branch_fwd_bwd(b)["yhat"]))
            assert abs(one["rss"] - float(r @ r)) < 1e-4 * float(r @ r)
        P.net.grouped_leapfrog(cfg, 5, finalize=True)
        hi2, hc2, st2 = P.net.grouped_state()
        assert np.all(np.isfinite(hc2)) and np.allclose(hi, hi2)
    finally:
        P.close()
