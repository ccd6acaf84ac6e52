"""k1_tc5 (csrc/k1_tc5.cuh: the <= 64-marker tensor-core kernel with a dedicated issuing warp and one-float cross-row sums)
against the oracle and against k1_tc, through the calls that take the LEAN launch it exists for: Net.gradient
(net/net.rs:520-527, branch_sampler.rs:743-782,813-875) and the launch-per-step HMC transition (branch_sampler.rs:1192-1299)."""
import numpy as np
import pytest

from oracle.branch import MCMCCfg as OCfg, REJECTED_EARLY

from test_gpu_parity import Problem, oracle_fwd_bwd, run_oracle_hmc, within

pytestmark = pytest.mark.gpu


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


# n, group sizes: [5,5,1] (the architecture k1_tc5 is instantiated for); 49..56 markers in every branch takes the NCT = 7 kernel
SHAPES = [
    (7, [5]),
    (255, [1, 8, 9]),
    (256, [50, 64]),
    (1030, [50, 50, 50, 33]),          # several super-tiles, ragged tail, overlapping groups, mixed chunk counts
    (1300, [50, 56, 49]),              # 7 chunks everywhere: the specialised kernel
    (2600, [64, 17]),
]


VARIANTS = [1, 2]       # bann_net_select_k1_tc_variant: TC_FIVE_WARPS (cross-row sums deferred, the default), TC_FIVE_WARPS_PLAIN


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("model", ["ridge_ard", "std_normal", "lasso_base"])
@pytest.mark.parametrize("shape", SHAPES)
def test_tc5_gradient_matches_oracle_and_k1_tc(rb, ctx, model, shape, variant):
    n, gs = shape
    P = Problem(rb, ctx, model, n, gs, 5, 5, seed=(sum(map(ord, model)) + 3 * n) % 1000, overlap=len(gs) > 1)
    try:
        net = P.net
        net.select_k1(net.K1_TENSOR)
        net.select_k1_tc_variant(variant)
        g5, r5 = net.gradient(y=P.y)
        assert "k1_tc5" in net.last_k1_kernel(), net.last_k1_kernel()
        g5b, r5b = net.gradient(y=P.y)
        assert np.array_equal(g5, g5b) and np.array_equal(r5, r5b)          # fixed-order sums: run-to-run identical
        net.select_k1_tc_variant(net.TC_FOUR_WARPS)
        g4, r4 = net.gradient(y=P.y)
        assert "k1_tc<" in net.last_k1_kernel()
        off = 0
        for b in range(len(gs)):
            Pn = P.cfgs[b].num_params
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(g5[off:off + Pn], t64["ldg"], t32["ldg"])
            within(r5[b], t64["rss"], t32["rss"])
            sc = np.max(np.abs(t64["ldg"]))
            assert np.max(np.abs(g5[off:off + Pn].astype(np.float64) - g4[off:off + Pn])) <= 4e-5 * sc
            assert abs(float(r5[b]) - float(r4[b])) <= 2e-5 * float(r4[b])
            off += Pn
    finally:
        P.close()


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "silu", "identity"])
def test_tc5_other_activations(rb, ctx, act):
    P = Problem(rb, ctx, "ridge_ard", 900, [40, 24, 56], 5, 5, act=act, seed=17)
    try:
        net = P.net
        net.select_k1(net.K1_TENSOR)
        net.select_k1_tc_variant(net.TC_FIVE_WARPS)
        g5, r5 = net.gradient(y=P.y)
        assert "k1_tc5" in net.last_k1_kernel()
        off = 0
        for b in range(3):
            Pn = P.cfgs[b].num_params
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(g5[off:off + Pn], t64["ldg"], t32["ldg"])
            within(r5[b], t64["rss"], t32["rss"])
            off += Pn
    finally:
        P.close()


# the other architectures k1_tc5 is instantiated for (at most one hidden layer): n, group sizes, hidden, summary, depth
ARCHS = [
    (515, [20, 33], 4, 3, 1),
    (700, [40, 24, 7], 2, 2, 1),
    (600, [20, 64], 3, 3, 1),
    (300, [33, 8], 4, 4, 1),
    (150, [12, 7], 2, 2, 0),          # no hidden layer: the summary layer reads the markers
    (900, [50, 56], 5, 5, 0),
]


@pytest.mark.parametrize("act", ["tanh", "silu"])
@pytest.mark.parametrize("arch", ARCHS)
def test_tc5_other_architectures(rb, ctx, arch, act):
    n, gs, h, s_, d = arch
    P = Problem(rb, ctx, "ridge_ard", n, gs, h, s_, depth=d, act=act, seed=n % 89, overlap=True)
    try:
        net = P.net
        net.select_k1(net.K1_TENSOR)
        net.select_k1_tc_variant(net.TC_FIVE_WARPS)
        g5, r5 = net.gradient(y=P.y)
        assert "k1_tc5" in net.last_k1_kernel(), net.last_k1_kernel()
        net.select_k1_tc_variant(net.TC_FOUR_WARPS)
        g4, r4 = net.gradient(y=P.y)
        off = 0
        for b in range(len(gs)):
            Pn = P.cfgs[b].num_params
            t64, t32 = oracle_fwd_bwd(P, b, P.y, np.float64), oracle_fwd_bwd(P, b, P.y, np.float32)
            within(g5[off:off + Pn], t64["ldg"], t32["ldg"])
            within(r5[b], t64["rss"], t32["rss"])
            sc = np.max(np.abs(t64["ldg"]))
            assert np.max(np.abs(g5[off:off + Pn].astype(np.float64) - g4[off:off + Pn])) <= 4e-5 * sc
            off += Pn
    finally:
        P.close()


def test_tc5_two_hidden_layers_keep_k1_tc(rb, ctx):
    P = Problem(rb, ctx, "ridge_ard", 515, [50, 64], 5, 5, depth=2, seed=5)
    try:
        net = P.net
        net.select_k1(net.K1_TENSOR)
        net.select_k1_tc_variant(net.TC_FIVE_WARPS)
        g, r = net.gradient(y=P.y)
        assert "k1_tc<" in net.last_k1_kernel()         # [5,5,5,1]: not instantiated for five warps, the four-warp kernel runs
        t64, t32 = oracle_fwd_bwd(P, 0, P.y, np.float64), oracle_fwd_bwd(P, 0, P.y, np.float32)
        within(g[:P.cfgs[0].num_params], t64["ldg"], t32["ldg"])
    finally:
        P.close()


@pytest.mark.parametrize("mode,factor,L", [("izmailov", 1.0, 12), ("uniform", 0.002, 10), ("uniform", 0.35, 20)])
def test_tc5_launch_per_step_transition(rb, ctx, mode, factor, L):
    """hmc_step on the launch-per-step path (K1 -> chunk reduction -> K2 per leapfrog) with k1_tc5 as K1: same injected momenta
    and uniforms as the oracle, decisions identical, Hamiltonians and parameters within FP32 tolerance; the large step of the
    last mode rejects early and must restore the parameters (branch_sampler.rs:1277)."""
    P = Problem(rb, ctx, "ridge_ard", 1300, [50, 31], 5, 5, seed=23)
    near_ties = 0
    try:
        net = P.net
        net.select_k1(net.K1_TENSOR)
        net.select_k1_tc_variant(net.TC_FIVE_WARPS)
        net.select_hmc_path(net.HMC_LAUNCHES)
        rng = np.random.default_rng(3)
        for b in range(2):
            cfg = rb.MCMCCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode=mode)
            ocfg = OCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode=mode)
            Pn = P.cfgs[b].num_params
            mom = rng.standard_normal(Pn).astype(np.float32)
            u = float(np.float32(rng.random(dtype=np.float32)))
            su = rng.random(Pn, dtype=np.float32)
            net.set_branch(b, P.cfgs[b].param_vec(), P.cfgs[b].precision_vec())
            got = net.hmc_step(b, cfg, momenta=mom, u=u, step_uniforms=su)    # (its last K1 launch is the prediction: k1_tc)
            after = net.get_branch(b)[0].copy()
            o64 = run_oracle_hmc(P, b, P.y, ocfg, mom, u, su, np.float64)
            o32 = run_oracle_hmc(P, b, P.y, ocfg, mom, u, su, np.float32)
            within(got.neg_h_init, o64["h_init"], o32["h_init"], scale=abs(o64["h_init"]))
            margin = 1e-3 * max(1.0, abs(o64["h_init"]) * 1e-3)
            if o64["status"] == REJECTED_EARLY or got.status == rb.HMC_REJECTED_EARLY:
                hs = np.array(o64["traj"]["hamiltonian"])
                if np.min(np.abs(np.abs(hs - hs[0]) - 10.0)) < margin:
                    near_ties += 1
                    continue
                assert got.status == o64["status"] and got.steps_done == o64["steps_done"]
                assert np.array_equal(after, P.cfgs[b].param_vec())
                continue
            if abs(min(o64["log_acc"], 0.0) - np.log(max(u, 1e-30))) < margin:
                near_ties += 1
                continue
            assert got.status == o64["status"] and got.steps_done == L
            within(got.neg_h_final, o64["h_final"], o32["h_final"], scale=max(abs(o64["h_init"]), abs(o64["h_final"])), rel=5e-5)
            within(after, o64["params_after"], o32["params_after"], rel=1e-4)
        assert near_ties <= 1
    finally:
        P.close()
