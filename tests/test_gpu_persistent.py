"""The persistent per-branch HMC kernel (csrc/k1_tcp.cuh: the whole L-step trajectory of BranchSampler::hmc_step,
branch_sampler.rs:1192-1299, in ONE cooperative launch with the branch's operands resident on chip) against the oracle and against
the launch-per-step path, same injected momenta and accept uniforms: Hamiltonians, final parameters, predictions and decisions."""
import numpy as np
import pytest

from oracle import net as onet
from oracle.branch import MCMCCfg as OCfg, REJECTED_EARLY

from test_gpu_parity import MODELS, Problem, mirror_net, run_oracle_hmc, within

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


SHAPES = [  # n, group sizes, hidden, summary  (architectures the persistent kernel is instantiated for)
    (600, [30, 11], 5, 5),          # 3 super-tiles: one per CTA, ragged last one
    (255, [50], 5, 5),              # a single partial super-tile
    (1030, [64, 1], 2, 2),          # 64 markers (the M = 64 backward tile full), 1 marker
    (700, [20, 33], 4, 3),
    (515, [8, 40], 3, 3),
]


MODES = (("izmailov", 1.0, 12), ("uniform", 0.002, 10), ("random", 0.01, 8))


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("shape", SHAPES)
def test_persistent_transition_matches_oracle_and_launch_path(rb, ctx, model, shape):
    n, gs, h, s = shape
    P = Problem(rb, ctx, model, n, gs, h, s, seed=(sum(map(ord, model)) + n) % 997)
    _compare_paths(rb, P, gs, MODES)


@pytest.mark.parametrize("act", ["relu", "leaky_relu", "silu", "identity"])
@pytest.mark.parametrize("shape", [(600, [30, 11], 5, 5, 1), (515, [8, 40], 5, 3, 2), (300, [12, 7], 2, 2, 0), (400, [20], 5, 5, 2)])
def test_persistent_other_activations_and_depths(rb, ctx, act, shape):
    """activation_functions.rs:23-45 and the depth-0 / depth-2 architectures through the persistent kernel."""
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, "ridge_ard", n, gs, h, s, depth=d, act=act, seed=n % 97)
    _compare_paths(rb, P, gs, (("izmailov", 0.05, 8), ("uniform", 0.001, 10)))      # mild steps: identity / ReLU nets are unbounded


def _compare_paths(rb, P, gs, modes):
    near_ties = 0
    try:
        rng = np.random.default_rng(5)
        for b in range(len(gs)):
            for mode, factor, L in modes:
                cfg = rb.MCMCCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode=mode)
                ocfg = OCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode=mode)
                Pn = P.cfgs[b].num_params
                mom = rng.standard_normal(Pn).astype(np.float32)
                u = float(np.float32(rng.random(dtype=np.float32)))
                su = rng.random(Pn, dtype=np.float32)
                res = {}
                for path in (P.net.HMC_LAUNCHES, P.net.HMC_PERSISTENT):
                    P.net.select_hmc_path(path)          # PERSISTENT fails loudly if the branch is not eligible
                    P.net.set_branch(b, P.cfgs[b].param_vec(), P.cfgs[b].precision_vec())
                    before = P.net.persistent_launches()
                    got = P.net.hmc_step(b, cfg, momenta=mom, u=u, step_uniforms=su)
                    assert P.net.persistent_launches() - before == (1 if path == P.net.HMC_PERSISTENT else 0)
                    res[path] = (got, P.net.get_branch(b)[0].copy())
                gl, gp = res[P.net.HMC_LAUNCHES][0], res[P.net.HMC_PERSISTENT][0]
                o64 = run_oracle_hmc(P, b, P.y, ocfg, mom, u, su, np.float64)
                o32 = run_oracle_hmc(P, b, P.y, ocfg, mom, u, su, np.float32)
                within(gp.neg_h_init, o64["h_init"], o32["h_init"], scale=abs(o64["h_init"]))
                margin = 1e-3 * max(1.0, abs(o64["h_init"]) * 1e-3)
                if o64["status"] == REJECTED_EARLY or gp.status == rb.HMC_REJECTED_EARLY or gl.status == rb.HMC_REJECTED_EARLY:
                    hs = np.array(o64["traj"]["hamiltonian"])
                    if np.min(np.abs(np.abs(hs - hs[0]) - 10.0)) < margin:
                        near_ties += 1
                        continue
                    assert gp.status == gl.status == o64["status"] and gp.steps_done == gl.steps_done == o64["steps_done"]
                    assert np.array_equal(res[P.net.HMC_PERSISTENT][1], P.cfgs[b].param_vec())       # restored (:1277)
                    continue
                la = o64["log_acc"]
                if abs(min(la, 0.0) - np.log(max(u, 1e-30))) < margin:
                    near_ties += 1
                    continue
                assert gp.status == gl.status == o64["status"], (gp.status, gl.status, o64["status"], la, u)
                assert gp.steps_done == gl.steps_done == L and gp.u_turn_step == gl.u_turn_step
                within(gp.neg_h_final, o64["h_final"], o32["h_final"], scale=max(abs(o64["h_init"]), abs(o64["h_final"])), rel=5e-5)
                within(res[P.net.HMC_PERSISTENT][1], o64["params_after"], o32["params_after"], rel=1e-4)
                # the two GPU paths differ only in the order of the fixed-order reductions
                assert np.allclose(res[P.net.HMC_PERSISTENT][1], res[P.net.HMC_LAUNCHES][1], rtol=2e-4, atol=2e-5)
                assert abs(gp.neg_h_final - gl.neg_h_final) <= 2e-4 * max(1.0, abs(gl.neg_h_final))
                if gp.status == rb.HMC_ACCEPTED:
                    within(gp.y_pred, o64["y_pred"], o32["y_pred"], rel=1e-4)
                    within(gp.log_density, o64["log_density"], o32["log_density"], scale=abs(o64["h_init"]), rel=5e-5)
        assert near_ties <= 2
    finally:
        P.close()


def test_persistent_early_rejection_and_explicit_target(rb, ctx):
    P = Problem(rb, ctx, "ridge_ard", 600, [30], 5, 5, seed=3)
    try:
        P.net.select_hmc_path(P.net.HMC_PERSISTENT)
        mom = np.random.default_rng(1).standard_normal(P.cfgs[0].num_params).astype(np.float32) * 5
        cfg = rb.MCMCCfg(hmc_step_size_factor=3.0, hmc_integration_length=30, hmc_max_hamiltonian_error=0.5)
        res = P.net.hmc_step(0, cfg, momenta=mom, u=0.5)
        assert res.status == rb.HMC_REJECTED_EARLY and 1 <= res.steps_done < 30
        assert np.array_equal(P.net.get_branch(0)[0], P.cfgs[0].param_vec())
        # an explicit target vector (bann_hmc_step's `target`): same decision and Hamiltonians as the launch path
        tgt = (P.y * 0.5 + 0.2).astype(np.float32)
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.3, hmc_integration_length=9)
        outs = []
        for path in (P.net.HMC_LAUNCHES, P.net.HMC_PERSISTENT):
            P.net.select_hmc_path(path)
            P.net.set_branch(0, P.cfgs[0].param_vec(), P.cfgs[0].precision_vec())
            outs.append(P.net.hmc_step(0, cfg, target=tgt, momenta=mom / 5, u=0.0))
        assert outs[0].status == outs[1].status == rb.HMC_ACCEPTED
        assert abs(outs[0].neg_h_init - outs[1].neg_h_init) <= 2e-5 * abs(outs[0].neg_h_init)
        assert abs(outs[0].neg_h_final - outs[1].neg_h_final) <= 2e-4 * abs(outs[0].neg_h_final)
        assert np.allclose(outs[0].y_pred, outs[1].y_pred, rtol=0, atol=2e-4)
    finally:
        P.close()


def test_persistent_several_tiles_per_cta_and_ineligible_branch(rb, ctx):
    """100k rows = 391 super-tiles on 148 co-resident CTAs: every CTA keeps three super-tiles (the configuration of
    `rs-bann train-new` at biobank scale).  Checked against the launch path (the oracle would need minutes here).  A branch with
    more than 64 markers is not eligible: AUTO falls back, PERSISTENT fails loudly."""
    from oracle import bed as obed
    from oracle.branch import make_cfg
    n, sizes = 100_000, [50, 100]
    rng = np.random.default_rng(2)
    g = obed.random_genotypes(n, sum(sizes), seed=4)
    gen = rb.Genotypes(ctx, obed.pack_columns(g), n, sum(sizes), [list(range(50)), list(range(50, 150))])
    cfgs = [make_cfg("ridge_ard", m, [5], 5, rng=rng) for m in sizes]
    net = rb.Net(ctx, gen, "ridge_ard", [c.layer_widths for c in cfgs])
    try:
        for b, c in enumerate(cfgs):
            c.bias_precisions = [np.ones(1, dtype=np.float32) for _ in c.bias_precisions]
            net.set_branch(b, c.param_vec(), c.precision_vec())
        net.set_targets(rng.normal(size=n).astype(np.float32))
        mom = rng.standard_normal(cfgs[0].num_params).astype(np.float32)
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.01, hmc_integration_length=10, hmc_max_hamiltonian_error=1e30)   # 100k rows: a sharp posterior
        outs = []
        for path in (net.HMC_LAUNCHES, net.HMC_PERSISTENT):
            net.select_hmc_path(path)
            net.set_branch(0, cfgs[0].param_vec(), cfgs[0].precision_vec())
            outs.append((net.hmc_step(0, cfg, momenta=mom, u=0.0), net.get_branch(0)[0].copy()))
        (a, pa), (b_, pb) = outs
        assert a.status == b_.status and a.steps_done == b_.steps_done == 10
        assert abs(a.neg_h_init - b_.neg_h_init) <= 2e-5 * abs(a.neg_h_init)
        assert abs(a.neg_h_final - b_.neg_h_final) <= 2e-4 * abs(a.neg_h_final)
        assert np.allclose(pa, pb, rtol=2e-4, atol=2e-5)
        if a.status == rb.HMC_ACCEPTED:
            assert np.allclose(a.y_pred, b_.y_pred, rtol=0, atol=2e-4)
        net.select_hmc_path(net.HMC_PERSISTENT)
        with pytest.raises(RuntimeError):
            net.hmc_step(1, cfg)                       # 100 markers: not eligible
        net.select_hmc_path(net.HMC_AUTO)
        before = net.persistent_launches()
        assert net.hmc_step(1, cfg, u=0.0).steps_done == 10 and net.persistent_launches() == before
    finally:
        net.close()
        gen.close()


@pytest.mark.parametrize("model", ["ridge_ard", "lasso_base"])
def test_chain_visits_through_the_persistent_kernel_match_the_oracle_chain(rb, ctx, model):
    """Net::train visits (net.rs:258-334) with AUTO = the persistent kernel for every transition: three sweeps against the oracle
    chain with injected draws, as tests/test_gpu_parity.py::test_train_visits_match_oracle."""
    P = Problem(rb, ctx, model, 500, [20, 15, 9], 4, 3, seed=31)
    try:
        onet_ = mirror_net(P)
        P.net.set_globals(2.0, 0.05, onet_.g_ow_reg_sum, onet_.g_ow_num_params, 0.0)
        P.net.select_hmc_path(P.net.HMC_PERSISTENT)
        ocfg = OCfg(hmc_step_size_factor=0.5, hmc_integration_length=6)
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.5, hmc_integration_length=6)
        draws = onet.Draws(seed=123)
        resid_o = onet.initialize_stats(onet_, P.payload, P.n, P.means, P.stds, P.y, np.float32)
        P.net.init_residual()
        xs = [P.x(b, np.float32) for b in range(3)]
        flips = 0
        for it in range(3):
            for b in draws.order(3):
                b = int(b)
                resid_o, res_o = onet.visit_branch(onet_, b, xs[b], resid_o, ocfg, draws)
                d = draws.log[-1]
                got = P.net.visit_branch(b, cfg, momenta=d["momenta"], u=d["u"], std_gammas=np.array(d["gammas"], dtype=np.float32))
                if got.status != res_o["status"]:
                    flips += 1
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
        assert flips <= 1
        assert P.net.persistent_launches() == 9
        st = P.net.stats()
        if flips == 0:
            assert st["num_accepted"] == onet_.num_accepted and st["num_early_rejected"] == onet_.num_early_rejected
            lo = onet.lpd_value(onet_)
            assert abs(st["lpd"] - lo) < 5e-4 * abs(lo), (st["lpd"], lo)
    finally:
        P.close()
