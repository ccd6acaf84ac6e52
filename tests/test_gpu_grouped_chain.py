"""The block-Jacobi (grouped) schedule as a CHAIN: bann_visit_group / bann_sweep(group_size = G) against
oracle/net.py:visit_group with every draw injected, G = 1 bit for bit the sequential chain, and posterior-predictive R^2
of the sequential and the grouped chain agreeing within Monte-Carlo error (the reference's R^2 is
1 - mse / variance, py-vis/vis.py:555-557, net/net.rs:597-610)."""
import numpy as np
import pytest

from oracle import bed as obed
from oracle import net as onet
from oracle.branch import MCMCCfg as OCfg

from test_gpu_parity import Problem, mirror_net

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


def _sync_device_from_oracle(P, onet_, resid_o):
    for bb, c in enumerate(onet_.cfgs):
        P.net.set_branch(bb, c.param_vec(), c.precision_vec())
    P.net.set_residual(resid_o)
    P.net.set_globals(onet_.g_error_precision, onet_.g_output_layer_precision, onet_.g_ow_reg_sum, onet_.g_ow_num_params,
                      onet_.output_bias)


@pytest.mark.parametrize("model", ["ridge_ard", "lasso_base", "std_normal", "ridge_base", "lasso_ard"])
@pytest.mark.parametrize("mode", ["izmailov", "random"])
def test_group_visits_match_oracle(rb, ctx, model, mode):
    """Three sweeps of groups of 3 + 2 branches (a ragged last group, overlapping marker sets) replayed draw for draw:
    parameters, precisions, residual, globals, counters and LPD of the CUDA chain follow the oracle's visit_group."""
    sizes = [20, 15, 9, 33, 50]
    P = Problem(rb, ctx, model, 700, sizes, 4, 3, seed=41, overlap=True)
    try:
        onet_ = mirror_net(P)
        P.net.set_globals(2.0, 0.05, onet_.g_ow_reg_sum, onet_.g_ow_num_params, 0.0)
        f = 0.5 if mode == "izmailov" else 0.02
        ocfg = OCfg(hmc_step_size_factor=f, hmc_integration_length=6, hmc_step_size_mode=mode)
        cfg = rb.MCMCCfg(hmc_step_size_factor=f, hmc_integration_length=6, hmc_step_size_mode=mode)
        draws = onet.Draws(seed=321)
        resid_o = onet.initialize_stats(onet_, P.payload, P.n, P.means, P.stds, P.y, np.float32)
        P.net.init_residual()
        xs = [P.x(b, np.float32) for b in range(len(sizes))]
        flips, visits = 0, 0
        for it in range(3):
            order = [int(b) for b in draws.order(len(sizes))]
            for grp in (order[:3], order[3:]):
                n0 = len(draws.log)
                resid_o, res_o = onet.visit_group(onet_, grp, xs, resid_o, ocfg, draws)
                inj = [dict(momenta=d["momenta"], u=d["u"], step_uniforms=d["step_uniforms"],
                            std_gammas=np.array(d["gammas"], dtype=np.float32)) for d in draws.log[n0:]]
                got = P.net.visit_group(grp, cfg, injections=inj)
                visits += len(grp)
                if [g.status for g in got] != [r["status"] for r in res_o]:
                    flips += 1          # near-tie in f32: resynchronise the device from the oracle
                    _sync_device_from_oracle(P, onet_, resid_o)
                    continue
                for b in grp:
                    pv, qv = P.net.get_branch(b)
                    assert np.allclose(pv, onet_.cfgs[b].param_vec(), rtol=2e-4, atol=2e-5)
                    assert np.allclose(qv, onet_.cfgs[b].precision_vec(), rtol=2e-4, atol=1e-6)
                assert np.allclose(P.net.residual(), resid_o, rtol=0, atol=5e-4)
                g = P.net.get_globals()
                assert abs(g["output_bias"] - onet_.output_bias) < 1e-5
                assert abs(g["ow_reg_sum"] - onet_.g_ow_reg_sum) < 1e-4 * max(1.0, onet_.g_ow_reg_sum)
                assert abs(g["error_precision"] - onet_.g_error_precision) < 1e-4 * onet_.g_error_precision
                assert abs(g["output_layer_precision"] - onet_.g_output_layer_precision) < 1e-4 * onet_.g_output_layer_precision
        assert flips <= 1
        st = P.net.stats()
        if flips == 0:
            assert st["num_samples"] == onet_.num_samples == visits
            assert st["num_accepted"] == onet_.num_accepted and st["num_early_rejected"] == onet_.num_early_rejected
            lo = onet.lpd_value(onet_)
            assert abs(st["lpd"] - lo) < 5e-4 * abs(lo), (st["lpd"], lo)
        r = P.net.residual()
        assert abs(st["mse_train"] - float(np.dot(r, r) / P.n)) < 1e-4
    finally:
        P.close()


def test_group_of_one_is_the_sequential_visit(rb, ctx):
    """bann_sweep(group_size = 1), bann_sweep(group_size = G) with one-member groups and bann_visit_group of single members take
    the same path as bann_visit_branch: identical bits."""
    outs = []
    for how in ("sweep1", "visit_group"):
        P = Problem(rb, ctx, "ridge_ard", 600, [20, 15, 9, 33], 5, 5, seed=7)
        try:
            cfg = rb.MCMCCfg(hmc_step_size_factor=0.4, hmc_integration_length=8)
            on = mirror_net(P)
            P.net.set_globals(2.0, 0.05, on.g_ow_reg_sum, on.g_ow_num_params, 0.0)
            P.net.init_residual()
            order = [2, 0, 3, 1]
            if how == "sweep1":
                P.net.sweep(cfg, order, seed=5, group_size=1)
            else:
                for b in order:
                    P.net.visit_group([b], cfg, seed=5)
            outs.append((P.net.residual().copy(), [P.net.get_branch(b)[0].copy() for b in range(4)], P.net.stats()))
        finally:
            P.close()
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, b in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, b)
    assert outs[0][2] == outs[1][2]


def test_sweep_rejects_duplicate_members_and_joint_modes(rb, ctx):
    P = Problem(rb, ctx, "ridge_ard", 300, [10, 12, 9], 4, 3, seed=3)
    try:
        on = mirror_net(P)
        P.net.set_globals(2.0, 0.05, on.g_ow_reg_sum, on.g_ow_num_params, 0.0)
        P.net.init_residual()
        with pytest.raises(RuntimeError):
            P.net.sweep(rb.MCMCCfg(), [0, 0, 1], group_size=3)
        with pytest.raises(RuntimeError):
            P.net.sweep(rb.MCMCCfg(joint_hmc=True), [0, 1, 2], group_size=3)
        P.net.sweep(rb.MCMCCfg(hmc_integration_length=3), [0, 1, 0], group_size=2)      # the same branch in two groups is fine
    finally:
        P.close()


def _simulate(n, n_test, sizes, seed, h2, causal):
    """Genotypes + a standardised linear-model phenotype with `causal` causal markers per branch and heritability h2
    (the recipe of the reference's simulator, data/linear_model.rs:46-90)."""
    rng = np.random.default_rng(seed)
    m = sum(sizes)
    g = obed.random_genotypes(n + n_test, m, seed=seed + 1).astype(np.float64)
    x = (g - g.mean(axis=0)) / g.std(axis=0)
    gv = np.zeros(n + n_test)
    start = 0
    for sz in sizes:
        idx = start + rng.choice(sz, size=causal, replace=False)
        gv += x[:, idx] @ rng.normal(0, 1, size=causal)
        start += sz
    gv = (gv - gv.mean()) / gv.std()
    y = np.sqrt(h2) * gv + np.sqrt(1 - h2) * rng.normal(size=n + n_test)
    return g.astype(np.uint8), y.astype(np.float32)


def _chain_r2(rb, ctx, gen, gte, y_tr, y_te, model, groups, hidden, summary, G, iters, L, step, seed):
    """Held-out R^2 = 1 - mse / variance (py-vis/vis.py:555-557) of the posterior-mean prediction over the kept models
    (first third = burn-in), chain run with group size G from the reference's default initial state."""
    from rs_bann_b200 import architectures
    nf = architectures.build_net(model, [len(c) for c in groups], 1, fixed_hidden=hidden, fixed_summary=summary, seed=5)
    net = rb.Net(ctx, gen, model, [c.layer_widths for c in nf.branch_cfgs], hyper=tuple(nf.hyper))
    for b, c in enumerate(nf.branch_cfgs):
        net.set_branch(b, c.param_vec(), c.precision_vec())
    net.set_globals(nf.g_error_precision, nf.g_output_layer_precision, nf.g_ow_reg_sum, nf.g_ow_num_params, 0.0)
    net.set_targets(y_tr)
    net.init_residual()
    cfg = rb.MCMCCfg(hmc_step_size_factor=step, hmc_integration_length=L)
    rng = np.random.default_rng(seed)
    pred, kept = np.zeros(y_te.size), 0
    for it in range(iters):
        st = net.sweep(cfg, rng.permutation(len(groups)), seed=1000 * seed + it, group_size=G)
        if it >= iters // 3:
            pred += net.predict(gte)
            kept += 1
    pred /= kept
    net.close()
    return 1.0 - float(np.mean((y_te - pred) ** 2)) / float(np.var(y_te)), st


R2_CASES = {
    # name: n, n_test, branch sizes, hidden = summary width, prior, sweeps, h2, causal markers per branch, chains per G,
    #       group sizes, asserted pair, tolerance on the difference of the chain-averaged R^2, floor for the sequential chain
    "ard_cfg2_like": (2000, 1000, [30] * 8, 3, "ridge_ard", 100, 0.6, 3, 3, (1, 8), (1, 8), 0.07, 0.35),
    "std_normal_cfg1_like": (2000, 1000, [30] * 8, 2, "std_normal", 100, 0.6, 3, 3, (1, 8), (1, 8), 0.09, 0.2),
    # BASELINE configs[0] at its real shape: 1000 individuals, 10 branches x 100 markers -- as many markers as individuals,
    # every branch can fit the whole residual on its own.  G = B is RECORDED here, not asserted (see the docstring).
    "std_normal_cfg1_overparameterised": (1000, 500, [100] * 10, 2, "std_normal", 300, 0.7, 1, 2, (1, 5, 10), None, None, None),
}


@pytest.mark.parametrize("case", sorted(R2_CASES))
def test_r2_of_sequential_and_grouped_chains_agree(rb, ctx, case):
    """North-star acceptance check: posterior-predictive R^2 on a held-out split, the sequential
    chain (G = 1, the reference's order) against the block-Jacobi chain with G = B (every branch concurrently: the schedule of
    the throughput benchmark), several independent chains each, on cfg2-like (ARD) and cfg1-like (StdNormal) simulated data.
    Calibrated with the oracle chain on the CPU (same simulation, tests/golden/r2_calibration.md): chain-to-chain spread of
    R^2 ~ 0.04 (ARD) / 0.1 (StdNormal), no systematic difference between the two schedules.

    The third case is BASELINE configs[0] at its real shape (1000 individuals x 10 branches x 100 markers).  There block-Jacobi
    with G = B is measurably WORSE than the sequential chain, on the GPU and in the oracle alike (oracle: R^2 0.17 sequential
    vs 0.01-0.08 fully Jacobi, training MSE 0.29 vs 0.43 after 200-800 sweeps): each of the ten over-parameterised branches
    fits the same frozen residual, the accepted moves add up and overshoot.  The numbers are recorded (printed), not
    asserted; DESIGN.md section 4 states the limitation and `--group-size` defaults to 1."""
    n, n_test, sizes, hidden, model, iters, h2, causal, chains, gsizes, pair, tol, floor = R2_CASES[case]
    g, y = _simulate(n, n_test, sizes, seed=11, h2=h2, causal=causal)
    groups, start = [], 0
    for sz in sizes:
        groups.append(list(range(start, start + sz)))
        start += sz
    m = sum(sizes)
    gen = rb.Genotypes(ctx, obed.pack_columns(g[:n]), n, m, groups)
    mu, sd = gen.col_stats()
    gte = rb.Genotypes(ctx, obed.pack_columns(g[n:]), n_test, m, groups, col_means=mu, col_stds=sd)
    r2 = {G: [] for G in gsizes}
    mse = {G: [] for G in gsizes}
    try:
        for G in gsizes:
            for chain in range(chains):
                v, st = _chain_r2(rb, ctx, gen, gte, y[:n], y[n:], model, groups, hidden, hidden, G, iters, 20, 0.3, seed=1 + chain)
                assert np.isfinite(v) and st["num_accepted"] > 0.05 * st["num_samples"], (G, chain, v, st)
                r2[G].append(v)
                mse[G].append(st["mse_train"])
    finally:
        gte.close()
        gen.close()
    print(f"\nR2 {case}: " + "; ".join(f"G={G}: " + ", ".join(f"{v:.3f}" for v in r2[G]) + f" (mean {np.mean(r2[G]):.3f}, "
                                      f"mse_train {np.mean(mse[G]):.3f})" for G in gsizes))
    if pair is not None:
        a, b = np.mean(r2[pair[0]]), np.mean(r2[pair[1]])
        assert a > floor, r2                               # the sequential chains learned something
        assert abs(a - b) < tol, r2
