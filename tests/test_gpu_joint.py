"""GPU parity of the flag-gated sampler modes against the CPU oracle, through the C ABI:
log_density_gradient_joint / log_density_joint (branch_sampler.rs:213-422), hmc_step_joint (:1070-1178),
gradient_descent (:964-1017), gradient_descent_joint (:1019-1066) and their dispatch inside a Net::train visit
(net.rs:268-290).  The joint densities and joint gradients of the oracle are pinned by the reference's golden vectors
(tests/test_oracle_golden.py); trajectories, line searches and decisions are "parity unpinned" (oracle only).

Tolerances as in test_gpu_parity.py: within TOL_K x the oracle's own f32-vs-f64 error + TOL_REL of the scale;
discrete decisions (accept / reject, the line search's step sizes) identical unless the f64 truth is a near-tie."""
import numpy as np
import pytest

from oracle import net as onet
from oracle.branch import ACCEPTED, REJECTED, REJECTED_EARLY, Branch, summary_stat_host
from oracle.branch import MCMCCfg as OCfg
from test_gpu_parity import HYPER, Problem, ctx, mirror_net, rb, within  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu

JOINT_MODELS = ["ridge_base", "ridge_ard", "lasso_base", "lasso_ard"]
OTHERS = 3.75      # output-weight statistic of the "other branches" of the net


def with_globals(P, b):
    """Give branch b's cfg a global output-weight statistic (own + others) and mirror it on the device."""
    c = P.cfgs[b]
    own = summary_stat_host(P.model, c.weights[-1].reshape(-1))
    c.ow_reg_sum = float(np.float32(own) + np.float32(OTHERS))
    c.ow_num_params = 3 * c.weights[-1].size
    P.net.set_globals(c.error_precision, float(c.weight_precisions[-1][0]), c.ow_reg_sum, c.ow_num_params, 0.0)
    P.net.set_branch(b, c.param_vec(), c.precision_vec())
    return c


@pytest.mark.parametrize("model", JOINT_MODELS)
@pytest.mark.parametrize("shape", [([30, 11], 5, 5, 1), ([9, 70], 4, 3, 2)])
def test_joint_density_and_gradient(rb, ctx, model, shape):
    sizes, hidden, summary, depth = shape
    P = Problem(rb, ctx, model, 500, sizes, hidden, summary, depth=depth, seed=21)
    try:
        for b in range(len(sizes)):
            c = with_globals(P, b)
            got = P.net.branch_joint(b)
            ref = {}
            for dt in (np.float32, np.float64):
                br = Branch(c, dt)
                rss, gw, gb, gwp, gbp, gep = br.log_density_gradient_joint(P.x(b, dt), P.y.astype(dt), HYPER)
                ref[dt] = dict(rss=rss, ldg=Branch.join_joint_vec(gw, gb, gwp, gbp, gep),
                               ldj=br.log_density_joint(rss, HYPER, P.n), ld=br.log_density(rss))
            t, m = ref[np.float64], ref[np.float32]
            within(got["rss"], t["rss"], m["rss"])
            within(got["log_density_joint"], t["ldj"], m["ldj"], scale=abs(t["ldj"]) + abs(t["rss"]))
            within(got["log_density"], t["ld"], m["ld"], scale=abs(t["ld"]) + abs(t["rss"]))
            Pn = c.num_params
            within(got["ldg"][:Pn], t["ldg"][:Pn], m["ldg"][:Pn])
            # precision gradients: differences of O(N) terms -- judge against the size of the terms
            within(got["ldg"][Pn:], t["ldg"][Pn:], m["ldg"][Pn:], scale=max(np.max(np.abs(t["ldg"][Pn:])), abs(t["rss"])))
    finally:
        P.close()


def test_joint_modes_refuse_std_normal(rb, ctx):
    P = Problem(rb, ctx, "std_normal", 200, [10], 3, 2, seed=1)
    try:
        with pytest.raises(rb.BannError, match="StdNormal"):
            P.net.branch_joint(0)
        with pytest.raises(rb.BannError, match="StdNormal"):
            P.net.hmc_step_joint(0, rb.MCMCCfg(hmc_integration_length=2))
        with pytest.raises(rb.BannError, match="StdNormal"):
            P.net.gradient_descent_joint(0, rb.MCMCCfg(hmc_integration_length=2))
    finally:
        P.close()


def run_oracle_joint(P, c, b, ocfg, mom, u, su, dt):
    br = Branch(c, dt)
    res = br.hmc_step_joint(P.x(b, dt), P.y.astype(dt), ocfg, HYPER, mom, u, su, record=True)
    res["params_after"], res["prec_after"] = br.param_vec(), br.precision_vec()
    return res


@pytest.mark.parametrize("model", JOINT_MODELS)
@pytest.mark.parametrize("factor,L", [(0.004, 10), (0.02, 8), (0.6, 6)])
def test_hmc_step_joint_trajectory_and_decision(rb, ctx, model, factor, L):
    P = Problem(rb, ctx, model, 600, [30, 11], 5, 5, seed=12)
    near_ties = 0
    try:
        rng = np.random.default_rng(77)
        for b in range(2):
            for trial in range(3):
                c = with_globals(P, b)
                T = c.num_params + c.precision_vec().size
                # the configured step-size mode is ignored by the joint sampler (always Random, :1094-1101)
                cfg = rb.MCMCCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_step_size_mode="izmailov")
                ocfg = OCfg(hmc_step_size_factor=factor, hmc_integration_length=L)
                mom = rng.standard_normal(T).astype(np.float32)
                u = float(np.float32(rng.random(dtype=np.float32)))
                su = rng.random(T, dtype=np.float32)
                got = P.net.hmc_step_joint(b, cfg, momenta=mom, u=u, step_uniforms=su, trajectory=True)
                o64 = run_oracle_joint(P, c, b, ocfg, mom, u, su, np.float64)
                o32 = run_oracle_joint(P, c, b, ocfg, mom, u, su, np.float32)
                hscale = abs(o64["h_init"]) + P.n
                within(got.neg_h_init, o64["h_init"], o32["h_init"], scale=hscale)
                nsteps = min(got.steps_done, o64["steps_done"], o32["steps_done"])
                assert nsteps >= 1
                hs = np.array(o64["traj"]["hamiltonian"])
                finite = np.all(np.isfinite(hs))
                for s in range(min(nsteps, 5) if finite else 0):
                    if abs(hs[s + 1] - hs[0]) > 50:      # blown up: rounding differences are amplified without bound
                        break
                    within(got.trajectory["params"][s], o64["traj"]["params"][s], o32["traj"]["params"][s], rel=5e-5)
                    within(got.trajectory["precisions"][s], o64["traj"]["precisions"][s], o32["traj"]["precisions"][s], rel=5e-5)
                    Pn = c.num_params
                    within(got.trajectory["ldg"][s][:Pn], o64["traj"]["ldg"][s][:Pn], o32["traj"]["ldg"][s][:Pn], rel=5e-5)
                    within(got.trajectory["ldg"][s][Pn:], o64["traj"]["ldg"][s][Pn:], o32["traj"]["ldg"][s][Pn:],
                           scale=np.max(np.abs(o64["traj"]["ldg"][s][Pn:])) + P.n, rel=5e-5)
                    within(got.trajectory["hamiltonian"][s + 1], hs[s + 1], o32["traj"]["hamiltonian"][s + 1],
                           scale=hscale + abs(hs[s + 1]), rel=5e-5)
                if not finite:                            # precisions crossed zero: NaN Hamiltonian -> Rejected (never early)
                    assert got.status == rb.HMC_REJECTED and o64["status"] == REJECTED
                    continue
                margin = 1e-3 * max(1.0, hscale * 1e-3)
                if o64["status"] == REJECTED_EARLY or got.status == rb.HMC_REJECTED_EARLY:
                    dist = np.min(np.abs(np.abs(hs - hs[0]) - 10.0))
                    if dist < margin:
                        near_ties += 1
                        continue
                    assert got.status == o64["status"] and got.steps_done == o64["steps_done"]
                    pv, qv = P.net.get_branch(b)
                    assert np.array_equal(pv, c.param_vec()) and np.array_equal(qv, c.precision_vec().astype(np.float32))
                    continue
                la = o64["log_acc"]
                if abs(min(la, 0.0) - np.log(max(u, 1e-30))) < margin:
                    near_ties += 1
                    continue
                assert got.status == o64["status"], (got.status, o64["status"], la, u)
                within(got.neg_h_final, o64["h_final"], o32["h_final"], scale=hscale, rel=1e-4)
                pv, qv = P.net.get_branch(b)
                within(pv, o64["params_after"], o32["params_after"], rel=1e-4)
                within(qv, o64["prec_after"], o32["prec_after"], rel=1e-4)
                if got.status == rb.HMC_ACCEPTED:
                    within(got.y_pred, o64["y_pred"], o32["y_pred"], rel=1e-4)
                    within(got.log_density, o64["log_density"], o32["log_density"], scale=hscale, rel=5e-5)
        assert near_ties <= 2
    finally:
        P.close()


@pytest.mark.parametrize("model", ["std_normal", "ridge_base", "ridge_ard", "lasso_base", "lasso_ard"])
def test_gradient_descent_line_search(rb, ctx, model):
    P = Problem(rb, ctx, model, 500, [20, 9], 4, 3, seed=5)
    try:
        mismatched = 0
        for b in range(2):
            c = P.cfgs[b]
            P.net.set_branch(b, c.param_vec(), c.precision_vec())
            cfg = rb.MCMCCfg(hmc_step_size_factor=2e-4, hmc_integration_length=4, gradient_descent=True)
            ocfg = OCfg(hmc_step_size_factor=2e-4, hmc_integration_length=4)
            got = P.net.gradient_descent(b, cfg)
            o = {}
            for dt in (np.float32, np.float64):
                br = Branch(c, dt)
                o[dt] = br.gradient_descent(P.x(b, dt), P.y.astype(dt), ocfg)
                o[dt]["params_after"] = br.param_vec()
            assert got.status == rb.HMC_ACCEPTED and got.steps_done == 4
            steps = got.trajectory["step_sizes"]
            if not np.allclose(steps, o[np.float64]["step_sizes"], rtol=1e-6):
                # a probe comparison decided differently: only legitimate when the f32 mimic disagrees with the truth too
                mismatched += 1
                assert not np.allclose(o[np.float32]["step_sizes"], o[np.float64]["step_sizes"], rtol=1e-6)
                continue
            assert got.trajectory["num_probes"] == o[np.float64]["num_probes"]
            pv, _ = P.net.get_branch(b)
            within(pv, o[np.float64]["params_after"], o[np.float32]["params_after"], rel=1e-4)
            within(got.y_pred, o[np.float64]["y_pred"], o[np.float32]["y_pred"], rel=1e-4)
            within(got.log_density, o[np.float64]["log_density"], o[np.float32]["log_density"], rel=1e-4)
        assert mismatched <= 1
    finally:
        P.close()


@pytest.mark.parametrize("model", JOINT_MODELS)
def test_gradient_descent_joint(rb, ctx, model):
    P = Problem(rb, ctx, model, 500, [20, 9], 4, 3, seed=6)
    try:
        seen = set()
        for b in range(2):
            # tiny steps: plain ascent; large steps: the error precision is driven below zero -> Rejected, state restored
            for factor, L in ((2e-5, 5), (0.02, 3), (0.05, 2)):
                c = with_globals(P, b)
                cfg = rb.MCMCCfg(hmc_step_size_factor=factor, hmc_integration_length=L, gradient_descent_joint=True)
                ocfg = OCfg(hmc_step_size_factor=factor, hmc_integration_length=L)
                got = P.net.gradient_descent_joint(b, cfg)
                o = {}
                for dt in (np.float32, np.float64):
                    br = Branch(c, dt)
                    o[dt] = br.gradient_descent_joint(P.x(b, dt), P.y.astype(dt), ocfg, HYPER)
                    o[dt]["params_after"], o[dt]["prec_after"] = br.param_vec(), br.precision_vec()
                t, m = o[np.float64], o[np.float32]
                if t["status"] != m["status"] or not (np.all(np.isfinite(t["params_after"])) and np.all(np.isfinite(m["params_after"]))
                                                      and np.all(np.isfinite(t["prec_after"])) and np.all(np.isfinite(m["prec_after"]))):
                    continue                                      # the f32 restatement itself leaves the truth: chaotic blow-up
                assert got.status == t["status"]
                seen.add(t["status"])
                pv, qv = P.net.get_branch(b)
                if t["status"] == REJECTED:                       # :1053-1058
                    assert np.array_equal(pv, c.param_vec()) and np.array_equal(qv, c.precision_vec().astype(np.float32))
                    continue
                within(pv, t["params_after"], m["params_after"], rel=1e-4)
                within(qv, t["prec_after"], m["prec_after"], rel=1e-4)
                within(got.y_pred, t["y_pred"], m["y_pred"], rel=1e-4)
                if np.isfinite(t["log_density"]):
                    within(got.log_density, t["log_density"], m["log_density"], scale=abs(t["log_density"]) + P.n, rel=1e-4)
        assert ACCEPTED in seen and REJECTED in seen, seen
    finally:
        P.close()


@pytest.mark.parametrize("mode,model", [("joint_hmc", "ridge_ard"), ("joint_hmc", "lasso_base"),
                                        ("gradient_descent_joint", "ridge_base"), ("gradient_descent_joint", "lasso_ard"),
                                        ("gradient_descent", "ridge_ard"), ("gradient_descent", "std_normal")])
def test_train_visits_flag_gated_modes(rb, ctx, mode, model):
    """Net::train's inner loop with --joint-hmc / --gradient-descent / --gradient-descent-joint (net.rs:268-290):
    no Gibbs draws in the joint modes, the stepper's result drives the residual / LPD / globals bookkeeping."""
    P = Problem(rb, ctx, model, 400, [12, 20, 7], 4, 3, seed=31)
    try:
        onet_ = mirror_net(P)
        # error precision < 1: hmc_step_joint's accept compares the NON-joint final density with the joint initial
        # Hamiltonian (:928-962,1164), whose (shape + (N - 2) / 2) log(lambda_e) term decides everything -- below 1 the
        # trajectories are accepted, above 1 rejected (covered by test_hmc_step_joint_trajectory_and_decision)
        onet_.g_error_precision = 0.8
        P.net.set_globals(onet_.g_error_precision, onet_.g_output_layer_precision, onet_.g_ow_reg_sum, onet_.g_ow_num_params,
                          onet_.output_bias)
        factor = {"joint_hmc": 0.003, "gradient_descent_joint": 1e-5, "gradient_descent": 1e-4}[mode]
        kw = dict(hmc_step_size_factor=factor, hmc_integration_length=4, hmc_step_size_mode="random", **{mode: True})
        cfg, ocfg = rb.MCMCCfg(**kw), OCfg(**kw)
        draws = onet.Draws(seed=3)
        resid_o = onet.initialize_stats(onet_, P.payload, P.n, P.means, P.stds, P.y, np.float32)
        P.net.init_residual()
        xs = [P.x(b, np.float32) for b in range(3)]
        flips = 0
        for it in range(2):
            for b in draws.order(3):
                b = int(b)
                resid_o, res_o = onet.visit_branch(onet_, b, xs[b], resid_o, ocfg, draws)
                d = draws.log[-1]
                gam = np.array(d["gammas"], dtype=np.float32) if d["gammas"] else None
                got = P.net.visit_branch(b, cfg, momenta=d["momenta"], u=d["u"], step_uniforms=d["step_uniforms"], std_gammas=gam)
                pv, qv = P.net.get_branch(b)
                ok = (got.status == res_o["status"]
                      and np.allclose(pv, onet_.cfgs[b].param_vec(), rtol=5e-4, atol=5e-5))
                if not ok:
                    flips += 1      # near-tie (accept / line search) in f32: resynchronise the device from the oracle
                    for bb, c in enumerate(onet_.cfgs):
                        P.net.set_branch(bb, c.param_vec(), c.precision_vec())
                    P.net.set_residual(resid_o)
                    P.net.set_globals(onet_.g_error_precision, onet_.g_output_layer_precision, onet_.g_ow_reg_sum,
                                      onet_.g_ow_num_params, onet_.output_bias)
                    continue
                assert np.allclose(qv, onet_.cfgs[b].precision_vec(), rtol=5e-4, atol=1e-6)
                assert np.allclose(P.net.residual(), resid_o, rtol=0, atol=1e-3)
                g = P.net.get_globals()
                assert abs(g["output_bias"] - onet_.output_bias) < 1e-5
                assert abs(g["ow_reg_sum"] - onet_.g_ow_reg_sum) < 2e-4 * max(1.0, onet_.g_ow_reg_sum)
                assert abs(g["error_precision"] - onet_.g_error_precision) < 5e-4 * abs(onet_.g_error_precision)
                assert abs(g["output_layer_precision"] - onet_.g_output_layer_precision) < 5e-4 * abs(onet_.g_output_layer_precision)
        assert flips <= 1
        st = P.net.stats()
        if flips == 0:
            assert st["num_samples"] == onet_.num_samples and st["num_accepted"] == onet_.num_accepted
            if model != "std_normal":
                lo = onet.lpd_value(onet_)
                assert abs(st["lpd"] - lo) < 1e-3 * abs(lo), (st["lpd"], lo)
    finally:
        P.close()


# ------------------------------------------------------------------ per-row diagnostics (Net::activations, effect sizes)
@pytest.mark.parametrize("model,act,shape", [("ridge_ard", "tanh", ([30, 11, 64], 5, 5, 1)), ("lasso_base", "silu", ([9, 70], 4, 3, 2)),
                                             ("std_normal", "relu", ([17, 5], 6, 2, 1))])
def test_activations_and_effect_sizes(rb, ctx, model, act, shape):
    from oracle import bed as obed
    sizes, hidden, summary, depth = shape
    P = Problem(rb, ctx, model, 391, sizes, hidden, summary, depth=depth, act=act, seed=41)
    try:
        pes_all = P.net.population_effect_sizes()
        off = 0
        for b in range(len(sizes)):
            c = P.cfgs[b]
            ref = {}
            for dt in (np.float32, np.float64):
                br = Branch(c, dt)
                ref[dt] = (br.forward_feed(P.x(b, dt))[1], br.effect_sizes(P.x(b, dt)))
            acts = P.net.branch_activations(b)
            assert len(acts) == len(ref[np.float64][0])                       # a_0 .. a_{last-1}, yhat (net.rs:509-518)
            for got, t, m in zip(acts, ref[np.float64][0], ref[np.float32][0]):
                assert got.shape == t.shape
                within(got, t, m, scale=max(1.0, np.max(np.abs(t))))
            es, pop = P.net.branch_effect_sizes(b)
            t, m = ref[np.float64][1], ref[np.float32][1]
            within(es, t, m, rel=5e-5)
            within(pop, t.sum(axis=0) / P.n, m.sum(axis=0, dtype=np.float32) / np.float32(P.n), scale=np.max(np.abs(t)), rel=5e-5)
            assert np.array_equal(pes_all[off:off + len(P.groups[b])], pop)
            off += len(P.groups[b])
        # a second store (test data) with its own column statistics
        g2 = obed.random_genotypes(77, P.m, seed=99)
        pl2 = obed.pack_columns(g2)
        mu2, sd2 = obed.col_stats(pl2, 77, P.m)
        test = rb.Genotypes(ctx, pl2, 77, P.m, P.groups)
        x2 = obed.submatrix_standardized(pl2, 77, P.groups[0], mu2, sd2, np.float64)
        acts2 = P.net.branch_activations(0, test)
        t2 = Branch(P.cfgs[0], np.float64).forward_feed(x2)[1]
        assert np.allclose(acts2[-1], t2[-1], rtol=0, atol=2e-5 * max(1.0, np.max(np.abs(t2[-1]))))
        test.close()
    finally:
        P.close()
