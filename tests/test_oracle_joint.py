"""CPU checks of the oracle's restatement of the flag-gated modes (joint HMC, gradient ascent): internal consistency in f64.

The joint densities and joint gradients themselves are pinned by the reference's golden vectors (test_oracle_golden.py); here
the two are checked AGAINST EACH OTHER by finite differences -- which also documents reference quirk Q8: for ARD priors the
gradient w.r.t. a weight precision counts the ROWS of the layer (`precisions.elements()`, ridge_ard.rs:231, lasso_ard.rs:229) where
the density counts its columns, so the analytic gradient differs from the derivative of the density by a known offset."""
import numpy as np
import pytest

from oracle import bed as obed
from oracle.branch import ACCEPTED, REJECTED, Branch, Hyper, MCMCCfg, make_cfg

HY = Hyper(dense=(3.0, 2.0), summary=(2.5, 1.5), output=(4.0, 5.0))


def problem(model, seed=0, n=120, m=9, hidden=(4,), summary=3):
    rng = np.random.default_rng(seed)
    g = obed.random_genotypes(n, m, seed=seed + 1)
    payload = obed.pack_columns(g)
    mu, sd = obed.col_stats(payload, n, m)
    x = obed.submatrix_standardized(payload, n, list(range(m)), mu, sd, np.float64)
    cfg = make_cfg(model, m, list(hidden), summary, rng=rng)
    cfg.biases = [rng.normal(0, 0.3, size=b.shape).astype(np.float32) for b in cfg.biases]
    cfg.weight_precisions = [rng.uniform(0.5, 3.0, size=p.shape).astype(np.float32) for p in cfg.weight_precisions]
    cfg.bias_precisions = [rng.uniform(0.5, 3.0, size=p.shape).astype(np.float32) for p in cfg.bias_precisions]
    cfg.error_precision = 1.3
    cfg.ow_reg_sum, cfg.ow_num_params = 7.5, 9
    y = rng.normal(size=n)
    return cfg, x, y


def joint_density(cfg, x, y, theta, prec):
    br = Branch(cfg, np.float64)
    br.load_param_vec(theta)
    br.load_precision_vec(prec)
    return float(br.log_density_joint(br.rss(x, y), HY, y.size))


@pytest.mark.parametrize("model", ["ridge_base", "lasso_base", "ridge_ard", "lasso_ard"])
def test_joint_gradient_is_the_derivative_of_the_joint_density(model):
    cfg, x, y = problem(model)
    br = Branch(cfg, np.float64)
    theta, prec = br.param_vec().copy(), br.precision_vec().copy()
    _, gw, gb, gwp, gbp, gep = br.log_density_gradient_joint(x, y, HY)
    g = Branch.join_joint_vec(gw, gb, gwp, gbp, gep)
    P, Q = theta.size, prec.size
    fd = np.zeros(P + Q)
    for k in range(P + Q):
        h = 1e-6 * max(1.0, abs(theta[k] if k < P else prec[k - P]))
        tp, tm, pp, pm = theta.copy(), theta.copy(), prec.copy(), prec.copy()
        if k < P:
            tp[k] += h; tm[k] -= h
        else:
            pp[k - P] += h; pm[k - P] -= h
        fd[k] = (joint_density(cfg, x, y, tp, pp) - joint_density(cfg, x, y, tm, pm)) / (2 * h)
    # the reference's backpropagation carries e, not 2e (Q4: d_rss is half the derivative of the rss) -- which is exactly what
    # the log density -lambda_e rss / 2 needs, so the parameter gradients match the finite differences
    assert np.allclose(g[:P], fd[:P], rtol=2e-5, atol=2e-6 * np.abs(fd[:P]).max())
    expected = fd[P:].copy()
    if model.endswith("ard"):          # Q8: rows instead of columns in the ARD precision gradient of the layers before the output
        off = 0
        for l in range(br.last):
            rows, cols = br.W[l].shape
            lam = br.wprec[l]
            expected[off:off + rows] += (rows - cols) / (2.0 * lam) if model == "ridge_ard" else (rows - cols) / lam
            off += rows
    assert np.allclose(g[P:], expected, rtol=2e-5, atol=2e-6 * np.abs(expected).max())


@pytest.mark.parametrize("model", ["ridge_base", "lasso_base"])   # (for ARD priors quirk Q8 breaks the conservation)
def test_joint_leapfrog_conserves_the_joint_hamiltonian_for_small_steps(model):
    cfg, x, y = problem(model, seed=3)
    br = Branch(cfg, np.float64)
    T = br.param_vec().size + br.precision_vec().size
    rng = np.random.default_rng(1)
    mom, su = rng.standard_normal(T), rng.random(T)
    res = br.hmc_step_joint(x, y, MCMCCfg(hmc_step_size_factor=2e-4, hmc_integration_length=25), HY, mom, 0.5, su, record=True)
    hs = np.array(res["traj"]["hamiltonian"])
    assert np.max(np.abs(hs - hs[0])) < 1e-4 * max(1.0, abs(hs[0]))
    # the accept step compares the NON-joint density with the joint Hamiltonian (reference behaviour, branch_sampler.rs:928-962,1164)
    assert res["status"] in (ACCEPTED, REJECTED)
    assert abs(res["h_final"] - hs[-1]) > 1.0          # the two differ by the hyper-prior terms, far more than the integration error


@pytest.mark.parametrize("model", ["std_normal", "ridge_ard"])
def test_gradient_descent_line_search_ascends_the_log_density(model):
    cfg, x, y = problem(model, seed=5)
    br = Branch(cfg, np.float64)
    before = float(br.log_density(br.rss(x, y)))
    res = br.gradient_descent(x, y, MCMCCfg(hmc_step_size_factor=1e-3, hmc_integration_length=6))
    assert res["status"] == ACCEPTED and res["log_density"] > before
    ratios = np.log2(np.array(res["step_sizes"]) / 1e-3)
    assert np.allclose(ratios, np.round(ratios))       # every step is the initial one doubled / halved an integer number of times


def test_gradient_descent_joint_rejects_a_negative_error_precision():
    cfg, x, y = problem("ridge_base", seed=7)
    br = Branch(cfg, np.float64)
    t0, p0 = br.param_vec().copy(), br.precision_vec().copy()
    res = br.gradient_descent_joint(x, y, MCMCCfg(hmc_step_size_factor=0.2, hmc_integration_length=2), HY)
    assert res["status"] == REJECTED
    assert np.array_equal(br.param_vec(), t0) and np.array_equal(br.precision_vec(), p0)     # :1053-1058 state restored
    ok = Branch(cfg, np.float64).gradient_descent_joint(x, y, MCMCCfg(hmc_step_size_factor=1e-5, hmc_integration_length=3), HY)
    assert ok["status"] == ACCEPTED
