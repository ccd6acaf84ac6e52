"""Pins the oracle's C restatement (the CPU-baseline kernel) against the NumPy oracle."""
import numpy as np
import pytest

from oracle import bed as obed
from oracle.branch import Branch, MCMCCfg, make_cfg
from oracle.cport import CPort


def lam_vector(cfg, model):
    br32 = Branch(cfg, np.float32)
    lam_w = []
    for l, w in enumerate(br32.W):
        p = br32.wprec[l]
        if model == "std_normal":
            lam_w.append(np.ones_like(w))
        elif model.endswith("ard") and l < br32.last:
            lam_w.append(np.repeat(p[:, None], w.shape[1], axis=1))
        else:
            lam_w.append(np.full_like(w, p[0]))
    return Branch.join_vec(lam_w, [np.zeros_like(b) for b in br32.b]).astype(np.float32)


@pytest.mark.parametrize("model,hidden,summary,depth", [("ridge_ard", 5, 5, 1), ("std_normal", 2, 2, 1),
                                                        ("lasso_base", 4, 3, 2)])
def test_cport_matches_numpy_oracle(model, hidden, summary, depth):
    cp = CPort()
    rng = np.random.default_rng(3)
    n, m = 700, 23
    g = obed.random_genotypes(n, m + 5, seed=9)
    payload = obed.pack_columns(g)
    mu, sd = obed.col_stats(payload, n, m + 5)
    cols = list(rng.permutation(m + 5)[:m])
    X = cp.decode_std(payload, n, cols, mu, sd)
    Xo = obed.submatrix_standardized(payload, n, cols, mu, sd)
    assert np.array_equal(X.reshape(n, m, order="F"), Xo)          # bit-exact decode + standardise
    cfg = make_cfg(model, m, [hidden] * depth, summary, rng=rng)
    cfg.biases = [rng.normal(0, 0.2, size=b.shape).astype(np.float32) for b in cfg.biases]
    cfg.bias_precisions = [np.ones(1, dtype=np.float32) for _ in cfg.biases]
    y = rng.normal(size=n).astype(np.float32)
    br = Branch(cfg, np.float64)
    rss, gW, gb = br.backpropagate(Xo, y)
    theta = cfg.param_vec().astype(np.float32)
    rss_c, d_c = cp.backprop(X, y, n, m, cfg.layer_widths, "tanh", theta)
    assert abs(rss_c - rss) < 2e-5 * rss
    exp = Branch.join_vec(gW, gb)
    assert np.allclose(d_c, exp, rtol=2e-4, atol=2e-5 * np.abs(exp).max())
    # leapfrog steps against the NumPy hmc_step trajectory (uniform step size, no early reject)
    L = 4
    ocfg = MCMCCfg(hmc_step_size_factor=1e-3, hmc_integration_length=L, hmc_step_size_mode="uniform",
                   hmc_max_hamiltonian_error=1e9)
    mom = rng.standard_normal(theta.size).astype(np.float32)
    res = Branch(cfg, np.float64).hmc_step(Xo, y, ocfg, mom, 0.0, record=True)
    ws, bs = Branch(cfg, np.float32).step_sizes(ocfg)
    eps = Branch.join_vec(ws, bs).astype(np.float32)
    lam = lam_vector(cfg, model)
    th, mm = theta.copy(), mom.copy()
    negh = cp.leapfrog(X, y, n, m, cfg.layer_widths, "tanh", model.startswith("lasso"), model == "std_normal", th, mm,
                       eps, lam, cfg.error_precision, L)
    assert np.allclose(th, res["traj"]["params"][-1], rtol=1e-4, atol=1e-6)
    assert abs(negh - res["traj"]["hamiltonian"][-1]) < 2e-5 * abs(res["traj"]["hamiltonian"][-1])
