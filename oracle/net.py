"""Oracle: the network-level chain driver (Gibbs over branches), residual bookkeeping,
log posterior density and prediction.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/net/net.rs, src/net/architectures.rs, src/net/log_posterior_density.rs,
src/net/params.rs (GlobalParams) and src/net/train_stats.rs of the reference.
Randomness is injected through a `Draws` object (branch orders, standard-gamma variates,
momenta, accept uniforms, step-size uniforms) so that the same draws can be replayed into
the CUDA library.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import bed as obed
from .branch import (ACCEPTED, REJECTED, REJECTED_EARLY, Branch, BranchCfg, Hyper, MCMCCfg,
                     is_lasso, make_cfg, summary_stat_host)

DEFAULT_INIT_OUTPUT_LAYER_PRECISION = 0.05  # architectures.rs:16


class Draws:
    """Deterministic source of injected randomness, recording everything it hands out."""

    def __init__(self, seed: int = 0):
        self.rng = np.random.default_rng(seed)
        self.log = []  # list of per-visit dicts

    def order(self, num_branches: int):
        o = np.arange(num_branches)
        self.rng.shuffle(o)
        return o

    def new_visit(self, branch_ix: int):
        self.cur = dict(branch=int(branch_ix), gammas=[], momenta=None, u=None, step_uniforms=None)
        self.log.append(self.cur)

    def std_gamma(self, shape: float) -> np.float32:
        g = np.float32(self.rng.standard_gamma(shape))
        self.cur["gammas"].append(float(g))
        return g

    def momenta(self, n: int):
        p = self.rng.standard_normal(n).astype(np.float32)
        self.cur["momenta"] = p
        return p

    def uniform(self) -> np.float32:
        u = np.float32(self.rng.random(dtype=np.float32))
        self.cur["u"] = float(u)
        return u

    def step_uniforms(self, n: int):
        v = self.rng.random(n, dtype=np.float32)
        self.cur["step_uniforms"] = v
        return v


@dataclass
class Net:
    """net.rs:74-85 (+ GlobalParams params.rs:13-56, OutputBias net.rs:29-72)."""
    model: str
    hyper: Hyper
    cfgs: List[BranchCfg]
    groups: List[List[int]]
    output_bias: float = 0.0
    # GlobalParams
    g_error_precision: float = 2.0
    g_output_layer_precision: float = DEFAULT_INIT_OUTPUT_LAYER_PRECISION
    g_ow_reg_sum: float = 0.0
    g_ow_num_params: int = 0
    # LogPosteriorDensity (log_posterior_density.rs:9-25)
    lpd_rss: float = float("-inf")
    lpd_out_w: float = float("-inf")
    lpd_local: Optional[np.ndarray] = None
    # TrainingStats (train_stats.rs:23-32)
    num_samples: int = 0
    num_accepted: int = 0
    num_early_rejected: int = 0
    mse_train: list = field(default_factory=list)
    lpd: list = field(default_factory=list)

    @property
    def num_branches(self):
        return len(self.cfgs)


def build_net(model: str, groups, hidden_layers: int, hidden_width: int, summary_width: int, hyper: Hyper,
              activation="tanh", seed=0, fixed_param_precision=None) -> Net:
    """BlockNetCfg::build_net (architectures.rs:187-237) with Fixed width rules and the default
    parameter initialisation.  The ChaCha20 init stream of the reference is third-party;
    weights come from a NumPy Generator instead (same distribution)."""
    rng = np.random.default_rng(seed)
    cfgs = []
    reg = np.float32(0)
    nparams = 0
    for g in groups:
        cfg = make_cfg(model, len(g), [hidden_width] * hidden_layers, summary_width,
                       activation=activation, rng=rng)
        if fixed_param_precision is not None:  # branch_cfg_builder.rs:254-262,276-283
            p = np.float32(fixed_param_precision)
            cfg.weight_precisions = [np.full(1, p, dtype=np.float32) for _ in cfg.layer_widths]
            cfg.bias_precisions = [np.full(1, p, dtype=np.float32) for _ in cfg.layer_widths[:-1]]
        cfgs.append(cfg)
        reg = np.float32(reg + np.float32(summary_stat_host(model, cfg.weights[-1])))  # :215-218
        nparams += summary_width  # :120-121
    # update_branch_cfgs_output_weight_precision, architectures.rs:175-185
    tot = np.float32(0)
    for c in cfgs:
        w = c.weights[-1].astype(np.float32)
        tot = np.float32(tot + np.float32(np.sum(w * w)))
    owp = np.float32(len(cfgs)) / tot
    for c in cfgs:
        c.weight_precisions[-1] = np.array([owp], dtype=np.float32)
    return Net(model=model, hyper=hyper, cfgs=cfgs, groups=[list(g) for g in groups], output_bias=0.0,
               g_error_precision=2.0,
               g_output_layer_precision=(fixed_param_precision if fixed_param_precision is not None
                                         else DEFAULT_INIT_OUTPUT_LAYER_PRECISION),
               g_ow_reg_sum=float(reg), g_ow_num_params=nparams,
               lpd_local=np.full(len(cfgs), -np.inf, dtype=np.float32))


def _update_cfg_globals(net: Net, cfg: BranchCfg):
    """branch_cfg.rs:59-63."""
    cfg.error_precision = net.g_error_precision
    cfg.weight_precisions[-1] = np.array([net.g_output_layer_precision], dtype=cfg.weight_precisions[-1].dtype)
    cfg.ow_reg_sum = net.g_ow_reg_sum
    cfg.ow_num_params = net.g_ow_num_params


def _update_globals_from_cfg(net: Net, cfg: BranchCfg):
    """params.rs:41-56."""
    assert cfg.error_precision >= 0.0
    net.g_error_precision = cfg.error_precision
    net.g_output_layer_precision = float(cfg.weight_precisions[-1][0])
    if cfg.ow_reg_sum < 0 or np.isnan(cfg.ow_reg_sum):
        raise RuntimeError("Invalid output weight summary statistic!")  # exit(SOFTWARE), params.rs:49-54
    net.g_ow_reg_sum = cfg.ow_reg_sum
    net.g_ow_num_params = cfg.ow_num_params


def _update_lpd(net: Net, branch_ix: int, branch: Branch, residual: np.ndarray):
    """log_posterior_density.rs:27-60 (StdNormal: unimplemented in the reference, Q6/H10;
    extension documented in DESIGN.md: local term = -1/2 sum theta^2, output term 0)."""
    dt = branch.dt
    if branch.model == "std_normal":
        local = dt.type(0)
        for w in branch.W:
            local = dt.type(local - dt.type(0.5) * dt.type(np.sum(w * w)))
        for b in branch.b:
            local = dt.type(local - dt.type(0.5) * dt.type(np.sum(b * b)))
        out_w = dt.type(0)
    else:
        out_w, local = branch.log_density_joint_components_curr_internal_state(net.hyper)
    net.lpd_local[branch_ix] = local
    net.lpd_out_w = float(out_w)
    k, s = net.hyper.output
    rss = branch.sum_sq(residual)
    n = dt.type(residual.size)
    net.lpd_rss = float(np.log(branch.eprec) * dt.type(dt.type(k) + dt.type(n - dt.type(2)) / dt.type(2))
                        - branch.eprec * dt.type(rss / dt.type(2) + dt.type(1) / dt.type(s)))


def lpd_value(net: Net) -> float:
    """log_posterior_density.rs:62-67 (sequential f32 sum of the local terms)."""
    acc = np.float32(0)
    for v in net.lpd_local:
        acc = np.float32(acc + np.float32(v))
    return float(np.float32(np.float32(np.float32(net.lpd_rss) + np.float32(net.lpd_out_w)) + acc))


def x_branch(net: Net, payload, n, means, stds, b, dtype):
    """data/genotypes.rs:44-48 -> io/bed.rs:325-355."""
    return obed.submatrix_standardized(payload, n, net.groups[b], means, stds, dtype)


def predict(net: Net, payload, n, means, stds, dtype=np.float32):
    """net.rs:545-559."""
    y = np.zeros(n, dtype=dtype) + dtype(net.output_bias)
    for b, cfg in enumerate(net.cfgs):
        y = y + Branch(cfg, dtype).predict(x_branch(net, payload, n, means, stds, b, dtype))
    return y


def initialize_stats(net: Net, payload, n, means, stds, y, dtype=np.float32):
    """net.rs:158-171."""
    residual = np.asarray(y, dtype=dtype) - dtype(net.output_bias)
    for b in range(net.num_branches):
        cfg = net.cfgs[b]
        _update_cfg_globals(net, cfg)
        br = Branch(cfg, dtype)
        residual = residual - br.predict(x_branch(net, payload, n, means, stds, b, dtype))
        _update_lpd(net, b, br, residual)
    return residual


def record_perf(net: Net, residual):
    """net.rs:597-606."""
    dt = residual.dtype
    net.lpd.append(lpd_value(net))
    net.mse_train.append(float(dt.type(np.dot(residual, residual)) / dt.type(residual.size)))


def visit_branch(net: Net, b: int, x, residual, mcmc: MCMCCfg, draws: Draws, dtype=np.float32, record=False):
    """One iteration of the inner loop of Net::train, net.rs:258-334. Returns (residual, hmc result)."""
    dt = np.dtype(dtype)
    cfg = net.cfgs[b]
    _update_cfg_globals(net, cfg)                       # :261-262
    br = Branch(cfg, dtype)                             # :268
    draws.new_visit(b)
    joint = mcmc.gradient_descent_joint or mcmc.joint_hmc
    if not joint:                                       # :268
        br.sample_error_precision(residual, net.hyper, draws.std_gamma)   # :272
        if not mcmc.fixed_param_precisions:
            br.sample_param_precisions(net.hyper, draws.std_gamma)        # :275
    prev_pred = br.predict(x)                           # :279
    residual = (residual + prev_pred).astype(dt)        # :280
    if mcmc.gradient_descent:                           # :282-290, in the reference's order
        res = br.gradient_descent(x, residual, mcmc)
    elif mcmc.gradient_descent_joint:
        res = br.gradient_descent_joint(x, residual, mcmc, net.hyper)
    elif mcmc.joint_hmc:
        nq = br.param_vec().size + br.precision_vec().size
        su = draws.step_uniforms(nq)
        res = br.hmc_step_joint(x, residual, mcmc, net.hyper, draws.momenta(nq), draws.uniform(), su, record=record)
    else:
        su = draws.step_uniforms(br.param_vec().size) if mcmc.hmc_step_size_mode == "random" else None
        res = br.hmc_step(x, residual, mcmc, draws.momenta(br.param_vec().size), draws.uniform(),
                          step_uniforms=su, record=record)  # :289
    net.num_samples += 1                                # train_stats.rs:48-56
    if res["status"] == ACCEPTED:
        net.num_accepted += 1
        residual = (residual - res["y_pred"]).astype(dt)   # :295
        _update_lpd(net, b, br, residual)               # :296
    else:
        if res["status"] == REJECTED_EARLY:
            net.num_early_rejected += 1
        residual = (residual - prev_pred).astype(dt)    # :299
    new_cfg = br.to_cfg()                               # :303
    _update_globals_from_cfg(net, new_cfg)              # :304
    net.cfgs[b] = new_cfg                               # :305
    # ML output bias, :320-332 + :43-45
    residual = (residual + dt.type(net.output_bias)).astype(dt)
    net.output_bias = float(dt.type(np.sum(residual)) / dt.type(residual.size))
    residual = (residual - dt.type(net.output_bias)).astype(dt)
    return residual, res


def visit_group(net: Net, members, xs, residual, mcmc: MCMCCfg, draws: Draws, dtype=np.float32):
    """One group visit of the block-Jacobi schedule (bann_visit_group / bann_sweep with group_size > 1).

    Every member runs the inner loop of Net::train (net.rs:258-334) -- globals -> cfg, Gibbs draws, prev_pred, HMC against
    residual + prev_pred -- against the residual and the global parameters FROZEN at group start.  Afterwards, once:
    residual -= sum over the accepted members (member order) of (y_new - (t - residual)), then per member in order the
    bookkeeping of :296-305 (counters, to_cfg against the running output-weight statistic, globals), the LPD terms of the last
    accepted member against the group's final residual, and the ML output bias (:320-332).  A group of one member is
    visit_branch.  Not in the reference (which only has the sequential order): this function is the specification the CUDA
    path is held to.  Returns (residual, [hmc results])."""
    dt = np.dtype(dtype)
    assert not (mcmc.joint_hmc or mcmc.gradient_descent or mcmc.gradient_descent_joint)
    r0 = np.asarray(residual, dtype=dt)
    runs = []
    for b in members:
        b = int(b)
        cfg = net.cfgs[b]
        _update_cfg_globals(net, cfg)                    # the globals are not written inside this loop: frozen
        br = Branch(cfg, dtype)
        own_old = dt.type(br.summary_stat(br.W[-1]))
        draws.new_visit(b)
        br.sample_error_precision(r0, net.hyper, draws.std_gamma)
        if not mcmc.fixed_param_precisions:
            br.sample_param_precisions(net.hyper, draws.std_gamma)
        prev = br.predict(xs[b])
        t = (r0 + prev).astype(dt)
        su = draws.step_uniforms(br.param_vec().size) if mcmc.hmc_step_size_mode == "random" else None
        res = br.hmc_step(xs[b], t, mcmc, draws.momenta(br.param_vec().size), draws.uniform(), step_uniforms=su)
        runs.append((b, br, own_old, t, res))
    r = r0.copy()
    for b, br, own_old, t, res in runs:
        if res["status"] == ACCEPTED:
            r = (r - (np.asarray(res["y_pred"], dtype=dt) - (t - r0))).astype(dt)
    results = []
    for b, br, own_old, t, res in runs:
        br.ow_reg_sum = dt.type(dt.type(net.g_ow_reg_sum) - own_old)     # from_cfg against the running global
        net.num_samples += 1
        if res["status"] == ACCEPTED:
            net.num_accepted += 1
            _update_lpd(net, b, br, r)
        elif res["status"] == REJECTED_EARLY:
            net.num_early_rejected += 1
        new_cfg = br.to_cfg()
        _update_globals_from_cfg(net, new_cfg)
        net.cfgs[b] = new_cfg
        results.append(res)
    r = (r + dt.type(net.output_bias)).astype(dt)
    net.output_bias = float(dt.type(np.sum(r)) / dt.type(r.size))
    r = (r - dt.type(net.output_bias)).astype(dt)
    return r, results


def train(net: Net, payload, n, means, stds, y, mcmc: MCMCCfg, chain_length: int, draws: Draws,
          dtype=np.float32, orders=None, group_size: int = 1):
    """Net::train, net.rs:201-358 (file output / test-set MSE omitted). Returns final residual.
    group_size > 1: the block-Jacobi schedule (visit_group) over consecutive groups of the order."""
    residual = initialize_stats(net, payload, n, means, stds, y, dtype)
    record_perf(net, residual)
    xs = [x_branch(net, payload, n, means, stds, b, dtype) for b in range(net.num_branches)]
    for it in range(chain_length):
        order = orders[it] if orders is not None else draws.order(net.num_branches)
        if group_size <= 1:
            for b in order:
                residual, _ = visit_branch(net, int(b), xs[int(b)], residual, mcmc, draws, dtype)
        else:
            for i in range(0, len(order), group_size):
                grp = [int(b) for b in order[i:i + group_size]]
                if len(grp) == 1:
                    residual, _ = visit_branch(net, grp[0], xs[grp[0]], residual, mcmc, draws, dtype)
                else:
                    residual, _ = visit_group(net, grp, xs, residual, mcmc, draws, dtype)
        record_perf(net, residual)
    return residual
