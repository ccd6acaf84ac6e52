"""Oracle: one branch of the BANN -- forward, backprop, priors, step sizes, leapfrog HMC,
Gibbs precision draws.  TEST INFRASTRUCTURE (see oracle/__init__.py).

Follows src/net/branch/branch_sampler.rs, the five prior files
(std_normal_branch.rs, ridge_base.rs, ridge_ard.rs, lasso_base.rs, lasso_ard.rs),
src/net/branch/momentum.rs, src/net/params.rs, src/net/gibbs_steps.rs,
src/net/activation_functions.rs and src/af_helpers.rs of the reference.

Every routine works in a caller-chosen dtype: np.float32 restates the reference's f32 op
order ("f32 mimic"), np.float64 is the "truth" against which GPU error is judged.
All randomness (momenta, uniforms, standard-gamma variates) is INJECTED.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MODEL_TYPES = ("std_normal", "ridge_base", "ridge_ard", "lasso_base", "lasso_ard")  # model_type.rs:6-13
ACTIVATIONS = ("tanh", "relu", "leaky_relu", "silu", "identity")  # activation_functions.rs:6-12
STEP_SIZE_MODES = ("uniform", "random", "std_scaled", "izmailov")  # mcmc_cfg.rs:265-270

REJECTED_EARLY, REJECTED, ACCEPTED = 0, 1, 2  # branch_sampler.rs:1310-1314


@dataclass
class Hyper:
    """NetworkPrecisionHyperparameters, params.rs:135-188. Gamma(shape, scale)."""
    dense: tuple = (0.001, 1000.0)
    summary: tuple = (0.001, 1000.0)
    output: tuple = (0.001, 1000.0)

    def layer_prior(self, layer_index: int, num_layers: int):
        if layer_index == num_layers - 1:
            return self.output
        if layer_index == num_layers - 2:
            return self.summary
        return self.dense


@dataclass
class MCMCCfg:
    """mcmc_cfg.rs:181-230 (the fields the hot path reads)."""
    hmc_step_size_factor: float = 1.0
    hmc_max_hamiltonian_error: float = 10.0
    hmc_integration_length: int = 100
    hmc_step_size_mode: str = "izmailov"
    fixed_param_precisions: bool = False
    joint_hmc: bool = False                 # mcmc_cfg.rs:21-26
    gradient_descent: bool = False
    gradient_descent_joint: bool = False


@dataclass
class BranchCfg:
    """Host snapshot of one branch (branch_cfg.rs:8-16, params.rs:191-199,467-476)."""
    model: str
    num_markers: int
    layer_widths: List[int]                 # hidden..., summary, 1 (branch_cfg_builder.rs:285-297)
    weights: List[np.ndarray]               # W_l [in_l, out_l]
    biases: List[np.ndarray]                # b_l [out_l] for l < last
    weight_precisions: List[np.ndarray]     # ARD: [in_l] for l<last, [1] for last; Base: [1]
    bias_precisions: List[np.ndarray]       # [1] per layer l<last
    error_precision: float = 2.0
    activation: str = "tanh"
    ow_reg_sum: float = 0.0                 # GLOBAL output-weight stat (own + others) while in a cfg
    ow_num_params: int = 0

    @property
    def num_layers(self):
        return len(self.layer_widths)

    @property
    def num_params(self):
        return sum(w.size for w in self.weights) + sum(b.size for b in self.biases)

    def param_vec(self) -> np.ndarray:
        """params.rs:700-715: all weights (column-major) layer by layer, then all biases."""
        parts = [np.asarray(w).reshape(-1, order="F") for w in self.weights]
        parts += [np.asarray(b).reshape(-1) for b in self.biases]
        return np.concatenate(parts) if parts else np.zeros(0)

    def load_param_vec(self, pv) -> None:
        """params.rs:673-698."""
        pv = np.asarray(pv)
        prev, ix = self.num_markers, 0
        dt = self.weights[0].dtype
        for l, w in enumerate(self.layer_widths):
            n = prev * w
            self.weights[l] = pv[ix:ix + n].reshape((prev, w), order="F").astype(dt).copy()
            ix += n
            prev = w
        for l, w in enumerate(self.layer_widths[:-1]):
            self.biases[l] = pv[ix:ix + w].astype(dt).copy()
            ix += w

    def precision_vec(self) -> np.ndarray:
        """params.rs:272-289: weight precisions, bias precisions, error precision."""
        parts = [np.asarray(p).reshape(-1) for p in self.weight_precisions]
        parts += [np.asarray(p).reshape(-1) for p in self.bias_precisions]
        parts.append(np.array([self.error_precision]))
        return np.concatenate(parts)

    def load_precision_vec(self, v) -> None:
        v = np.asarray(v, dtype=np.float64)
        ix = 0
        for l in range(len(self.weight_precisions)):
            n = self.weight_precisions[l].size
            self.weight_precisions[l] = v[ix:ix + n].astype(self.weight_precisions[l].dtype).copy()
            ix += n
        for l in range(len(self.bias_precisions)):
            self.bias_precisions[l] = v[ix:ix + 1].astype(self.bias_precisions[l].dtype).copy()
            ix += 1
        self.error_precision = float(v[ix])

    def astype(self, dt) -> "BranchCfg":
        c = copy.deepcopy(self)
        c.weights = [np.asarray(w, dtype=dt) for w in c.weights]
        c.biases = [np.asarray(b, dtype=dt) for b in c.biases]
        c.weight_precisions = [np.asarray(p, dtype=dt) for p in c.weight_precisions]
        c.bias_precisions = [np.asarray(p, dtype=dt) for p in c.bias_precisions]
        return c


def is_ard(model: str) -> bool:
    return model in ("ridge_ard", "lasso_ard")


def is_lasso(model: str) -> bool:
    return model in ("lasso_base", "lasso_ard")


def summary_stat_host(model: str, vals) -> float:
    """BranchSampler::summary_stat_fn_host: ridge -> sum sq, lasso -> sum abs,
    std normal -> constant 1.0 (std_normal_branch.rs:35-37)."""
    v = np.asarray(vals, dtype=np.float32)
    if model == "std_normal":
        return 1.0
    if is_lasso(model):
        return float(np.add.reduce(np.abs(v), dtype=np.float32))
    return float(np.add.reduce(v * v, dtype=np.float32))


def make_cfg(model, num_markers, hidden_widths, summary_width, weights=None, biases=None,
             precision=None, activation="tanh", rng=None, dtype=np.float32) -> BranchCfg:
    """Build a BranchCfg. With explicit weights/biases and `precision` it mirrors the
    test-only BranchBuilder (branch_builder.rs:224-246,505-527: all precisions = `precision`,
    ARD weight precisions are vectors of length in_l, output stats = (0, #output weights)).
    Without weights it mirrors the default init of BranchCfgBuilder
    (branch_cfg_builder.rs:180-186: W ~ N(0, 1/m), b = 0; ML precisions :237-283,308-328)."""
    widths = list(hidden_widths) + [summary_width, 1]
    ins = [num_markers] + widths[:-1]
    if weights is None:
        rng = rng or np.random.default_rng(0)
        sd = math.sqrt(1.0 / num_markers)
        weights = [rng.normal(0.0, sd, size=(i, o)).astype(dtype) for i, o in zip(ins, widths)]
        biases = [np.zeros(o, dtype=dtype) for o in widths[:-1]]
    weights = [np.asarray(w, dtype=dtype).reshape(i, o) for w, i, o in zip(weights, ins, widths)]
    biases = [np.asarray(b, dtype=dtype).reshape(o) for b, o in zip(biases, widths[:-1])]
    nl = len(widths)
    if precision is not None:
        p = dtype(precision)
        if is_ard(model):
            wp = [np.full(i, p, dtype=dtype) for i in ins[:-1]] + [np.full(1, p, dtype=dtype)]
            # branch_builder.rs:432,525 builds one vector per entry of `widths` = [m, h.., s]
        else:
            wp = [np.full(1, p, dtype=dtype) for _ in range(nl)]
        bp = [np.full(1, p, dtype=dtype) for _ in range(nl - 1)]
        err = float(p)
    else:
        with np.errstate(divide="ignore"):
            if is_ard(model):  # branch_cfg_builder.rs:308-328 (last layer placeholder 1.0)
                wp = []
                for l in range(nl - 1):
                    ss = np.add.reduce((weights[l] * weights[l]).astype(np.float32), axis=1, dtype=np.float32)
                    wp.append((np.float32(widths[l]) / ss).astype(dtype))
                wp.append(np.ones(1, dtype=dtype))
            else:  # :237-252
                wp = [np.array([np.float32(w.size) / np.float32(np.sum((w * w).astype(np.float32)))], dtype=dtype)
                      for w in weights]
            bp = [np.array([np.float32(b.size) / np.float32(np.sum((b * b).astype(np.float32)))], dtype=dtype)
                  for b in biases]  # :264-274 (zero biases -> +inf)
        err = 2.0  # :394
    return BranchCfg(model=model, num_markers=num_markers, layer_widths=widths, weights=weights,
                     biases=biases, weight_precisions=wp, bias_precisions=bp, error_precision=err,
                     activation=activation, ow_reg_sum=0.0, ow_num_params=summary_width)


# ------------------------------------------------------------------ activations
def _af_sign_neg(x):
    """arrayfire::sign: 1 for negative, else 0 (lasso_ard.rs:328-337)."""
    return (x < 0).astype(x.dtype)


def act_h(name: str, x):
    """activation_functions.rs:23-31."""
    dt = x.dtype
    if name == "tanh":
        return np.tanh(x)
    if name == "relu":
        return x * (x > 0).astype(dt)
    if name == "leaky_relu":
        return x * (x > 0).astype(dt) + x * _af_sign_neg(x) * dt.type(0.01)
    if name == "silu":
        return x * (dt.type(1) / (dt.type(1) + np.exp(-x)))
    if name == "identity":
        return dt.type(1) * x
    raise ValueError(name)


def act_dhdx(name: str, x):
    """activation_functions.rs:33-45 -- evaluated on the PRE-activation."""
    dt = x.dtype
    if name == "tanh":
        t = np.tanh(x)
        return dt.type(1) - t * t
    if name == "relu":
        return (x > 0).astype(dt) * dt.type(1)
    if name == "leaky_relu":
        return (x > 0).astype(dt) * dt.type(1) + _af_sign_neg(x) * dt.type(0.01)
    if name == "silu":
        sg = dt.type(1) / (dt.type(1) + np.exp(-x))
        fx = x * sg
        return fx + sg * (dt.type(1) - fx)
    if name == "identity":
        return np.ones_like(x)
    raise ValueError(name)


def sign3(x):
    """af_helpers.rs:53-58: -1 / 0 / +1."""
    return np.sign(x).astype(x.dtype)


# ------------------------------------------------------------------ branch object
class Branch:
    """Device-side branch of the reference (branch_struct.rs:12-29): built from a cfg,
    own output-weight statistic subtracted from the global one on construction."""

    def __init__(self, cfg: BranchCfg, dtype=np.float32):
        self.dt = np.dtype(dtype)
        c = cfg.astype(dtype)
        self.model = c.model
        self.num_markers = c.num_markers
        self.layer_widths = list(c.layer_widths)
        self.num_layers = len(self.layer_widths)
        self.W = c.weights
        self.b = c.biases
        self.wprec = c.weight_precisions
        self.bprec = c.bias_precisions
        self.eprec = self.dt.type(c.error_precision)
        self.activation = c.activation
        self.ow_num_params = self.dt.type(c.ow_num_params)
        # branch_struct.rs:27 + branch_sampler.rs:1180-1183
        self.ow_reg_sum = self.dt.type(c.ow_reg_sum) - self.summary_stat(self.W[-1])

    # ---- helpers
    @property
    def last(self):
        return self.num_layers - 1

    def f(self, v):
        return self.dt.type(v)

    def sum_sq(self, a):
        """af_helpers.rs:29-39 (dot(flat, flat))."""
        a = np.asarray(a, dtype=self.dt).reshape(-1)
        return self.dt.type(np.dot(a, a))

    def l1(self, a):
        """af_helpers.rs:45-47."""
        return self.dt.type(np.sum(np.abs(np.asarray(a, dtype=self.dt))))

    def summary_stat(self, a):
        """summary_stat_fn (device variant: StdNormal uses sum of squares,
        std_normal_branch.rs:39-41)."""
        return self.l1(a) if is_lasso(self.model) else self.sum_sq(a)

    def param_vec(self):
        parts = [w.reshape(-1, order="F") for w in self.W] + [b.reshape(-1) for b in self.b]
        return np.concatenate(parts)

    def load_param_vec(self, pv):
        pv = np.asarray(pv, dtype=self.dt)
        prev, ix = self.num_markers, 0
        for l, w in enumerate(self.layer_widths):
            n = prev * w
            self.W[l] = pv[ix:ix + n].reshape((prev, w), order="F").copy()
            ix += n
            prev = w
        for l, w in enumerate(self.layer_widths[:-1]):
            self.b[l] = pv[ix:ix + w].copy()
            ix += w

    def split_vec(self, v):
        """param_vec-ordered vector -> (per-layer weight arrays, per-layer bias arrays)."""
        v = np.asarray(v, dtype=self.dt)
        prev, ix, ws, bs = self.num_markers, 0, [], []
        for w in self.layer_widths:
            n = prev * w
            ws.append(v[ix:ix + n].reshape((prev, w), order="F").copy())
            ix += n
            prev = w
        for w in self.layer_widths[:-1]:
            bs.append(v[ix:ix + w].copy())
            ix += w
        return ws, bs

    @staticmethod
    def join_vec(ws, bs):
        return np.concatenate([w.reshape(-1, order="F") for w in ws] + [b.reshape(-1) for b in bs])

    def to_cfg(self) -> BranchCfg:
        """branch_sampler.rs:155-171: own output-weight stat added back for the snapshot."""
        reg = self.ow_reg_sum + self.summary_stat(self.W[-1])
        return BranchCfg(model=self.model, num_markers=self.num_markers,
                         layer_widths=list(self.layer_widths),
                         weights=[w.copy() for w in self.W], biases=[b.copy() for b in self.b],
                         weight_precisions=[p.copy() for p in self.wprec],
                         bias_precisions=[p.copy() for p in self.bprec],
                         error_precision=float(self.eprec), activation=self.activation,
                         ow_reg_sum=float(reg), ow_num_params=int(self.ow_num_params))

    # ---- forward / backward
    def forward_feed(self, x):
        """branch_sampler.rs:743-782. Returns (pre_activations, activations);
        the last activation is y_hat [N,1]."""
        x = np.asarray(x, dtype=self.dt)
        pre, acts = [], []
        inp = x
        for l in range(self.num_layers - 1):
            z = inp @ self.W[l] + self.b[l][None, :]
            a = act_h(self.activation, z)
            pre.append(z)
            acts.append(a)
            inp = a
        acts.append(inp @ self.W[self.last])
        return pre, acts

    def predict(self, x):
        """branch_sampler.rs:915-918."""
        return self.forward_feed(x)[1][-1][:, 0].copy()

    def rss(self, x, y):
        """branch_sampler.rs:905-909."""
        r = self.predict(x) - np.asarray(y, dtype=self.dt)
        return self.dt.type(np.sum(r * r))

    def backpropagate(self, x, y):
        """branch_sampler.rs:813-875. Returns (rss, d_rss_wrt_weights, d_rss_wrt_biases).
        Note Q4: uses e, not 2e."""
        x = np.asarray(x, dtype=self.dt)
        y = np.asarray(y, dtype=self.dt).reshape(-1, 1)
        pre, acts = self.forward_feed(x)
        nl = self.num_layers
        error = acts[-1] - y
        rss = self.dt.type(np.dot(error[:, 0], error[:, 0]))
        gW = [None] * nl
        gb = [None] * (nl - 1)
        gW[nl - 1] = acts[nl - 2].T @ error
        error = error @ self.W[nl - 1].T
        for l in range(nl - 2, -1, -1):
            inp = acts[l - 1] if l > 0 else x
            delta = act_dhdx(self.activation, pre[l]) * error
            gb[l] = np.sum(delta, axis=0)
            gW[l] = (delta.T @ inp).T
            if l > 0:
                error = delta @ self.W[l].T
        return rss, gW, gb

    # ---- densities (non-joint)
    def log_density_wrt_weights(self):
        m = self.model
        ld = self.f(0)
        if m == "std_normal":  # std_normal_branch.rs:132-145
            for l in range(self.num_layers):
                ld = ld - self.f(self.sum_sq(self.W[l]) / self.f(2))
        elif m == "ridge_base":  # ridge_base.rs:159-173
            for l in range(self.num_layers):
                ld = ld - self.f(self.sum_sq(self.W[l]) / self.f(2)) * self.wprec[l][0]
        elif m == "lasso_base":  # lasso_base.rs:160-173
            for l in range(self.num_layers):
                ld = ld - self.l1(self.W[l]) * self.wprec[l][0]
        elif m == "ridge_ard":  # ridge_ard.rs:171-194
            for l in range(self.last):
                rows = np.sum(self.W[l] * self.W[l], axis=1)
                ld = ld - self.f(np.dot(self.f(0.5) * rows, self.wprec[l]))
            ld = ld - self.f(0.5) * self.sum_sq(self.W[self.last]) * self.wprec[self.last][0]
        elif m == "lasso_ard":  # lasso_ard.rs:173-194
            for l in range(self.last):
                rows = np.sum(np.abs(self.W[l]), axis=1)
                ld = ld - self.f(np.dot(rows, self.wprec[l]))
            ld = ld - self.l1(self.W[self.last]) * self.wprec[self.last][0]
        else:
            raise ValueError(m)
        return self.f(ld)

    def log_density_wrt_rss(self, rss):
        """branch_sampler.rs:100-102."""
        return self.f(self.f(-1.0) * self.eprec * self.f(self.f(rss) / self.f(2)))

    def log_density_wrt_biases_l2(self):
        """branch_sampler.rs:115-128 (only used by the joint path / golden test)."""
        ld = self.f(0)
        for l in range(self.last):
            ld = ld - self.bprec[l][0] * self.f(self.sum_sq(self.b[l]) / self.f(2))
        return self.f(ld)

    def log_density(self, rss):
        """branch_sampler.rs:72-78 (bias term 0, :106-112); StdNormal override
        std_normal_branch.rs:147-158 adds -1/2 sum b^2 (Q5)."""
        if self.model == "std_normal":
            ld = self.f(self.f(-0.5) * self.eprec * self.f(rss))
            for l in range(self.num_layers):
                ld = self.f(ld - self.f(0.5) * self.f(np.sum(self.W[l] * self.W[l])))
            for l in range(self.last):
                ld = self.f(ld - self.f(0.5) * self.f(np.sum(self.b[l] * self.b[l])))
            return ld
        wrt_w = self.log_density_wrt_weights()
        wrt_e = self.log_density_wrt_rss(rss)
        wrt_b = self.f(0)
        return self.f(self.f(wrt_w + wrt_b) + wrt_e)

    # ---- gradients (non-joint)
    def ldg_wrt_weights(self, gW):
        m = self.model
        out = []
        for l in range(self.num_layers):
            if m == "std_normal":  # std_normal_branch.rs:160-169
                out.append(-(self.eprec * gW[l] + self.W[l]))
            elif m == "ridge_base":  # ridge_base.rs:175-184
                out.append(-(self.eprec * gW[l] + self.wprec[l][0] * self.W[l]))
            elif m == "lasso_base":  # lasso_base.rs:175-185
                out.append(-(self.eprec * gW[l] + self.wprec[l][0] * sign3(self.W[l])))
            elif m == "ridge_ard":  # ridge_ard.rs:196-219
                if l < self.last:
                    out.append(-(self.eprec * gW[l] + self.wprec[l][:, None] * self.W[l]))
                else:
                    out.append(-(self.eprec * gW[l] + self.wprec[l][0] * self.W[l]))
            elif m == "lasso_ard":  # lasso_ard.rs:196-218
                if l < self.last:
                    out.append(-(self.eprec * gW[l] + self.wprec[l][:, None] * sign3(self.W[l])))
                else:
                    out.append(-(self.eprec * gW[l] + self.wprec[l][0] * sign3(self.W[l])))
        return [o.astype(self.dt) for o in out]

    def ldg_wrt_biases(self, gb):
        """branch_sampler.rs:322-331: no prior term."""
        return [(self.f(-1.0) * self.eprec * g).astype(self.dt) for g in gb]

    def ldg_wrt_biases_l2(self, gb):
        """branch_sampler.rs:334-345."""
        return [(self.f(-1.0) * self.bprec[l][0] * self.b[l] - self.eprec * gb[l]).astype(self.dt)
                for l in range(self.last)]

    def log_density_gradient(self, x, y):
        """branch_sampler.rs:380-391. Returns (rss, ldg_w list, ldg_b list)."""
        rss, gW, gb = self.backpropagate(x, y)
        return rss, self.ldg_wrt_weights(gW), self.ldg_wrt_biases(gb)

    # ---- joint densities / gradients (LPD bookkeeping + golden pins)
    def log_density_joint_wrt_local_weights(self, hyper: Hyper):
        nl = self.num_layers
        ld = self.f(0)
        m = self.model
        for l in range(self.last):
            shape, scale = (self.f(v) for v in hyper.layer_prior(l, nl))
            W, lam = self.W[l], self.wprec[l]
            if m == "ridge_base":  # ridge_base.rs:117-136
                ld = ld - self.f(self.sum_sq(W) / self.f(2) + self.f(1) / scale) * lam[0]
                ld = ld + self.f(shape + self.f(self.f(W.size) - self.f(2)) / self.f(2)) * np.log(lam[0])
            elif m == "lasso_base":  # lasso_base.rs:119-138
                ld = ld - self.f(self.l1(W) + self.f(1) / scale) * lam[0]
                ld = ld + self.f(shape + self.f(W.size) - self.f(1)) * np.log(lam[0])
            elif m == "ridge_ard":  # ridge_ard.rs:119-148
                rows = np.sum(W * W, axis=1)
                ld = ld - self.f(np.dot(rows / self.f(2) + self.f(1) / scale, lam))
                ncols = self.f(W.shape[1])
                ld = ld + self.f(np.dot(self.f(shape + (ncols - self.f(2)) / self.f(2)) * np.ones(W.shape[0], dtype=self.dt),
                                        np.log(lam)))
            elif m == "lasso_ard":  # lasso_ard.rs:123-151
                rows = np.sum(np.abs(W), axis=1)
                ld = ld - self.f(np.dot(rows + self.f(1) / scale, lam))
                ncols = self.f(W.shape[1])
                ld = ld + self.f(np.dot(self.f(shape + ncols - self.f(1)) * np.ones(W.shape[0], dtype=self.dt),
                                        np.log(lam)))
            else:
                raise NotImplementedError("joint density unimplemented for std_normal (Q6)")
        return self.f(ld)

    def log_density_joint_wrt_output_weights(self, hyper: Hyper):
        if self.model == "std_normal":
            raise NotImplementedError("joint density unimplemented for std_normal (Q6)")
        l = self.last
        shape, scale = (self.f(v) for v in hyper.layer_prior(l, self.num_layers))
        lam = self.wprec[l][0]
        ld = self.f(0)
        if is_lasso(self.model):  # lasso_base.rs:140-158, lasso_ard.rs:153-171
            g = self.l1(self.W[l]) + self.ow_reg_sum
            ld = ld - self.f(g + self.f(1) / scale) * lam
            ld = ld + self.f(shape + self.ow_num_params - self.f(1)) * np.log(lam)
        else:  # ridge_base.rs:138-157, ridge_ard.rs:150-169
            g = self.sum_sq(self.W[l]) + self.ow_reg_sum
            ld = ld - self.f(self.f(0.5) * g + self.f(1) / scale) * lam
            ld = ld + self.f(shape + self.f(self.ow_num_params - self.f(2)) / self.f(2)) * np.log(lam)
        return self.f(ld)

    def log_density_joint_wrt_weights(self, hyper):
        """branch_sampler.rs:229-237."""
        return self.f(self.log_density_joint_wrt_local_weights(hyper) + self.log_density_joint_wrt_output_weights(hyper))

    def log_density_joint_wrt_rss(self, rss, hyper: Hyper, n: int):
        """branch_sampler.rs:240-257."""
        shape, scale = (self.f(v) for v in hyper.output)
        ld = self.f(0)
        ld = ld + self.f(shape + self.f(self.f(n) - self.f(2)) / self.f(2)) * np.log(self.eprec)
        ld = ld - self.eprec * self.f(self.f(rss) / self.f(2) + self.f(1) / scale)
        return self.f(ld)

    def log_density_joint_wrt_biases(self, hyper: Hyper):
        """branch_sampler.rs:260-279."""
        ld = self.f(0)
        for l in range(self.last):
            shape, scale = (self.f(v) for v in hyper.layer_prior(l, self.num_layers))
            ld = ld - self.bprec[l][0] * self.f(self.sum_sq(self.b[l]) / self.f(2) + self.f(1) / scale)
            nvar = self.f(self.b[l].size)
            ld = ld + self.f(shape + self.f(nvar - self.f(2)) / self.f(2)) * np.log(self.bprec[l][0])
        return self.f(ld)

    def log_density_joint(self, rss, hyper, n):
        """branch_sampler.rs:292-305."""
        w = self.log_density_joint_wrt_weights(hyper)
        e = self.log_density_joint_wrt_rss(rss, hyper, n)
        b = self.log_density_joint_wrt_biases(hyper)
        return self.f(self.f(w + b) + e)

    def log_density_joint_components_curr_internal_state(self, hyper):
        """branch_sampler.rs:307-318 -> (wrt_output_weights, wrt_local_params)."""
        out_w = self.log_density_joint_wrt_output_weights(hyper)
        local = self.f(self.log_density_joint_wrt_biases(hyper) + self.log_density_joint_wrt_local_weights(hyper))
        return out_w, local

    def ldg_wrt_weight_precisions(self, hyper: Hyper):
        nl, m = self.num_layers, self.model
        out = []
        for l in range(self.last):
            shape, scale = (self.f(v) for v in hyper.layer_prior(l, nl))
            W, lam = self.W[l], self.wprec[l]
            if m == "ridge_base":  # ridge_base.rs:186-200
                out.append((self.f(2) * shape + self.f(W.size) - self.f(2)) / (self.f(2) * lam)
                           - self.f(1) / scale - self.sum_sq(W) / self.f(2))
            elif m == "lasso_base":  # lasso_base.rs:187-201
                out.append((shape + self.f(W.size) - self.f(1)) / lam - self.f(1) / scale - self.l1(W))
            elif m == "ridge_ard":  # ridge_ard.rs:221-236 (Q8: precisions.elements() == #rows)
                out.append((self.f(2) * shape + self.f(lam.size) - self.f(2)) / (self.f(2) * lam)
                           - self.f(1) / scale - np.sum(W * W, axis=1) / self.f(2))
            elif m == "lasso_ard":  # lasso_ard.rs:220-234
                out.append((shape + self.f(lam.size) - self.f(1)) / lam - self.f(1) / scale
                           - np.sum(np.abs(W), axis=1))
            else:
                raise NotImplementedError
        l = self.last
        shape, scale = (self.f(v) for v in hyper.layer_prior(l, nl))
        lam = self.wprec[l]
        if is_lasso(m):
            out.append((shape + self.ow_num_params - self.f(1)) / lam - self.f(1) / scale
                       - (self.l1(self.W[l]) + self.ow_reg_sum))
        else:
            out.append((self.f(2) * shape + self.ow_num_params - self.f(2)) / (self.f(2) * lam)
                       - self.f(1) / scale - (self.sum_sq(self.W[l]) + self.ow_reg_sum) / self.f(2))
        return [np.asarray(o, dtype=self.dt).reshape(-1) for o in out]

    def ldg_wrt_bias_precisions(self, hyper: Hyper):
        """branch_sampler.rs:348-367."""
        out = []
        for l in range(self.last):
            shape, scale = (self.f(v) for v in hyper.layer_prior(l, self.num_layers))
            nvar = self.f(self.b[l].size)
            out.append(self.f((self.f(2) * shape + (nvar - self.f(2))) / (self.f(2) * self.bprec[l][0])
                              - self.f(1) / scale - self.sum_sq(self.b[l]) / self.f(2)))
        return out

    def ldg_wrt_error_precision(self, last_rss, n: int, hyper: Hyper):
        """branch_sampler.rs:369-378."""
        shape, scale = (self.f(v) for v in hyper.output)
        return self.f((self.f(2) * shape + self.f(n) - self.f(2)) / (self.f(2) * self.eprec)
                      - self.f(1) / scale - self.f(last_rss) / self.f(2))

    # ---- step sizes
    def step_sizes(self, cfg: MCMCCfg, uniforms=None):
        """Per-parameter step sizes as (weights list, biases list).
        uniform: branch_sampler.rs:706-732; random: :654-681 (uniforms injected, param_vec order);
        std_scaled: ridge_base.rs:52-80 (ARD variants return empty vectors, ridge_ard.rs:56-68);
        izmailov: ridge_base.rs:82-115, ridge_ard.rs:70-117, lasso_base.rs:84-117,
        lasso_ard.rs:77-121, std_normal_branch.rs:81-112 (no factor, Q7)."""
        f = self.f(cfg.hmc_step_size_factor)
        L = self.f(cfg.hmc_integration_length)
        pi = self.f(np.float32(np.pi)) if self.dt == np.float32 else self.f(np.float64(np.float32(np.pi)))
        mode = cfg.hmc_step_size_mode
        nl = self.num_layers
        ws, bs = [], []
        with np.errstate(divide="ignore", invalid="ignore"):
            if mode == "uniform":
                ws = [np.full(w.shape, f, dtype=self.dt) for w in self.W]
                bs = [np.full(b.shape, f, dtype=self.dt) for b in self.b]
            elif mode == "random":
                assert uniforms is not None, "random step sizes need injected uniforms"
                prop = self.f(self.f(self.param_vec().size) ** self.f(-0.25)) * f
                uw, ub = self.split_vec(uniforms)
                ws = [(u * prop).astype(self.dt) for u in uw]
                bs = [(u * prop).astype(self.dt) for u in ub]
            elif mode == "std_scaled":
                if is_ard(self.model):
                    raise IndexError("ARD std_scaled_step_sizes returns empty vectors in the reference")
                for l in range(nl):
                    v = f * np.sqrt(self.f(1) / self.wprec[l][0])
                    ws.append(np.full(self.W[l].shape, v, dtype=self.dt))
                for l in range(nl - 1):
                    v = f * (self.f(1) / np.sqrt(self.bprec[l][0]))
                    bs.append(np.full(self.b[l].shape, v, dtype=self.dt))
            elif mode == "izmailov":
                m = self.model
                for l in range(nl):
                    lam = self.wprec[l]
                    ard_layer = is_ard(m) and l < self.last
                    if m == "std_normal":
                        v = pi / (self.f(2) * np.sqrt(lam[0]) * L)
                    elif is_lasso(m):
                        if ard_layer:
                            v = f * (self.f(1) / (self.f(4) * lam * L))
                        else:
                            v = f / (self.f(4) * lam[0] * L) if m == "lasso_base" else (f * self.f(1)) / (self.f(4) * lam[0] * L)
                    else:
                        if ard_layer:
                            v = f * (pi / (self.f(2) * np.sqrt(lam) * L))
                        else:
                            v = (f * pi) / (self.f(2) * np.sqrt(lam[0]) * L)
                    if ard_layer:
                        ws.append(np.repeat(np.asarray(v, dtype=self.dt)[:, None], self.W[l].shape[1], axis=1))
                    else:
                        ws.append(np.full(self.W[l].shape, v, dtype=self.dt))
                for l in range(nl - 1):
                    core = pi / (self.f(2) * np.sqrt(self.bprec[l][0]) * L)
                    if m == "std_normal":
                        v = self.f(1) * core
                    elif m == "lasso_ard":
                        v = self.f(1) * ((f * pi) / (self.f(2) * np.sqrt(self.bprec[l][0]) * L))
                    else:
                        v = f * core
                    bs.append(np.full(self.b[l].shape, v, dtype=self.dt))
            else:
                raise ValueError(mode)
        return ws, bs

    # ---- HMC
    @staticmethod
    def kinetic(pw, pb, dt):
        """momentum.rs:147-158: 0.5 * (sum over arrays of sum_all(p*p))."""
        acc = dt.type(0)
        for a in pw:
            acc = dt.type(acc + dt.type(np.sum(a * a)))
        for a in pb:
            acc = dt.type(acc + dt.type(np.sum(a * a)))
        return dt.type(dt.type(0.5) * acc)

    def neg_hamiltonian(self, pw, pb, x, y):
        """branch_sampler.rs:878-883."""
        return self.f(self.log_density(self.rss(x, y)) - self.kinetic(pw, pb, self.dt))

    def net_movement(self, init_W, init_b, pw, pb):
        """branch_sampler.rs:551-588: (theta - theta0) . p."""
        acc = self.f(0)
        for l in range(self.num_layers):
            acc = acc + self.f(np.sum((self.W[l] - init_W[l]) * pw[l]))
        for l in range(self.last):
            acc = acc + self.f(np.sum((self.b[l] - init_b[l]) * pb[l]))
        return self.f(acc)

    def hmc_step(self, x, y, cfg: MCMCCfg, momenta, u, step_uniforms=None, record=False):
        """One HMC transition, branch_sampler.rs:1192-1299 + :928-962.
        momenta: param_vec-ordered N(0,1) draws; u: the f32 uniform of is_accepted (:546-548).
        Returns dict(status, log_density, y_pred, h_init, h_final, steps_done, u_turn_step,
        traj (if record): lists of params / ldg / hamiltonian per step as in trajectory.rs)."""
        dt = self.dt
        x = np.asarray(x, dtype=dt)
        y = np.asarray(y, dtype=dt)
        init_W = [w.copy() for w in self.W]
        init_b = [b.copy() for b in self.b]
        ew, eb = self.step_sizes(cfg, step_uniforms)
        pw, pb = self.split_vec(momenta)
        h_init = self.neg_hamiltonian(pw, pb, x, y)
        traj = dict(params=[], ldg=[], hamiltonian=[float(h_init)])
        _, gw, gb = self.log_density_gradient(x, y)
        u_turn_step = -1
        half = dt.type(0.5)
        h_curr = h_init
        for step in range(cfg.hmc_integration_length):
            for l in range(self.num_layers):  # momentum.rs:129-136
                pw[l] = (pw[l] + half * ew[l] * gw[l]).astype(dt)
            for l in range(self.last):
                pb[l] = (pb[l] + eb[l] * half * gb[l]).astype(dt)
            for l in range(self.num_layers):  # params.rs:728-738
                self.W[l] = (self.W[l] + ew[l] * pw[l]).astype(dt)
            for l in range(self.last):
                self.b[l] = (self.b[l] + eb[l] * pb[l]).astype(dt)
            _, gw, gb = self.log_density_gradient(x, y)
            for l in range(self.num_layers):
                pw[l] = (pw[l] + half * ew[l] * gw[l]).astype(dt)
            for l in range(self.last):
                pb[l] = (pb[l] + eb[l] * half * gb[l]).astype(dt)
            h_curr = self.neg_hamiltonian(pw, pb, x, y)
            if record:
                traj["params"].append(self.param_vec().astype(np.float64))
                traj["ldg"].append(self.join_vec(gw, gb).astype(np.float64))
                traj["hamiltonian"].append(float(h_curr))
            if abs(h_curr - h_init) > dt.type(cfg.hmc_max_hamiltonian_error):  # :1264-1279
                self.W, self.b = init_W, init_b
                return dict(status=REJECTED_EARLY, log_density=None, y_pred=None, h_init=float(h_init),
                            h_final=float(h_curr), steps_done=step + 1, u_turn_step=u_turn_step, traj=traj)
            if u_turn_step < 0 and self.net_movement(init_W, init_b, pw, pb) < 0:  # :1281-1284
                u_turn_step = step
        # accept_or_reject_hmc_state, :928-962
        y_pred = self.predict(x)
        r = y_pred - y
        rss = dt.type(np.sum(r * r))
        log_density = self.log_density(rss)
        h_final = dt.type(log_density - self.kinetic(pw, pb, dt))
        log_acc = dt.type(h_final - h_init)
        with np.errstate(over="ignore", invalid="ignore"):
            acc_prob = dt.type(1) if log_acc >= 0 else np.exp(log_acc)
        accepted = bool(dt.type(u) < acc_prob)
        out = dict(h_init=float(h_init), h_final=float(h_final), steps_done=cfg.hmc_integration_length,
                   u_turn_step=u_turn_step, traj=traj, log_acc=float(log_acc))
        if accepted:
            out.update(status=ACCEPTED, log_density=float(log_density), y_pred=y_pred)
        else:
            self.W, self.b = init_W, init_b  # :1293-1296
            out.update(status=REJECTED, log_density=None, y_pred=None)
        return out

    def numerical_ldg(self, x, y):
        """branch_sampler.rs:480-504 (NUMERICAL_DELTA = 0.001, :30): forward differences of log_density, the perturbed
        vector walked exactly as the reference does (+= delta, evaluate, -= delta), original parameters reloaded at the end."""
        dt = self.dt
        x = np.asarray(x, dtype=dt)
        y = np.asarray(y, dtype=dt)
        delta = dt.type(0.001)
        curr_pv = self.param_vec().astype(dt).copy()
        next_pv = curr_pv.copy()
        curr_ld = dt.type(self.log_density(self.rss(x, y)))
        res = []
        for pix in range(curr_pv.size):
            next_pv[pix] = dt.type(next_pv[pix] + delta)
            self.load_param_vec(next_pv)
            res.append(dt.type(dt.type(dt.type(self.log_density(self.rss(x, y))) - curr_ld) / delta))
            next_pv[pix] = dt.type(next_pv[pix] - delta)
        self.load_param_vec(curr_pv)
        return np.array(res, dtype=dt)

    def effect_sizes(self, x):
        """branch_sampler.rs:784-811: back-propagation of the prediction to the (standardised) input, seeded with
        yhat W_last^T (the reference multiplies by the prediction itself).  Returns [n, m]."""
        x = np.asarray(x, dtype=self.dt)
        pre, acts = self.forward_feed(x)
        error = acts[-1] @ self.W[self.last].T
        for l in range(self.num_layers - 2, -1, -1):
            delta = act_dhdx(self.activation, pre[l]) * error
            error = delta @ self.W[l].T
        return error

    # ---- joint HMC / gradient descent (flag-gated modes)
    def precision_vec(self):
        """BranchPrecisions::param_vec (params.rs:272-289)."""
        return np.concatenate([np.asarray(p, dtype=self.dt).reshape(-1) for p in self.wprec]
                              + [np.asarray(p, dtype=self.dt).reshape(-1) for p in self.bprec]
                              + [np.array([self.eprec], dtype=self.dt)])

    def load_precision_vec(self, v):
        v = np.asarray(v, dtype=self.dt)
        ix = 0
        for l in range(self.num_layers):
            n = self.wprec[l].size
            self.wprec[l] = v[ix:ix + n].copy()
            ix += n
        for l in range(self.last):
            self.bprec[l] = v[ix:ix + 1].copy()
            ix += 1
        self.eprec = self.dt.type(v[ix])

    def split_precision_vec(self, v):
        """precision param_vec-ordered vector -> (per-layer weight-precision arrays, per-layer bias-precision arrays, error)."""
        v = np.asarray(v, dtype=self.dt)
        ix, wp, bp = 0, [], []
        for l in range(self.num_layers):
            n = self.wprec[l].size
            wp.append(v[ix:ix + n].copy())
            ix += n
        for l in range(self.last):
            bp.append(v[ix:ix + 1].copy())
            ix += 1
        return wp, bp, self.dt.type(v[ix])

    def log_density_gradient_joint(self, x, y, hyper: Hyper):
        """branch_sampler.rs:406-422: backpropagate, weights under the prior, l2-regularised biases (:334-345),
        precisions of weights (per prior), biases (:348-367) and error (:369-378, last_rss of this pass).
        Returns (rss, gw, gb, gwp, gbp, gep)."""
        rss, gW, gb = self.backpropagate(x, y)
        n = np.asarray(y).size
        return (rss, self.ldg_wrt_weights(gW), self.ldg_wrt_biases_l2(gb), self.ldg_wrt_weight_precisions(hyper),
                [np.asarray(v, dtype=self.dt).reshape(-1) for v in self.ldg_wrt_bias_precisions(hyper)],
                self.ldg_wrt_error_precision(rss, n, hyper))

    @staticmethod
    def join_joint_vec(gw, gb, gwp, gbp, gep):
        """BranchLogDensityGradientJoint::param_vec order (gradient.rs:66-97): weights, biases, weight precisions,
        bias precisions, error precision."""
        return np.concatenate([w.reshape(-1, order="F") for w in gw] + [b.reshape(-1) for b in gb]
                              + [p.reshape(-1) for p in gwp] + [p.reshape(-1) for p in gbp]
                              + [np.asarray(gep).reshape(1)])

    def kinetic_joint(self, pw, pb, pwp, pbp, pep):
        """momentum.rs:83-104."""
        dt = self.dt
        acc = dt.type(0)
        for grp in (pw, pb, pwp, pbp):
            for a in grp:
                acc = dt.type(acc + dt.type(np.sum(a * a)))
        acc = dt.type(acc + dt.type(pep * pep))
        return dt.type(dt.type(0.5) * acc)

    def hmc_step_joint(self, x, y, cfg: MCMCCfg, hyper: Hyper, momenta, u, step_uniforms, record=False):
        """hmc_step_joint (branch_sampler.rs:1070-1178).  momenta / step_uniforms: P + Q draws, parameters in param_vec
        order followed by the precisions in their param_vec order.  Step sizes are always Random with the joint
        proportionality factor (P + Q)^(-1/4) * f (:654-704).  The final accept uses the NON-joint log density
        (accept_or_reject_hmc_state, :928-962) against the joint initial Hamiltonian -- as the reference does."""
        dt = self.dt
        x = np.asarray(x, dtype=dt)
        y = np.asarray(y, dtype=dt)
        n = y.size
        P = self.param_vec().size
        Q = self.precision_vec().size
        momenta = np.asarray(momenta, dtype=dt)
        step_uniforms = np.asarray(step_uniforms, dtype=dt)
        init_theta, init_prec = self.param_vec().copy(), self.precision_vec().copy()
        prop = self.f(self.f(self.f(P) + self.f(Q)) ** self.f(-0.25)) * self.f(cfg.hmc_step_size_factor)
        ew, eb = self.split_vec((step_uniforms[:P] * prop).astype(dt))
        ewp, ebp, eep = self.split_precision_vec((step_uniforms[P:] * prop).astype(dt))
        pw, pb = self.split_vec(momenta[:P])
        pwp, pbp, pep = self.split_precision_vec(momenta[P:])

        def neg_h():  # neg_hamiltonian_joint, :886-903
            return self.f(self.log_density_joint(self.rss(x, y), hyper, n) - self.kinetic_joint(pw, pb, pwp, pbp, pep))

        h_init = neg_h()
        traj = dict(params=[], precisions=[], ldg=[], hamiltonian=[float(h_init)])
        _, gw, gb, gwp, gbp, gep = self.log_density_gradient_joint(x, y, hyper)
        half = dt.type(0.5)
        h_curr = h_init

        def half_step():  # momentum.rs:32-58
            nonlocal pep
            for l in range(self.num_layers):
                pw[l] = (pw[l] + half * ew[l] * gw[l]).astype(dt)
            for l in range(self.last):
                pb[l] = (pb[l] + eb[l] * half * gb[l]).astype(dt)
            for l in range(self.num_layers):
                pwp[l] = (pwp[l] + ewp[l] * half * gwp[l]).astype(dt)
            for l in range(self.last):
                pbp[l] = (pbp[l] + ebp[l] * half * gbp[l]).astype(dt)
            pep = dt.type(pep + eep * half * gep)

        err_state = np.errstate(all="ignore")   # precisions may cross zero / blow up: NaN and inf propagate as in the reference
        err_state.__enter__()
        for step in range(cfg.hmc_integration_length):
            half_step()
            for l in range(self.num_layers):  # params.rs:728-738
                self.W[l] = (self.W[l] + ew[l] * pw[l]).astype(dt)
            for l in range(self.last):
                self.b[l] = (self.b[l] + eb[l] * pb[l]).astype(dt)
            for l in range(self.num_layers):  # params.rs:344-355
                self.wprec[l] = (self.wprec[l] + ewp[l] * pwp[l]).astype(dt)
            for l in range(self.last):
                self.bprec[l] = (self.bprec[l] + ebp[l] * pbp[l]).astype(dt)
            self.eprec = dt.type(self.eprec + eep * pep)
            _, gw, gb, gwp, gbp, gep = self.log_density_gradient_joint(x, y, hyper)
            half_step()
            h_curr = neg_h()
            if record:
                traj["params"].append(self.param_vec().astype(np.float64))
                traj["precisions"].append(self.precision_vec().astype(np.float64))
                traj["ldg"].append(self.join_joint_vec(gw, gb, gwp, gbp, gep).astype(np.float64))
                traj["hamiltonian"].append(float(h_curr))
            if not (abs(h_curr - h_init) <= dt.type(cfg.hmc_max_hamiltonian_error)):  # :1146-1162 (NaN: `>` is false in Rust)
                if not np.isnan(h_curr):
                    self.load_param_vec(init_theta)
                    self.load_precision_vec(init_prec)
                    err_state.__exit__(None, None, None)
                    return dict(status=REJECTED_EARLY, log_density=None, y_pred=None, h_init=float(h_init),
                                h_final=float(h_curr), steps_done=step + 1, traj=traj)
        err_state.__exit__(None, None, None)
        with np.errstate(all="ignore"):
            y_pred = self.predict(x)
            r = y_pred - y
            rss = dt.type(np.sum(r * r))
            log_density = self.log_density(rss)
            h_final = dt.type(log_density - self.kinetic_joint(pw, pb, pwp, pbp, pep))
            log_acc = dt.type(h_final - h_init)
            acc_prob = dt.type(1) if log_acc >= 0 else np.exp(log_acc)
        accepted = bool(dt.type(u) < acc_prob)
        out = dict(h_init=float(h_init), h_final=float(h_final), steps_done=cfg.hmc_integration_length, traj=traj,
                   log_acc=float(log_acc))
        if accepted:
            out.update(status=ACCEPTED, log_density=float(log_density), y_pred=y_pred)
        else:
            self.load_param_vec(init_theta)
            self.load_precision_vec(init_prec)
            out.update(status=REJECTED, log_density=None, y_pred=None)
        return out

    def _descend(self, step, gw, gb):
        """BranchParams::descend_gradient (params.rs:740-749)."""
        s = self.f(step)
        for l in range(self.num_layers):
            self.W[l] = (self.W[l] + s * gw[l]).astype(self.dt)
        for l in range(self.last):
            self.b[l] = (self.b[l] + s * gb[l]).astype(self.dt)

    def gradient_descent(self, x, y, cfg: MCMCCfg):
        """gradient_descent (branch_sampler.rs:964-1003): `hmc_integration_length` ascent steps on the log density, each
        with the doubling / halving line search on the RSS of probe steps (:1005-1017).  Always Accepted.
        Returns dict(status, log_density, y_pred, step_sizes (the accepted step size of every iteration), num_probes)."""
        dt = self.dt
        x = np.asarray(x, dtype=dt)
        y = np.asarray(y, dtype=dt)

        def probe(gw, gb, s):
            keep = self.param_vec().copy()
            self._descend(s, gw, gb)
            res = self.rss(x, y)
            self.load_param_vec(keep)
            return res

        _, gw, gb = self.log_density_gradient(x, y)
        taken, nprobe = [], 0
        for _ in range(cfg.hmc_integration_length):
            step = self.f(cfg.hmc_step_size_factor)
            prev = probe(gw, gb, step)
            fac = self.f(2.0) if probe(gw, gb, self.f(2.0) * step) < prev else self.f(0.5)
            step = self.f(step * fac)
            curr = probe(gw, gb, step)
            nprobe += 3
            while curr < prev:
                prev = curr
                step = self.f(step * fac)
                curr = probe(gw, gb, step)
                nprobe += 1
            step = self.f(step / fac)
            self._descend(step, gw, gb)
            taken.append(float(step))
            _, gw, gb = self.log_density_gradient(x, y)
        y_pred = self.predict(x)
        r = y_pred - y
        rss = dt.type(np.sum(r * r))
        return dict(status=ACCEPTED, log_density=float(self.log_density(rss)), y_pred=y_pred, step_sizes=taken,
                    num_probes=nprobe)

    def gradient_descent_joint(self, x, y, cfg: MCMCCfg, hyper: Hyper):
        """gradient_descent_joint (branch_sampler.rs:1019-1066): fixed step size `hmc_step_size_factor` on parameters and
        precisions; Rejected (state restored) when the error precision ends <= 0."""
        dt = self.dt
        x = np.asarray(x, dtype=dt)
        y = np.asarray(y, dtype=dt)
        init_theta, init_prec = self.param_vec().copy(), self.precision_vec().copy()
        s = self.f(cfg.hmc_step_size_factor)
        with np.errstate(all="ignore"):
            _, gw, gb, gwp, gbp, gep = self.log_density_gradient_joint(x, y, hyper)
            for _ in range(cfg.hmc_integration_length):
                self._descend(s, gw, gb)
                for l in range(self.num_layers):  # params.rs:357-367
                    self.wprec[l] = (self.wprec[l] + s * gwp[l]).astype(dt)
                for l in range(self.last):
                    self.bprec[l] = (self.bprec[l] + s * gbp[l]).astype(dt)
                self.eprec = dt.type(self.eprec + s * gep)
                _, gw, gb, gwp, gbp, gep = self.log_density_gradient_joint(x, y, hyper)
            y_pred = self.predict(x)
            r = y_pred - y
            rss = dt.type(np.sum(r * r))
            log_density = self.log_density_joint(rss, hyper, y.size)
        if not (self.eprec > 0):  # `<= 0.0` in the reference; NaN compares false there -> Accepted
            if not np.isnan(self.eprec):
                self.load_param_vec(init_theta)
                self.load_precision_vec(init_prec)
                return dict(status=REJECTED, log_density=None, y_pred=None)
        return dict(status=ACCEPTED, log_density=float(log_density), y_pred=y_pred)

    # ---- Gibbs precision draws (standard-gamma variates injected through `gam`)
    def _ridge_post(self, k, s, stat, n, gam):
        """gibbs_steps.rs:76-94,115-129: Gamma(k + n/2, 2s / (2 + s*stat))."""
        shape = self.f(self.f(k) + self.f(n) / self.f(2))
        scale = self.f(self.f(2) * self.f(s) / (self.f(2) + self.f(s) * self.f(stat)))
        return self.f(self.f(gam(float(shape))) * scale)

    def _lasso_post(self, k, s, stat, n, gam):
        """gibbs_steps.rs:25-57: Gamma(k + n, s / (1 + s*stat))."""
        shape = self.f(self.f(k) + self.f(n))
        scale = self.f(self.f(s) / (self.f(1) + self.f(s) * self.f(stat)))
        return self.f(self.f(gam(float(shape))) * scale)

    def sample_error_precision(self, residual, hyper: Hyper, gam):
        """branch_sampler.rs:190-202 (Q9: output-layer hyperparameters)."""
        k, s = hyper.output
        self.eprec = self._ridge_post(k, s, self.sum_sq(residual), np.asarray(residual).size, gam)

    def sample_prior_precisions(self, hyper: Hyper, gam):
        """ridge_base.rs:235-253, ridge_ard.rs:271-301, lasso_base.rs:235-258,
        lasso_ard.rs:268-311, std_normal_branch.rs:189 (no-op)."""
        m = self.model
        if m == "std_normal":
            return
        for l in range(self.last):
            k, s = hyper.layer_prior(l, self.num_layers)
            W = self.W[l]
            if m == "ridge_base":
                self.wprec[l] = np.array([self._ridge_post(k, s, self.sum_sq(W), W.size, gam)], dtype=self.dt)
            elif m == "lasso_base":
                self.wprec[l] = np.array([self._lasso_post(k, s, self.l1(W), W.size, gam)], dtype=self.dt)
            elif m == "ridge_ard":
                rows = np.sum(W * W, axis=1)
                gsz = self.layer_widths[l]
                self.wprec[l] = np.array([self._ridge_post(k, s, r, gsz, gam) for r in rows], dtype=self.dt)
            elif m == "lasso_ard":
                rows = np.sum(np.abs(W), axis=1)
                gsz = self.layer_widths[l]
                self.wprec[l] = np.array([self._lasso_post(k, s, r, gsz, gam) for r in rows], dtype=self.dt)
            self.bprec[l] = np.array([self._ridge_post(k, s, self.sum_sq(self.b[l]), self.b[l].size, gam)],
                                     dtype=self.dt)

    def sample_output_weight_precisions(self, hyper: Hyper, gam):
        """branch_sampler.rs:178-188 + precision_posterior_host of each prior."""
        if self.model == "std_normal":
            self.wprec[self.last] = np.array([1.0], dtype=self.dt)  # std_normal_branch.rs:178-186
            return
        k, s = hyper.output
        stat = self.f(self.ow_reg_sum + self.summary_stat(self.W[self.last]))
        n = int(self.ow_num_params)
        post = self._lasso_post if is_lasso(self.model) else self._ridge_post
        self.wprec[self.last] = np.array([post(k, s, stat, n, gam)], dtype=self.dt)

    def sample_param_precisions(self, hyper, gam):
        """branch_sampler.rs:173-176."""
        self.sample_prior_precisions(hyper, gam)
        self.sample_output_weight_precisions(hyper, gam)

    def gibbs_shapes(self, hyper: Hyper, n: int, fixed_param_precisions=False):
        """Shapes of the Gamma draws of one visit in consumption order (for injection)."""
        shapes = []
        rec = lambda sh: (shapes.append(sh), 1.0)[1]
        b = copy.deepcopy(self)
        b.sample_error_precision(np.zeros(n, dtype=self.dt), hyper, rec)
        if not fixed_param_precisions:
            b.sample_param_precisions(hyper, rec)
        return shapes
