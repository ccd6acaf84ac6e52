"""Net construction on the host: BlockNetCfg + BranchCfgBuilder of the reference
(net/architectures.rs:58-237, net/branch/branch_cfg_builder.rs:180-186,237-398).

The reference draws the initial weights from `thread_rng()` (unseedable); here a numpy Generator is used, the
distributions are the same: W ~ N(0, 1/m_b) (or N(0, v)), biases 0, maximum-likelihood precisions."""
from typing import List, Optional, Sequence

import numpy as np

from .files import BranchCfgFile, NetFile

DEFAULT_INIT_OUTPUT_LAYER_PRECISION = 0.05      # architectures.rs:16
MODEL_TYPES = ["ridge_ard", "ridge_base", "lasso_ard", "lasso_base", "std_normal"]


def is_ard(model: str) -> bool:
    return model.endswith("ard")


def hidden_width(num_markers: int, fixed: Optional[int], fraction: float) -> int:
    """HiddenLayerWidthRule (architectures.rs:93-101): Fixed(w) | FractionOfInput(f), never below 1."""
    return int(fixed) if fixed is not None else max(int(np.float32(num_markers) * np.float32(fraction)), 1)


def summary_width(hidden: int, fixed: Optional[int], fraction: Optional[float]) -> int:
    """SummaryLayerWidthRule (architectures.rs:103-118): Fixed | FractionOfHiddenLayerWidth | LikeHiddenLayerWidth."""
    if fixed is not None:
        if fixed == 0:
            raise ValueError("Branch cannot be initiated with summary layer width = 0.")
        return int(fixed)
    if fraction is None:
        return hidden
    return max(int(np.float32(hidden) * np.float32(fraction)), 1)


def _ss(a) -> np.float32:
    a = np.asarray(a, dtype=np.float32)
    return np.float32(np.sum(a * a, dtype=np.float32))


def build_branch_cfg(model: str, num_markers: int, depth: int, hidden: int, summary: int, activation: str,
                     rng: np.random.Generator, fixed_param_precision: Optional[float] = None,
                     init_param_variance: Optional[float] = None) -> BranchCfgFile:
    """BranchCfgBuilder::build_base / build_ard (branch_cfg_builder.rs:338-398)."""
    widths = [hidden] * depth + [summary, 1]
    ins = [num_markers] + widths[:-1]
    var = init_param_variance if init_param_variance is not None else 1.0 / num_markers   # :180-186
    weights = [rng.normal(0.0, np.sqrt(var), size=i * o).astype(np.float32) for i, o in zip(ins, widths)]   # column-major [in x out]
    biases = [np.zeros(o, dtype=np.float32) for o in widths[:-1]]
    if init_param_variance is not None:
        biases = [rng.normal(0.0, np.sqrt(var), size=o).astype(np.float32) for o in widths[:-1]]
    nl = len(widths)
    with np.errstate(divide="ignore"):
        if fixed_param_precision is not None:
            if is_ard(model):
                raise NotImplementedError("ARD type models with fixed param precisions are not implemented. "
                                          "Use a Base type model with fixed precisions instead.")   # :324-326
            wp = [np.full(1, fixed_param_precision, dtype=np.float32) for _ in range(nl)]
            bp = [np.full(1, fixed_param_precision, dtype=np.float32) for _ in range(nl - 1)]
        else:
            if is_ard(model):       # :308-328: one precision per input row of every layer but the last
                wp = []
                for l in range(nl - 1):
                    w = weights[l].reshape((ins[l], widths[l]), order="F")
                    wp.append(np.array([np.float32(widths[l]) / _ss(w[r]) for r in range(ins[l])], dtype=np.float32))
                wp.append(np.ones(1, dtype=np.float32))
            else:                   # :237-252
                wp = [np.array([np.float32(w.size) / _ss(w)], dtype=np.float32) for w in weights]
            bp = [np.array([np.float32(b.size) / _ss(b)], dtype=np.float32) for b in biases]    # :264-274 (zero biases: +inf)
    num_weights = sum(w.size for w in weights)
    return BranchCfgFile(num_params=num_weights + sum(b.size for b in biases), num_weights=num_weights,
                         num_markers=num_markers, layer_widths=widths, weights=weights, biases=biases,
                         ow_reg_sum=0.0, ow_num_params=0, weight_precisions=wp, bias_precisions=bp,
                         error_precision=[2.0], activation=activation)


def output_stat(model: str, w) -> np.float32:
    """summary_stat_fn_host: sum of squares (ridge, std normal) or of absolute values (lasso)."""
    w = np.asarray(w, dtype=np.float32)
    return np.float32(np.sum(np.abs(w), dtype=np.float32)) if model.startswith("lasso") else _ss(w)


def build_net(model: str, markers_per_branch: Sequence[int], depth: int, activation: str = "tanh",
              fixed_hidden: Optional[int] = None, rel_hidden: float = 0.5, fixed_summary: Optional[int] = None,
              rel_summary: Optional[float] = 1.0, hyper=(0.001, 1000.0, 0.001, 1000.0, 0.001, 1000.0),
              fixed_param_precision: Optional[float] = None, init_param_variance: Optional[float] = None,
              seed: Optional[int] = None) -> NetFile:
    """BlockNetCfg::build_net (architectures.rs:187-237)."""
    rng = np.random.default_rng(seed)
    cfgs: List[BranchCfgFile] = []
    reg_sum, num_out = np.float32(0.0), 0
    for m in markers_per_branch:
        h = hidden_width(m, fixed_hidden, rel_hidden)
        s = summary_width(h, fixed_summary, rel_summary)
        cfgs.append(build_branch_cfg(model, int(m), depth, h, s, activation, rng, fixed_param_precision, init_param_variance))
        reg_sum = np.float32(reg_sum + output_stat(model, cfgs[-1].weights[-1]))
        num_out += s
    with np.errstate(divide="ignore"):
        ow_prec = np.float32(len(cfgs)) / np.float32(sum(_ss(c.weights[-1]) for c in cfgs))    # architectures.rs:175-185
    for c in cfgs:
        c.weight_precisions[-1] = np.array([ow_prec], dtype=np.float32)
    return NetFile(hyper=[float(x) for x in hyper], branch_cfgs=cfgs, output_bias=[2.0, 1.0, 0.0],
                   lpd_local=[float("-inf")] * len(cfgs), g_error_precision=2.0,
                   g_output_layer_precision=float(fixed_param_precision if fixed_param_precision is not None
                                                  else DEFAULT_INIT_OUTPUT_LAYER_PRECISION),
                   g_ow_reg_sum=float(reg_sum), g_ow_num_params=int(num_out))
