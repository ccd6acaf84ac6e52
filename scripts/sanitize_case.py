"""Smallest launches of the three tensor-core kernel families (k1_tc, k1_tcw, k1_tcx) and of a block-Jacobi group visit, for
`compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_case.py` (one tool per gpurun call, B200_PROFILING.md)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rs_bann_b200 as rb  # noqa: E402
from oracle import bed as obed  # noqa: E402
from oracle.branch import make_cfg  # noqa: E402


def case(ctx, n, sizes, hidden, summary, depth, want):
    rng = np.random.default_rng(1)
    m = sum(sizes)
    g = obed.random_genotypes(n, m, seed=2)
    groups, start = [], 0
    for sz in sizes:
        groups.append(list(range(start, start + sz)))
        start += sz
    gen = rb.Genotypes(ctx, obed.pack_columns(g), n, m, groups)
    cfgs = [make_cfg("ridge_ard", sz, [hidden] * depth, summary, rng=rng) for sz in sizes]
    net = rb.Net(ctx, gen, "ridge_ard", [c.layer_widths for c in cfgs])
    for b, c in enumerate(cfgs):
        c.bias_precisions = [np.ones(1, dtype=np.float32) for _ in c.bias_precisions]
        net.set_branch(b, c.param_vec(), c.precision_vec())
    ow = sum(float(np.sum(c.weights[-1] ** 2)) for c in cfgs)
    net.set_globals(2.0, 0.05, ow, sum(c.layer_widths[-2] for c in cfgs), 0.0)
    net.set_targets(rng.normal(size=n).astype(np.float32))
    net.select_k1(net.K1_TENSOR)
    out = net.branch_fwd_bwd(0)
    assert want in net.last_k1_kernel(), net.last_k1_kernel()
    assert np.all(np.isfinite(out["ldg"]))
    net.init_residual()
    res = net.visit_group(list(range(len(sizes))), rb.MCMCCfg(hmc_integration_length=2, hmc_step_size_factor=0.1), seed=3)
    print(want, "ok: rss", out["rss"], "group statuses", [r.status for r in res], flush=True)
    net.close(); gen.close()


ctx = rb.Context(0)
case(ctx, 300, [20, 9], 5, 5, 1, "k1_tc<")
case(ctx, 300, [100, 70], 5, 5, 1, "k1_tcw")
case(ctx, 300, [100, 70], 16, 16, 2, "k1_tcx")
ctx.close()
print("SANITIZE_CASE_DONE")
