"""Wide branches (BASELINE.json configs[3]: 1000 markers, widths [16,16,16,1]) on the three-pass tensor-core kernel
(k1_tcx.cuh: tcgen05 first-layer contractions with 48-column split operands, FP32 tail with shared-memory staged cross-row
products): parity against the oracle and against the shape-agnostic kernel, through the C ABI."""
import numpy as np
import pytest

from oracle import net as onet
from oracle.branch import MCMCCfg as OCfg
from test_gpu_parity import (Problem, ctx, mirror_net, oracle_fwd_bwd, rb, run_oracle_hmc, within)  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu

WIDE_SHAPES = [
    # n, group sizes, hidden, summary, depth
    (700, [1000], 16, 16, 2),            # configs[3] branch shape: 16 marker blocks (last one ragged), 4 slabs
    (300, [70, 129, 33], 16, 16, 2),     # 1..3 marker blocks, overlapping groups, ragged last super-tile
    (515, [256, 257], 16, 16, 1),
    (257, [64, 65], 8, 8, 1),
    (1030, [300, 8], 8, 4, 1),           # several super-tiles; a branch with a single 8-marker chunk
    (128, [2048], 16, 16, 1),            # the tensor-core store's marker limit
    (400, [90, 700], 12, 12, 2),         # first-layer width 12: 36 of 48 accumulator columns used
    (300, [130], 12, 12, 1),
    (600, [64, 200], 8, 8, 2),
    (257, [40, 600], 4, 4, 1),
]


@pytest.mark.parametrize("model", ["ridge_ard", "lasso_base", "std_normal"])
@pytest.mark.parametrize("shape", WIDE_SHAPES)
def test_wide_fwd_bwd_parity(rb, ctx, model, shape):
    n, gs, h, s, d = shape
    P = Problem(rb, ctx, model, n, gs, h, s, depth=d, seed=(sum(map(ord, model)) + 3 * n) % 1000, overlap=len(gs) > 1)
    try:
        assert P.gen.has_tc_store()
        for b in range(len(gs)):
            tgt = P.y if b % 2 == 0 else (P.y * 0.5 + 0.1).astype(np.float32)
            P.net.select_k1(P.net.K1_TENSOR)          # fails loudly if the launch is not eligible
            got = P.net.branch_fwd_bwd(b, target=None if b % 2 == 0 else tgt)
            P.net.select_k1(P.net.K1_GENERIC)
            gen = P.net.branch_fwd_bwd(b, target=None if b % 2 == 0 else tgt)
            t64, t32 = oracle_fwd_bwd(P, b, tgt, np.float64), oracle_fwd_bwd(P, b, tgt, np.float32)
            for key in ("yhat", "rss", "d_rss", "ldg"):
                within(got[key], t64[key], t32[key])
                sc = np.max(np.abs(t64[key]))
                assert np.max(np.abs(np.asarray(got[key], dtype=np.float64) - gen[key])) <= 1e-4 * sc + 1e-30, key
    finally:
        P.close()


def test_wide_net_gradient_and_predict(rb, ctx):
    """Grouped launches (every branch in one launch, several row chunks), Net::predict (forward only) and the per-branch
    entry point agree; more than 2048 markers per branch fall back to the shape-agnostic kernel."""
    P = Problem(rb, ctx, "ridge_ard", 2000, [200, 1000, 64, 300], 16, 16, 2, seed=77)
    try:
        P.net.select_k1(P.net.K1_TENSOR)
        grads, rss = P.net.gradient(y=P.y)
        off = 0
        for b, c in enumerate(P.cfgs):
            one = P.net.branch_fwd_bwd(b)
            scale = np.max(np.abs(one["ldg"]))
            # different row chunking (grouped: 2 CTAs per SM over 4 entries; single: over 1 entry): fixed order inside each
            assert np.max(np.abs(one["ldg"] - grads[off:off + c.num_params])) <= 2e-5 * scale
            assert abs(one["rss"] - rss[b]) <= 2e-5 * rss[b]
            off += c.num_params
        onet_ = mirror_net(P)
        exp64 = onet.predict(onet_, P.payload, P.n, P.means, P.stds, np.float64)
        exp32 = onet.predict(onet_, P.payload, P.n, P.means, P.stds, np.float32)
        P.net.set_globals(2.0, 0.05, 0.0, 0, 0.0)
        within(P.net.predict(), exp64, exp32)
    finally:
        P.close()
    Q = Problem(rb, ctx, "ridge_ard", 130, [2049], 16, 16, 1, seed=5)
    try:
        assert not Q.gen.has_tc_store()
        Q.net.select_k1(Q.net.K1_TENSOR)
        with pytest.raises(RuntimeError):
            Q.net.branch_fwd_bwd(0)
        Q.net.select_k1(Q.net.K1_AUTO)
        got = Q.net.branch_fwd_bwd(0)
        t64, t32 = oracle_fwd_bwd(Q, 0, Q.y, np.float64), oracle_fwd_bwd(Q, 0, Q.y, np.float32)
        within(got["ldg"], t64["ldg"], t32["ldg"])
    finally:
        Q.close()


def test_wide_hmc_step_matches_oracle(rb, ctx):
    """A whole HMC transition of a wide branch (AUTO mode picks the three-pass kernel): trajectory and decision."""
    P = Problem(rb, ctx, "ridge_ard", 600, [300], 16, 16, 2, seed=19)
    try:
        rng = np.random.default_rng(5)
        kw = dict(hmc_step_size_factor=0.002, hmc_integration_length=6, hmc_step_size_mode="uniform", hmc_max_hamiltonian_error=10.0)
        Pn = P.cfgs[0].num_params
        mom = rng.standard_normal(Pn).astype(np.float32)
        got = P.net.hmc_step(0, rb.MCMCCfg(**kw), momenta=mom, u=0.5, trajectory=True)
        o64 = run_oracle_hmc(P, 0, P.y, OCfg(**kw), mom, 0.5, None, np.float64)
        o32 = run_oracle_hmc(P, 0, P.y, OCfg(**kw), mom, 0.5, None, np.float32)
        within(got.neg_h_init, o64["h_init"], o32["h_init"], scale=abs(o64["h_init"]))
        for s in range(min(got.steps_done, o64["steps_done"], 4)):
            within(got.trajectory["params"][s], o64["traj"]["params"][s], o32["traj"]["params"][s], rel=5e-5)
            within(got.trajectory["ldg"][s], o64["traj"]["ldg"][s], o32["traj"]["ldg"][s], rel=5e-5)
        if abs(min(o64.get("log_acc", 0.0), 0.0) - np.log(0.5)) > 1e-2 and o64["status"] != 0:
            assert got.status == o64["status"]
    finally:
        P.close()
