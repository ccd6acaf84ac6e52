"""Parity at BASELINE.json's full size (configs[2]: 100k individuals x 10 000 branches x 50 markers, widths [5,5,1],
RidgeARD) -- the workload bench.py times -- through properties that do not need the oracle to scan 12.5 GB:

  * sampled branches against the oracle: the columns of a few branches are decoded (bit-exact test hook), the oracle's
    f64 forward / backward runs on all 100k rows of those branches, the full-network launch must agree;
  * two independent kernels, two store layouts: the tensor-core kernel (tcgen05, bf16-subnormal store) against the FFMA
    kernel (byte-tile store) on EVERY branch;
  * affinity in the targets: the raw gradient sums are affine in t, so g(t1) + g(t2) - g(0) == g(t1 + t2) for the whole net;
  * a checksum of checksums: per-branch rss summed over the net is invariant to the row order of the shards
    (rows [0, N/2) + rows [N/2, N) stores against the whole store; the synthetic generator is keyed by global row);
  * run-to-run determinism (fixed-order reductions): bit-identical gradients on a repeated launch.
"""
import numpy as np
import pytest

from oracle.branch import Branch, make_cfg

pytestmark = pytest.mark.gpu

N, B, PER, WIDTHS = 100_000, 10_000, 50, [5, 5, 1]


@pytest.fixture(scope="module")
def full(request):
    import rs_bann_b200 as rb
    if not rb.cuda_available():
        pytest.skip("no CUDA device")
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200 (stores of the full configuration: 27 GB)")
    from bench import WORKLOADS, default_params
    wl = WORKLOADS["cfg3"]
    assert (wl["n"], wl["B"], wl["per"], wl["widths"]) == (N, B, PER, WIDTHS)
    ctx = rb.Context(0)
    gen = rb.Genotypes.random(ctx, N, B * PER, None, seed=42, row_offset=0, n_total=N, uniform_groups=(B, PER))
    mu, sd = rb.global_col_stats(gen.col_counts(), N, None)
    gen.set_col_stats(mu, sd)
    net = rb.Net(ctx, gen, wl["model"], [WIDTHS] * B)
    pv, qv = default_params(wl)
    # biases away from zero so that every term of the backward pass is exercised
    rng = np.random.default_rng(7)
    P = pv.size // B
    pvm = pv.reshape(B, P).copy()
    pvm[:, -10:] = rng.normal(0, 0.3, size=(B, 10)).astype(np.float32)
    pv = np.ascontiguousarray(pvm.reshape(-1))
    net.set_all_params(pv, qv)
    y = rng.normal(size=N).astype(np.float32)
    net.set_targets(y)
    assert gen.has_tc_store()

    class F:
        pass
    f = F()
    f.rb, f.ctx, f.gen, f.net, f.pv, f.qv, f.y, f.mu, f.sd, f.P = rb, ctx, gen, net, pv, qv, y, mu, sd, P
    yield f
    net.close(); gen.close(); ctx.close()


def test_sampled_branches_match_the_oracle_on_all_rows(full):
    f = full
    f.net.select_k1(f.net.K1_TENSOR)
    grads, rss = f.net.gradient(y=f.y)
    g = grads.reshape(B, f.P)
    for b in (0, 4321, B - 1):
        x = f.gen.x_group(b, standardized=True).astype(np.float64)          # bit-exact decode (tests/test_gpu_parity.py)
        assert x.shape == (N, PER)
        cfg = make_cfg("ridge_ard", PER, [5], 5)
        cfg.load_param_vec(f.pv.reshape(B, f.P)[b])
        Q = f.qv.size // B
        cfg.load_precision_vec(f.qv.reshape(B, Q)[b])
        o = {}
        for dt in (np.float32, np.float64):
            br = Branch(cfg, dt)
            r, lw, lb = br.log_density_gradient(x.astype(dt), f.y.astype(dt))
            o[dt] = (float(r), Branch.join_vec(lw, lb).astype(np.float64))
        t, m = o[np.float64], o[np.float32]
        assert abs(rss[b] - t[0]) <= 8 * abs(m[0] - t[0]) + 2e-5 * t[0]
        tol = 8 * np.abs(m[1] - t[1]) + 2e-5 * np.max(np.abs(t[1]))
        assert np.all(np.abs(g[b] - t[1]) <= tol), np.max(np.abs(g[b] - t[1]) / tol)


def test_tensor_core_and_ffma_kernels_agree_on_every_branch(full):
    f = full
    f.net.select_k1(f.net.K1_TENSOR)
    g_tc, r_tc = f.net.gradient(y=f.y)
    g_tc2, r_tc2 = f.net.gradient(y=f.y)
    assert np.array_equal(g_tc, g_tc2) and np.array_equal(r_tc, r_tc2)       # fixed-order reductions: deterministic
    f.net.select_k1(f.net.K1_FFMA)
    g_ff, r_ff = f.net.gradient(y=f.y)
    f.net.select_k1(f.net.K1_AUTO)
    assert np.allclose(r_tc, r_ff, rtol=2e-5, atol=0)
    a, b = g_tc.reshape(B, f.P), g_ff.reshape(B, f.P)
    scale = np.max(np.abs(b), axis=1, keepdims=True)
    assert np.max(np.abs(a - b) / scale) < 1e-4
    # checksum of checksums over the whole net, f64
    assert abs(r_tc.astype(np.float64).sum() - r_ff.astype(np.float64).sum()) < 1e-6 * r_ff.astype(np.float64).sum()


def test_five_warp_kernel_matches_the_oracle_and_k1_tc(full):
    """k1_tc5 (dedicated issuing warp, deferred cross-row sums: csrc/k1_tc5.cuh, the default) at full size: sampled branches
    against the oracle on all rows, every branch against k1_tc, run-to-run identical."""
    f = full
    f.net.select_k1(f.net.K1_TENSOR)
    f.net.select_k1_tc_variant(f.net.TC_FIVE_WARPS)
    try:
        g5, r5 = f.net.gradient(y=f.y)
        assert "k1_tc5" in f.net.last_k1_kernel()
        g5b, r5b = f.net.gradient(y=f.y)
        assert np.array_equal(g5, g5b) and np.array_equal(r5, r5b)
        f.net.select_k1_tc_variant(f.net.TC_FOUR_WARPS)
        g4, r4 = f.net.gradient(y=f.y)
        assert "k1_tc<" in f.net.last_k1_kernel()
    finally:
        f.net.select_k1_tc_variant(f.net.TC_FIVE_WARPS)
    f.net.select_k1(f.net.K1_AUTO)
    assert np.allclose(r5, r4, rtol=2e-5, atol=0)
    a, b = g5.reshape(B, f.P), g4.reshape(B, f.P)
    scale = np.max(np.abs(b), axis=1, keepdims=True)
    assert np.max(np.abs(a - b) / scale) < 1e-4
    for br_ix in (17, 7777):
        x = f.gen.x_group(br_ix, standardized=True).astype(np.float64)
        cfg = make_cfg("ridge_ard", PER, [5], 5)
        cfg.load_param_vec(f.pv.reshape(B, f.P)[br_ix])
        cfg.load_precision_vec(f.qv.reshape(B, f.qv.size // B)[br_ix])
        o = {}
        for dt in (np.float32, np.float64):
            r, lw, lb = Branch(cfg, dt).log_density_gradient(x.astype(dt), f.y.astype(dt))
            o[dt] = (float(r), Branch.join_vec(lw, lb).astype(np.float64))
        t, m = o[np.float64], o[np.float32]
        assert abs(r5[br_ix] - t[0]) <= 8 * abs(m[0] - t[0]) + 2e-5 * t[0]
        tol = 8 * np.abs(m[1] - t[1]) + 2e-5 * np.max(np.abs(t[1]))
        assert np.all(np.abs(a[br_ix] - t[1]) <= tol), np.max(np.abs(a[br_ix] - t[1]) / tol)


def test_raw_sums_are_affine_in_the_targets(full):
    f = full
    rng = np.random.default_rng(11)
    t1 = rng.normal(size=N).astype(np.float32)
    t2 = rng.normal(size=N).astype(np.float32)
    t12 = (t1.astype(np.float64) + t2.astype(np.float64))
    # ldg = -(lambda_e * d_rss(t) + prior term): affine in t for fixed parameters, the prior term cancels in the combination
    g0, _ = f.net.gradient(y=np.zeros(N, dtype=np.float32))
    g1, _ = f.net.gradient(y=t1)
    g2, _ = f.net.gradient(y=t2)
    g12, _ = f.net.gradient(y=t12.astype(np.float32))
    lhs = g1.astype(np.float64) + g2.astype(np.float64) - g0.astype(np.float64)
    a, b = lhs.reshape(B, f.P), g12.astype(np.float64).reshape(B, f.P)
    scale = np.max(np.abs(b), axis=1, keepdims=True)
    assert np.max(np.abs(a - b) / scale) < 2e-4      # f32 sums over 100k rows + the rounding of t1 + t2 to f32


def test_rss_checksum_is_invariant_to_row_sharding(full):
    f = full
    rb = f.rb
    _, rss_whole = f.net.gradient(y=f.y)
    total = 0.0
    per_branch = np.zeros(B, dtype=np.float64)
    half = (N // 2 // 128) * 128                      # shards on 128-row tile boundaries
    for r0, r1 in ((0, half), (half, N)):
        gen = rb.Genotypes.random(f.ctx, r1 - r0, B * PER, None, seed=42, row_offset=r0, n_total=N, uniform_groups=(B, PER))
        gen.set_col_stats(f.mu, f.sd)                 # GLOBAL column statistics
        net = rb.Net(f.ctx, gen, "ridge_ard", [WIDTHS] * B)
        net.set_all_params(f.pv, f.qv)
        net.set_targets(np.ascontiguousarray(f.y[r0:r1]))
        _, rss = net.gradient(y=np.ascontiguousarray(f.y[r0:r1]))
        per_branch += rss.astype(np.float64)
        total += float(rss.astype(np.float64).sum())
        net.close(); gen.close()
    assert np.allclose(per_branch, rss_whole, rtol=2e-5, atol=0)
    assert abs(total - float(rss_whole.astype(np.float64).sum())) < 1e-6 * total
