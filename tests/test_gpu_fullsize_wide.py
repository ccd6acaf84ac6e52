"""Parity at BASELINE.json's full size for the other two tensor-core kernel families:

  * configs[1]: 10k individuals x 1000 branches x 500 markers, widths [5,5,1]  -> k1_tcw (K-blocked, 8 marker blocks per branch,
    40 super-tiles per branch, several CTAs per branch: the cross-super-tile accumulation in tensor memory and the multi-CTA
    chunk reduction at the shape the benchmark runs);
  * configs[3]: 50k individuals x 2000 branches x 1000 markers, widths [16,16,16,1] -> k1_tcx (three passes, two 512-marker
    slabs per branch, 196 super-tiles per branch).

Same size-independent properties as tests/test_gpu_fullsize.py: sampled branches against the oracle on ALL rows (bit-exact decode
hook -> oracle f64 / f32), the tensor-core kernel against an independent kernel (FFMA where instantiated, else the
shape-agnostic one) on EVERY branch, run-to-run determinism, a checksum of the per-branch rss over the whole net."""
import numpy as np
import pytest

from oracle.branch import Branch, make_cfg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["cfg2", "cfg4"])
def full(request):
    import rs_bann_b200 as rb
    if not rb.cuda_available():
        pytest.skip("no CUDA device")
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs a 180 GB B200")
    from bench import WORKLOADS, default_params
    wl = WORKLOADS[request.param]
    N, B, PER, W = wl["n"], wl["B"], wl["per"], wl["widths"]
    ctx = rb.Context(0)
    gen = rb.Genotypes.random(ctx, N, B * PER, None, seed=42, row_offset=0, n_total=N, uniform_groups=(B, PER))
    mu, sd = rb.global_col_stats(gen.col_counts(), N, None)
    gen.set_col_stats(mu, sd)
    net = rb.Net(ctx, gen, wl["model"], [W] * B)
    pv, qv = default_params(wl)
    rng = np.random.default_rng(7)
    P = pv.size // B
    nb = sum(W[:-1])
    pvm = pv.reshape(B, P).copy()
    pvm[:, -nb:] = rng.normal(0, 0.3, size=(B, nb)).astype(np.float32)        # biases away from zero
    pv = np.ascontiguousarray(pvm.reshape(-1))
    net.set_all_params(pv, qv)
    y = rng.normal(size=N).astype(np.float32)
    net.set_targets(y)
    assert gen.has_tc_store()

    class F:
        pass
    f = F()
    f.name, f.rb, f.ctx, f.gen, f.net, f.pv, f.qv, f.y, f.P, f.N, f.B, f.PER, f.W = (request.param, rb, ctx, gen, net, pv, qv, y, P,
                                                                                   N, B, PER, W)
    yield f
    net.close(); gen.close(); ctx.close()


def test_tensor_core_family_is_the_one_that_runs(full):
    f = full
    f.net.select_k1(f.net.K1_TENSOR)           # fails loudly if the launch is not eligible
    f.net.gradient(y=f.y)
    assert ("k1_tcw" if f.name == "cfg2" else "k1_tcx") in f.net.last_k1_kernel()


def test_sampled_branches_match_the_oracle_on_all_rows(full):
    f = full
    f.net.select_k1(f.net.K1_TENSOR)
    grads, rss = f.net.gradient(y=f.y)
    g = grads.reshape(f.B, f.P)
    depth, hidden, summary = len(f.W) - 2, f.W[0], f.W[-2]
    for b in (0, f.B // 3 + 1, f.B - 1):
        x = f.gen.x_group(b, standardized=True).astype(np.float64)          # bit-exact decode (tests/test_gpu_parity.py)
        assert x.shape == (f.N, f.PER)
        cfg = make_cfg("ridge_ard", f.PER, [hidden] * depth, summary)
        cfg.load_param_vec(f.pv.reshape(f.B, f.P)[b])
        Q = f.qv.size // f.B
        cfg.load_precision_vec(f.qv.reshape(f.B, Q)[b])
        o = {}
        for dt in (np.float32, np.float64):
            br = Branch(cfg, dt)
            r, lw, lb = br.log_density_gradient(x.astype(dt), f.y.astype(dt))
            o[dt] = (float(r), Branch.join_vec(lw, lb).astype(np.float64))
        t, m = o[np.float64], o[np.float32]
        assert abs(rss[b] - t[0]) <= 8 * abs(m[0] - t[0]) + 2e-5 * t[0]
        tol = 8 * np.abs(m[1] - t[1]) + 2e-5 * np.max(np.abs(t[1]))
        assert np.all(np.abs(g[b] - t[1]) <= tol), np.max(np.abs(g[b] - t[1]) / tol)


def test_tensor_core_and_independent_kernel_agree_on_every_branch(full):
    f = full
    f.net.select_k1(f.net.K1_TENSOR)
    g_tc, r_tc = f.net.gradient(y=f.y)
    g_tc2, r_tc2 = f.net.gradient(y=f.y)
    assert np.array_equal(g_tc, g_tc2) and np.array_equal(r_tc, r_tc2)       # fixed-order reductions: deterministic
    f.net.select_k1(f.net.K1_FFMA)             # FFMA kernel where instantiated for the shape, else the shape-agnostic kernel
    g_ff, r_ff = f.net.gradient(y=f.y)
    other = f.net.last_k1_kernel()
    f.net.select_k1(f.net.K1_AUTO)
    assert "k1_tc" not in other, other
    assert np.allclose(r_tc, r_ff, rtol=2e-5, atol=0)
    a, b = g_tc.reshape(f.B, f.P), g_ff.reshape(f.B, f.P)
    scale = np.max(np.abs(b), axis=1, keepdims=True)
    assert np.max(np.abs(a - b) / scale) < 1e-4
    assert abs(r_tc.astype(np.float64).sum() - r_ff.astype(np.float64).sum()) < 1e-6 * r_ff.astype(np.float64).sum()
