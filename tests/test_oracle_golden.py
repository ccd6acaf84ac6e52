"""Pins the CPU oracle against the reference's OWN golden vectors.

Sources of the expected values (reference repo paths):
  io/bed.rs:413-497 (byte packing, small.bed decode, col means/stds, standardised
  sub-matrix), resources/test/README.md:6-32 (matrix + byte string),
  net/branch/{ridge_base,ridge_ard,lasso_base,lasso_ard}.rs `mod tests`
  (forward_feed, log_density_gradient, log_density_joint, log_density_gradient_joint,
  lasso_base::log_density), net/params.rs:791 (param_vec order),
  net/branch/branch_cfg_builder.rs:407-418 and net/architectures.rs:246-256 (param counts).

Tolerance: the reference asserts exact f32 equality against its ArrayFire
backend; values that pass through 1 - tanh^2 near saturation differ between tanh
implementations in the 3rd-4th digit, O(1) values agree to ~2e-6.
"""
import os

import numpy as np
import pytest

from oracle import bed as obed
from oracle.branch import Branch, Hyper, MCMCCfg, make_cfg

SMALL_MATRIX = np.array([
    [0, 1, 0, 0, 0, 0, 2, 1, 0, 0, 1], [0, 0, 0, 1, 0, 2, 0, 1, 0, 1, 1], [1, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0],
    [0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1], [1, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1], [0, 2, 0, 1, 0, 1, 0, 1, 2, 2, 0],
    [0, 0, 0, 1, 0, 2, 1, 1, 0, 0, 1], [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0], [0, 1, 0, 0, 0, 1, 0, 1, 0, 0, 0],
    [0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 2], [1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1], [0, 1, 0, 0, 0, 1, 1, 2, 1, 1, 1],
    [0, 0, 0, 0, 0, 2, 1, 2, 0, 1, 1], [0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 1], [0, 0, 1, 1, 0, 0, 0, 1, 0, 1, 0],
    [0, 1, 0, 0, 0, 1, 0, 1, 2, 1, 0], [1, 0, 0, 0, 0, 2, 0, 2, 0, 1, 1], [0, 0, 0, 0, 0, 1, 1, 1, 0, 1, 1],
    [2, 1, 0, 1, 0, 0, 1, 1, 0, 1, 0], [0, 0, 0, 1, 0, 1, 1, 1, 0, 0, 0]])  # resources/test/README.md:8-29

SMALL_BYTES = (b'\xef\xbe\xef\xff\xce\xee\xf3\xaa\xbf\xef\xff\xff\xff\xef\xff\xfb\xeb\xef\xef\xaf\xff\xff\xff\xff\xff'
               b'\xb3\xca\xaa\xbc\xb8\xec\xef\xbf\xfe\xab\xba\xea.\xa8\xa8\xff\xf3\xbf?\xff\xbb\xf2\xaf\xaa\xea\xba\xee'
               b'\xa3\xfa\xfa')  # README.md:32


@pytest.fixture(scope="module")
def small(golden_dir):
    raw = open(os.path.join(golden_dir, "small.bed"), "rb").read()
    assert raw[:3] == obed.BED_SIGNATURE
    return np.frombuffer(raw[3:], dtype=np.uint8), 20, 11


def test_chunk_to_byte():
    # io/bed.rs:413-415
    assert obed.pack_columns(np.array([[1], [0], [1], [1]]))[0] == 174


def test_small_bed_bytes_and_decode(small):
    payload, n, m = small
    assert bytes(payload) == SMALL_BYTES
    dec = obed.decode_columns(payload, n, range(m))
    assert np.array_equal(dec, SMALL_MATRIX.astype(np.float32))       # io/bed.rs:431-448
    # packing reproduces the file wherever the padding bits are defined by the rule
    # (the last byte of each column of small.bed pads with 11, ours with 00)
    repacked = obed.pack_columns(SMALL_MATRIX)
    assert np.array_equal(obed.decode_columns(repacked, n, range(m)), dec)
    exp0 = [0, 0, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1, 0, 2, 0]
    exp5 = [0, 2, 0, 1, 1, 1, 2, 0, 1, 1, 1, 1, 2, 0, 0, 1, 2, 1, 0, 1]
    assert np.array_equal(obed.decode_columns(payload, n, [0, 5]).T, np.array([exp0, exp5], dtype=np.float32))  # :451-465


def test_small_bed_col_stats(small):
    payload, n, m = small
    means, stds = obed.col_stats(payload, n, m)
    exp_means = np.array([0.35, 0.5, 0.05, 0.35, 0., 0.9, 0.45, 1., 0.25, 0.7, 0.65], dtype=np.float32)   # :468-472
    exp_stds = np.array([0.5722761, 0.591608, 0.21794495, 0.47696957, 0.0, 0.70000005, 0.58949125, 0.5477226,
                         0.622495, 0.55677646, 0.5722762], dtype=np.float32)                               # :475-482
    assert np.array_equal(means, exp_means)
    assert np.array_equal(stds, exp_stds)


def test_small_bed_standardized(small):
    payload, n, m = small
    means, stds = obed.col_stats(payload, n, m)
    sub = obed.submatrix_standardized(payload, n, [0, 5], means, stds)
    exp = np.array([
        -0.6115929, -0.6115929, 1.1358153, -0.6115929, 1.1358153, -0.6115929, -0.6115929, 1.1358153, -0.6115929,
        -0.6115929, 1.1358153, -0.6115929, -0.6115929, -0.6115929, -0.6115929, -0.6115929, 1.1358153, -0.6115929,
        2.8832235, -0.6115929, -1.2857141, 1.5714285, -1.2857141, 0.14285716, 0.14285716, 0.14285716, 1.5714285,
        -1.2857141, 0.14285716, 0.14285716, 0.14285716, 0.14285716, 1.5714285, -1.2857141, -1.2857141, 0.14285716,
        1.5714285, 0.14285716, -1.2857141, 0.14285716], dtype=np.float32)                                  # :485-497
    assert np.array_equal(sub.reshape(-1, order="F"), exp)


def test_four_by_two_and_random(golden_dir):
    raw = open(os.path.join(golden_dir, "four_by_two.bed"), "rb").read()
    dec = obed.decode_columns(np.frombuffer(raw, dtype=np.uint8), 4, [0, 1])
    assert dec.shape == (4, 2) and set(np.unique(dec)) <= {0.0, 1.0, 2.0}
    payload, n, m = obed.read_bed(os.path.join(golden_dir, "random"))
    assert (n, m) == (100, 20)
    g = obed.decode_columns(payload, n, range(m))
    # padding-free fixture: repacking is the identity (bed.rs:418-428 round trip)
    assert np.array_equal(obed.pack_columns(g.astype(np.int64)), payload)
    means, stds = obed.col_stats(payload, n, m)
    assert np.all(stds > 0)


def test_gene_grouping_fixture(golden_dir):
    groups = obed.read_grouping(os.path.join(golden_dir, "small.gene_grouping"))
    assert groups == [[0, 1, 2, 3], [1, 2, 3, 5], [5, 6, 7, 8, 9, 10]]   # overlapping, marker 4 in none (Q15)
    offs, ids = obed.groups_to_csr(groups)
    assert list(offs) == [0, 4, 8, 14] and len(ids) == 14


# ------------------------------------------------------------------ branch micro-fixture
X = np.array([1., 0., 0., 2., 1., 1., 2., 0., 0., 2., 0., 1.], dtype=np.float32).reshape(4, 3, order="F")
Y = np.array([0.0, 2.0, 1.0, 1.5], dtype=np.float32)
W = [np.arange(6, dtype=np.float32).reshape(3, 2, order="F"), np.array([[1.], [2.]], dtype=np.float32),
     np.array([[2.]], dtype=np.float32)]
B = [np.array([0., 1.], dtype=np.float32), np.array([2.], dtype=np.float32)]
HYPER = Hyper(dense=(3.0, 2.0), summary=(3.0, 2.0), output=(4.0, 5.0))


def fixture_branch(model, precision, dtype=np.float32):
    cfg = make_cfg(model, 3, [2], 1, weights=W, biases=B, precision=precision)
    cfg.ow_reg_sum = float(Branch(cfg, dtype).summary_stat(W[-1]))  # builder stores reg_sum 0 AFTER construction
    br = Branch(cfg, dtype)
    assert float(br.ow_reg_sum) == 0.0 and float(br.ow_num_params) == 1.0  # new_single_branch(0.0, 1)
    return br


def close(a, b, rtol=2e-6, atol=2e-6):
    return np.allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("model", ["ridge_base", "ridge_ard", "lasso_base", "lasso_ard"])
def test_forward_feed(model):
    # ridge_base.rs:371, ridge_ard.rs:453, lasso_base.rs:371, lasso_ard.rs:452
    br = fixture_branch(model, 1.0)
    pre, acts = br.forward_feed(X)
    assert len(acts) == 3 and [a.shape for a in acts] == [(4, 2), (4, 1), (4, 1)]
    exp0 = np.array([0.7615942, 0.9999092, 0.9640276, 0.9640276, 0.99999976, 0.9999999999998128, 0.99999994,
                     0.9999999999244973], dtype=np.float32).reshape(4, 2, order="F")
    assert close(acts[0], exp0)
    assert close(acts[1][:, 0], [0.99985373, 0.99990916, 0.9999024, 0.9999024])
    assert close(acts[2][:, 0], [1.9997075, 1.9998183, 1.9998049, 1.9998049])


def _check_grad(gw, gb, exp_w, exp_b):
    for g, e in zip(gw, exp_w):
        g = g.reshape(-1, order="F")
        e = np.asarray(e)
        big = np.abs(e) > 1e-2
        assert close(g[big], e[big], rtol=3e-6)
        # saturated-tanh entries: different tanh implementations (Q16)
        assert np.allclose(g[~big], e[~big], rtol=5e-3, atol=1e-8)
    for g, e in zip(gb, exp_b):
        g = np.asarray(g).reshape(-1)
        e = np.asarray(e)
        big = np.abs(e) > 1e-2
        assert close(g[big], e[big], rtol=3e-6)
        assert np.allclose(g[~big], e[~big], rtol=5e-3, atol=1e-8)


def test_log_density_gradient_ridge():
    # ridge_base.rs:545-590 / ridge_ard.rs:658-709 (precision 1.0)
    exp_w = [[-0.0005189283, -1.0005465, -2.0000138, -3.0, -4.0, -5.0], [-1.0014552, -2.0017552], [-5.4986963]]
    exp_b = [[-0.00053271546, -1.2088213e-9], [-0.0017552058]]
    for model in ("ridge_base", "ridge_ard"):
        _, gw, gb = fixture_branch(model, 1.0).log_density_gradient(X, Y)
        _check_grad(gw, gb, exp_w, exp_b)


def test_log_density_gradient_lasso():
    # lasso_ard.rs:620-672 (precision 1.0)
    exp_w = [[-0.0005189283, -1.0005465, -1.0000138, -1.0, -1.0, -1.0], [-1.0014552, -1.0017552], [-4.4986963]]
    exp_b = [[-0.00053271546, -1.2088213e-9], [-0.0017552058]]
    _, gw, gb = fixture_branch("lasso_ard", 1.0).log_density_gradient(X, Y)
    _check_grad(gw, gb, exp_w, exp_b)
    # lasso_base.rs:574-606 (precision 2.0)
    exp_w = [[-0.0010378566, -2.001093, -2.0000277, -2.0, -2.0, -2.0], [-2.0029104, -2.0035105], [-8.997393]]
    exp_b = [[-0.0010654309, -2.4176425e-9], [-0.0035104116]]
    _, gw, gb = fixture_branch("lasso_base", 2.0).log_density_gradient(X, Y)
    _check_grad(gw, gb, exp_w, exp_b)


JOINT = {  # model: (wrt_w, total, d/dlambda weights)
    "ridge_base": (-58.428806, -63.799007, [[-25.5], [-1.5], [-0.45000005]]),       # ridge_base.rs:430-540
    "ridge_ard": (-57.269924, -62.640125, [[-3.25, -7.25, -13.25], [0.5, -1.0], [-0.45000005]]),  # ridge_ard.rs:521-642
    "lasso_base": (-31.309645111040876, -36.67984440609501, [[-11.5], [-1.5], [-0.20000005]]),  # lasso_base.rs:430-536
    "lasso_ard": (-30.150764, -35.520966, [[-1.0, -3.0, -5.0], [0.5, -0.5], [-0.20000005]]),    # lasso_ard.rs:521-618
}


@pytest.mark.parametrize("model", list(JOINT))
def test_log_density_joint(model):
    br = fixture_branch(model, 2.0)
    rss = br.rss(X, Y)
    assert close(rss, 5.248245, rtol=3e-7)
    assert close(br.log_density_joint_wrt_rss(rss, HYPER, 4), -2.182509)
    assert close(br.log_density_joint_wrt_weights(HYPER), JOINT[model][0])
    assert close(br.log_density_joint_wrt_biases(HYPER), -3.1876905)
    assert close(br.log_density_joint(rss, HYPER, 4), JOINT[model][1])


@pytest.mark.parametrize("model", list(JOINT))
def test_log_density_gradient_joint(model):
    br = fixture_branch(model, 2.0)
    rss, gW, gb = br.backpropagate(X, Y)
    gw = br.ldg_wrt_weights(gW)
    gbl2 = br.ldg_wrt_biases_l2(gb)
    if model.startswith("ridge"):
        exp_w = [[-0.0010378566, -2.00109287, -4.00002756, -6.0, -8.0, -10.0], [-2.0029104, -4.0035105], [-10.997393]]
    else:
        exp_w = [[-0.0010378566, -2.001093, -2.0000277, -2.0, -2.0, -2.0], [-2.0029104, -2.0035105], [-8.997393]]
    exp_b = [[-0.0010654309, -2.0], [-4.0035105]]
    _check_grad(gw, gbl2, exp_w, exp_b)
    assert close(br.ldg_wrt_error_precision(rss, 4, HYPER), -0.32412243)
    for g, e in zip(br.ldg_wrt_weight_precisions(HYPER), JOINT[model][2]):
        assert close(g, e)
    assert close(br.ldg_wrt_bias_precisions(HYPER), [0.5, -1.25])


def test_lasso_base_log_density():
    # lasso_base.rs:539-571
    br = fixture_branch("lasso_base", 2.0)
    rss = br.rss(X, Y)
    assert close(br.log_density_wrt_rss(rss), -5.24824469)
    assert close(br.log_density_wrt_weights(), -40.0)
    assert close(br.log_density_wrt_biases_l2(), -5.0)
    assert close(br.log_density(rss), -45.24824469)


def test_numerical_gradient_cross_check():
    # ridge_ard.rs:712-771 / lasso_ard.rs:674 (finite differences, delta 1e-3, tol 1e-2), in f64
    for model in ("ridge_ard", "lasso_ard", "std_normal"):
        br = fixture_branch(model, 1.0, np.float64) if model != "std_normal" else Branch(
            make_cfg("std_normal", 3, [2], 1, weights=W, biases=B, precision=1.0), np.float64)
        _, gw, gb = br.log_density_gradient(X, Y)
        ana = Branch.join_vec(gw, gb)
        pv = br.param_vec().copy()
        base = br.log_density(br.rss(X, Y))
        for i in range(pv.size):
            p2 = pv.copy()
            p2[i] += 1e-6
            br.load_param_vec(p2)
            num = (br.log_density(br.rss(X, Y)) - base) / 1e-6
            nb = sum(w.size for w in br.W)
            if model == "std_normal" and i >= nb:
                continue  # Q5: StdNormal density has -b^2/2 but its gradient omits it
            if model == "lasso_ard" and abs(pv[i]) < 1e-9:
                continue  # |w| kink at 0
            assert abs(num - ana[i]) < 1e-3 * max(1.0, abs(ana[i])), (model, i, num, ana[i])
        br.load_param_vec(pv)


def test_param_vec_order_and_counts():
    # params.rs:791-795
    cfg = make_cfg("ridge_base", 2, [], 1, weights=[[0.1, 0.2], [0.3]], biases=[[0.4]], precision=1.0)
    assert np.allclose(cfg.param_vec(), [0.1, 0.2, 0.3, 0.4])
    # branch_cfg_builder.rs:407-418: m=3, one hidden layer 3, summary 1 -> 17 params
    assert make_cfg("ridge_base", 3, [3], 1).num_params == 17
    # architectures.rs:246-256: summary width 2 -> 22
    assert make_cfg("ridge_base", 3, [3], 2).num_params == 22
    br = fixture_branch("ridge_ard", 1.0)
    pv = br.param_vec()
    assert np.allclose(pv, [0, 1, 2, 3, 4, 5, 1, 2, 2, 0, 1, 2])
    br.load_param_vec(pv[::-1].copy())
    assert np.allclose(br.param_vec(), pv[::-1])


def test_uniform_step_sizes_and_net_movement():
    # ridge_base.rs:592-622
    br = fixture_branch("ridge_base", 1.0)
    ws, bs = br.step_sizes(MCMCCfg(hmc_step_size_factor=1.0, hmc_step_size_mode="uniform"))
    assert all(np.all(w == 1.0) for w in ws) and all(np.all(b == 1.0) for b in bs)
    ones_w = [np.ones_like(w) for w in br.W]
    ones_b = [np.ones_like(b) for b in br.b]
    assert br.net_movement([0 * w for w in br.W], [0 * b for b in br.b], ones_w, ones_b) > 0
    assert br.net_movement([0 * w + 100 for w in br.W], [0 * b + 100 for b in br.b], ones_w, ones_b) < 0
