"""Host side of the kept rs-bann surface: file formats, net construction, command line parsing.
CPU only; the GPU end-to-end run of the commands is tests/test_cli_gpu.py."""
import json
import os
import struct

import numpy as np
import pytest

from oracle import bed as obed
from oracle.branch import make_cfg

import rs_bann_b200 as rb  # noqa: F401
from rs_bann_b200 import architectures as arch
from rs_bann_b200 import files
from rs_bann_b200.cli import _fmt, build_parser, r2


def test_read_bed_small_fixture_matches_oracle(golden_dir):
    payload, n, m = files.read_bed(os.path.join(golden_dir, "small"))
    opayload, on, om = obed.read_bed(os.path.join(golden_dir, "small"))
    assert (n, m) == (on, om) == (20, 11)
    assert np.array_equal(payload, opayload)


def test_bed_round_trip_and_packing(tmp_path):
    rng = np.random.default_rng(0)
    for n in (1, 4, 7, 100):
        g = rng.integers(0, 3, size=(n, 5))
        payload = files.pack_genotypes(g)
        assert np.array_equal(payload, obed.pack_columns(g))                   # io/bed.rs:378-395 (e.g. [1,0,1,1] -> 174)
        files.write_bed(str(tmp_path / f"x{n}"), payload, n, 5)
        p2, n2, m2 = files.read_bed(str(tmp_path / f"x{n}"))
        assert (n2, m2) == (n, 5) and np.array_equal(p2, payload)
    assert files.pack_genotypes(np.array([[1], [0], [1], [1]]))[0] == 174      # bed.rs:413 golden


def test_bed_rejects_sample_major_and_truncated_files(tmp_path):
    stem = str(tmp_path / "bad")
    open(stem + ".dims", "w").write("4\t2")
    open(stem + ".bed", "wb").write(bytes([0x6C, 0x1B, 0x00, 1, 2]))          # sample-major flag (bed.rs:200-202)
    with pytest.raises(ValueError):
        files.read_bed(stem)
    open(stem + ".bed", "wb").write(bytes([0x6C, 0x1B, 0x01, 1]))
    with pytest.raises(ValueError):
        files.read_bed(stem)


def test_dims_fallback_to_fam_and_bim_line_counts(tmp_path):
    stem = str(tmp_path / "fb")
    open(stem + ".fam", "w").write("a\nb\nc\n")
    open(stem + ".bim", "w").write("x\ny\n")
    assert files.read_dims(stem) == (3, 2)                                      # io/dims.rs:28-34


def test_grouping_fixture_overlapping_groups(golden_dir, tmp_path):
    groups = files.read_grouping(os.path.join(golden_dir, "small.gene_grouping"))
    assert groups == obed.read_grouping(os.path.join(golden_dir, "small.gene_grouping"))
    assert set(groups[0]) & set(groups[1])                                      # markers 1-3 sit in groups 0 and 1 (Q15)
    files.write_grouping(str(tmp_path / "g.groups"), groups)
    assert files.read_grouping(str(tmp_path / "g.groups")) == groups
    open(tmp_path / "gap.groups", "w").write("0\t0\n1\t2\n")
    with pytest.raises(ValueError):
        files.read_grouping(str(tmp_path / "gap.groups"))


def test_phen_is_bincode_vec_f32(tmp_path):
    y = np.array([0.5, -1.25, 3.0], dtype=np.float32)
    files.write_phen(str(tmp_path / "y.phen"), y)
    raw = open(tmp_path / "y.phen", "rb").read()
    assert raw == struct.pack("<Q3f", 3, 0.5, -1.25, 3.0)                      # u64 LE length + f32 LE values
    assert np.array_equal(files.read_phen(str(tmp_path / "y.phen")), y)


def test_net_file_layout_and_round_trip(tmp_path):
    nf = arch.build_net("ridge_ard", [3, 5], depth=1, fixed_hidden=3, fixed_summary=2, seed=3,
                        hyper=(3.0, 2.0, 2.5, 1.5, 4.0, 5.0))
    nf.mse_train, nf.mse_test, nf.lpd = [1.5, 1.25], [2.0, 1.75], [-10.0, -9.0]
    nf.num_samples, nf.num_accepted, nf.num_early_rejected = 4, 3, 1
    path = str(tmp_path / "m.bin")
    files.write_net(path, nf)
    raw = open(path, "rb").read()
    # field order of Net (net/net.rs:76-85): hyperparams (6 x f32), num_branches u64, Vec<BranchCfg> length u64, ...
    assert struct.unpack("<6f", raw[:24]) == (3.0, 2.0, 2.5, 1.5, 4.0, 5.0)
    assert struct.unpack("<QQ", raw[24:40]) == (2, 2)
    # first BranchCfg: num_params, num_weights, num_markers, layer_widths = Vec<usize> [3, 2, 1]
    assert struct.unpack("<QQQ", raw[40:64]) == (22, 17, 3)                    # architectures.rs:246-256: 22 params for m=3,h=3,s=2
    assert struct.unpack("<QQQQ", raw[64:96]) == (3, 3, 2, 1)
    # the file ends with GlobalParams: error precision 2.0, output precision 0.05, reg_sum, num_params (u64) = 2 + 2
    gep, gop, _reg = struct.unpack("<3f", raw[-20:-8])
    assert gep == 2.0 and gop == pytest.approx(0.05)
    assert struct.unpack("<Q", raw[-8:])[0] == 4
    back = files.read_net(path)
    assert back.training_stats_json() == nf.training_stats_json()
    for a, b in zip(back.branch_cfgs, nf.branch_cfgs):
        assert np.array_equal(a.param_vec(), b.param_vec()) and np.array_equal(a.precision_vec(), b.precision_vec())
        assert a.layer_widths == b.layer_widths and a.activation == b.activation
    nf.mse_test = None                                                          # Option<Vec<f32>> None tag
    files.write_net(path, nf)
    assert files.read_net(path).mse_test is None
    with pytest.raises(ValueError):
        open(path, "ab").write(b"\0")
        files.read_net(path)


def test_param_vec_order_matches_reference_layout():
    """params.rs:700-715 via the oracle (pinned by the reference's param_vec test, params.rs:791)."""
    nf = arch.build_net("ridge_base", [4], depth=2, fixed_hidden=3, fixed_summary=2, seed=5)
    c = nf.branch_cfgs[0]
    o = make_cfg("ridge_base", 4, [3, 3], 2,
                 weights=[w.reshape((i, w.size // i), order="F") for w, i in zip(c.weights, [4, 3, 3, 2])], biases=c.biases,
                 precision=1.0)
    assert np.array_equal(c.param_vec(), o.param_vec())
    assert c.num_params == o.num_params == 4 * 3 + 3 * 3 + 3 * 2 + 2 + 3 + 3 + 2


def test_build_net_initial_state():
    nf = arch.build_net("ridge_ard", [50, 8], depth=1, rel_hidden=0.5, rel_summary=1.0, seed=0)
    c0, c1 = nf.branch_cfgs
    assert c0.layer_widths == [25, 25, 1] and c1.layer_widths == [4, 4, 1]     # FractionOfInput / FractionOfHiddenLayerWidth
    assert c0.num_params == 50 * 25 + 25 * 25 + 25 + 25 + 25
    assert [len(p) for p in c0.weight_precisions] == [50, 25, 1]               # ARD: one precision per input row, output shared
    w0 = c0.weights[0].reshape((50, 25), order="F")
    assert np.allclose(c0.weight_precisions[0], 25.0 / np.sum(w0 * w0, axis=1), rtol=1e-5)   # branch_cfg_builder.rs:308-328
    assert np.all(np.isinf(c0.bias_precisions[0]))                             # zero biases: ML precision = inf (Q7)
    ow = 2.0 / (np.sum(c0.weights[-1] ** 2) + np.sum(c1.weights[-1] ** 2))
    assert np.isclose(c0.weight_precisions[-1][0], ow, rtol=1e-5) and c0.weight_precisions[-1] == c1.weight_precisions[-1]
    assert nf.g_output_layer_precision == pytest.approx(0.05) and nf.g_error_precision == 2.0   # architectures.rs:16,229-235
    assert nf.g_ow_num_params == 25 + 4
    assert nf.g_ow_reg_sum == pytest.approx(float(np.sum(c0.weights[-1] ** 2) + np.sum(c1.weights[-1] ** 2)), rel=1e-5)
    base = arch.build_net("lasso_base", [6], depth=0, fixed_summary=3, fixed_param_precision=1.5, seed=1)
    assert base.branch_cfgs[0].layer_widths == [3, 1]
    assert all(p[0] == 1.5 for p in base.branch_cfgs[0].weight_precisions[:-1])
    assert base.g_output_layer_precision == 1.5
    assert base.g_ow_reg_sum == pytest.approx(float(np.sum(np.abs(base.branch_cfgs[0].weights[-1]))), rel=1e-6)
    with pytest.raises(NotImplementedError):
        arch.build_net("ridge_ard", [6], depth=1, fixed_param_precision=1.0)   # branch_cfg_builder.rs:324-326
    assert arch.hidden_width(1, None, 0.5) == 1 and arch.summary_width(3, None, 0.1) == 1


def test_cli_argument_surface_and_directory_names():
    ap = build_parser()
    a = ap.parse_args("train-new tr tr.phen g.groups 10 20 ridge-ard tanh 1 --fixed-hidden-layer-width 5 --burn-in 2 "
                      "--step-size 0.1 --trace".split())
    assert (a.bfile_train, a.p_train, a.groups, a.chain_length, a.integration_length) == ("tr", "tr.phen", "g.groups", 10, 20)
    assert (a.model_type, a.activation_function, a.branch_depth, a.step_size_mode) == ("ridge_ard", "tanh", 1, "izmailov")
    assert a.dpk == 0.001 and a.ops == 1000.0 and a.max_hamiltonian_error == 10.0 and a.report_interval == 1
    b = ap.parse_args("train --step-size-mode std-scaled tr tr.phen g 5 7 StdNormal models/4.bin".split())
    assert b.model_type == "std_normal" and b.model_file == "models/4.bin" and b.step_size_mode == "std_scaled"
    p = ap.parse_args("predict te g.groups -m out/models".split())
    assert p.model_path == "out/models"
    s = ap.parse_args("simulate-xy std-normal tanh 100 10 1000 2 1 0.5 -o sim".split())
    assert (s.num_markers_per_branch, s.num_branches, s.num_individuals, s.hidden_layer_width, s.branch_depth,
            s.heritability) == (100, 10, 1000, 2, 1, 0.5)
    assert _fmt(1.0) == "1" and _fmt(0.001) == "0.001" and _fmt(1000.0) == "1000" and _fmt(0.5) == "0.5"
    with pytest.raises(SystemExit):
        ap.parse_args("train-new tr p g 1 1 linear tanh 1".split())


def test_r2_definition():
    y = np.array([1.0, 2.0, 3.0, 4.0])
    assert r2(y, y) == 1.0 and r2(y, np.full(4, 2.5)) == pytest.approx(0.0)    # 1 - mse / variance (py-vis/vis.py:555-557)
