"""End-to-end run of the kept command surface on the GPU: simulate-xy -> train-new -> predict -> train (resume),
checking every output file of the reference and that the chain learns (posterior predictive R^2)."""
import glob
import io
import json
import os
from contextlib import redirect_stdout

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rb():
    import rs_bann_b200 as rb
    if not rb.cuda_available():
        pytest.skip("no CUDA device")
    return rb


def run(argv):
    from rs_bann_b200.cli import main
    buf = io.StringIO()
    with redirect_stdout(buf):
        main(argv)
    return buf.getvalue()


def test_simulate_train_predict_resume(rb, tmp_path):
    from rs_bann_b200 import files
    from rs_bann_b200.cli import r2
    sim = run(["simulate-xy", "-o", str(tmp_path), "--seed", "1", "ridge-base", "tanh", "20", "5", "1200", "3", "1", "0.6"]).strip()
    for f in ("model.bin", "model.params", "args.json", "train.bed", "train.dims", "train.groups", "train.phen",
              "test.bed", "test.phen", "train_phen_stats.json", "test_phen_stats.json"):
        assert os.path.exists(os.path.join(sim, f)), f
    assert os.path.basename(sim) == "RidgeBase_Tanh_b5_wh3_ws3_d1_m20_n1200_h0.6_rep1"        # rs-bann.rs:803-814,776-787
    stats = json.load(open(os.path.join(sim, "train_phen_stats.json")))
    assert stats["env_variance"] == pytest.approx(stats["variance"] * 0.4, rel=0.15)           # h2 = 0.6
    truth = files.read_net(os.path.join(sim, "model.bin"))
    assert len(truth.branch_cfgs) == 5 and truth.branch_cfgs[0].layer_widths == [3, 3, 1]

    tr, te = os.path.join(sim, "train"), os.path.join(sim, "test")
    out = run(["train-new", tr, tr + ".phen", tr + ".groups", "30", "20", "ridge-base", "tanh", "1",
               "--fixed-hidden-layer-width", "3", "--bfile-test", te, "--p-test", te + ".phen", "-o", str(tmp_path / "fit"),
               "--burn-in", "20", "--step-size", "0.3", "--trace", "--seed", "7", "--report-interval", "10"]).strip()
    assert os.path.basename(out).startswith("RidgeBase_Tanh_d1_cl30_il20_Izmailov_st0.3_dpk0.001_dps1000_spk0.001_sps1000"
                                            "_opk0.001_ops1000_fhlw3_rslw1_rep1")
    args = json.load(open(os.path.join(out, "args.json")))
    assert args["model_type"] == "RidgeBase" and args["branch_depth"] == 1
    hyper = json.load(open(os.path.join(out, "hyperparams")))
    assert len(hyper["branch_hyperparams"]) == 5 and hyper["precision_hyperparams"]["dense"] == {"shape": 0.001, "scale": 1000.0}
    ts = json.load(open(os.path.join(out, "training_stats")))
    assert ts["num_samples"] == 30 * 5 and len(ts["mse_train"]) == 31 and len(ts["mse_test"]) == 31 and len(ts["lpd"]) == 31
    assert ts["num_accepted"] > 0.3 * ts["num_samples"]
    assert ts["mse_train"][-1] < 0.9 * ts["mse_train"][0]                                      # the chain fits
    assert sum(1 for _ in open(os.path.join(out, "trace"))) == 31                              # one line per iteration incl. 0
    models = sorted(glob.glob(os.path.join(out, "models", "*.bin")))
    assert [os.path.basename(m) for m in models] == sorted(f"{i}.bin" for i in range(20, 31))  # chain_ix >= burn_in
    last = files.read_net(os.path.join(out, "models", "30.bin"))
    assert last.num_samples == 150 and len(last.lpd_local) == 5 and np.isfinite(last.lpd_rss)
    assert last.mse_train == pytest.approx(ts["mse_train"])

    csv = run(["predict", te, te + ".groups", "-m", os.path.join(out, "models")]).strip().splitlines()
    assert len(csv) == len(models)
    preds = np.array([[float(v) for v in row.split(",")] for row in csv])
    assert preds.shape == (len(models), 1200)
    y_te = files.read_phen(te + ".phen")
    r2_fit = r2(y_te, preds.mean(axis=0))
    assert r2_fit > 0.15, r2_fit                                                               # h2 = 0.6 upper bound
    # mse_test recorded during training = mse of the same model on the same data (net.rs:637-646)
    assert np.mean((y_te - preds[-1]) ** 2) == pytest.approx(ts["mse_test"][-1], rel=1e-4)

    rows = run(["branch-r2", te, te + ".phen", te + ".groups", "-m", os.path.join(out, "models")]).strip().splitlines()
    assert len(rows) == len(models) and len(rows[0].split(",")) == 5
    assert all(float(v) < 1.0 for v in rows[-1].split(","))
    gdir = run(["gradients", tr, tr + ".phen", tr + ".groups", "-m", os.path.join(out, "models")]).strip()
    gj = json.load(open(os.path.join(gdir, "30.json")))
    assert len(gj) == 5 and [len(w) for w in gj[0]["wrt_weights"]] == [60, 9, 3] and [len(b) for b in gj[0]["wrt_biases"]] == [3, 3]
    assert np.all(np.isfinite(np.concatenate([np.concatenate([np.array(w) for w in br["wrt_weights"]]) for br in gj])))

    pdir = run(["population-effect-sizes", tr, tr + ".phen", tr + ".groups", "-m", os.path.join(out, "models")]).strip()
    pes = json.load(open(os.path.join(pdir, "30.json")))
    assert os.path.basename(pdir) == "population_effect_sizes" and len(pes) == 5 * 20 and np.all(np.isfinite(pes))
    adir = run(["activations", te, te + ".groups", "-m", os.path.join(out, "models")]).strip()
    aj = json.load(open(os.path.join(adir, "30.json")))["activations"]
    assert len(aj) == 5 and [a["dims"] for a in aj[0]] == [[1200, 3], [1200, 3], [1200, 1]]
    yhat_sum = np.sum([np.array(br[-1]["data"]) for br in aj], axis=0) + last.output_bias[2]
    assert np.allclose(yhat_sum, preds[-1], rtol=0, atol=1e-4)                                  # sum of branch predictions + bias

    res = run(["train", tr, tr + ".phen", tr + ".groups", "3", "10", "ridge-base", os.path.join(out, "models", "30.bin"),
               "-o", str(tmp_path / "resume"), "--burn-in", "0", "--seed", "3"]).strip()
    assert os.path.basename(res) == "30_cl3_il10_Izmailov_st1_dtheta0_dlambda0"                # rs-bann.rs:1152-1160
    assert sorted(os.path.basename(m) for m in glob.glob(os.path.join(res, "models", "*.bin"))) == ["0.bin", "1.bin", "2.bin", "3.bin"]
    first = files.read_net(os.path.join(res, "models", "0.bin"))
    assert np.array_equal(first.branch_cfgs[2].param_vec(), last.branch_cfgs[2].param_vec())   # resumed from the saved state
    assert first.mse_train[-1] == pytest.approx(last.mse_train[-1], rel=1e-4)


@pytest.mark.parametrize("flag,step", [("--gradient-descent", "1e-4"), ("--gradient-descent-joint", "1e-5"), ("--joint-hmc", "0.003")])
def test_train_new_flag_gated_modes(rb, tmp_path, flag, step):
    """`train-new --gradient-descent | --gradient-descent-joint | --joint-hmc` (cli.rs mcmc args, net.rs:282-290) run through
    the same chain driver and write the same files; the ascent modes must lower the training error."""
    model = "ridge-ard" if flag == "--gradient-descent" else "ridge-base"   # fixed precisions exist for Base models only (:324-326)
    sim = run(["simulate-xy", "-o", str(tmp_path), "--seed", "2", model, "tanh", "12", "4", "600", "3", "1", "0.6"]).strip()
    tr = os.path.join(sim, "train")
    out = run(["train-new", tr, tr + ".phen", tr + ".groups", "4", "5", model, "tanh", "1", "--fixed-hidden-layer-width", "3",
               "-o", str(tmp_path / "fit"), "--burn-in", "0", "--step-size", step, "--seed", "5", flag]
              # the default initialisation has zero biases, i.e. infinite maximum-likelihood bias precisions
              # (branch_cfg_builder.rs:237-283): the joint gradient is NaN there, in the reference as well
              + (["--fixed-param-precision", "1.0"] if flag != "--gradient-descent" else [])).strip()
    tag = {"--gradient-descent": "_gd_fhlw3", "--gradient-descent-joint": "_gdj_fp1_fhlw3", "--joint-hmc": "_joint_fp1_fhlw3"}[flag]
    assert f"_ops1000{tag}" in os.path.basename(out)             # rs-bann.rs:1036-1050
    ts = json.load(open(os.path.join(out, "training_stats")))
    assert ts["num_samples"] == 4 * 4 and len(ts["mse_train"]) == 5 and np.all(np.isfinite(ts["mse_train"]))
    assert len(glob.glob(os.path.join(out, "models", "*.bin"))) == 5
    if flag != "--joint-hmc":
        assert ts["num_accepted"] == ts["num_samples"]           # the ascent modes always accept (error precision stays > 0)
        assert ts["mse_train"][-1] < ts["mse_train"][0]
    else:
        assert ts["num_accepted"] + ts["num_early_rejected"] <= ts["num_samples"]


def test_train_new_trajectories_file(rb, tmp_path):
    """--trajectories: one JSON line per HMC transition in <outdir>/traj (trajectory.rs:4-43, mcmc_cfg.rs:247-249), and the
    chain itself is the one `bann_sweep` samples with the same seed."""
    sim = run(["simulate-xy", "-o", str(tmp_path), "--seed", "4", "ridge-base", "tanh", "10", "3", "500", "3", "1", "0.5"]).strip()
    tr = os.path.join(sim, "train")
    common = [tr, tr + ".phen", tr + ".groups", "3", "7", "ridge-base", "tanh", "1", "--fixed-hidden-layer-width", "3",
              "--burn-in", "0", "--step-size", "0.3", "--seed", "11"]
    out_t = run(["train-new"] + common + ["-o", str(tmp_path / "a"), "--trajectories"]).strip()
    out_s = run(["train-new"] + common + ["-o", str(tmp_path / "b")]).strip()
    lines = [json.loads(l) for l in open(os.path.join(out_t, "traj"))]
    assert len(lines) == 3 * 3                                    # chain_length x branches
    P = 10 * 3 + 3 * 3 + 3 + 3 + 3
    for t in lines:
        assert set(t) == {"params", "precisions", "ldg", "num_ldg", "hamiltonian"}
        n = len(t["params"])
        assert 1 <= n <= 7 and len(t["hamiltonian"]) == n + 1 and len(t["ldg"]) == n
        assert all(len(r) == P for r in t["params"]) and all(len(r) == P for r in t["ldg"])
        assert np.all(np.isfinite(t["hamiltonian"]))
    a = json.load(open(os.path.join(out_t, "training_stats")))
    b = json.load(open(os.path.join(out_s, "training_stats")))
    assert a["num_accepted"] == b["num_accepted"] and a["mse_train"] == pytest.approx(b["mse_train"], rel=1e-6)
    # joint mode: precisions ride along
    out_j = run(["train-new"] + common + ["-o", str(tmp_path / "c"), "--trajectories", "--joint-hmc", "--fixed-param-precision", "1.0"]).strip()
    tj = [json.loads(l) for l in open(os.path.join(out_j, "traj"))]
    Q = 3 + 2 + 1
    assert len(tj) == 9 and all(len(r) == Q for t in tj for r in t["precisions"]) and all(len(r) == P + Q for t in tj for r in t["ldg"])


def test_train_new_group_size(rb, tmp_path):
    """`train-new --group-size G` (extension): block-Jacobi groups through bann_sweep; --group-size 1 is the default chain bit for
    bit, the grouped chains write the same files, visit every branch once per iteration and fit the data."""
    sim = run(["simulate-xy", "-o", str(tmp_path), "--seed", "6", "ridge-ard", "tanh", "15", "6", "1500", "3", "1", "0.6"]).strip()
    tr = os.path.join(sim, "train")
    common = [tr, tr + ".phen", tr + ".groups", "25", "15", "ridge-ard", "tanh", "1", "--fixed-hidden-layer-width", "3",
              "--burn-in", "24", "--step-size", "0.3", "--seed", "9"]
    outs = {}
    for tag, extra in (("default", []), ("g1", ["--group-size", "1"]), ("g4", ["--group-size", "4"]), ("all", ["--group-size", "0"])):
        out = run(["train-new"] + common + ["-o", str(tmp_path / tag)] + extra).strip()
        outs[tag] = json.load(open(os.path.join(out, "training_stats")))
        assert outs[tag]["num_samples"] == 25 * 6 and len(outs[tag]["mse_train"]) == 26
        assert len(glob.glob(os.path.join(out, "models", "*.bin"))) == 2
    assert outs["default"] == outs["g1"]
    for tag in ("g4", "all"):
        assert np.all(np.isfinite(outs[tag]["mse_train"])) and outs[tag]["num_accepted"] > 0.3 * outs[tag]["num_samples"]
        assert outs[tag]["mse_train"][-1] < 0.9 * outs[tag]["mse_train"][0]
    with pytest.raises(SystemExit):
        run(["train-new"] + common + ["-o", str(tmp_path / "bad"), "--group-size", "3", "--joint-hmc"])


def test_train_new_numerical_gradient_trajectories(rb, tmp_path):
    """--trajectories --num-grad-traj (branch_sampler.rs:1259-1261): every recorded step carries numerical_ldg next to the
    analytical gradient; --num-grad integrates with it (:1232-1247) and the chain still fits."""
    sim = run(["simulate-xy", "-o", str(tmp_path), "--seed", "8", "ridge-base", "tanh", "6", "2", "400", "2", "1", "0.5"]).strip()
    tr = os.path.join(sim, "train")
    common = [tr, tr + ".phen", tr + ".groups", "2", "4", "ridge-base", "tanh", "1", "--fixed-hidden-layer-width", "2",
              "--burn-in", "0", "--step-size", "0.2", "--seed", "13"]
    out = run(["train-new"] + common + ["-o", str(tmp_path / "a"), "--trajectories", "--num-grad-traj"]).strip()
    lines = [json.loads(l) for l in open(os.path.join(out, "traj"))]
    assert len(lines) == 2 * 2
    for t in lines:
        n = len(t["params"])
        assert len(t["num_ldg"]) == n and all(len(r) == len(t["ldg"][0]) for r in t["num_ldg"])
        a, b = np.array(t["ldg"]), np.array(t["num_ldg"])
        assert np.max(np.abs(a - b)) < 0.05 * np.max(np.abs(a)) + 1.0        # forward differences in f32: noisy, but the same gradient
    out2 = run(["train-new"] + common + ["-o", str(tmp_path / "b"), "--num-grad"]).strip()
    ts = json.load(open(os.path.join(out2, "training_stats")))
    assert ts["num_samples"] == 4 and np.all(np.isfinite(ts["mse_train"]))
