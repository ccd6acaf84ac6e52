"""Row-sharded sequential-exact chain (BASELINE.json configs[4]): branch visits of Net::train
(net/net.rs:258-334) with the individuals split over ranks and every cross-row sum exchanged through
peer-mapped inboxes inside the reduction kernels (csrc/comm.cuh).

On one GPU the ranks are emulated by contexts of the same process (own streams, one host thread per rank --
the kernels of different ranks must be able to wait for each other); tests/multirank_worker.py runs the same
check with one process per GPU under torchrun.

Bar: all ranks bit-identical to each other (rank-ordered sums); against the single-rank chain within the FP32
tolerance below (only the order of the cross-row sums differs); accept / reject decisions identical.
"""
import threading

import numpy as np
import pytest

from oracle import bed as obed
from oracle.branch import Hyper, make_cfg

pytestmark = pytest.mark.gpu

HYPER = Hyper(dense=(3.0, 2.0), summary=(2.5, 1.5), output=(4.0, 5.0))
RTOL, ATOL = 2e-4, 2e-5      # parameters / precisions / residual after the sweeps vs the single-rank chain


@pytest.fixture(scope="module")
def rb():
    import rs_bann_b200 as rb
    if not rb.cuda_available():
        pytest.skip("no CUDA device")
    return rb


def build_problem(model, n, group_sizes, hidden, summary, seed):
    rng = np.random.default_rng(seed)
    m = sum(group_sizes)
    g = obed.random_genotypes(n, m, seed=seed + 1)
    payload = obed.pack_columns(g)
    means, stds = obed.col_stats(payload, n, m)
    groups, start = [], 0
    for sz in group_sizes:
        groups.append(list(range(start, start + sz)))
        start += sz
    cfgs = []
    for cols in groups:
        cfg = make_cfg(model, len(cols), [hidden], summary, activation="tanh", rng=rng)
        cfg.biases = [rng.normal(0, 0.3, size=b.shape).astype(np.float32) for b in cfg.biases]
        cfg.weight_precisions = [rng.uniform(0.5, 3.0, size=p.shape).astype(np.float32) for p in cfg.weight_precisions]
        cfg.bias_precisions = [rng.uniform(0.5, 3.0, size=p.shape).astype(np.float32) for p in cfg.bias_precisions]
        cfgs.append(cfg)
    y = rng.normal(0, 1, size=n).astype(np.float32)
    return dict(n=n, m=m, payload=payload, means=means, stds=stds, groups=groups, cfgs=cfgs, y=y, model=model)


def make_net(rb, ctx, P, r0, r1):
    payload = P["payload"] if (r0, r1) == (0, P["n"]) else rb.shard_payload(P["payload"], P["n"], P["m"], r0, r1)
    gen = rb.Genotypes(ctx, payload, r1 - r0, P["m"], P["groups"], col_means=P["means"], col_stds=P["stds"],
                       n_total=P["n"])
    net = rb.Net(ctx, gen, P["model"], [c.layer_widths for c in P["cfgs"]],
                 hyper=(*HYPER.dense, *HYPER.summary, *HYPER.output))
    for b, c in enumerate(P["cfgs"]):
        net.set_branch(b, c.param_vec(), c.precision_vec())
    ow = sum(float(np.sum(np.abs(c.weights[-1]) if "lasso" in P["model"] else c.weights[-1] ** 2)) for c in P["cfgs"])
    net.set_globals(2.0, 0.05, ow, sum(c.layer_widths[-2] for c in P["cfgs"]), 0.0)
    return gen, net


def run_chain(net, rb, y_local, B, sweeps, L, seed=7):
    cfg = rb.MCMCCfg(hmc_step_size_factor=0.5, hmc_integration_length=L)
    net.set_targets(y_local)
    net.init_residual()
    rng = np.random.default_rng(seed)
    st = None
    for _ in range(sweeps):
        st = net.sweep(cfg, rng.permutation(B), seed=seed)
    pv, qv = net.get_all_params()
    return dict(pv=pv, qv=qv, resid=net.residual(), stats=st, globals=net.get_globals(), yhat=net.predict())


@pytest.mark.parametrize("model,world", [("ridge_ard", 2), ("lasso_base", 2), ("ridge_base", 3), ("std_normal", 2)])
def test_sharded_sweeps_match_single_rank(rb, model, world):
    P = build_problem(model, 1000, [20, 50, 9, 33], 5, 5, seed=11)
    B = len(P["groups"])
    # single rank: the reference chain (also loads every kernel before ranks start waiting for each other)
    ctx1 = rb.Context(0)
    gen1, net1 = make_net(rb, ctx1, P, 0, P["n"])
    ref = run_chain(net1, rb, P["y"], B, sweeps=2, L=8)
    net1.close(); gen1.close(); ctx1.close()

    ctxs = [rb.Context(0, rank=r, world=world) for r in range(world)]
    handles = [c.comm_handle() for c in ctxs]
    for c in ctxs:
        c.comm_connect(handles)
        assert c.comm_connected()
    shards = [rb.row_shard(P["n"], r, world) for r in range(world)]
    built = [make_net(rb, ctxs[r], P, *shards[r]) for r in range(world)]
    out, errs = [None] * world, []
    start = threading.Barrier(world)

    def worker(r):
        try:
            start.wait()
            out[r] = run_chain(built[r][1], rb, P["y"][shards[r][0]:shards[r][1]], B, sweeps=2, L=8)
        except Exception as ex:     # noqa: BLE001
            errs.append((r, ex))

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for gen, net in built:
        net.close(); gen.close()
    for c in ctxs:
        c.close()
    assert not errs, errs
    # replicas: bit-identical state on every rank
    for r in range(1, world):
        assert np.array_equal(out[r]["pv"], out[0]["pv"]) and np.array_equal(out[r]["qv"], out[0]["qv"])
        assert out[r]["globals"] == out[0]["globals"]
        assert out[r]["stats"] == out[0]["stats"]
    # against the single-rank chain
    s, s1 = out[0]["stats"], ref["stats"]
    assert (s["num_samples"], s["num_accepted"], s["num_early_rejected"]) == \
           (s1["num_samples"], s1["num_accepted"], s1["num_early_rejected"])
    assert s["num_samples"] == 2 * B
    assert np.allclose(out[0]["pv"], ref["pv"], rtol=RTOL, atol=ATOL)
    assert np.allclose(out[0]["qv"], ref["qv"], rtol=RTOL, atol=ATOL)
    resid = np.concatenate([o["resid"] for o in out])
    assert np.allclose(resid, ref["resid"], rtol=0, atol=5e-4)
    yhat = np.concatenate([o["yhat"] for o in out])
    assert np.allclose(yhat, ref["yhat"], rtol=0, atol=5e-4)
    assert abs(s["mse_train"] - s1["mse_train"]) < 1e-4 * max(1.0, s1["mse_train"])
    assert abs(s["lpd"] - s1["lpd"]) < 5e-4 * abs(s1["lpd"])
    assert abs(s["output_bias"] - s1["output_bias"]) < 1e-5


def test_sharded_fwd_bwd_sums_over_ranks(rb):
    """backpropagate (branch_sampler.rs:813-875) on sharded rows: rss and raw gradient sums are totals over ranks."""
    P = build_problem("ridge_ard", 700, [40, 13], 4, 3, seed=3)
    ctx1 = rb.Context(0)
    gen1, net1 = make_net(rb, ctx1, P, 0, P["n"])
    net1.set_targets(P["y"])
    ref = [net1.branch_fwd_bwd(b) for b in range(2)]
    net1.close(); gen1.close(); ctx1.close()
    world = 2
    ctxs = [rb.Context(0, rank=r, world=world) for r in range(world)]
    handles = [c.comm_handle() for c in ctxs]
    for c in ctxs:
        c.comm_connect(handles)
    shards = [rb.row_shard(P["n"], r, world) for r in range(world)]
    built = [make_net(rb, ctxs[r], P, *shards[r]) for r in range(world)]
    out, errs = [None] * world, []

    def worker(r):
        try:
            net = built[r][1]
            net.set_targets(P["y"][shards[r][0]:shards[r][1]])
            out[r] = [net.branch_fwd_bwd(b) for b in range(2)]
        except Exception as ex:     # noqa: BLE001
            errs.append((r, ex))

    ts = [threading.Thread(target=worker, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for gen, net in built:
        net.close(); gen.close()
    for c in ctxs:
        c.close()
    assert not errs, errs
    for b in range(2):
        assert out[0][b]["rss"] == out[1][b]["rss"] and np.array_equal(out[0][b]["ldg"], out[1][b]["ldg"])
        assert abs(out[0][b]["rss"] - ref[b]["rss"]) < 1e-5 * ref[b]["rss"]
        sc = np.max(np.abs(ref[b]["ldg"]))
        assert np.max(np.abs(out[0][b]["ldg"] - ref[b]["ldg"])) < 2e-5 * sc
        yh = np.concatenate([out[r][b]["yhat"] for r in range(world)])
        assert np.allclose(yh, ref[b]["yhat"], rtol=0, atol=1e-5)


def test_sharded_visit_without_comm_fails_loudly(rb):
    P = build_problem("ridge_base", 256, [8], 2, 2, seed=5)
    ctx = rb.Context(0, rank=0, world=2)
    gen, net = make_net(rb, ctx, P, 0, 128)
    with pytest.raises(rb.BannError, match="bann_ctx_comm"):
        net.visit_branch(0, rb.MCMCCfg(hmc_integration_length=2))
    net.close(); gen.close(); ctx.close()


def test_sharded_chain_one_process_per_gpu():
    """The same check with real peers: one process per GPU under torchrun, CUDA IPC handles (needs >= 2 GPUs)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(here, "multirank_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTIRANK_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]


def test_train_new_cli_under_torchrun_matches_single_gpu(tmp_path):
    """`rs-bann train-new` launched by torchrun shards the individuals over the GPUs; with the same --seed the
    recorded chain (mse, lpd, acceptance counts, saved model) follows the single-GPU run (needs >= 2 GPUs)."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from rs_bann_b200 import files
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "rs-bann")
    sim = subprocess.run([sys.executable, exe, "simulate-xy", "-o", str(tmp_path), "--seed", "1", "ridge-ard", "tanh", "20", "6",
                          "2000", "3", "1", "0.6"], capture_output=True, text=True, timeout=300, check=True).stdout.strip()
    tr, te = os.path.join(sim, "train"), os.path.join(sim, "test")
    common = ["train-new", tr, tr + ".phen", tr + ".groups", "6", "15", "ridge-ard", "tanh", "1", "--fixed-hidden-layer-width", "3",
              "--bfile-test", te, "--p-test", te + ".phen", "--burn-in", "5", "--step-size", "0.3", "--seed", "7"]
    one = subprocess.run([sys.executable, exe] + common + ["-o", str(tmp_path / "one")], capture_output=True, text=True, timeout=300)
    assert one.returncode == 0, one.stdout[-2000:] + one.stderr[-3000:]
    one = one.stdout.strip()
    two = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29633", exe] + common + ["-o", str(tmp_path / "two")],
                         capture_output=True, text=True, timeout=600)
    assert two.returncode == 0, two.stdout[-2000:] + two.stderr[-3000:]
    two_dir = [ln for ln in two.stdout.strip().splitlines() if os.path.isdir(ln)][-1]
    a = json.load(open(os.path.join(one, "training_stats")))
    b = json.load(open(os.path.join(two_dir, "training_stats")))
    assert (a["num_samples"], a["num_accepted"], a["num_early_rejected"]) == (b["num_samples"], b["num_accepted"], b["num_early_rejected"])
    assert np.allclose(a["mse_train"], b["mse_train"], rtol=1e-4) and np.allclose(a["mse_test"], b["mse_test"], rtol=1e-4)
    assert np.allclose(a["lpd"][1:], b["lpd"][1:], rtol=1e-3, equal_nan=True)   # NaN while a never-accepted branch keeps its infinite ML bias precision (Q7)
    ma, mb = files.read_net(os.path.join(one, "models", "6.bin")), files.read_net(os.path.join(two_dir, "models", "6.bin"))
    for ca, cb in zip(ma.branch_cfgs, mb.branch_cfgs):
        assert np.allclose(ca.param_vec(), cb.param_vec(), rtol=2e-3, atol=2e-4)
