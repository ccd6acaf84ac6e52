"""Row-sharded sequential-exact chain (BASELINE.json configs[4]): branch visits of Net::train
(net/net.rs:258-334) with the individuals split over ranks and every cross-row sum exchanged through
peer-mapped inboxes inside the reduction kernels (csrc/comm.cuh).

The kernels of different ranks wait for each other, so the ranks must sit on DIFFERENT GPUs (separately launched
kernels of one GPU are not guaranteed to be co-resident: no cross-launch spin-waits on one device).  The checks
therefore run with one process per GPU under torchrun (tests/multirank_worker.py, needs >= 2 GPUs: `gpurun --gpus 2`);
the host-side logic of the N > 1 path is covered on the CPU with gloo (tests/test_dist_gloo.py).

Bar: all ranks bit-identical to each other (rank-ordered sums); against the single-rank chain within the FP32
tolerance below (only the order of the cross-row sums differs); accept / reject decisions identical.
"""
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


def run_chain(net, rb, y_local, B, sweeps, L, seed=7, group_size=1, max_h_err=10.0, factor=0.5):
    cfg = rb.MCMCCfg(hmc_step_size_factor=factor, hmc_integration_length=L, hmc_max_hamiltonian_error=max_h_err)
    net.set_targets(y_local)
    net.init_residual()
    rng = np.random.default_rng(seed)
    st = None
    for _ in range(sweeps):
        st = net.sweep(cfg, rng.permutation(B), seed=seed, group_size=group_size)
    pv, qv = net.get_all_params()
    return dict(pv=pv, qv=qv, resid=net.residual(), stats=st, globals=net.get_globals(), yhat=net.predict())


def test_sharded_visit_without_comm_fails_loudly(rb):
    P = build_problem("ridge_base", 256, [8], 2, 2, seed=5)
    ctx = rb.Context(0, rank=0, world=2)
    gen, net = make_net(rb, ctx, P, 0, 128)
    with pytest.raises(rb.BannError, match="bann_ctx_comm"):
        net.visit_branch(0, rb.MCMCCfg(hmc_integration_length=2))
    net.close(); gen.close(); ctx.close()


def test_sharded_group_visit_without_bulk_exchange_fails_loudly(rb):
    P = build_problem("ridge_base", 256, [8, 9], 2, 2, seed=5)
    ctx = rb.Context(0, rank=0, world=2)
    gen, net = make_net(rb, ctx, P, 0, 128)
    with pytest.raises(rb.BannError, match="comm"):
        net.visit_group([0, 1], rb.MCMCCfg(hmc_integration_length=2))
    with pytest.raises(rb.BannError, match="bann_net_comm"):
        net.gradient()
    net.close(); gen.close(); ctx.close()


def test_sharded_chain_one_process_per_gpu():
    """One process per GPU under torchrun, CUDA IPC handles (needs >= 2 GPUs): sequential sweeps (also with early rejections),
    block-Jacobi sweeps over the bulk exchange, Net.gradient with 1 / world host slices -- replicas bit-identical, results
    equal to the single-rank run within the FP32 tolerance."""
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
    # null (serde_json's spelling of a non-finite f32) while a never-accepted branch keeps its infinite ML bias precision (Q7)
    la, lb = (np.array([np.nan if v is None else v for v in x["lpd"][1:]], dtype=np.float64) for x in (a, b))
    assert np.allclose(la, lb, rtol=1e-3, equal_nan=True)
    ma, mb = files.read_net(os.path.join(one, "models", "6.bin")), files.read_net(os.path.join(two_dir, "models", "6.bin"))
    for ca, cb in zip(ma.branch_cfgs, mb.branch_cfgs):
        assert np.allclose(ca.param_vec(), cb.param_vec(), rtol=2e-3, atol=2e-4)
