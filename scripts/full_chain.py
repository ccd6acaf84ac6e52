"""BASELINE.json configs[4], shortened: 500k individuals x 100k markers (2000 branches x 50 markers, widths [5,5,1], RidgeARD) on
N GPUs (rows sharded, one process per GPU), alternating Gibbs precision draws and HMC branch updates (`bann_sweep`), with the
posterior-predictive R^2 of the posterior-mean prediction on a held-out split.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 \
      scripts/full_chain.py --iterations 50 --group-size 64 [--individuals 500000 --branches 2000 --markers 50]

Data: device-generated synthetic genotypes (the generator is keyed by (column, global row): independent of the sharding), the
phenotype is the prediction of a random "true" net of the same architecture on those genotypes plus Gaussian noise, scaled to
heritability h2 (the recipe of `rs-bann simulate-xy`, rs-bann.rs:896-909); the held-out individuals are further rows of the
same generator.  Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rs_bann_b200 as rb  # noqa: E402
from rs_bann_b200 import architectures  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--individuals", type=int, default=500_000)
    ap.add_argument("--test-individuals", type=int, default=20_000)
    ap.add_argument("--branches", type=int, default=2000)
    ap.add_argument("--markers", type=int, default=50)
    ap.add_argument("--iterations", type=int, default=50)
    ap.add_argument("--integration-length", type=int, default=100)
    ap.add_argument("--group-size", type=int, default=64)
    ap.add_argument("--step-size", type=float, default=0.3)
    ap.add_argument("--h2", type=float, default=0.5)
    ap.add_argument("--causal-branches", type=int, default=200, help="branches of the true net with non-zero output weights")
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rb.Context(local, rank=rank, world=world)
    rb.connect_ranks(ctx)
    N, NT, B, per = a.individuals, a.test_individuals, a.branches, a.markers
    M = B * per
    widths = [5, 5, 1]
    t_setup = time.perf_counter()
    r0, r1 = rb.row_shard(N, rank, world)
    gen = rb.Genotypes.random(ctx, r1 - r0, M, None, seed=42, row_offset=r0, n_total=N + NT, uniform_groups=(B, per))

    def allreduce_counts(c):
        t = torch.from_numpy(c).cuda()
        if world > 1:
            dist.all_reduce(t)
        return t.cpu().numpy()

    mu, sd = rb.global_col_stats(gen.col_counts(), N, allreduce_counts)
    gen.set_col_stats(mu, sd)
    # held-out individuals: rows [N, N + NT) of the same generator, on rank 0 only (prediction is row-local)
    test = None
    if rank == 0:
        test = rb.Genotypes.random(rb.Context(local) if world > 1 else ctx, NT, M, None, seed=42, row_offset=N, n_total=N + NT,
                                   uniform_groups=(B, per))
        test.set_col_stats(mu, sd)

    # ---- phenotype from a random true net: a sparse set of causal branches, tanh units, heritability h2
    rng = np.random.default_rng(a.seed)
    true = architectures.build_net("ridge_ard", [per] * B, 1, fixed_hidden=5, fixed_summary=5, seed=a.seed + 100)
    causal = set(rng.choice(B, size=min(a.causal_branches, B), replace=False).tolist())
    for b, c in enumerate(true.branch_cfgs):
        c.weights[0] = (c.weights[0] * 3.0).astype(np.float32)
        if b not in causal:
            c.weights[-1][:] = 0.0
    def load(net, nf):
        pv = np.concatenate([c.param_vec() for c in nf.branch_cfgs]).astype(np.float32)
        qv = np.concatenate([np.where(np.isfinite(c.precision_vec()), c.precision_vec(), 1.0) for c in nf.branch_cfgs]).astype(np.float32)
        net.set_all_params(pv, qv)
        net.set_globals(nf.g_error_precision, nf.g_output_layer_precision, nf.g_ow_reg_sum, nf.g_ow_num_params, 0.0)

    net = rb.Net(ctx, gen, "ridge_ard", [widths] * B)
    rb.connect_net(net)
    load(net, true)
    gv_local = net.predict()                                   # this rank's rows
    s = torch.tensor([gv_local.astype(np.float64).sum(), (gv_local.astype(np.float64) ** 2).sum()], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(s)
    gmean = float(s[0]) / N
    gsd = float(np.sqrt(max(float(s[1]) / N - gmean ** 2, 1e-30)))
    noise = np.random.default_rng(1000 + rank).normal(size=r1 - r0)
    y_local = (np.sqrt(a.h2) * (gv_local - gmean) / gsd + np.sqrt(1 - a.h2) * noise).astype(np.float32)
    y_test = gv_test = None
    if rank == 0:
        tnet = rb.Net(test.ctx, test, "ridge_ard", [widths] * B)
        load(tnet, true)
        gv_test = (tnet.predict() - gmean) / gsd
        y_test = (np.sqrt(a.h2) * gv_test + np.sqrt(1 - a.h2) * np.random.default_rng(7).normal(size=NT)).astype(np.float32)
        tnet.close()

    # ---- the chain: reference default initial state, bann_sweep per iteration
    init = architectures.build_net("ridge_ard", [per] * B, 1, fixed_hidden=5, fixed_summary=5, seed=a.seed + 200)
    load(net, init)
    net.set_targets(y_local)
    net.init_residual()
    cfg = rb.MCMCCfg(hmc_step_size_factor=a.step_size, hmc_integration_length=a.integration_length)
    G = B if a.group_size == 0 else a.group_size
    order_rng = np.random.default_rng(a.seed + 300)              # same order on every rank
    ctx.sync()
    if world > 1:
        dist.barrier()
    t_setup = time.perf_counter() - t_setup
    hist, pred, kept = [], None, 0
    burn = a.iterations // 2
    ptest = None
    if rank == 0:
        ptest = rb.Net(test.ctx, test, "ridge_ard", [widths] * B)
    t_chain = 0.0
    for it in range(a.iterations):
        t0 = time.perf_counter()
        st = net.sweep(cfg, order_rng.permutation(B), seed=a.seed * 100000 + it, group_size=G)
        ctx.sync()
        t_chain += time.perf_counter() - t0
        hist.append((st["mse_train"], st["num_accepted"] / max(st["num_samples"], 1)))
        if it >= burn:                                           # held-out prediction of this posterior sample (rank 0)
            pv, qv = net.get_all_params()
            if rank == 0:
                ptest.set_all_params(pv, qv)
                ptest.set_globals(1.0, 1.0, 1.0, 1, st["output_bias"])
                p = ptest.predict()
                pred = p.astype(np.float64) if pred is None else pred + p
                kept += 1
    tt = torch.tensor([t_chain], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        pred /= max(kept, 1)
        r2 = 1.0 - float(np.mean((y_test - pred) ** 2)) / float(np.var(y_test))
        r2_true = 1.0 - float(np.mean((y_test - np.sqrt(a.h2) * gv_test) ** 2)) / float(np.var(y_test))
        print(json.dumps(dict(
            what="BASELINE configs[4] shortened: alternating Gibbs + HMC chain, rows sharded",
            individuals=N, test_individuals=NT, branches=B, markers_per_branch=per, widths=widths, prior="ridge_ard", n_gpus=world,
            iterations=a.iterations, integration_length=a.integration_length, group_size=G, step_size=a.step_size, h2=a.h2,
            causal_branches=len(causal), setup_s=round(t_setup, 2), chain_s=round(float(tt[0]), 3),
            s_per_iteration=round(float(tt[0]) / a.iterations, 4),
            branch_leapfrogs_per_s=round(a.iterations * B * a.integration_length / float(tt[0]), 1),
            mse_train_first_last=[round(hist[0][0], 4), round(hist[-1][0], 4)],
            acceptance_last=round(hist[-1][1], 3), posterior_mean_over=kept,
            r2_test=round(r2, 4), r2_test_of_true_genetic_value=round(r2_true, 4))), flush=True)
        ptest.close()
    net.close(); gen.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
