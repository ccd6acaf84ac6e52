"""One process per GPU (torchrun): the row-sharded sequential-exact chain over the peer-memory exchange.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tests/multirank_worker.py [--rate]

Check: every rank runs the same sweeps on its row shard -- sequential (also with early rejections) and block-Jacobi
groups -- and Net.gradient with 1 / world host slices; rank 0 also runs the unsharded chain on its own GPU and compares.  --rate: visits / s of the sharded chain at 100k rows.
Prints one line starting with MULTIRANK_OK or MULTIRANK_FAIL (rank 0)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rs_bann_b200 as rb  # noqa: E402
from test_gpu_sharded_chain import build_problem, make_net, run_chain  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = rb.Context(local, rank=rank, world=world)
    rb.connect_ranks(ctx)
    ok, msgs = True, []
    # (model, sweep kwargs): sequential-exact sweeps (push exchange inside KR), the same with a tight Hamiltonian-error bound so
    # that trajectories are rejected EARLY (every epoch the host hands out must still be exchanged, comm.cuh), and block-Jacobi
    # sweeps in groups of 3 (bulk exchange: reduce-scatter + all-gather over peer memory)
    # hmc_path: 0 = the persistent per-branch kernel where eligible (its cross-rank sums travel inside the kernel), 1 = a launch per step
    cases = [("ridge_ard", dict()), ("lasso_base", dict()),
             ("ridge_ard", dict(max_h_err=0.05, factor=1.5)),
             ("ridge_ard", dict(hmc_path=1)), ("ridge_ard", dict(hmc_path=1, max_h_err=0.05, factor=1.5)),
             ("ridge_ard", dict(group_size=3)), ("std_normal", dict(group_size=4)),
             # a branch with more than 64 markers takes the launch path (push exchange inside KR) between persistent transitions
             ("ridge_base", dict(groups=[20, 80, 9, 33]))]
    for model, kw_all in cases:
        kw = {k: v for k, v in kw_all.items() if k not in ("hmc_path", "groups")}
        P = build_problem(model, 3000, kw_all.get("groups", [20, 50, 9, 33]), 5, 5, seed=11)
        B = len(P["groups"])
        r0, r1 = rb.row_shard(P["n"], rank, world)
        gen, net = make_net(rb, ctx, P, r0, r1)
        rb.connect_net(net)
        net.select_hmc_path(kw_all.get("hmc_path", 0))
        out = run_chain(net, rb, P["y"][r0:r1], B, sweeps=2, L=8, **kw)
        n_persistent = net.persistent_launches()
        # Net.gradient on sharded rows: every rank passes only its 1 / world slice of the parameters and receives its slice
        pv_all = out["pv"].copy()
        plo, phi, olo, ohi = net.gradient_slice()
        pv_in = np.full_like(pv_all, np.nan)
        pv_in[plo:phi] = pv_all[plo:phi]                 # everything outside the slice is never read
        grads = np.full(net.num_params(), np.nan, dtype=np.float32)
        rss = np.full(B, np.nan, dtype=np.float32)
        net.gradient(pv_in, P["y"][r0:r1], out=(grads, rss))
        gr = np.concatenate([grads, rss])
        gparts = [None] * world
        dist.all_gather_object(gparts, (olo, ohi, gr[olo:ohi].copy(), bool(np.all(np.isnan(np.delete(gr, np.s_[olo:ohi]))))))
        net.close(); gen.close()
        # replicas identical: compare a digest over ranks
        dig = torch.tensor([float(np.sum(out["pv"].astype(np.float64))), float(np.sum(out["qv"].astype(np.float64))),
                            out["stats"]["num_accepted"], out["stats"]["num_early_rejected"], out["globals"]["output_bias"]],
                           dtype=torch.float64, device="cuda")
        lo, hi = dig.clone(), dig.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(lo, hi))
        resid = [None] * world
        dist.all_gather_object(resid, out["resid"])
        if rank == 0:
            c1 = rb.Context(local)
            g1, n1 = make_net(rb, c1, P, 0, P["n"])
            ref = run_chain(n1, rb, P["y"], B, sweeps=2, L=8, **kw)
            g_ref, rss_ref = n1.gradient(pv_all, P["y"])
            n1.close(); g1.close(); c1.close()
            s, s1 = out["stats"], ref["stats"]
            gfull = np.concatenate([p[2] for p in sorted(gparts, key=lambda p: p[0])])
            gexp = np.concatenate([g_ref, rss_ref])
            grad_ok = (gfull.size == gexp.size and all(p[3] for p in gparts)
                       and np.allclose(gfull, gexp, rtol=2e-4, atol=2e-4 * float(np.max(np.abs(gexp)))))
            good = (same and s["num_accepted"] == s1["num_accepted"] and s["num_early_rejected"] == s1["num_early_rejected"]
                    and np.allclose(out["pv"], ref["pv"], rtol=2e-4, atol=2e-5)
                    and np.allclose(out["qv"], ref["qv"], rtol=2e-4, atol=2e-5)
                    and np.allclose(np.concatenate(resid), ref["resid"], rtol=0, atol=5e-4)
                    and abs(s["lpd"] - s1["lpd"]) < 5e-4 * abs(s1["lpd"]) and grad_ok)
            if "max_h_err" in kw:
                good = good and s["num_early_rejected"] > 0        # the case exists to exercise early rejections
            if kw.get("group_size", 1) == 1:                       # sequential sweeps: every visit through the path asked for
                eligible = sum(1 for g in P["groups"] if len(g) <= 64)
                good = good and n_persistent == (0 if kw_all.get("hmc_path", 0) == 1 else 2 * eligible)
            ok = ok and good
            msgs.append(f"{model} {kw_all}: persistent launches {n_persistent}, replicas_identical={same} accepted {s['num_accepted']}/{s['num_samples']} early "
                        f"{s['num_early_rejected']} (single rank {s1['num_accepted']}, {s1['num_early_rejected']}) "
                        f"max|dtheta|={np.max(np.abs(out['pv'] - ref['pv'])):.2e} lpd {s['lpd']:.4f} vs {s1['lpd']:.4f} "
                        f"sliced_gradient_ok={grad_ok}")
    if "--rate" in sys.argv:
        from bench import default_params
        n, B, per, L = 100000, 64, 50, 100
        r0, r1 = rb.row_shard(n, rank, world)
        gen = rb.Genotypes.random(ctx, r1 - r0, B * per, None, seed=42, row_offset=r0, n_total=n, uniform_groups=(B, per))

        def allreduce_counts(c):
            t = torch.from_numpy(c).cuda()
            dist.all_reduce(t)
            return t.cpu().numpy()

        mu, sd = rb.global_col_stats(gen.col_counts(), n, allreduce_counts)
        gen.set_col_stats(mu, sd)
        wl = dict(B=B, per=per, widths=[5, 5, 1], model="ridge_ard")
        net = rb.Net(ctx, gen, "ridge_ard", [[5, 5, 1]] * B)
        pv, qv = default_params(wl)
        net.set_all_params(pv, qv)
        w_out = pv.reshape(B, -1)[:, per * 5 + 25:per * 5 + 30]
        net.set_globals(2.0, 0.05, float(np.sum(w_out ** 2)), B * 5)
        net.set_targets(np.random.default_rng(1).normal(size=n).astype(np.float32)[r0:r1])
        net.init_residual()
        rb.connect_net(net)
        cfg = rb.MCMCCfg(hmc_step_size_factor=0.1, hmc_integration_length=L, hmc_max_hamiltonian_error=1e30)
        for path, name in ((1, "launch per step"), (0, "persistent kernel")):
            net.select_hmc_path(path)
            net.sweep(cfg, np.arange(B), seed=1)
            ctx.sync()
            dist.barrier()
            before = net.persistent_launches()
            t0 = time.perf_counter()
            st = net.sweep(cfg, np.random.default_rng(2).permutation(B), seed=2)
            ctx.sync()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt[0])
            if rank == 0:
                msgs.append(f"rate ({name}, {net.persistent_launches() - before} persistent launches): world={world} n={n} B={B} "
                            f"m_b={per} L={L}: {B / dt:.1f} visits/s, {dt / B / L * 1e6:.1f} us per leapfrog, "
                            f"accepted {st['num_accepted']}/{st['num_samples']}")
        net.close(); gen.close()
    ctx.close()
    if rank == 0:
        print(("MULTIRANK_OK " if ok else "MULTIRANK_FAIL ") + " | ".join(msgs), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
