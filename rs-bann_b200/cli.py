"""`rs-bann` command line surface kept by this build: train-new, train, predict, simulate-xy, branch-r2.

Same positional / optional arguments, output directory naming and output files as the reference
(src/bin/cli/cli.rs:62-153,257-316,325-404,425-435; src/bin/rs-bann.rs:276-312,793-964,1006-1218): `args.json`,
`hyperparams`, `training_stats`, `trace`, `models/<ix>.bin` (bincode, readable by the reference) and the predict CSV.
The chain itself runs on the GPU through the C ABI (include/bann.h); there is no CPU path.
"""
import argparse
import glob
import json
import os
import sys
from typing import List, Optional

import numpy as np

from . import files
from .architectures import build_net
from .api import Context, Genotypes, MCMCCfg, Net

MODEL_CLI = {"ridge-ard": "ridge_ard", "ridge-base": "ridge_base", "lasso-ard": "lasso_ard", "lasso-base": "lasso_base",
             "std-normal": "std_normal"}
MODEL_JSON = {"ridge_ard": "RidgeARD", "ridge_base": "RidgeBase", "lasso_ard": "LassoARD", "lasso_base": "LassoBase",
              "std_normal": "StdNormal"}                       # net/model_type.rs:6-13 (serde / Display names)
ACT_CLI = {"tanh": "tanh", "re-lu": "relu", "relu": "relu", "leaky-re-lu": "leaky_relu", "leaky-relu": "leaky_relu",
           "si-lu": "silu", "silu": "silu", "identity": "identity"}
STEP_CLI = {"uniform": "uniform", "random": "random", "std-scaled": "std_scaled", "izmailov": "izmailov"}
STEP_DISPLAY = {"uniform": "Uniform", "random": "Random", "std_scaled": "StdScaled", "izmailov": "Izmailov"}


def _fmt(v) -> str:
    """Rust `{}` formatting of the numbers that end up in directory names (1.0 -> "1", 0.001 -> "0.001")."""
    if isinstance(v, float) and v == int(v) and abs(v) < 1e15:
        return str(int(v))
    return repr(v) if isinstance(v, float) else str(v)


def _model_type(s: str) -> str:
    k = s.lower()
    if k in MODEL_CLI:
        return MODEL_CLI[k]
    if k.replace("-", "_") in MODEL_JSON:
        return k.replace("-", "_")
    for internal, js in MODEL_JSON.items():
        if js.lower() == k:
            return internal
    raise argparse.ArgumentTypeError(f"unknown model type {s!r} (Linear models are not supported)")


def _activation(s: str) -> str:
    k = s.lower()
    if k not in ACT_CLI:
        raise argparse.ArgumentTypeError(f"unknown activation function {s!r}")
    return ACT_CLI[k]


def _step_mode(s: str) -> str:
    k = s.lower()
    if k not in STEP_CLI:
        raise argparse.ArgumentTypeError(f"unknown step size mode {s!r}")
    return STEP_CLI[k]


# ------------------------------------------------------------------ argument groups (clap structs of the reference)
def add_train_io(p):
    p.add_argument("bfile_train"); p.add_argument("p_train"); p.add_argument("groups")
    p.add_argument("--bfile-test"); p.add_argument("--p-test")
    p.add_argument("-o", "--outpath", default="./")


def add_mcmc(p):
    p.add_argument("chain_length", type=int); p.add_argument("integration_length", type=int)
    p.add_argument("--max-hamiltonian-error", type=float, default=10.0)
    p.add_argument("--step-size", type=float, default=1.0)
    p.add_argument("--report-interval", type=int, default=1)
    p.add_argument("--fixed-param-precision", type=float)
    p.add_argument("--step-size-mode", type=_step_mode, default="izmailov")
    p.add_argument("-d", "--debug-prints", action="store_true")
    p.add_argument("--trace", action="store_true")
    p.add_argument("--burn-in", type=int)
    for unsupported in ("--trajectories", "--num-grad-traj", "--num-grad", "--gradient-descent", "--gradient-descent-joint"):
        p.add_argument(unsupported, action="store_true")
    p.add_argument("-j", "--joint-hmc", action="store_true")
    p.add_argument("--seed", type=int, default=None, help="(extension) seed of the chain's counter-based RNG; the reference is unseedable")
    p.add_argument("--device", type=int, default=0, help="(extension) CUDA device")
    p.add_argument("--group-size", type=int, default=1,
                   help="(extension) branches advanced concurrently per group visit: 1 = the reference's sequential order; "
                        "G > 1 = block-Jacobi groups of G branches against the residual frozen at group start; 0 = all branches")


def _check_supported(a):
    if (a.num_grad or a.num_grad_traj) and getattr(a, "group_size", 1) not in (1, None):
        sys.exit("--num-grad / --num-grad-traj are per-branch debugging aids of the sequential order (group size 1)")
    if a.num_grad_traj and not a.trajectories:
        pass        # mcmc_cfg.rs: num_grad_traj only matters while trajectories are recorded (branch_sampler.rs:1259)


# ------------------------------------------------------------------ device <-> file state
def net_to_device(ctx, gen, model: str, nf: files.NetFile) -> Net:
    act = nf.branch_cfgs[0].activation
    net = Net(ctx, gen, model, [c.layer_widths for c in nf.branch_cfgs], hyper=tuple(nf.hyper), activation=act)
    for b, c in enumerate(nf.branch_cfgs):
        if c.num_markers != len(gen.groups[b]):
            sys.exit(f"branch {b}: the model expects {c.num_markers} markers, the grouping has {len(gen.groups[b])}")
        net.set_branch(b, c.param_vec(), c.precision_vec())
    net.set_globals(nf.g_error_precision, nf.g_output_layer_precision, nf.g_ow_reg_sum, nf.g_ow_num_params, nf.output_bias[2])
    return net


def net_from_device(net: Net, nf: files.NetFile) -> files.NetFile:
    """to_cfg of every branch + global state (branch_sampler.rs:155-171, params.rs:41-56) into the file structure."""
    g = net.get_globals()
    for b, c in enumerate(nf.branch_cfgs):
        pv, qv = net.get_branch(b)
        c.load_param_vec(pv)
        c.load_precision_vec(qv)
        c.ow_reg_sum, c.ow_num_params = float(g["ow_reg_sum"]), int(g["ow_num_params"])
    st = net.stats()
    nf.output_bias = [float(g["error_precision"]), nf.output_bias[1], float(g["output_bias"])]
    nf.num_samples, nf.num_accepted, nf.num_early_rejected = st["num_samples"], st["num_accepted"], st["num_early_rejected"]
    a, o, loc = net.lpd_terms()
    nf.lpd_rss, nf.lpd_out_w, nf.lpd_local = a, o, [float(x) for x in loc]
    nf.g_error_precision, nf.g_output_layer_precision = float(g["error_precision"]), float(g["output_layer_precision"])
    nf.g_ow_reg_sum, nf.g_ow_num_params = float(g["ow_reg_sum"]), int(g["ow_num_params"])
    return nf


def _load_data(ctx, bfile, groups_path, phen_path=None, shard=False):
    """BedVM::from_file + grouping + phenotypes.  shard=True (training data under torchrun): this rank keeps rows
    [r0, r1) of every column (128-row tile boundaries); column statistics come from all-reduced value counts."""
    payload, n, m = files.read_bed(bfile)
    groups = files.read_grouping(groups_path)
    y = files.read_phen(phen_path) if phen_path else None
    if y is not None and y.size != n:
        sys.exit(f"{phen_path}: {y.size} phenotypes for {n} individuals")
    if shard and ctx.world > 1:
        import torch
        import torch.distributed as dist
        from .dist import global_col_stats, row_shard, shard_payload
        r0, r1 = row_shard(n, ctx.rank, ctx.world)
        if r1 <= r0:
            sys.exit(f"{bfile}: {n} individuals are too few for {ctx.world} ranks (shards are whole 128-row tiles)")
        one = np.ones(m, dtype=np.float32)
        gen = Genotypes(ctx, shard_payload(payload, n, m, r0, r1), r1 - r0, m, groups, col_means=0 * one, col_stds=one,
                        n_total=n)

        def allreduce(c):
            t = torch.from_numpy(c).cuda()
            dist.all_reduce(t)
            return t.cpu().numpy()

        gen.set_col_stats(*global_col_stats(gen.col_counts(), n, allreduce))
        return gen, (y[r0:r1] if y is not None else None)
    gen = Genotypes(ctx, payload, n, m, groups)
    return gen, y


def _dist_context(device: int):
    """One process per GPU when launched by torchrun (RANK / WORLD_SIZE / LOCAL_RANK); otherwise a single rank."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world == 1:
        return Context(device), 0, 1
    import torch
    import torch.distributed as dist
    from .dist import connect_ranks
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", os.environ["RANK"]))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local, rank=rank, world=world)
    connect_ranks(ctx)           # peer-mapped inboxes: cross-rank sums inside the reduction kernels (csrc/comm.cuh)
    return ctx, rank, world


def _prepare_dist(a):
    """Ranks must agree on every random choice made on the host (initial weights, branch orders)."""
    a._dist = _dist_context(a.device)
    if a._dist[2] > 1:
        a.seed = _bcast_int(a.seed if a.seed is not None else int.from_bytes(os.urandom(4), "little"), a._dist[2])


def _bcast_int(v: int, world: int) -> int:
    if world == 1:
        return v
    import torch.distributed as dist
    box = [v]
    dist.broadcast_object_list(box, src=0)
    return int(box[0])


def run_chain(a, model: str, nf: files.NetFile, outdir: str, args_json: dict):
    """Net::train (net/net.rs:201-358) on the device: the reference's sequential order, or block-Jacobi groups (--group-size)."""
    ctx, rank, world = a._dist if getattr(a, "_dist", None) else _dist_context(a.device)
    gen, y = _load_data(ctx, a.bfile_train, a.groups, a.p_train, shard=True)
    test = None
    if a.bfile_test and a.p_test:
        test = _load_data(ctx, a.bfile_test, a.groups, a.p_test)                  # replicated: predict is row-local
    lead = rank == 0              # replicas hold identical state; rank 0 writes the files
    burn_in = a.burn_in if a.burn_in is not None else a.chain_length - 1          # mcmc_cfg.rs:152-156
    net = net_to_device(ctx, gen, model, nf)
    if world > 1:                 # block-Jacobi groups on sharded rows sum over ranks through the net's bulk exchange region
        from .dist import connect_net
        connect_net(net)
    if lead:
        os.makedirs(outdir, exist_ok=True)
        with open(os.path.join(outdir, "args.json"), "w") as f:
            json.dump(args_json, f, indent=2)
        with open(os.path.join(outdir, "hyperparams"), "w") as f:                 # net.rs:149-156
            json.dump(files.hyperparams_json(nf), f)
        if a.chain_length > burn_in:
            os.makedirs(os.path.join(outdir, "models"), exist_ok=True)
            os.makedirs(os.path.join(outdir, "effect_sizes"), exist_ok=True)
    net.set_targets(y)
    net.init_residual()
    cfg = MCMCCfg(hmc_step_size_factor=a.step_size, hmc_max_hamiltonian_error=a.max_hamiltonian_error,
                  hmc_integration_length=a.integration_length, hmc_step_size_mode=a.step_size_mode,
                  fixed_param_precisions=a.fixed_param_precision is not None, joint_hmc=a.joint_hmc,
                  gradient_descent=a.gradient_descent, gradient_descent_joint=a.gradient_descent_joint,   # net.rs:282-290
                  num_grad=a.num_grad, num_grad_traj=a.num_grad_traj)
    seed = _bcast_int(a.seed if a.seed is not None else int.from_bytes(os.urandom(4), "little"), world)
    rng = np.random.default_rng(seed)
    group_size = net.num_branches if a.group_size == 0 else max(1, min(a.group_size, net.num_branches))
    if group_size > 1 and (a.joint_hmc or a.gradient_descent or a.gradient_descent_joint or a.trajectories):
        sys.exit("--group-size > 1 runs the HMC sampler; --joint-hmc / --gradient-descent* / --trajectories are sequential modes")
    trace = open(os.path.join(outdir, "trace"), "w") if (a.trace and lead) else None
    traj_file = open(os.path.join(outdir, "traj"), "a") if (a.trajectories and lead) else None   # mcmc_cfg.rs:247-249, appended

    def record_perf(st):                                                          # net.rs:597-610
        nf.lpd.append(float(st["lpd"]))
        nf.mse_train.append(float(st["mse_train"]))
        if test is not None:
            r = test[1] - net.predict(test[0])
            nf.mse_test = (nf.mse_test or []) + [float(np.sum(r.astype(np.float32) ** 2, dtype=np.float32) / np.float32(r.size))]

    def report(i, st):                                                            # net.rs:670-693
        if not lead:
            return
        ns = max(st["num_samples"], 1)
        acc, early = st["num_accepted"] / ns, st["num_early_rejected"] / ns
        line = (f"i: {i} \t | acc: {acc:.2f} \t | early_rej: {early:.2f} \t | end_rej: {1 - acc - early:.2f} \t | "
                f"mse(trn): {nf.mse_train[-1]:.4f}")
        if nf.mse_test is not None:
            line += f" \t | mse(tst): {nf.mse_test[-1]:.4f}"
        print(line + f" | lpd: {nf.lpd[-1]:.4f}", file=sys.stderr)

    def dump_trace():
        net_from_device(net, nf)
        trace.write(files.json_dumps([c.to_json() for c in nf.branch_cfgs]) + "\n")

    def save_model(ix):
        if lead:
            files.write_net(os.path.join(outdir, "models", f"{ix}.bin"), net_from_device(net, nf))

    if lead:
        print(f"Training net with {net.num_branches} branches, {net.num_params()} params"
              + (f", rows sharded over {world} GPUs" if world > 1 else ""), file=sys.stderr)
    st = net.stats()
    record_perf(st)
    report(0, st)
    if trace:
        dump_trace()
    if burn_in == 0:
        save_model(0)
    for chain_ix in range(1, a.chain_length + 1):
        order = rng.permutation(net.num_branches)                                 # net.rs:257
        if a.trajectories:                                                        # branch_sampler.rs:1198-1207,1286-1289: one JSON line per transition
            for b in order:
                res = net.visit_branch_traj(int(b), cfg, seed=seed + chain_ix)
                if traj_file is not None:
                    t = res.trajectory
                    traj_file.write(files.json_dumps(dict(params=[[float(v) for v in r] for r in t["params"]],
                                                    precisions=[[float(v) for v in r] for r in t["precisions"]],
                                                    ldg=[[float(v) for v in r] for r in t["ldg"]],
                                                    num_ldg=[[float(v) for v in r] for r in t["num_ldg"]],
                                                    hamiltonian=[float(v) for v in t["hamiltonian"]])) + "\n")
            st = net.stats()
        else:
            st = net.sweep(cfg, order, seed=seed + chain_ix, group_size=group_size)
        record_perf(st)
        if chain_ix >= burn_in:
            save_model(chain_ix)
        if chain_ix % a.report_interval == 0:
            report(chain_ix, st)
        if trace:
            dump_trace()
    net_from_device(net, nf)
    if lead:
        with open(os.path.join(outdir, "training_stats"), "w") as f:              # train_stats.rs:83-87
            files.json_dump(nf.training_stats_json(), f)
        if trace:
            trace.close()
        if traj_file:
            traj_file.close()
        print("Completed training", file=sys.stderr)
    net.close(); gen.close()
    if test is not None:
        test[0].close()
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    return outdir if lead else None


def _replicate_dir(parent: str, outdir: str) -> str:
    """set_replicate_ix (rs-bann.rs:776-787)."""
    rep = 1
    while os.path.exists(os.path.join(parent, f"{outdir}_rep{rep}")):
        rep += 1
    return os.path.join(parent, f"{outdir}_rep{rep}")


# ------------------------------------------------------------------ subcommands
def cmd_train_new(a):
    _check_supported(a)
    _prepare_dist(a)
    model = a.model_type
    outdir = (f"{MODEL_JSON[model]}_{files.ACTIVATION_JSON[files.ACTIVATIONS.index(a.activation_function)]}_d{a.branch_depth}"
              f"_cl{a.chain_length}_il{a.integration_length}_{STEP_DISPLAY[a.step_size_mode]}_st{_fmt(a.step_size)}"
              f"_dpk{_fmt(a.dpk)}_dps{_fmt(a.dps)}_spk{_fmt(a.spk)}_sps{_fmt(a.sps)}_opk{_fmt(a.opk)}_ops{_fmt(a.ops)}")
    outdir += ("_joint" if a.joint_hmc else "") + ("_gd" if a.gradient_descent else "") + ("_gdj" if a.gradient_descent_joint else "")   # rs-bann.rs:1036-1046
    if a.fixed_param_precision is not None:
        outdir += f"_fp{_fmt(a.fixed_param_precision)}"
    outdir += f"_fhlw{a.fixed_hidden_layer_width}" if a.fixed_hidden_layer_width is not None else \
        f"_rhlw{_fmt(a.relative_hidden_layer_width)}"
    outdir += f"_fslw{a.fixed_summary_layer_width}" if a.fixed_summary_layer_width is not None else \
        f"_rslw{_fmt(a.relative_summary_layer_width)}"
    path = _replicate_dir(a.outpath, outdir)
    groups = files.read_grouping(a.groups)
    nf = build_net(model, [len(g) for g in groups], a.branch_depth, a.activation_function,
                   fixed_hidden=a.fixed_hidden_layer_width, rel_hidden=a.relative_hidden_layer_width,
                   fixed_summary=a.fixed_summary_layer_width, rel_summary=a.relative_summary_layer_width,
                   hyper=(a.dpk, a.dps, a.spk, a.sps, a.opk, a.ops), fixed_param_precision=a.fixed_param_precision,
                   seed=a.seed)
    args_json = dict(model_type=MODEL_JSON[model],
                     activation_function=files.ACTIVATION_JSON[files.ACTIVATIONS.index(a.activation_function)],
                     branch_depth=a.branch_depth, relative_hidden_layer_width=a.relative_hidden_layer_width,
                     fixed_hidden_layer_width=a.fixed_hidden_layer_width,
                     relative_summary_layer_width=a.relative_summary_layer_width,
                     fixed_summary_layer_width=a.fixed_summary_layer_width, dpk=a.dpk, dps=a.dps, spk=a.spk, sps=a.sps,
                     opk=a.opk, ops=a.ops)                                         # cli.rs:350-404
    out = run_chain(a, model, nf, path, args_json)
    if out:
        print(out)


def cmd_train(a):
    _check_supported(a)
    if not os.path.isfile(a.model_file):
        sys.exit("Specified model: No such file found")                           # rs-bann.rs:1147-1150
    _prepare_dist(a)
    stem = os.path.splitext(os.path.basename(a.model_file))[0]
    outdir = (f"{stem}_cl{a.chain_length}_il{a.integration_length}_{STEP_DISPLAY[a.step_size_mode]}_st{_fmt(a.step_size)}"
              f"_dtheta{_fmt(a.perturb_params or 0.0)}_dlambda{_fmt(a.perturb_precisions or 0.0)}")
    outdir += ("_joint" if a.joint_hmc else "") + ("_gd" if a.gradient_descent else "") + ("_gdj" if a.gradient_descent_joint else "")   # rs-bann.rs:1164-1174
    if a.fixed_param_precision is not None:
        outdir += "_fp"
    nf = files.read_net(a.model_file)
    rng = np.random.default_rng(a.seed)
    if a.perturb_params:                                                           # net.rs:187-199
        for c in nf.branch_cfgs:
            c.weights = [w + rng.normal(0, a.perturb_params, size=w.shape).astype(np.float32) for w in c.weights]
            c.biases = [b + rng.normal(0, a.perturb_params, size=b.shape).astype(np.float32) for b in c.biases]
    if a.perturb_precisions:
        for c in nf.branch_cfgs:
            c.weight_precisions = [np.abs(p + rng.normal(0, a.perturb_precisions, size=p.shape)).astype(np.float32)
                                   for p in c.weight_precisions]
    args_json = dict(model_type=MODEL_JSON[a.model_type], model_file=a.model_file, perturb_params=a.perturb_params,
                     perturb_precisions=a.perturb_precisions)                     # cli.rs:325-338
    out = run_chain(a, a.model_type, nf, os.path.join(a.outpath, outdir) if a.outpath != "./" else outdir, args_json)
    if out:
        print(out)


def _model_files(model_path: str) -> List[str]:
    return sorted(p for p in glob.glob(os.path.join(model_path, "*")) if os.path.isfile(p))   # rs-bann.rs:291-298


def _read_model_type(model_path: str) -> str:
    args_path = os.path.join(os.path.dirname(os.path.normpath(model_path)), "args.json")        # rs-bann.rs:282-286
    return _model_type(json.load(open(args_path))["model_type"])


def cmd_predict(a):
    ctx = Context(a.device)
    gen, _ = _load_data(ctx, a.bfile, a.groups)
    model = _read_model_type(a.model_path)
    for path in _model_files(a.model_path):
        nf = files.read_net(path)
        net = net_to_device(ctx, gen, model, nf)
        print(",".join(repr(float(v)) for v in net.predict()))                    # csv row per model (rs-bann.rs:300-311)
        net.close()
    gen.close(); ctx.close()


def r2(y, yhat) -> float:
    """1 - mse / variance (py-vis/vis.py:555-557), population variance."""
    y, yhat = np.asarray(y, dtype=np.float64), np.asarray(yhat, dtype=np.float64)
    return float(1.0 - np.mean((y - yhat) ** 2) / np.var(y))


def cmd_r2(a):
    """(extension) posterior predictive R^2 of the saved models on a phenotype file: per model and of the mean prediction."""
    ctx = Context(a.device)
    gen, y = _load_data(ctx, a.bfile, a.groups, a.phen)
    model = _read_model_type(a.model_path)
    preds = []
    for path in _model_files(a.model_path):
        net = net_to_device(ctx, gen, model, files.read_net(path))
        preds.append(net.predict())
        net.close()
        print(f"{os.path.basename(path)}\t{r2(y, preds[-1]):.6f}")
    print(f"posterior_mean\t{r2(y, np.mean(preds, axis=0)):.6f}")
    gen.close(); ctx.close()


def cmd_branch_r2(a):
    """branch_r2 (rs-bann.rs:128-170, net.rs:648-656): per model a CSV row of 1 - rss_b / sum(y^2) for every branch."""
    ctx = Context(a.device)
    gen, y = _load_data(ctx, a.bfile, a.groups, a.phen)
    model = _read_model_type(a.model_path)
    yy = float(np.sum(y.astype(np.float32) ** 2, dtype=np.float32))
    for path in _model_files(a.model_path):
        net = net_to_device(ctx, gen, model, files.read_net(path))
        net.set_targets(y)
        print(",".join(repr(1.0 - float(net.branch_fwd_bwd(b, want_yhat=False)["rss"]) / yy) for b in range(net.num_branches)))
        net.close()
    gen.close(); ctx.close()


def cmd_gradients(a):
    """gradients (rs-bann.rs:226-274, net.rs:520-527): log-density gradient of every branch against y, one JSON file per
    model under <model dir>/../gradients/.  (The reference serialises ArrayFire arrays through afserde; here each branch
    is {"wrt_weights": [[...] per layer, column-major], "wrt_biases": [[...] per layer]}.)"""
    ctx = Context(a.device)
    gen, y = _load_data(ctx, a.bfile, a.groups, a.phen)
    model = _read_model_type(a.model_path)
    outdir = os.path.join(os.path.dirname(os.path.normpath(a.model_path)), "gradients")
    os.makedirs(outdir, exist_ok=True)
    for path in _model_files(a.model_path):
        nf = files.read_net(path)
        net = net_to_device(ctx, gen, model, nf)
        grads, _ = net.gradient(y=y)                                              # one fused full-network launch
        out, off = [], 0
        for c in nf.branch_cfgs:
            g = grads[off:off + c.num_params]
            off += c.num_params
            ws, bs, ix, prev = [], [], 0, c.num_markers
            for w in c.layer_widths:
                ws.append([float(v) for v in g[ix:ix + prev * w]]); ix += prev * w; prev = w
            for w in c.layer_widths[:-1]:
                bs.append([float(v) for v in g[ix:ix + w]]); ix += w
            out.append(dict(wrt_weights=ws, wrt_biases=bs))
        stem = os.path.splitext(os.path.basename(path))[0]
        json.dump(out, open(os.path.join(outdir, stem + ".json"), "w"))
        net.close()
    gen.close(); ctx.close()
    print(outdir)


def cmd_population_effect_sizes(a):
    """population-effect-sizes (rs-bann.rs:314-372, net.rs:529-543): column means of every branch's effect sizes, one JSON
    list (all branches concatenated in branch order) per model under <model dir>/../population_effect_sizes/."""
    ctx = Context(a.device)
    gen, _ = _load_data(ctx, a.bfile, a.groups, a.phen)
    model = _read_model_type(a.model_path)
    outdir = os.path.join(os.path.dirname(os.path.normpath(a.model_path)), "population_effect_sizes")
    os.makedirs(outdir, exist_ok=True)
    for path in _model_files(a.model_path):
        net = net_to_device(ctx, gen, model, files.read_net(path))
        pes = net.population_effect_sizes()
        stem = os.path.splitext(os.path.basename(path))[0]
        json.dump([float(v) for v in pes], open(os.path.join(outdir, stem + ".json"), "w"))
        net.close()
    gen.close(); ctx.close()
    print(outdir)


def cmd_activations(a):
    """activations (rs-bann.rs:175-224, net.rs:509-518): the forward pass of every branch, one JSON file per model under
    <model dir>/../activations/.  (The reference serialises ArrayFire arrays through afserde; here
    {"activations": [[{"dims": [n, w], "data": [column-major]} per layer incl. the prediction] per branch]}.)"""
    ctx = Context(a.device)
    gen, _ = _load_data(ctx, a.bfile, a.groups)
    model = _read_model_type(a.model_path)
    outdir = os.path.join(os.path.dirname(os.path.normpath(a.model_path)), "activations")
    os.makedirs(outdir, exist_ok=True)
    for path in _model_files(a.model_path):
        nf = files.read_net(path)
        net = net_to_device(ctx, gen, model, nf)
        out = [[dict(dims=list(arr.shape), data=[float(v) for v in arr.reshape(-1, order="F")])
                for arr in net.branch_activations(b)] for b in range(len(nf.branch_cfgs))]
        stem = os.path.splitext(os.path.basename(path))[0]
        json.dump(dict(activations=out), open(os.path.join(outdir, stem + ".json"), "w"))
        net.close()
    gen.close(); ctx.close()
    print(outdir)


def cmd_simulate_xy(a):
    """simulate_xy (rs-bann.rs:793-964): random genotypes (io/bed.rs:136-188), a random net, y = net(X) + noise."""
    if not 0.0 <= a.heritability <= 1.0:
        sys.exit("Heritability must be within [0, 1].")
    ctx = Context(a.device)
    model, B, per, n = a.model_type, a.num_branches, a.num_markers_per_branch, a.num_individuals
    name = (f"{MODEL_JSON[model]}_{files.ACTIVATION_JSON[files.ACTIVATIONS.index(a.activation_function)]}_b{B}"
            f"_wh{a.hidden_layer_width}_ws{a.summary_layer_width or a.hidden_layer_width}"
            f"_d{a.branch_depth}_m{per}_n{n}_h{_fmt(a.heritability)}")            # rs-bann.rs:803-814
    if a.init_param_variance is not None:
        name += f"_v{a.init_param_variance!r}"
    path = _replicate_dir(a.outdir, name)
    os.makedirs(path, exist_ok=True)
    rng = np.random.default_rng(a.seed)
    groups = [list(range(b * per, (b + 1) * per)) for b in range(B)]             # UniformGrouping (group/uniform.rs:4-41)
    m = B * per
    while True:
        nf = build_net(model, [per] * B, a.branch_depth, a.activation_function, fixed_hidden=a.hidden_layer_width,
                       fixed_summary=a.summary_layer_width, rel_summary=None, init_param_variance=a.init_param_variance,
                       seed=int(rng.integers(1 << 31)))
        mafs = rng.uniform(0.0, 0.5, size=m)                                      # rs-bann.rs:877-880

        def random_bed():
            g = rng.binomial(2, mafs[None, :], size=(n, m)).astype(np.uint8)
            for j in np.nonzero(g.min(axis=0) == g.max(axis=0))[0]:               # no monomorphic columns (bed.rs:156-185)
                while g[:, j].min() == g[:, j].max():
                    g[:, j] = rng.binomial(2, max(mafs[j], 0.05), size=n)
            return files.pack_genotypes(g)
        out = {}
        for split in ("train", "test"):
            payload = random_bed()
            gen = Genotypes(ctx, payload, n, m, groups)
            net = net_to_device(ctx, gen, model, nf)
            gv = net.predict()
            net.close(); gen.close()
            y, resid_var = gv.copy(), 0.0
            if a.heritability != 1.0:
                resid_var = float(np.var(gv.astype(np.float64), ddof=1)) * (1.0 / a.heritability - 1.0)
                y = (gv + rng.normal(0.0, np.sqrt(resid_var), size=n)).astype(np.float32)
            out[split] = (payload, gv, y, resid_var)
        if a.heritability != 1.0 and min(out["train"][3], out["test"][3]) < 0.01:
            continue                                                               # rs-bann.rs:899-907
        break
    files.write_net(os.path.join(path, "model.bin"), nf)
    with open(os.path.join(path, "model.params"), "w") as f:
        f.write(files.json_dumps([c.to_json() for c in nf.branch_cfgs]) + "\n")
    for split, (payload, gv, y, resid_var) in out.items():
        stem = os.path.join(path, split)
        files.write_bed(stem, payload, n, m)
        files.write_grouping(stem + ".groups", groups)
        files.write_phen(stem + ".phen", y)
        y64 = y.astype(np.float64)
        files.write_phen_stats(os.path.join(path, f"{split}_phen_stats.json"), y64.mean(), y64.var(ddof=1), resid_var)
        if a.json_data:
            json.dump(dict(y=[float(v) for v in gv]), open(os.path.join(path, f"genetic_values_{split}.json"), "w"))
            json.dump(dict(y=[float(v) for v in y]), open(os.path.join(path, f"phen_{split}.json"), "w"))
    json.dump({k: v for k, v in vars(a).items() if k not in ("func", "device", "seed")} |
              dict(model_type=MODEL_JSON[model]), open(os.path.join(path, "args.json"), "w"), indent=2)
    ctx.close()
    print(path)


def build_parser():
    ap = argparse.ArgumentParser(prog="rs-bann", description="rs-bann hot path on B200 (train / predict surface)")
    sub = ap.add_subparsers(dest="cmd", required=True)
    p = sub.add_parser("train-new", help="Train new model on data in .bed format")
    add_train_io(p); add_mcmc(p)
    p.add_argument("model_type", type=_model_type); p.add_argument("activation_function", type=_activation)
    p.add_argument("branch_depth", type=int)
    p.add_argument("--relative-hidden-layer-width", type=float, default=0.5)
    p.add_argument("--fixed-hidden-layer-width", type=int)
    p.add_argument("--relative-summary-layer-width", type=float, default=1.0)
    p.add_argument("--fixed-summary-layer-width", type=int)
    for k, d in (("dpk", 0.001), ("dps", 1000.0), ("spk", 0.001), ("sps", 1000.0), ("opk", 0.001), ("ops", 1000.0)):
        p.add_argument(f"--{k}", type=float, default=d)
    p.set_defaults(func=cmd_train_new)
    p = sub.add_parser("train", help="Train prespecified model.")
    add_train_io(p); add_mcmc(p)
    p.add_argument("model_type", type=_model_type); p.add_argument("model_file")
    p.add_argument("--perturb-params", type=float); p.add_argument("--perturb-precisions", type=float)
    p.set_defaults(func=cmd_train)
    p = sub.add_parser("predict", help="Use trained model to predict phenotypes.")
    p.add_argument("bfile"); p.add_argument("groups"); p.add_argument("-m", "--model-path", default="./models")
    p.add_argument("--device", type=int, default=0)
    p.set_defaults(func=cmd_predict)
    p = sub.add_parser("r2", help="(extension) posterior predictive R^2 of saved models")
    p.add_argument("bfile"); p.add_argument("phen"); p.add_argument("groups"); p.add_argument("-m", "--model-path", default="./models")
    p.add_argument("--device", type=int, default=0)
    p.set_defaults(func=cmd_r2)
    for name, fn, hlp in (("branch-r2", cmd_branch_r2, "Use trained model to compute r2 values for each model branch."),
                          ("gradients", cmd_gradients, "Report gradient wrt to params in trained model."),
                          ("population-effect-sizes", cmd_population_effect_sizes, "Report population level effect sizes in trained model.")):
        p = sub.add_parser(name, help=hlp)
        p.add_argument("bfile"); p.add_argument("phen"); p.add_argument("groups"); p.add_argument("-m", "--model-path", default="./models")
        p.add_argument("--device", type=int, default=0)
        p.set_defaults(func=fn)
    p = sub.add_parser("activations", help="Report activations in trained model.")
    p.add_argument("bfile"); p.add_argument("groups"); p.add_argument("-m", "--model-path", default="./models")
    p.add_argument("--device", type=int, default=0)
    p.set_defaults(func=cmd_activations)
    p = sub.add_parser("simulate-xy", help="Simulate marker and phenotype data under a network model.")
    p.add_argument("-o", "--outdir", default="./")
    p.add_argument("model_type", type=_model_type); p.add_argument("activation_function", type=_activation)
    p.add_argument("num_markers_per_branch", type=int); p.add_argument("num_branches", type=int)
    p.add_argument("num_individuals", type=int); p.add_argument("hidden_layer_width", type=int)
    p.add_argument("branch_depth", type=int)
    p.add_argument("heritability", type=float, nargs="?", default=1.0)
    p.add_argument("--summary-layer-width", type=int); p.add_argument("--init-param-variance", type=float)
    p.add_argument("--json-data", action="store_true")
    p.add_argument("--seed", type=int, default=None); p.add_argument("--device", type=int, default=0)
    p.set_defaults(func=cmd_simulate_xy)
    return ap


def main(argv: Optional[List[str]] = None):
    a = build_parser().parse_args(argv)
    a.func(a)


if __name__ == "__main__":
    main()
