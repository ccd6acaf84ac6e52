#!/usr/bin/env python
"""bench.py -- full-network HMC leapfrog steps / second (BASELINE.json metric).

One "step" = one full-network leapfrog step = B branch-leapfrogs (momentum half step, position
step, fused forward+backward of every branch over all N rows against its own target vector, prior
gradient, second half step, Hamiltonian), schedule G = B (all branches advance concurrently,
SURVEY H1).  Rows are sharded over ranks; the per-step [gW | gb | rss] sums are all-reduced INSIDE the library over
NVLink peer memory (bann_net_comm_connect; torch.distributed only carries the 128-byte handles and the timing max).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl ours|reference]

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# BASELINE.json configs.  widths = hidden..., summary, 1
WORKLOADS = {
    "cfg1": dict(desc="configs[0] tiny: 1k individuals x 10 branches x 100 markers, widths [2,2,1], StdNormal",
                 n=1000, B=10, per=100, widths=[2, 2, 1], model="std_normal"),
    "cfg2": dict(desc="configs[1]: 10k individuals x 1000 branches x 500 markers, widths [5,5,1], RidgeARD",
                 n=10000, B=1000, per=500, widths=[5, 5, 1], model="ridge_ard"),
    "cfg3": dict(desc="configs[2] biobank: 100k individuals x 10000 branches x 50 markers (500k markers), "
                      "widths [5,5,1], RidgeARD", n=100000, B=10000, per=50, widths=[5, 5, 1], model="ridge_ard"),
    "cfg4": dict(desc="configs[3] wide branches: 50k individuals x 2000 branches x 1000 markers, widths [16,16,16,1], RidgeARD",
                 n=50000, B=2000, per=1000, widths=[16, 16, 16, 1], model="ridge_ard"),
    "cfg4s": dict(desc="configs[3] at 1/20 of the branches (smoke size)", n=50000, B=100, per=1000, widths=[16, 16, 16, 1],
                  model="ridge_ard"),
    "cfg3r8": dict(desc="configs[2] with 1/8 of the rows: the per-GPU shard of the 8-GPU run on one GPU (prologue / epilogue share of K1)",
                   n=12544, B=10000, per=50, widths=[5, 5, 1], model="ridge_ard"),
    "cfg3s": dict(desc="configs[2] at 1/10 of the branches (smoke size)", n=100000, B=1000, per=50, widths=[5, 5, 1],
                  model="ridge_ard"),
    "cfg3s_321": dict(desc="cfg3s with widths [3,2,1]: no kernel is instantiated for it, runs zero-padded on the [5,5,1] kernel",
                      n=100000, B=1000, per=50, widths=[3, 2, 1], model="ridge_ard"),
}
METRIC = "full_network_hmc_leapfrog_steps_per_sec"
UNIT = "steps/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` captures
# (NOT measured in this run -- a bench run is never profiled); None where no capture exists for the workload
NCU_TRAFFIC = {
    "cfg3s": dict(bytes=1.8090e9, source="profiles/r2_k1_tc5_ncu_summary.md (ncu, not this run)"),
    "cfg3": dict(bytes=1.8090e10, source="profiles/r2_k1_tc5_ncu_summary.md: cfg3s capture x 10 branches (ncu, not this run)"),
}


def cpu_threads():
    """Host threads the CPU arm may use: the cores this process is allowed on (torchrun's OMP_NUM_THREADS=1 is overridden
    explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def default_params(wl, seed=42):
    """Reference default init (branch_cfg_builder.rs:180-186,308-328): W ~ N(0, 1/m_b), b = 0, ML ARD
    precisions out_l / sum_c W[r,c]^2, output precision placeholder, error precision 2.0.  Bias
    precisions: the ML value for zero biases is +inf (step size 0, Q7); a chain resamples them at
    the first Gibbs visit, the bench uses 1.0 so that every parameter moves."""
    rng = np.random.default_rng(seed)
    B, m, widths, model = wl["B"], wl["per"], wl["widths"], wl["model"]
    ins = [m] + widths[:-1]
    ard = model.endswith("ard")
    Ws = [rng.normal(0.0, np.sqrt(1.0 / m), size=(B, o, i)).astype(np.float32) for i, o in zip(ins, widths)]  # [B][col][row]
    nb = sum(widths[:-1])
    pv = np.concatenate([w.reshape(B, -1) for w in Ws] + [np.zeros((B, nb), dtype=np.float32)], axis=1)
    precs = []
    for l, w in enumerate(Ws):
        if ard and l < len(widths) - 1:
            precs.append((np.float32(widths[l]) / np.sum(w * w, axis=1)).astype(np.float32))       # per input row
        elif model == "std_normal" or l == len(widths) - 1:
            precs.append(np.ones((B, 1), dtype=np.float32))
        else:
            precs.append((np.float32(w[0].size) / np.sum(w * w, axis=(1, 2)))[:, None].astype(np.float32))
    precs.append(np.ones((B, len(widths) - 1), dtype=np.float32))          # bias precisions
    precs.append(np.full((B, 1), 2.0, dtype=np.float32))                  # error precision
    qv = np.concatenate(precs, axis=1)
    return np.ascontiguousarray(pv.reshape(-1)), np.ascontiguousarray(qv.reshape(-1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.stop_flag, self.idx = [], False, gpu_index
        self.nv = self.nv_handle = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.nv_handle = nv, nv.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run_nvml(self):
        """Fast path: NVML (nvidia_ml_py) polled every 10 ms -- a 0.3 s timed region still gets ~30 samples."""
        nv, h = self.nv, self.nv_handle          # initialised in the constructor: the first sample lands inside the timed region
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))
        while not self.stop_flag:
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            self.rows.append([str(self.idx), str(sm), str(mx), "0"] + ["Active" if r & b else "Not Active" for b, _ in bits])
            time.sleep(0.01)

    def run(self):
        if self.nv is not None:
            try:
                self.run_nvml()
                return
            except Exception:
                pass                  # NVML query failed: fall back to polling nvidia-smi
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.idx)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t.start()

    def stop(self):
        self.stop_flag = True
        self.t.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 8 for k in range(4) if r[4 + k].lower() == "active"})
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


def cpu_sample_branches(wl):
    """Branches in the fixed CPU sample: ~6.4e8 row x marker products per leapfrog evaluation (128 branches at config 3)."""
    return int(min(256, max(4, round(6.4e8 / (wl["n"] * wl["per"])))))


def cpu_reference_rate(wl, leapfrogs=12):
    """Times the oracle's C restatement of the reference's op sequence (host decode + dense f32 + 2 forwards and 1 backward
    per leapfrog) on a FIXED sample: all N rows, cpu_sample_branches(wl) branches, `leapfrogs` leapfrog steps each, every
    allowed host thread (pinned explicitly).  The same sample serves `cpu_baseline` and `--impl reference`; the full-network
    rate is the sample's branch-leapfrog rate divided by B (extrapolation factor B / sample branches, stated in the line)."""
    from oracle import bed as obed
    from oracle.cport import CPort
    cp = CPort(threads=cpu_threads())
    n, m, widths, B = wl["n"], wl["per"], wl["widths"], wl["B"]
    K = cpu_sample_branches(wl)
    rng = np.random.default_rng(1)
    ncols = m * 4
    g = rng.binomial(2, rng.uniform(0.01, 0.5, size=ncols)[None, :], size=(n, ncols)).astype(np.uint8)
    payload = obed.pack_columns(g)
    gm = g.astype(np.float32)
    mu = gm.mean(axis=0).astype(np.float32)
    sd = np.maximum(gm.std(axis=0), 1e-3).astype(np.float32)
    y = rng.normal(size=n).astype(np.float32)
    ins = [m] + widths[:-1]
    P = sum(i * o for i, o in zip(ins, widths)) + sum(widths[:-1])
    L_ref = 100.0                       # decode happens once per visit of L = 100 leapfrogs (mcmc_cfg.rs:38)
    t_decode, t_leap, t_wall = 0.0, 0.0, 0.0
    for nbr in range(K + 1):            # branch 0 = warm-up (thread pool, page faults), not timed
        cols = np.arange((nbr % 4) * m, (nbr % 4 + 1) * m)
        theta = rng.normal(0, np.sqrt(1.0 / m), size=P).astype(np.float32)
        mom = rng.normal(size=P).astype(np.float32)
        eps = np.full(P, 1e-3, dtype=np.float32)
        lam = np.ones(P, dtype=np.float32)
        t0 = time.perf_counter()
        X = cp.decode_std(payload, n, cols, mu, sd)
        t1 = time.perf_counter()
        cp.leapfrog(X, y, n, m, widths, "tanh", False, wl["model"] == "std_normal", theta, mom, eps, lam, 2.0, leapfrogs)
        t2 = time.perf_counter()
        if nbr > 0:
            t_wall += t2 - t0
            t_decode += t1 - t0
            t_leap += (t2 - t1) / (leapfrogs + 0.5)   # + initial gradient evaluation (half a leapfrog)
    per_branch_leapfrog = t_leap / K + (t_decode / K) / L_ref
    steps_per_s = 1.0 / (per_branch_leapfrog * B)
    sample = (f"fixed sample: {K} branches x {leapfrogs} leapfrogs over all {n} rows (m_b={m}, widths {widths}), "
              f"{cp.threads} OpenMP threads (nproc {os.cpu_count()}); host decode per visit amortised over L=100; "
              f"extrapolated x{B / K:.1f} to B={B} branches")
    return dict(value=steps_per_s, unit=UNIT, cores=cp.threads, kind="port", sample=sample,
                sample_branches=K, extrapolation_factor=B / K, sample_seconds=t_wall), per_branch_leapfrog


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    per = []
    base = None
    for s in range(args.warmup + steps):
        base, t = cpu_reference_rate(wl)
        if s >= args.warmup:
            per.append(t)
    t_bl = statistics.mean(per)
    value = 1.0 / (t_bl * wl["B"])
    base["value"] = value
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=args.warmup,
                ms_per_step=t_bl * wl["B"] * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None, dtype="f32",
                data="synthetic", impl="reference",
                config=dict(workload=f"{args.workload}: {wl['desc']}", schedule="sequential branch visits (reference order)",
                            note="CPU restatement of the reference's op sequence (oracle C port), not the reference binary: "
                                 "rs-bann needs cargo + ArrayFire, neither is in this image"),
                cpu_baseline=base, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sequential", action="store_true", help="skip the informational timing of the sequential schedule")
    ap.add_argument("--generic", action="store_true", help="force the shape-agnostic kernel (debug)")
    ap.add_argument("--k1", default="auto", choices=["auto", "tensor", "ffma", "generic"],
                    help="which fused forward+backward kernel may run (auto: tensor-core where eligible)")
    ap.add_argument("--k1-tc-variant", default="default", choices=["default", "four", "five", "five-plain"],
                    help="<= 64-marker tensor-core kernel: k1_tc (compute warps issue the MMAs) or k1_tc5 (dedicated issuing warp)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist

    import rs_bann_b200 as rb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL writes its version / debug lines to the process's stdout: route fd 1 to stderr until the JSON line is printed
        sys.stdout.flush()
        _saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = rb.Context(local, stream=stream, rank=rank, world=world)

    N, B, per, widths = wl["n"], wl["B"], wl["per"], wl["widths"]
    M = B * per
    r0, r1 = rb.row_shard(N, rank, world)          # row shards on 128-row tile boundaries
    n_local = r1 - r0
    gen = rb.Genotypes.random(ctx, n_local, M, None, seed=42, row_offset=r0, n_total=N, uniform_groups=(B, per))
    def allreduce_counts(c):
        t = torch.from_numpy(c).to(dev)
        dist.all_reduce(t)
        return t.cpu().numpy()

    mu, sd = rb.global_col_stats(gen.col_counts(), N, allreduce_counts if world > 1 else None)
    gen.set_col_stats(mu, sd)
    net = rb.Net(ctx, gen, wl["model"], [widths] * B)
    pv, qv = default_params(wl)
    net.set_all_params(pv, qv)
    y = np.random.default_rng(42).normal(size=N).astype(np.float32)
    y_local = np.ascontiguousarray(y[r0:r1])
    net.set_targets(y_local)
    if args.generic:
        net.force_generic(True)
    elif args.k1 != "auto":
        net.select_k1(dict(tensor=net.K1_TENSOR, ffma=net.K1_FFMA, generic=net.K1_GENERIC)[args.k1])
    if args.k1_tc_variant != "default":
        net.select_k1_tc_variant({"four": net.TC_FOUR_WARPS, "five": net.TC_FIVE_WARPS,
                                  "five-plain": net.TC_FIVE_WARPS_PLAIN}[args.k1_tc_variant])
    # cross-rank sums: INSIDE the library over NVLink peer memory (bann_net_comm_connect: reduce-scatter + all-gather kernels
    # on the same stream, csrc/comm.cuh).  torch.distributed only carries the 128-byte region handles once.
    rb.connect_net(net)

    def allreduce():
        if world > 1:
            net.grouped_allreduce()

    cfg = rb.MCMCCfg(hmc_step_size_factor=0.1, hmc_integration_length=100, hmc_max_hamiltonian_error=1e30)
    net.grouped_begin(cfg, seed=42, per_branch_targets=True)      # t_b = r + yhat_b (net.rs:279-280), momenta, H_init (+ exchange)

    def step(ev=None):
        if ev:
            ev[0].record()
        net.grouped_phase_a()          # K1: fused fwd+bwd of every branch (+ chunk reduction)
        if ev:
            ev[1].record()
        allreduce()
        net.grouped_phase_b(cfg)       # K2: gradient, half steps, position step, Hamiltonian
        if ev:
            ev[2].record()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # the tensor-core kernels read their own copy of the genotypes: give the byte-tile copy back (13 GB of 27 at cfg3)
    byte_store_released = False
    if args.warmup > 0 and gen.has_tc_store() and net.last_k1_kernel().startswith(("k1_tc", "k1_tcw", "k1_tcx")):
        gen.release_byte_store()
        byte_store_released = True
    if world > 1:
        dist.barrier()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    rb.launch_count(reset=True)
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        step(evs[s])
    end = torch.cuda.Event(enable_timing=True)
    end.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    launches = rb.launch_count()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    total_ms = evs[0][0].elapsed_time(end)
    k1_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    tm = torch.tensor([total_ms, k1_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms, k1_ms = float(tm[0]), float(tm[1])
    hi, hc, st = net.grouped_state()
    active = int(np.sum(st == 3))
    assert active == B, f"{B - active} branches stopped during the timed region: number invalid"
    assert np.all(np.isfinite(hc)), "non-finite Hamiltonian"
    alg_bytes = net.algorithmic_bytes()                  # this rank's rows
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
    value = args.steps / (total_ms * 1e-3)

    # ---- end to end through the host-facing call (Net::gradient = full-network fwd+grad), HOST buffers
    # page-locked HOST buffers (bann_pinned_alloc): every step copies the inputs from them and the results into them
    pv_h = rb.pinned_empty(pv.size); pv_h[:] = pv
    y_h = rb.pinned_empty(y_local.size); y_h[:] = y_local
    grads = rb.pinned_empty(net.num_params())
    rss = rb.pinned_empty(B)
    pv, y_local = pv_h, y_h
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        net.gradient(pv, y_local, out=(grads, rss))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        net.gradient(pv, y_local, out=(grads, rss))
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = e2e_steps / float(te[0])
    # bytes every step moves between HOST and device, summed over ranks: each rank reads its 1 / world slice of the
    # (replicated) parameters plus its own rows of y, and writes its slice of [grads | rss] (bann_net_gradient_slice)
    plo, phi, olo, ohi = net.gradient_slice()
    hb = torch.tensor([4 * (phi - plo) + 4 * n_local, 4 * (ohi - olo)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(hb)
    h2d, d2h = int(hb[0]), int(hb[1])
    k1_name = net.last_k1_kernel()
    traffic = NCU_TRAFFIC.get(args.workload) if (world == 1 and k1_name.startswith(("k1_tc<", "k1_tc5<"))) else None

    # ---- informational: the reference's own schedule (Net::train, net.rs:258-334: one branch at a time, Gibbs draws + one HMC
    #      transition of L = 100 leapfrog steps per visit) on the same net, 64 visits after 64 warm-up visits
    seq = None
    if not args.no_sequential:
        if world > 1:
            rb.connect_ranks(ctx)
        pv0, qv0 = default_params(wl)
        net.set_all_params(pv0, qv0)
        ins = [per] + widths[:-1]
        nw = sum(i * o for i, o in zip(ins, widths))
        w_out = pv0.reshape(B, -1)[:, nw - widths[-2]:nw]
        net.set_globals(2.0, 0.05, float(np.sum(w_out.astype(np.float64) ** 2)), B * widths[-2])   # architectures.rs:209-235
        net.set_targets(y_local)
        net.init_residual()
        nv = min(B, 64)
        scfg = rb.MCMCCfg(hmc_step_size_factor=0.1, hmc_integration_length=100, hmc_max_hamiltonian_error=1e30)
        net.sweep(scfg, np.arange(nv), seed=1)
        ctx.sync()
        if world > 1:
            dist.barrier()
        before = net.persistent_launches()
        t0 = time.perf_counter()
        sst = net.sweep(scfg, np.random.default_rng(2).permutation(nv), seed=2)
        ctx.sync()
        ts = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        ts = float(ts[0])
        seq = dict(visits_per_s=nv / ts, us_per_leapfrog=ts / nv / 100 * 1e6, branch_leapfrogs_per_s=nv * 100 / ts, visits=nv,
                   integration_length=100, transitions_through_persistent_kernel=int(net.persistent_launches() - before),
                   accepted=int(sst["num_accepted"]), kernel=net.last_k1_kernel(),
                   note="sequential-exact schedule (bann_sweep, group size 1), wall clock around the sweep, not part of `value`")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="strong", vs_baseline=None,
                    dtype="f32", data="synthetic",
                    config=dict(workload=f"{args.workload}: {wl['desc']}", schedule="grouped G=B (all branches per launch)",
                                individuals=N, branches=B, markers_per_branch=per, widths=widths, prior=wl["model"],
                                rows_per_gpu=n_local,
                                parallelism=(f"row-sharded x{world}, all-reduce of [gW|gb|rss] inside the library over NVLink peer "
                                             f"memory (reduce-scatter + all-gather kernels, no NCCL on the data path)"),
                                l2="working set (packed genotypes + per-branch targets) >> 126 MB L2, no flush needed",
                                init="reference default init, seed 42; bias precisions 1.0",
                                branch_leapfrogs_per_step=B, active_branches=active, byte_tile_store_released=byte_store_released),
                    k1_ms=k1_ms, wall_s=t_wall, gpu_launches=int(launches) * world, sequential_schedule=seq,
                    roofline=dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                                  traffic=traffic["bytes"] if traffic else None,
                                  traffic_source=traffic["source"] if traffic else None,
                                  peak_source=peak_src,
                                  algorithmic_bytes_per_launch=alg_bytes, kernel=k1_name,
                                  kernel_ms=k1_ms),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                             call="Net.gradient / bann_net_gradient (host params + targets in, gradients + rss out; on sharded rows "
                                  "every rank moves its 1 / world slice, the ranks all-gather / all-reduce on the device)"),
                    clocks=clocks)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_reference_rate(wl)
            except Exception as ex:   # the oracle is test infrastructure; never let it break the GPU number
                line["cpu_baseline"] = dict(error=str(ex))
        sys.stdout.flush()
        if world > 1:
            os.dup2(_saved_stdout, 1)
        print(json.dumps(line), flush=True)
    net.close()
    gen.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
