"""Sequential-exact schedule (Net::train order, bann_sweep): branch visits / s and branch-leapfrogs / s on one GPU."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import rs_bann_b200 as rb  # noqa: E402
from bench import default_params  # noqa: E402

n, B, per, L = 100000, 64, 50, 100
if len(sys.argv) > 1:
    n, B, per, L = map(int, sys.argv[1:5])
ctx = rb.Context(0)
gen = rb.Genotypes.random(ctx, n, B * per, None, seed=42, uniform_groups=(B, per))
wl = dict(B=B, per=per, widths=[5, 5, 1], model="ridge_ard")
net = rb.Net(ctx, gen, "ridge_ard", [[5, 5, 1]] * B)
pv, qv = default_params(wl)
net.set_all_params(pv, qv)
P = 50 * 5 + 5 * 5 + 5 + 5 + 5 if per == 50 else per * 5 + 40
w_out = pv.reshape(B, -1)[:, per * 5 + 25:per * 5 + 30]
net.set_globals(2.0, 0.05, float(np.sum(w_out ** 2)), B * 5)          # architectures.rs:209-235
net.set_targets(np.random.default_rng(1).normal(size=n).astype(np.float32))
net.init_residual()
import os
if os.environ.get('K1'):
    net.select_k1(int(os.environ['K1']))
if os.environ.get('HMC_PATH'):          # 0 auto (persistent kernel where eligible), 1 launch per step, 2 persistent or fail
    net.select_hmc_path(int(os.environ['HMC_PATH']))
cfg = rb.MCMCCfg(hmc_step_size_factor=0.1, hmc_integration_length=L, hmc_max_hamiltonian_error=1e30)
net.sweep(cfg, np.arange(B), seed=1)
ctx.sync()
rb.launch_count(reset=True)
t0 = time.perf_counter()
st = net.sweep(cfg, np.random.default_rng(2).permutation(B), seed=2)
ctx.sync()
dt = time.perf_counter() - t0
print(f"n={n} B={B} m_b={per} L={L}: {B / dt:.1f} visits/s, {B * L / dt:.0f} branch-leapfrogs/s, {dt / B / L * 1e6:.1f} us per leapfrog, "
      f"{rb.launch_count() / B:.0f} launches per visit, accepted {st['num_accepted']}/{st['num_samples']}, "
      f"transitions through the persistent kernel: {net.persistent_launches()}, last kernel: {net.last_k1_kernel()}")
