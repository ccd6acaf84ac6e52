"""CPU oracle for the rs-bann HMC/Gibbs hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in NumPy (and a small C port used
only for CPU timing), the arithmetic of the reference's branch sampler so that the
CUDA path can be checked against it.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.  Nothing under
`rs-bann_b200/` imports or links anything from here.

Why a restatement and not the reference itself: the reference is a Rust crate whose array
math lives in the third-party crate `arrayfire` 3.8.0 (Cargo.lock:75-76; Cargo.toml:18)
over the system ArrayFire 3.8 C library.  Neither `cargo`/`rustc` nor ArrayFire exist in
this image and there is no network, so `oracle/_ref` cannot be built ("unbuildable").

Pinning status (see tests/test_oracle_golden.py):
  * pinned by the reference's own golden vectors: .bed decode, column means / stds,
    standardised sub-matrix (io/bed.rs:431-497), byte packing (bed.rs:413), forward
    activations, non-joint gradients, joint densities / gradients for the four
    ridge/lasso priors, lasso_base::log_density, param counts, param_vec order.
  * PARITY UNPINNED by any reference test (the oracle is the only pin): leapfrog
    trajectories, Hamiltonians, accept/reject, Gibbs draws, Izmailov / Random /
    StdScaled step sizes, Net::train bookkeeping, Net::predict, LPD, StdNormal.

Every function cites the reference file:line it follows (paths relative to the
reference repository root).
"""
