"""Host-side mirror of the reference's interface for the HMC/Gibbs hot path, above the C ABI.

Names follow the reference: `Net.train / predict / gradient` (net/net.rs:201,545,520),
`hmc_step`, `sample_error_precision + sample_param_precisions` (-> `gibbs_branch`),
`BranchCfg` param/precision vectors (net/params.rs:272-289,700-715), `MCMCCfg`
(net/mcmc_cfg.rs:181-204), `HMCStepResult` (net/branch/branch_sampler.rs:1310-1314).
All compute happens in libbann_b200.so on the GPU; this layer only marshals buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import (ACT_NAMES, HMC_ACCEPTED, HMC_REJECTED, HMC_REJECTED_EARLY, MODEL_NAMES, STEP_NAMES, BannError,
                   check, lib)

_fp = C.POINTER(C.c_float)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    return a.ctypes.data_as(_fp) if a is not None else None


@dataclass
class MCMCCfg:
    """net/mcmc_cfg.rs:181-230 (fields read by the hot path; same defaults)."""
    hmc_step_size_factor: float = 1.0
    hmc_max_hamiltonian_error: float = 10.0
    hmc_integration_length: int = 100
    hmc_step_size_mode: str = "izmailov"
    fixed_param_precisions: bool = False
    joint_hmc: bool = False                 # net/mcmc_cfg.rs:21-26: the flag-gated modes of Net::train (net.rs:268-290)
    gradient_descent: bool = False
    gradient_descent_joint: bool = False
    num_grad: bool = False                  # the reference's debugging aids (branch_sampler.rs:1232-1261)
    num_grad_traj: bool = False

    def c(self) -> _lib.McmcCfg:
        return _lib.McmcCfg(self.hmc_step_size_factor, self.hmc_max_hamiltonian_error, self.hmc_integration_length,
                            STEP_NAMES[self.hmc_step_size_mode], int(self.fixed_param_precisions), int(self.joint_hmc),
                            int(self.gradient_descent), int(self.gradient_descent_joint), int(self.num_grad),
                            int(self.num_grad_traj))


@dataclass
class HMCStepResult:
    status: int
    log_density: float
    neg_h_init: float
    neg_h_final: float
    steps_done: int
    u_turn_step: int
    y_pred: Optional[np.ndarray] = None
    trajectory: Optional[dict] = None

    @property
    def accepted(self):
        return self.status == HMC_ACCEPTED


def cuda_available() -> bool:
    return bool(lib.bann_cuda_available())


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib.bann_pinned_free(self.ptr)
        except Exception:
            pass


def pinned_empty(n: int, dtype=np.float32) -> np.ndarray:
    """Page-locked host array (bann_pinned_alloc): `Net.gradient` copies such buffers by DMA without staging."""
    dt = np.dtype(dtype)
    p = C.c_void_p()
    check(lib.bann_pinned_alloc(int(n) * dt.itemsize, C.byref(p)))
    owner = _PinnedOwner(p)
    buf = (C.c_char * (int(n) * dt.itemsize)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(n))
    _PINNED_KEEP[id(buf)] = (owner, buf)        # the allocation lives as long as the module (arrays may be viewed anywhere)
    return arr


_PINNED_KEEP = {}


class Context:
    """One per process / GPU. `stream`: a raw cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream)."""

    def __init__(self, device: int = 0, stream: Optional[int] = None, rank: int = 0, world: int = 1):
        h = C.c_void_p()
        if stream is not None and stream == 0:
            stream = 1          # torch's default stream is handle 0 == the legacy default stream: cudaStreamLegacy
        check(lib.bann_ctx_create(device, C.c_void_p(stream) if stream is not None else None, rank, world, C.byref(h)))
        self.h, self.rank, self.world, self.device = h, rank, world, device

    def sync(self):
        check(lib.bann_ctx_sync(self.h))

    # ---- peer-memory exchange of the sequential-exact schedule on sharded rows (include/bann.h, comm.cuh)
    def comm_handle(self) -> bytes:
        """Allocates this rank's inbox and returns its handle (to be all-gathered in rank order)."""
        buf = (C.c_uint8 * _lib.COMM_HANDLE_BYTES)()
        check(lib.bann_ctx_comm_handle(self.h, buf))
        return bytes(buf)

    def comm_connect(self, handles: Sequence[bytes]):
        """Maps every peer's inbox; `handles[r]` = rank r's `comm_handle()`."""
        assert len(handles) == self.world and all(len(h) == _lib.COMM_HANDLE_BYTES for h in handles)
        blob = b"".join(handles)
        check(lib.bann_ctx_comm_connect(self.h, (C.c_uint8 * len(blob)).from_buffer_copy(blob)))

    def comm_connected(self) -> bool:
        return bool(lib.bann_ctx_comm_connected(self.h))

    def close(self):
        if self.h:
            lib.bann_ctx_destroy(self.h)
            self.h = None


class _UniformGroups:
    """UniformGrouping (group/uniform.rs:11-24) without materialising the index lists."""

    def __init__(self, nb, per):
        self.nb, self.per = nb, per

    def __len__(self):
        return self.nb

    def __getitem__(self, b):
        return range(b * self.per, (b + 1) * self.per)


def stats_from_counts(counts: np.ndarray, n_total: int):
    """Column mean / population std from per-value counts (global over all row shards).
    mean = (n1 + 2 n2) / N is exact; the std sum is evaluated in f64 (the reference's sequential f32
    sum, io/bed.rs:231-238, is order dependent and cannot be reproduced from shard partials)."""
    c = counts.astype(np.float64)
    nf = float(n_total)
    mean = ((c[:, 1] + 2.0 * c[:, 2]).astype(np.float32) / np.float32(n_total)).astype(np.float32)  # f32 division as bed.rs:233
    m64 = mean.astype(np.float64)
    ss = c[:, 0] * (0 - m64) ** 2 + c[:, 1] * (1 - m64) ** 2 + c[:, 2] * (2 - m64) ** 2
    return mean, np.sqrt(ss / nf).astype(np.float32)


class Genotypes:
    """Device-resident packed genotype store = BedVM + MarkerGrouping (CompressedGenotypes,
    data/genotypes.rs:14-48).  `payload`: PLINK variant-major bytes without signature for the
    rows this rank holds; `groups`: list of marker-index lists (may overlap)."""

    def __init__(self, ctx: Context, payload, n: int, m: int, groups: Sequence[Sequence[int]],
                 col_means=None, col_stds=None, n_total: Optional[int] = None):
        self.ctx, self.n, self.m = ctx, int(n), int(m)
        self.groups = [list(map(int, g)) for g in groups]
        payload = np.ascontiguousarray(np.frombuffer(bytes(payload), dtype=np.uint8)
                                       if not isinstance(payload, np.ndarray) else payload, dtype=np.uint8)
        assert payload.size == m * ((n + 3) // 4), "payload size does not match n, m"
        offs = np.zeros(len(self.groups) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(g) for g in self.groups])
        ids = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.uint64) for g in self.groups]))
        mu = _f32(col_means) if col_means is not None else None
        sd = _f32(col_stds) if col_stds is not None else None
        h = C.c_void_p()
        check(lib.bann_genotypes_create(ctx.h, payload.ctypes.data_as(C.c_void_p), n, n_total or n, m, _ptr(mu), _ptr(sd),
                                        len(self.groups), offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        ids.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(h)))
        self.h = h

    @classmethod
    def random(cls, ctx: Context, n: int, m: int, groups, seed: int = 42, row_offset: int = 0,
               n_total: Optional[int] = None, maf_lo: float = 0.01, maf_hi: float = 0.5, uniform_groups=None):
        """Device-generated synthetic store in the spirit of BedVM::random (io/bed.rs:136-188).
        `uniform_groups=(B, per)` avoids materialising Python lists for very large groupings."""
        self = cls.__new__(cls)
        self.ctx, self.n, self.m = ctx, int(n), int(m)
        if uniform_groups is not None:
            nb, per = uniform_groups
            offs = (np.arange(nb + 1, dtype=np.uint64) * np.uint64(per))
            ids = np.arange(nb * per, dtype=np.uint64)
            self.groups = _UniformGroups(nb, per)
        else:
            self.groups = [list(map(int, g)) for g in groups]
            offs = np.zeros(len(self.groups) + 1, dtype=np.uint64)
            offs[1:] = np.cumsum([len(g) for g in self.groups])
            ids = np.ascontiguousarray(np.concatenate([np.asarray(g, dtype=np.uint64) for g in self.groups]))
        h = C.c_void_p()
        check(lib.bann_genotypes_random(ctx.h, n, row_offset, n_total or n, m, seed, maf_lo, maf_hi, len(self.groups),
                                        offs.ctypes.data_as(C.POINTER(C.c_uint64)),
                                        ids.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(h)))
        self.h = h
        return self

    @property
    def num_groups(self):
        return len(self.groups)

    def num_markers_per_group(self):
        return [len(g) for g in self.groups]

    def col_stats(self):
        mu = np.empty(self.m, dtype=np.float32)
        sd = np.empty(self.m, dtype=np.float32)
        check(lib.bann_genotypes_col_stats(self.h, _ptr(mu), _ptr(sd)))
        return mu, sd

    def col_counts(self):
        out = np.empty(3 * self.m, dtype=np.uint64)
        check(lib.bann_genotypes_col_counts(self.h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return out.reshape(self.m, 3)

    def set_col_stats(self, mu, sd):
        mu, sd = _f32(mu), _f32(sd)
        check(lib.bann_genotypes_set_col_stats(self.h, _ptr(mu), _ptr(sd)))

    def x_group(self, b: int, standardized: bool = True) -> np.ndarray:
        """GroupedGenotypes::x_group_af (data/genotypes.rs:44-48): [n, m_b] f32 (test hook)."""
        out = np.empty(self.n * len(self.groups[b]), dtype=np.float32)
        check(lib.bann_genotypes_decode_branch(self.h, b, int(standardized), _ptr(out)))
        return out.reshape((self.n, len(self.groups[b])), order="F")

    def x_group_tc(self, b: int, standardized: bool = True) -> np.ndarray:
        """The same matrix decoded from the tensor-core store (test hook)."""
        out = np.empty(self.n * len(self.groups[b]), dtype=np.float32)
        check(lib.bann_genotypes_decode_branch_tc(self.h, b, int(standardized), _ptr(out)))
        return out.reshape((self.n, len(self.groups[b])), order="F")

    def has_tc_store(self) -> bool:
        return bool(lib.bann_genotypes_has_tc_store(self.h))

    def has_byte_store(self) -> bool:
        return bool(lib.bann_genotypes_has_byte_store(self.h))

    def release_byte_store(self) -> None:
        """Give back the byte-tile copy of the genotypes (read only by the FFMA / shape-agnostic / probe kernels)."""
        check(lib.bann_genotypes_release_byte_store(self.h))

    def close(self):
        if self.h:
            lib.bann_genotypes_destroy(self.h)
            self.h = None


class Net:
    """Device-resident Net<B> (net/net.rs:74-85). Branch state stays on the GPU between visits."""

    def __init__(self, ctx: Context, gen: Genotypes, model_type: str, layer_widths: Sequence[Sequence[int]],
                 hyper=(0.001, 1000.0, 0.001, 1000.0, 0.001, 1000.0), activation: str = "tanh"):
        self.ctx, self.gen = ctx, gen
        self.model_type, self.activation = model_type, activation
        self.layer_widths = [list(map(int, w)) for w in layer_widths]
        assert len(self.layer_widths) == gen.num_groups
        lay = (_lib.BranchLayout * len(self.layer_widths))()
        for i, w in enumerate(self.layer_widths):
            lay[i].num_layers = len(w)
            for k, v in enumerate(w):
                lay[i].widths[k] = v
        hy = _f32(hyper)
        h = C.c_void_p()
        check(lib.bann_net_create(ctx.h, gen.h, MODEL_NAMES[model_type], ACT_NAMES[activation], lay, _ptr(hy),
                                  C.byref(h)))
        self.h = h
        self.num_branches = len(self.layer_widths)
        self._sizes = []
        for b in range(self.num_branches):
            p, q = C.c_uint64(), C.c_uint64()
            check(lib.bann_net_branch_sizes(h, b, C.byref(p), C.byref(q)))
            self._sizes.append((p.value, q.value))

    # ---- state
    def num_branch_params(self, b):
        return self._sizes[b][0]

    def num_branch_precisions(self, b):
        return self._sizes[b][1]

    def num_params(self):
        return sum(p for p, _ in self._sizes)

    def set_branch(self, b, param_vec=None, precision_vec=None):
        pv = _f32(param_vec) if param_vec is not None else None
        qv = _f32(precision_vec) if precision_vec is not None else None
        if pv is not None:
            assert pv.size == self._sizes[b][0]
        if qv is not None:
            assert qv.size == self._sizes[b][1]
        check(lib.bann_net_set_branch(self.h, b, _ptr(pv), _ptr(qv)))

    def get_branch(self, b):
        pv = np.empty(self._sizes[b][0], dtype=np.float32)
        qv = np.empty(self._sizes[b][1], dtype=np.float32)
        check(lib.bann_net_get_branch(self.h, b, _ptr(pv), _ptr(qv)))
        return pv, qv

    def set_all_params(self, param_vecs, precision_vecs=None):
        pv = _f32(param_vecs)
        assert pv.size == self.num_params()
        qv = _f32(precision_vecs) if precision_vecs is not None else None
        check(lib.bann_net_set_all_params(self.h, _ptr(pv), _ptr(qv)))

    def get_all_params(self):
        pv = np.empty(self.num_params(), dtype=np.float32)
        qv = np.empty(sum(q for _, q in self._sizes), dtype=np.float32)
        check(lib.bann_net_get_all_params(self.h, _ptr(pv), _ptr(qv)))
        return pv, qv

    def set_globals(self, error_precision, output_layer_precision, ow_reg_sum, ow_num_params, output_bias=0.0):
        g = _f32([error_precision, output_layer_precision, ow_reg_sum, ow_num_params, output_bias])
        check(lib.bann_net_set_globals(self.h, _ptr(g)))

    def get_globals(self):
        g = np.empty(5, dtype=np.float32)
        check(lib.bann_net_get_globals(self.h, _ptr(g)))
        return dict(error_precision=float(g[0]), output_layer_precision=float(g[1]), ow_reg_sum=float(g[2]),
                    ow_num_params=float(g[3]), output_bias=float(g[4]))

    def set_targets(self, y):
        y = _f32(y)
        assert y.size == self.gen.n
        check(lib.bann_net_set_targets(self.h, _ptr(y)))

    def residual(self):
        r = np.empty(self.gen.n, dtype=np.float32)
        check(lib.bann_net_get_residual(self.h, _ptr(r)))
        return r

    def set_residual(self, r):
        r = _f32(r)
        check(lib.bann_net_set_residual(self.h, _ptr(r)))

    def init_residual(self):
        """initialize_stats, net/net.rs:158-171."""
        check(lib.bann_net_init_residual(self.h))

    # ---- hot path, one branch
    def branch_fwd_bwd(self, b, target=None, want_yhat=True):
        """backpropagate + log_density_gradient (branch_sampler.rs:813-875,380-391)."""
        P = self._sizes[b][0]
        t = _f32(target) if target is not None else None
        rss = np.empty(1, dtype=np.float32)
        ldg = np.empty(P, dtype=np.float32)
        drss = np.empty(P, dtype=np.float32)
        yh = np.empty(self.gen.n, dtype=np.float32) if want_yhat else None
        check(lib.bann_branch_fwd_bwd(self.h, b, _ptr(t), _ptr(rss), _ptr(ldg), _ptr(drss), _ptr(yh)))
        return dict(rss=float(rss[0]), ldg=ldg, d_rss=drss, yhat=yh)

    def branch_log_density(self, b, rss):
        out = np.empty(1, dtype=np.float32)
        check(lib.bann_branch_log_density(self.h, b, float(rss), _ptr(out)))
        return float(out[0])

    def branch_numerical_ldg(self, b, target=None) -> np.ndarray:
        """numerical_ldg (branch_sampler.rs:480-504): forward differences of log_density, delta = 0.001."""
        t = _f32(target) if target is not None else None
        out = np.empty(self._sizes[b][0], dtype=np.float32)
        check(lib.bann_branch_numerical_ldg(self.h, b, _ptr(t), _ptr(out)))
        return out

    def branch_step_sizes(self, b, cfg: MCMCCfg, step_uniforms=None):
        out = np.empty(self._sizes[b][0], dtype=np.float32)
        su = _f32(step_uniforms) if step_uniforms is not None else None
        c = cfg.c()
        check(lib.bann_branch_step_sizes(self.h, b, C.byref(c), _ptr(su), _ptr(out)))
        return out

    @staticmethod
    def _inject(momenta=None, u=None, step_uniforms=None, std_gammas=None):
        keep = []
        inj = _lib.RngInject()
        if momenta is not None:
            a = _f32(momenta); keep.append(a); inj.momenta = _ptr(a)
        if u is not None:
            a = _f32([u]); keep.append(a); inj.accept_uniform = _ptr(a)
        if step_uniforms is not None:
            a = _f32(step_uniforms); keep.append(a); inj.step_uniforms = _ptr(a)
        if std_gammas is not None:
            a = _f32(std_gammas); keep.append(a); inj.std_gammas = _ptr(a); inj.num_std_gammas = a.size
        return inj, keep

    def hmc_step(self, b, cfg: MCMCCfg, target=None, momenta=None, u=None, step_uniforms=None, trajectory=False,
                 want_yhat=True) -> HMCStepResult:
        """BranchSampler::hmc_step (branch_sampler.rs:1192-1299)."""
        P, L = self._sizes[b][0], cfg.hmc_integration_length
        inj, keep = self._inject(momenta, u, step_uniforms)
        t = _f32(target) if target is not None else None
        res = _lib.HmcResult()
        traj = _lib.Trajectory()
        tp = tl = th = None
        if trajectory:
            tp = np.zeros(L * P, dtype=np.float32); tl = np.zeros(L * P, dtype=np.float32)
            th = np.zeros(L + 1, dtype=np.float32)
            traj.params, traj.ldg, traj.hamiltonian = _ptr(tp), _ptr(tl), _ptr(th)
        yh = np.empty(self.gen.n, dtype=np.float32) if want_yhat else None
        c = cfg.c()
        check(lib.bann_hmc_step(self.h, b, _ptr(t), C.byref(c), C.byref(inj), C.byref(res),
                                C.byref(traj) if trajectory else None, _ptr(yh)))
        out = HMCStepResult(res.status, res.log_density, res.neg_h_init, res.neg_h_final, res.steps_done,
                            res.u_turn_step, yh)
        if trajectory:
            out.trajectory = dict(params=tp.reshape(L, P), ldg=tl.reshape(L, P), hamiltonian=th)
        return out

    # ---- flag-gated sampler modes
    def branch_joint(self, b, target=None) -> dict:
        """log_density_gradient_joint + log_density_joint + log_density (branch_sampler.rs:406-422,292-305,72-78).
        `ldg`: P + Q values, parameters then precisions (gradient.rs:66-97)."""
        P, Q = self._sizes[b]
        t = _f32(target) if target is not None else None
        rss, ldj, ld = C.c_float(), C.c_float(), C.c_float()
        g = np.empty(P + Q, dtype=np.float32)
        check(lib.bann_branch_joint(self.h, b, _ptr(t), C.byref(rss), C.byref(ldj), C.byref(ld), _ptr(g)))
        return dict(rss=rss.value, log_density_joint=ldj.value, log_density=ld.value, ldg=g)

    def hmc_step_joint(self, b, cfg: MCMCCfg, target=None, momenta=None, u=None, step_uniforms=None, trajectory=False,
                       want_yhat=True) -> HMCStepResult:
        """BranchSampler::hmc_step_joint (branch_sampler.rs:1070-1178); momenta / step_uniforms: P + Q values."""
        (P, Q), L = self._sizes[b], cfg.hmc_integration_length
        inj, keep = self._inject(momenta, u, step_uniforms)
        t = _f32(target) if target is not None else None
        res = _lib.HmcResult()
        traj = _lib.TrajectoryJoint()
        tp = tq = tl = th = None
        if trajectory:
            tp = np.zeros(L * P, dtype=np.float32); tq = np.zeros(L * Q, dtype=np.float32)
            tl = np.zeros(L * (P + Q), dtype=np.float32); th = np.zeros(L + 1, dtype=np.float32)
            traj.params, traj.precisions, traj.ldg, traj.hamiltonian = _ptr(tp), _ptr(tq), _ptr(tl), _ptr(th)
        yh = np.empty(self.gen.n, dtype=np.float32) if want_yhat else None
        c = cfg.c()
        check(lib.bann_hmc_step_joint(self.h, b, _ptr(t), C.byref(c), C.byref(inj), C.byref(res),
                                      C.byref(traj) if trajectory else None, _ptr(yh)))
        out = HMCStepResult(res.status, res.log_density, res.neg_h_init, res.neg_h_final, res.steps_done,
                            res.u_turn_step, yh)
        if trajectory:
            out.trajectory = dict(params=tp.reshape(L, P), precisions=tq.reshape(L, Q), ldg=tl.reshape(L, P + Q),
                                  hamiltonian=th)
        return out

    def gradient_descent(self, b, cfg: MCMCCfg, target=None, want_yhat=True) -> HMCStepResult:
        """BranchSampler::gradient_descent (branch_sampler.rs:964-1017).  `trajectory['step_sizes']`: the step taken in
        every iteration, `trajectory['num_probes']`: probe evaluations."""
        L = cfg.hmc_integration_length
        t = _f32(target) if target is not None else None
        res = _lib.HmcResult()
        steps = np.zeros(max(L, 1), dtype=np.float32)
        nprobe = C.c_uint32()
        yh = np.empty(self.gen.n, dtype=np.float32) if want_yhat else None
        c = cfg.c()
        check(lib.bann_gradient_descent(self.h, b, _ptr(t), C.byref(c), C.byref(res), _ptr(steps), C.byref(nprobe), _ptr(yh)))
        return HMCStepResult(res.status, res.log_density, res.neg_h_init, res.neg_h_final, res.steps_done, res.u_turn_step,
                             yh, dict(step_sizes=steps[:L], num_probes=nprobe.value))

    def gradient_descent_joint(self, b, cfg: MCMCCfg, target=None, want_yhat=True) -> HMCStepResult:
        """BranchSampler::gradient_descent_joint (branch_sampler.rs:1019-1066)."""
        t = _f32(target) if target is not None else None
        res = _lib.HmcResult()
        yh = np.empty(self.gen.n, dtype=np.float32) if want_yhat else None
        c = cfg.c()
        check(lib.bann_gradient_descent_joint(self.h, b, _ptr(t), C.byref(c), C.byref(res), _ptr(yh)))
        return HMCStepResult(res.status, res.log_density, res.neg_h_init, res.neg_h_final, res.steps_done, res.u_turn_step, yh)

    def gibbs_branch(self, b, cfg: MCMCCfg, std_gammas=None):
        """sample_error_precision + sample_param_precisions (branch_sampler.rs:173-202)."""
        inj, keep = self._inject(std_gammas=std_gammas)
        c = cfg.c()
        check(lib.bann_gibbs_branch(self.h, b, C.byref(c), C.byref(inj) if std_gammas is not None else None))

    def visit_branch(self, b, cfg: MCMCCfg, momenta=None, u=None, step_uniforms=None, std_gammas=None) -> HMCStepResult:
        """One iteration of the inner loop of Net::train (net/net.rs:258-334)."""
        inj, keep = self._inject(momenta, u, step_uniforms, std_gammas)
        res = _lib.HmcResult()
        c = cfg.c()
        any_inj = any(v is not None for v in (momenta, u, step_uniforms, std_gammas))
        check(lib.bann_visit_branch(self.h, b, C.byref(c), C.byref(inj) if any_inj else None, C.byref(res)))
        return HMCStepResult(res.status, res.log_density, res.neg_h_init, res.neg_h_final, res.steps_done,
                             res.u_turn_step)

    def visit_branch_traj(self, b, cfg: MCMCCfg, seed: int = 0) -> HMCStepResult:
        """One visit with the built-in RNG and the trajectory of its HMC transition (trajectory.rs:4-43): `trajectory` holds
        only the steps that ran (params / ldg / hamiltonian, plus precisions in the joint mode)."""
        (P, Q), L = self._sizes[b], cfg.hmc_integration_length
        jt = cfg.joint_hmc and not (cfg.gradient_descent or cfg.gradient_descent_joint)
        Q = Q if jt else 0
        res = _lib.HmcResult()
        traj = _lib.TrajectoryJoint()
        tp = np.zeros(L * P, dtype=np.float32); tq = np.zeros(max(L * Q, 1), dtype=np.float32)
        tl = np.zeros(L * (P + Q), dtype=np.float32); th = np.zeros(L + 1, dtype=np.float32)
        traj.params, traj.precisions, traj.ldg, traj.hamiltonian = _ptr(tp), _ptr(tq), _ptr(tl), _ptr(th)
        tn = np.zeros(L * P, dtype=np.float32) if (cfg.num_grad_traj and not jt) else None
        if tn is not None:
            traj.num_ldg = _ptr(tn)
        c = cfg.c()
        check(lib.bann_visit_branch_traj(self.h, b, C.byref(c), seed, C.byref(res), C.byref(traj)))
        n = 0 if (cfg.gradient_descent or cfg.gradient_descent_joint) else min(res.steps_done, L)
        out = HMCStepResult(res.status, res.log_density, res.neg_h_init, res.neg_h_final, res.steps_done, res.u_turn_step)
        out.trajectory = dict(params=tp.reshape(L, P)[:n], precisions=tq[:L * Q].reshape(L, Q)[:n] if Q else np.zeros((0, 0), np.float32),
                              ldg=tl.reshape(L, P + Q)[:n], hamiltonian=th[:n + 1],
                              num_ldg=tn.reshape(L, P)[:n] if tn is not None else np.zeros((0, P), np.float32))
        return out

    def visit_group(self, members, cfg: MCMCCfg, injections=None, seed: int = 0) -> List[HMCStepResult]:
        """One block-Jacobi group visit (bann_visit_group): every member runs the inner loop of Net::train against the residual
        and globals frozen at group start.  injections: None or one dict(momenta=, u=, step_uniforms=, std_gammas=) per member."""
        members = np.ascontiguousarray(members, dtype=np.uint64)
        res = (_lib.HmcResult * members.size)()
        c = cfg.c()
        inj_arr, keep = None, []
        if injections is not None:
            assert len(injections) == members.size
            inj_arr = (_lib.RngInject * members.size)()
            for i, d in enumerate(injections):
                one, k = self._inject(d.get("momenta"), d.get("u"), d.get("step_uniforms"), d.get("std_gammas"))
                inj_arr[i] = one
                keep.append(k)
        check(lib.bann_visit_group(self.h, members.ctypes.data_as(C.POINTER(C.c_uint64)), members.size, C.byref(c), inj_arr, seed,
                                   res))
        return [HMCStepResult(r.status, r.log_density, r.neg_h_init, r.neg_h_final, r.steps_done, r.u_turn_step) for r in res]

    def sweep(self, cfg: MCMCCfg, order, seed: int = 0, group_size: int = 1):
        """One pass over `order`: group_size 1 = the reference's sequential order, G > 1 = block-Jacobi groups of G branches."""
        order = np.ascontiguousarray(order, dtype=np.uint64)
        st = _lib.SweepStats()
        c = cfg.c()
        check(lib.bann_sweep(self.h, C.byref(c), order.ctypes.data_as(C.POINTER(C.c_uint64)), order.size, int(group_size), seed,
                             C.byref(st)))
        return self._stats(st)

    @staticmethod
    def _stats(st):
        return dict(num_samples=st.num_samples, num_accepted=st.num_accepted, num_early_rejected=st.num_early_rejected,
                    mse_train=st.mse_train, lpd=st.lpd, output_bias=st.output_bias, error_precision=st.error_precision,
                    output_layer_precision=st.output_layer_precision)

    def stats(self):
        st = _lib.SweepStats()
        check(lib.bann_net_stats(self.h, C.byref(st)))
        return self._stats(st)

    def lpd_terms(self):
        """LogPosteriorDensity fields (net/log_posterior_density.rs:9-16): (rss term, output-weight term, per-branch terms)."""
        a, b = np.empty(1, dtype=np.float32), np.empty(1, dtype=np.float32)
        loc = np.empty(self.num_branches, dtype=np.float32)
        check(lib.bann_net_lpd_terms(self.h, _ptr(a), _ptr(b), _ptr(loc)))
        return float(a[0]), float(b[0]), loc

    def train(self, cfg: MCMCCfg, chain_length: int, seed: int = 0, orders=None, group_size: int = 1):
        """Net::train (net/net.rs:201-358) with the built-in RNG; group_size 1 = sequential-exact schedule."""
        self.init_residual()
        hist = [self.stats()]
        rng = np.random.default_rng(seed)
        for it in range(chain_length):
            order = orders[it] if orders is not None else rng.permutation(self.num_branches)
            hist.append(self.sweep(cfg, order, seed=seed + 1 + it, group_size=group_size))
        return hist

    def predict(self, test: Optional[Genotypes] = None) -> np.ndarray:
        """Net::predict (net/net.rs:545-559)."""
        g = test or self.gen
        out = np.empty(g.n, dtype=np.float32)
        check(lib.bann_predict(self.h, test.h if test is not None else None, _ptr(out)))
        return out

    # ---- per-row diagnostics of saved models
    def branch_activations(self, b, genotypes: Optional[Genotypes] = None) -> List[np.ndarray]:
        """forward_feed of branch b as Net::activations collects it (net/net.rs:509-518): [a_0 .. a_{last-1}, yhat],
        arrays of shape [n, w_l] (the reference's column-major Array<f32>)."""
        g = genotypes or self.gen
        widths = self.layer_widths[b][:-1] + [1]
        out = np.empty(g.n * sum(widths), dtype=np.float32)
        check(lib.bann_branch_activations(self.h, b, genotypes.h if genotypes is not None else None, _ptr(out)))
        res, off = [], 0
        for w in widths:
            res.append(out[off:off + g.n * w].reshape((g.n, w), order="F"))
            off += g.n * w
        return res

    def branch_effect_sizes(self, b, genotypes: Optional[Genotypes] = None, per_row=True, population=True):
        """BranchSampler::effect_sizes [n, m_b] and its column means (Net::population_effect_sizes, net/net.rs:529-543)."""
        g = genotypes or self.gen
        m = len(self.gen.groups[b])
        es = np.empty(g.n * m, dtype=np.float32) if per_row else None
        pop = np.empty(m, dtype=np.float32) if population else None
        check(lib.bann_branch_effect_sizes(self.h, b, genotypes.h if genotypes is not None else None, _ptr(es), _ptr(pop)))
        return (es.reshape((g.n, m), order="F") if per_row else None), pop

    def population_effect_sizes(self, genotypes: Optional[Genotypes] = None) -> np.ndarray:
        """Net::population_effect_sizes (net/net.rs:529-543): all branches, concatenated in branch order."""
        return np.concatenate([self.branch_effect_sizes(b, genotypes, per_row=False)[1] for b in range(self.num_branches)])

    # ---- full network (grouped) operations
    def gradient(self, param_vecs=None, y=None, allreduce=None, out=None):
        """Net::gradient (net/net.rs:520-527) through HOST buffers: returns (grads, rss per branch).
        `allreduce`: callable run between the two halves when rows are sharded over ranks and the caller brings its own
        collective; with the bulk exchange connected (dist.connect_net) leave it None: the library sums over ranks itself and
        every rank reads / writes only its `gradient_slice()` of the host vectors."""
        pv = _f32(param_vecs) if param_vecs is not None else None
        yy = _f32(y) if y is not None else None
        grads, rss = out if out is not None else (np.empty(self.num_params(), dtype=np.float32),
                                                  np.empty(self.num_branches, dtype=np.float32))
        if allreduce is None:     # one rank, or sharded rows with the bulk exchange connected (fails loudly otherwise)
            check(lib.bann_net_gradient(self.h, _ptr(pv), _ptr(yy), _ptr(grads), _ptr(rss)))
        else:                     # sharded rows, the caller's own collective between the two halves
            check(lib.bann_net_gradient_begin(self.h, _ptr(pv), _ptr(yy)))
            allreduce()
            check(lib.bann_net_gradient_end(self.h, _ptr(grads), _ptr(rss)))
        return grads, rss

    def gradient_slice(self):
        """(param_lo, param_hi, out_lo, out_hi): the element ranges of `param_vecs` / of [grads | rss] this rank reads / writes
        in `gradient` (everything on one rank; 1 / world each on sharded rows with the bulk exchange connected)."""
        v = [C.c_uint64() for _ in range(4)]
        check(lib.bann_net_gradient_slice(self.h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)

    # ---- bulk peer-memory exchange of the grouped schedule on sharded rows (include/bann.h, comm.cuh: XgComm)
    def comm_handle(self) -> bytes:
        buf = (C.c_uint8 * _lib.COMM_HANDLE_BYTES)()
        check(lib.bann_net_comm_handle(self.h, buf))
        return bytes(buf)

    def comm_connect(self, handles: Sequence[bytes]):
        assert all(len(h) == _lib.COMM_HANDLE_BYTES for h in handles)
        blob = b"".join(handles)
        check(lib.bann_net_comm_connect(self.h, (C.c_uint8 * len(blob)).from_buffer_copy(blob)))

    def comm_connected(self) -> bool:
        return bool(lib.bann_net_comm_connected(self.h))

    def grouped_begin(self, cfg: MCMCCfg, seed: int = 0, per_branch_targets: bool = False):
        c = cfg.c()
        check(lib.bann_grouped_begin(self.h, C.byref(c), seed, int(per_branch_targets)))

    def grouped_leapfrog(self, cfg: MCMCCfg, num_steps: int, finalize: bool = False):
        c = cfg.c()
        check(lib.bann_grouped_leapfrog(self.h, C.byref(c), num_steps, int(finalize)))

    def grouped_phase_a(self):
        check(lib.bann_grouped_phase_a(self.h))

    def grouped_phase_b(self, cfg: MCMCCfg, is_init=False, is_last=False):
        c = cfg.c()
        check(lib.bann_grouped_phase_b(self.h, C.byref(c), int(is_init), int(is_last)))

    def grouped_allreduce(self):
        """Sum of the step's [gW | gb | rss] over ranks inside the library (NVLink peer memory); no-op on one rank."""
        check(lib.bann_grouped_allreduce(self.h))

    def last_k1_kernel(self) -> str:
        return lib.bann_net_last_k1_kernel(self.h).decode()

    def grouped_finish(self, seed: int = 0):
        a, e = C.c_uint64(), C.c_uint64()
        check(lib.bann_grouped_finish(self.h, seed, C.byref(a), C.byref(e)))
        return a.value, e.value

    def grouped_state(self):
        B = self.num_branches
        hi = np.empty(B, dtype=np.float32); hc = np.empty(B, dtype=np.float32); st = np.empty(B, dtype=np.int32)
        check(lib.bann_grouped_state(self.h, _ptr(hi), _ptr(hc), st.ctypes.data_as(C.POINTER(C.c_int32))))
        return hi, hc, st

    def allreduce_buffer(self):
        p, n = C.c_void_p(), C.c_uint64()
        check(lib.bann_allreduce_buffer(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def force_generic(self, on=True):
        check(lib.bann_net_force_generic(self.h, int(on)))

    K1_AUTO, K1_TENSOR, K1_FFMA, K1_GENERIC = 0, 1, 2, 3

    HMC_AUTO, HMC_LAUNCHES, HMC_PERSISTENT = 0, 1, 2

    def select_hmc_path(self, which: int):
        """Per-branch transitions: persistent cooperative kernel where eligible (AUTO) / launch per step / persistent or fail."""
        check(lib.bann_net_select_hmc_path(self.h, which))

    def persistent_launches(self) -> int:
        return int(lib.bann_net_persistent_launches(self.h))

    def select_k1(self, which: int):
        """Which fused forward+backward kernel may run: auto / tensor-core / FFMA / shape-agnostic."""
        check(lib.bann_net_select_k1(self.h, int(which)))

    TC_FOUR_WARPS, TC_FIVE_WARPS, TC_FIVE_WARPS_PLAIN = 0, 1, 2

    def select_k1_tc_variant(self, which: int):
        """Variant of the <= 64-marker tensor-core kernel: k1_tc (compute warps issue the MMAs) / k1_tc5 (dedicated issuing warp,
        cross-row sums deferred under the next super-tile's MUFU phase; the default) / k1_tc5 without the deferral."""
        check(lib.bann_net_select_k1_tc_variant(self.h, int(which)))

    def algorithmic_bytes(self) -> int:
        v = C.c_uint64()
        check(lib.bann_net_algorithmic_bytes(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            lib.bann_net_destroy(self.h)
            self.h = None
