"""On-disk formats of the kept rs-bann surface: PLINK .bed (+ .dims / .bim / .fam),
grouping TSV, bincode phenotypes and model files, JSON side files.

bincode 1.3 with its default options, which is what `serialize_into` / `deserialize_from` use in the reference
(net/net.rs:107-115, data/phenotypes.rs:28-36): little endian, fixed-width integers, `usize` as u64, `Vec<T>` as a
u64 length followed by the elements, `Option<T>` as a u8 tag, unit enum variants as a u32 index, structs as their
fields in declaration order, `PhantomData` as nothing.
"""
import json
import math
import os
import struct
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np


def _finite(o):
    """serde_json writes `null` for non-finite f32 (a fresh net has +inf bias precisions and -inf / NaN LPD terms);
    Python's json would emit the invalid tokens Infinity / NaN."""
    if isinstance(o, float):
        return o if math.isfinite(o) else None
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    return o


def json_dumps(obj, **kw):
    return json.dumps(_finite(obj), allow_nan=False, **kw)


def json_dump(obj, fp, **kw):
    fp.write(json_dumps(obj, **kw))


BED_SIGNATURE = bytes([0x6C, 0x1B, 0x01])            # io/bed.rs:193-213 (variant-major only)
ACTIVATIONS = ["tanh", "relu", "leaky_relu", "silu", "identity"]       # activation_functions.rs:6-12 (variant order)
ACTIVATION_JSON = ["Tanh", "ReLU", "LeakyReLU", "SiLU", "Identity"]


# ------------------------------------------------------------------ bincode primitives
class _Reader:
    def __init__(self, data: bytes):
        self.d, self.o = data, 0

    def take(self, n):
        if self.o + n > len(self.d):
            raise ValueError("bincode: unexpected end of file")
        b = self.d[self.o:self.o + n]
        self.o += n
        return b

    def u8(self):
        return self.take(1)[0]

    def u32(self):
        return struct.unpack("<I", self.take(4))[0]

    def u64(self):
        return struct.unpack("<Q", self.take(8))[0]

    def f32(self):
        return struct.unpack("<f", self.take(4))[0]

    def vec_f32(self):
        n = self.u64()
        return np.frombuffer(self.take(4 * n), dtype="<f4").astype(np.float32)

    def vec_u64(self):
        n = self.u64()
        return [int(v) for v in np.frombuffer(self.take(8 * n), dtype="<u8")]

    def vec_vec_f32(self):
        return [self.vec_f32() for _ in range(self.u64())]


class _Writer:
    def __init__(self):
        self.parts = []

    def u8(self, v):
        self.parts.append(struct.pack("<B", v))

    def u32(self, v):
        self.parts.append(struct.pack("<I", v))

    def u64(self, v):
        self.parts.append(struct.pack("<Q", int(v)))

    def f32(self, v):
        self.parts.append(struct.pack("<f", float(v)))

    def vec_f32(self, a):
        a = np.ascontiguousarray(a, dtype="<f4").reshape(-1)
        self.u64(a.size)
        self.parts.append(a.tobytes())

    def vec_u64(self, a):
        self.u64(len(a))
        self.parts.append(np.asarray(a, dtype="<u8").tobytes())

    def vec_vec_f32(self, vs):
        self.u64(len(vs))
        for v in vs:
            self.vec_f32(v)

    def bytes(self):
        return b"".join(self.parts)


# ------------------------------------------------------------------ genotypes, groupings, phenotypes
def read_dims(stem: str):
    """io/dims.rs:15-34: `<stem>.dims` = "N\\tM"; fallback: line counts of .fam / .bim."""
    p = stem + ".dims"
    if os.path.exists(p):
        n, m = open(p).read().split()
        return int(n), int(m)

    def lines(path):
        with open(path, "rb") as f:
            return sum(1 for _ in f)
    return lines(stem + ".fam"), lines(stem + ".bim")


def read_bed(stem: str):
    """BedVM::from_file (io/bed.rs:193-213): returns (payload u8 [M * ceil(N/4)], N, M)."""
    n, m = read_dims(stem)
    with open(stem + ".bed", "rb") as f:
        sig = f.read(3)
        if sig != BED_SIGNATURE:
            raise ValueError(f"{stem}.bed: not a variant-major PLINK .bed file (signature {sig.hex()})")
        payload = np.frombuffer(f.read(), dtype=np.uint8)
    bpc = (n + 3) // 4
    if payload.size != m * bpc:
        raise ValueError(f"{stem}.bed: expected {m * bpc} payload bytes for N={n}, M={m}, found {payload.size}")
    return payload.copy(), n, m


def write_bed(stem: str, payload, n: int, m: int):
    """BedVM::to_file (io/bed.rs:248-263): .bed + .dims."""
    with open(stem + ".bed", "wb") as f:
        f.write(BED_SIGNATURE)
        f.write(np.ascontiguousarray(payload, dtype=np.uint8).tobytes())
    with open(stem + ".dims", "w") as f:
        f.write(f"{n}\t{m}")


def pack_genotypes(g: np.ndarray) -> np.ndarray:
    """[N, M] values in {0,1,2} -> PLINK payload (io/bed.rs:378-395: 2 -> 00, 1 -> 10, 0 -> 11, LSB first)."""
    g = np.asarray(g)
    n, m = g.shape
    bpc = (n + 3) // 4
    code = np.where(g == 2, 0, np.where(g == 1, 2, 3)).astype(np.uint8)
    pad = np.zeros((bpc * 4 - n, m), dtype=np.uint8)
    code = np.concatenate([code, pad], axis=0).reshape(bpc, 4, m)
    byte = code[:, 0] | (code[:, 1] << 2) | (code[:, 2] << 4) | (code[:, 3] << 6)
    return np.ascontiguousarray(byte.T).reshape(-1)


def read_grouping(path: str) -> List[List[int]]:
    """ExternalGrouping::from_file (group/external.rs:15-58): TSV `marker_ix \\t group_ix`, group ids 0..G-1."""
    groups = {}
    with open(path) as f:
        for line in f:
            if not line.strip():
                continue
            mk, gr = line.split()[:2]
            groups.setdefault(int(gr), []).append(int(mk))
    ids = sorted(groups)
    if ids != list(range(len(ids))):
        raise ValueError(f"{path}: group ids must be contiguous and start at 0")
    return [groups[i] for i in ids]


def write_grouping(path: str, groups: Sequence[Sequence[int]]):
    """MarkerGrouping::to_file (group/grouping.rs:18-31)."""
    with open(path, "w") as f:
        for gi, cols in enumerate(groups):
            for c in cols:
                f.write(f"{c}\t{gi}\n")


def read_phen(path: str) -> np.ndarray:
    """Phenotypes::from_file (data/phenotypes.rs:28-31): bincode of struct { y: Vec<f32> }."""
    r = _Reader(open(path, "rb").read())
    return r.vec_f32()


def write_phen(path: str, y):
    w = _Writer()
    w.vec_f32(y)
    with open(path, "wb") as f:
        f.write(w.bytes())


def write_phen_stats(path: str, mean: float, variance: float, env_variance: float):
    """data/phen_stats.rs:6-24 (pretty JSON)."""
    with open(path, "w") as f:
        json.dump(dict(mean=float(mean), variance=float(variance), env_variance=float(env_variance)), f, indent=2)


# ------------------------------------------------------------------ model files
@dataclass
class BranchCfgFile:
    """BranchCfg (net/branch/branch_cfg.rs:8-16) with BranchParamsHost / BranchPrecisionsHost
    (net/params.rs:467-476,191-199).  Weight matrices are flat, column-major [in_l x out_l]."""
    num_params: int
    num_weights: int
    num_markers: int
    layer_widths: List[int]
    weights: List[np.ndarray]
    biases: List[np.ndarray]
    ow_reg_sum: float                    # OutputWeightSummaryStatsHost (params.rs:370-374)
    ow_num_params: int
    weight_precisions: List[np.ndarray]
    bias_precisions: List[np.ndarray]
    error_precision: List[float]
    activation: str = "tanh"

    def param_vec(self) -> np.ndarray:
        """params.rs:700-715: all weights layer by layer, then all biases."""
        return np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in self.weights] +
                              [np.asarray(b, dtype=np.float32).reshape(-1) for b in self.biases])

    def precision_vec(self) -> np.ndarray:
        """params.rs:272-289: weight precisions, bias precisions, error precision."""
        return np.concatenate([np.asarray(p, dtype=np.float32).reshape(-1) for p in self.weight_precisions] +
                              [np.asarray(p, dtype=np.float32).reshape(-1) for p in self.bias_precisions] +
                              [np.asarray(self.error_precision, dtype=np.float32).reshape(-1)])

    def load_param_vec(self, pv):
        pv = np.asarray(pv, dtype=np.float32)
        ix, prev = 0, self.num_markers
        for l, w in enumerate(self.layer_widths):
            self.weights[l] = pv[ix:ix + prev * w].copy()
            ix += prev * w
            prev = w
        for l, w in enumerate(self.layer_widths[:-1]):
            self.biases[l] = pv[ix:ix + w].copy()
            ix += w

    def load_precision_vec(self, qv):
        qv = np.asarray(qv, dtype=np.float32)
        ix = 0
        for l, p in enumerate(self.weight_precisions):
            self.weight_precisions[l] = qv[ix:ix + len(p)].copy()
            ix += len(p)
        for l, p in enumerate(self.bias_precisions):
            self.bias_precisions[l] = qv[ix:ix + len(p)].copy()
            ix += len(p)
        self.error_precision = [float(qv[ix])]

    def to_json(self):
        """serde_json form used by the `trace` file (net/net.rs:241-244)."""
        f = lambda vs: [[float(x) for x in np.asarray(v).reshape(-1)] for v in vs]   # noqa: E731
        return dict(num_params=self.num_params, num_weights=self.num_weights, num_markers=self.num_markers,
                    layer_widths=list(self.layer_widths),
                    params=dict(weights=f(self.weights), biases=f(self.biases), layer_widths=list(self.layer_widths),
                                num_markers=self.num_markers,
                                output_weight_summary_stats=dict(reg_sum=float(self.ow_reg_sum),
                                                                 num_params=int(self.ow_num_params))),
                    precisions=dict(weight_precisions=f(self.weight_precisions), bias_precisions=f(self.bias_precisions),
                                    error_precision=[float(x) for x in self.error_precision]),
                    activation_function=ACTIVATION_JSON[ACTIVATIONS.index(self.activation)])


@dataclass
class NetFile:
    """Net<B> as serialised (net/net.rs:74-85); the model type is NOT in the file, it lives in args.json."""
    hyper: List[float]                   # dense(shape, scale), summary(shape, scale), output(shape, scale)
    branch_cfgs: List[BranchCfgFile]
    output_bias: List[float] = field(default_factory=lambda: [2.0, 1.0, 0.0])   # error_precision, precision, bias
    num_samples: int = 0
    num_accepted: int = 0
    num_early_rejected: int = 0
    mse_train: List[float] = field(default_factory=list)
    mse_test: Optional[List[float]] = None
    lpd: List[float] = field(default_factory=list)
    lpd_rss: float = float("-inf")       # LogPosteriorDensity (log_posterior_density.rs:9-25)
    lpd_out_w: float = float("-inf")
    lpd_local: List[float] = field(default_factory=list)
    g_error_precision: float = 2.0       # GlobalParams (params.rs:13-18)
    g_output_layer_precision: float = 0.05
    g_ow_reg_sum: float = 0.0
    g_ow_num_params: int = 0

    def training_stats_json(self):
        """train_stats.rs:23-32,83-87."""
        return dict(num_samples=self.num_samples, num_accepted=self.num_accepted,
                    num_early_rejected=self.num_early_rejected, mse_train=[float(x) for x in self.mse_train],
                    mse_test=None if self.mse_test is None else [float(x) for x in self.mse_test],
                    lpd=[float(x) for x in self.lpd])


def _write_branch_cfg(w: _Writer, c: BranchCfgFile):
    w.u64(c.num_params); w.u64(c.num_weights); w.u64(c.num_markers); w.vec_u64(c.layer_widths)
    w.vec_vec_f32(c.weights); w.vec_vec_f32(c.biases); w.vec_u64(c.layer_widths); w.u64(c.num_markers)
    w.f32(c.ow_reg_sum); w.u64(c.ow_num_params)
    w.vec_vec_f32(c.weight_precisions); w.vec_vec_f32(c.bias_precisions); w.vec_f32(c.error_precision)
    w.u32(ACTIVATIONS.index(c.activation))


def _read_branch_cfg(r: _Reader) -> BranchCfgFile:
    num_params, num_weights, num_markers, widths = r.u64(), r.u64(), r.u64(), r.vec_u64()
    weights, biases = r.vec_vec_f32(), r.vec_vec_f32()
    widths2, markers2 = r.vec_u64(), r.u64()
    if widths2 != widths or markers2 != num_markers:
        raise ValueError("model file: BranchCfg and BranchParamsHost disagree on the architecture")
    reg_sum, ow_n = r.f32(), r.u64()
    wp, bp, ep = r.vec_vec_f32(), r.vec_vec_f32(), r.vec_f32()
    act = r.u32()
    if act >= len(ACTIVATIONS):
        raise ValueError(f"model file: unknown activation function tag {act}")
    return BranchCfgFile(num_params, num_weights, num_markers, widths, weights, biases, reg_sum, ow_n, wp, bp,
                         [float(x) for x in ep], ACTIVATIONS[act])


def write_net(path: str, net: NetFile):
    w = _Writer()
    for v in net.hyper:
        w.f32(v)
    w.u64(len(net.branch_cfgs))
    w.u64(len(net.branch_cfgs))
    for c in net.branch_cfgs:
        _write_branch_cfg(w, c)
    for v in net.output_bias:
        w.f32(v)
    w.u64(net.num_samples); w.u64(net.num_accepted); w.u64(net.num_early_rejected)
    w.vec_f32(net.mse_train)
    if net.mse_test is None:
        w.u8(0)
    else:
        w.u8(1); w.vec_f32(net.mse_test)
    w.vec_f32(net.lpd)
    w.f32(net.lpd_rss); w.f32(net.lpd_out_w); w.vec_f32(net.lpd_local)
    w.f32(net.g_error_precision); w.f32(net.g_output_layer_precision); w.f32(net.g_ow_reg_sum); w.u64(net.g_ow_num_params)
    with open(path, "wb") as f:
        f.write(w.bytes())


def read_net(path: str) -> NetFile:
    r = _Reader(open(path, "rb").read())
    hyper = [r.f32() for _ in range(6)]
    nb = r.u64()
    ncfg = r.u64()
    if ncfg != nb:
        raise ValueError("model file: num_branches != len(branch_cfgs)")
    cfgs = [_read_branch_cfg(r) for _ in range(ncfg)]
    ob = [r.f32() for _ in range(3)]
    ns, na, ne = r.u64(), r.u64(), r.u64()
    mse_train = [float(x) for x in r.vec_f32()]
    mse_test = [float(x) for x in r.vec_f32()] if r.u8() else None
    lpd = [float(x) for x in r.vec_f32()]
    lpd_rss, lpd_out = r.f32(), r.f32()
    lpd_local = [float(x) for x in r.vec_f32()]
    gep, gop, grs, gnp = r.f32(), r.f32(), r.f32(), r.u64()
    if r.o != len(r.d):
        raise ValueError(f"model file: {len(r.d) - r.o} trailing bytes")
    return NetFile(hyper, cfgs, ob, ns, na, ne, mse_train, mse_test, lpd, lpd_rss, lpd_out, lpd_local, gep, gop, grs, gnp)


def hyperparams_json(net: NetFile):
    """Net::write_hyperparams (net/net.rs:149-156, params.rs:66-103)."""
    ph = lambda a, b: dict(shape=float(a), scale=float(b))   # noqa: E731
    return dict(branch_hyperparams=[dict(num_params=c.num_params, num_markers=c.num_markers,
                                         layer_widths=list(c.layer_widths)) for c in net.branch_cfgs],
                precision_hyperparams=dict(dense=ph(*net.hyper[0:2]), summary=ph(*net.hyper[2:4]),
                                           output=ph(*net.hyper[4:6])))
