"""Oracle: PLINK .bed payload handling (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/io/bed.rs, src/io/bed_lookup_tables.rs, src/io/dims.rs,
src/group/{grouping,uniform,external}.rs of the reference.
"""
from __future__ import annotations

import os
import numpy as np

BED_SIGNATURE = bytes([0x6C, 0x1B, 0x01])  # io/bed.rs:13 (variant-major)

# io/bed_lookup_tables.rs:4 -- 2-bit code (LSB first) -> genotype value.
# 00 -> 2, 01 -> 0 (missing, no NA handling: bed.rs:15-16,121), 10 -> 1, 11 -> 0
CODE_TO_VALUE = np.array([2.0, 0.0, 1.0, 0.0], dtype=np.float32)
# io/bed.rs:16 BED_VALUE_MAPPING: value -> code used when packing
VALUE_TO_CODE = np.array([0x03, 0x02, 0x00], dtype=np.uint8)


def _build_lut() -> np.ndarray:
    b = np.arange(256, dtype=np.uint16)
    lut = np.empty((256, 4), dtype=np.float32)
    for k in range(4):
        lut[:, k] = CODE_TO_VALUE[(b >> (2 * k)) & 3]
    return lut


BED_LOOKUP_GENOTYPE = _build_lut()  # [256,4] == the 1024-entry table of the reference


def bytes_per_col(n: int) -> int:
    """io/bed.rs:215-219."""
    return (n + 3) // 4


def pack_columns(g: np.ndarray) -> np.ndarray:
    """Pack an integer genotype matrix [N, M] (values 0/1/2) into the variant-major
    payload (no signature).  io/bed.rs:378-395 (`chunkf32_to_byte`, `vecf32_to_bed`):
    first individual of each chunk of four goes to the two LOW bits; a short last chunk
    leaves the high bits 0 (== code 00 == value 2, only ever truncated away)."""
    g = np.asarray(g)
    n, m = g.shape
    bpc = bytes_per_col(n)
    codes = VALUE_TO_CODE[g.astype(np.int64)]  # [N, M]
    pad = bpc * 4 - n
    if pad:
        codes = np.concatenate([codes, np.zeros((pad, m), dtype=np.uint8)], axis=0)
    codes = codes.reshape(bpc, 4, m)
    by = (codes[:, 0] | (codes[:, 1] << 2) | (codes[:, 2] << 4) | (codes[:, 3] << 6)).astype(np.uint8)
    return np.ascontiguousarray(by.T).reshape(-1)  # column after column


def decode_columns(payload: np.ndarray, n: int, cols) -> np.ndarray:
    """LUT-decode columns to f32 [N, len(cols)] (raw 0/1/2).  io/bed.rs:287-302."""
    payload = np.frombuffer(bytes(payload), dtype=np.uint8) if not isinstance(payload, np.ndarray) else payload
    bpc = bytes_per_col(n)
    cols = np.asarray(cols, dtype=np.int64)
    out = np.empty((n, len(cols)), dtype=np.float32)
    for k, c in enumerate(cols):
        col = payload[c * bpc:(c + 1) * bpc]
        out[:, k] = BED_LOOKUP_GENOTYPE[col].reshape(-1)[:n]
    return out


def col_stats(payload: np.ndarray, n: int, m: int):
    """Per-column mean and POPULATION std in sequential f32, io/bed.rs:231-238:
        mean = sum_f32(vals) / n ; std = sqrt(sum_f32((v-mean)^2) / n)
    `iter().sum::<f32>()` is a left-to-right f32 accumulation, restated with
    np.add.accumulate (sequential by definition)."""
    means = np.empty(m, dtype=np.float32)
    stds = np.empty(m, dtype=np.float32)
    nf = np.float32(n)
    for j in range(m):
        v = decode_columns(payload, n, [j])[:, 0]
        s = np.add.accumulate(v, dtype=np.float32)[-1] if n else np.float32(0)
        mean = np.float32(s / nf)
        d = (v - mean).astype(np.float32)
        sq = (d * d).astype(np.float32)
        ss = np.add.accumulate(sq, dtype=np.float32)[-1] if n else np.float32(0)
        means[j] = mean
        stds[j] = np.sqrt(np.float32(ss / nf), dtype=np.float32)
    return means, stds


def submatrix_standardized(payload, n, cols, means, stds, dtype=np.float32) -> np.ndarray:
    """(raw - mean_j) / std_j, subtract then divide.  io/bed.rs:325-355.
    In the f64 "truth" variant the f32 means/stds are used as given (they are data)."""
    raw = decode_columns(payload, n, cols).astype(dtype)
    cols = np.asarray(cols, dtype=np.int64)
    mu = np.asarray(means)[cols].astype(dtype)
    sd = np.asarray(stds)[cols].astype(dtype)
    with np.errstate(divide="ignore", invalid="ignore"):  # Q3: std==0 divides by zero
        return ((raw - mu[None, :]) / sd[None, :]).astype(dtype)


def read_bed(stem: str):
    """io/bed.rs:193-245 + io/dims.rs:15-34 (.dims, else .fam/.bim line counts)."""
    dims = stem + ".dims"
    if os.path.exists(dims):
        n, m = (int(x) for x in open(dims).readline().split()[:2])
    else:
        n = sum(1 for _ in open(stem + ".fam"))
        m = sum(1 for _ in open(stem + ".bim"))
    raw = open(stem + ".bed", "rb").read()
    if raw[:2] != BED_SIGNATURE[:2]:
        raise ValueError("bad .bed signature")
    if raw[2] != 1:
        raise ValueError("SampleMajor .bed not supported (bed.rs:200-202)")
    payload = np.frombuffer(raw[3:], dtype=np.uint8).copy()
    assert payload.size == m * bytes_per_col(n)
    return payload, n, m


def write_bed(stem: str, payload: np.ndarray, n: int, m: int) -> None:
    """io/bed.rs:248-264."""
    with open(stem + ".bed", "wb") as f:
        f.write(BED_SIGNATURE)
        f.write(bytes(np.asarray(payload, dtype=np.uint8)))
    with open(stem + ".dims", "w") as f:
        f.write(f"{n}\t{m}")


# ---------------------------------------------------------------- groupings
def uniform_grouping(num_groups: int, per_group: int):
    """group/uniform.rs:11-24."""
    return [list(range(g * per_group, (g + 1) * per_group)) for g in range(num_groups)]


def read_grouping(path: str):
    """group/external.rs:15-58: TSV `marker_ix<TAB>group_ix`, 0-based contiguous ids,
    markers kept in file order, groups may overlap / be non-contiguous (Q15)."""
    groups = {}
    for line in open(path):
        f = line.split()
        if not f:
            continue
        groups.setdefault(int(f[1]), []).append(int(f[0]))
    assert not any(k >= len(groups) for k in groups), "group ids must be 0-based contiguous"
    return [groups[k] for k in range(len(groups))]


def write_grouping(path: str, groups) -> None:
    """group/grouping.rs:18-31."""
    with open(path, "w") as f:
        for gi, g in enumerate(groups):
            for mk in g:
                f.write(f"{mk}\t{gi}\n")


def groups_to_csr(groups):
    offs = np.zeros(len(groups) + 1, dtype=np.uint64)
    for i, g in enumerate(groups):
        offs[i + 1] = offs[i] + len(g)
    ids = np.array([c for g in groups for c in g], dtype=np.uint64)
    return offs, ids


def random_genotypes(n: int, m: int, seed: int = 42, maf_lo=0.01, maf_hi=0.5) -> np.ndarray:
    """Synthetic genotypes in the spirit of BedVM::random (io/bed.rs:136-188):
    maf_j ~ U(0.01, 0.5), g ~ Binomial(2, maf_j), monomorphic columns redrawn.
    (The reference's ChaCha20 stream is third-party and not reproduced.)"""
    rng = np.random.default_rng(seed)
    g = np.empty((n, m), dtype=np.uint8)
    for j in range(m):
        while True:
            maf = rng.uniform(maf_lo, maf_hi)
            col = rng.binomial(2, maf, size=n).astype(np.uint8)
            if col.min() != col.max():
                g[:, j] = col
                break
    return g
