"""Host-side logic of the row-sharded (multi-GPU) path: one process per GPU, torch.distributed
for the plumbing.  Individuals (rows) are sharded on 128-row tile boundaries; parameters,
momenta and precisions are replicated; the only data-path collectives are
  * an all-reduce(sum) of the per-column genotype counts at load (global column statistics),
  * an all-reduce(sum) of the per-step [gW | gb | rss] sums between K1 and K2 -- inside the library over
    NVLink peer memory once `connect_net` has wired the ranks' exchange regions (csrc/comm.cuh: XgComm).
Every replica then applies the identical update (same Philox keys), so nothing is broadcast.
The sequential-exact schedule (branch visits of Net::train) sums over ranks INSIDE the library's
reduction kernels through peer-mapped inboxes (csrc/comm.cuh); `connect_ranks` wires them up."""
from typing import Callable, Optional, Tuple

import numpy as np

TILE_ROWS = 128


def row_shard(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[r0, r1) of this rank: whole 128-row tiles, remainder to the last non-empty rank."""
    tiles = (n_total + TILE_ROWS - 1) // TILE_ROWS
    tpr = (tiles + world - 1) // world
    r0 = min(n_total, rank * tpr * TILE_ROWS)
    r1 = min(n_total, (rank + 1) * tpr * TILE_ROWS)
    if r1 <= r0:
        # an empty shard would mean zero-block kernel launches and zero-sized stores in the library: refuse up front
        raise ValueError(f"rank {rank} of {world} would hold no individuals: {n_total} individuals are {tiles} row tile(s) of "
                         f"{TILE_ROWS}, which fill only {(tiles + tpr - 1) // tpr} rank(s); use fewer ranks")
    return r0, r1


def shard_payload(payload: np.ndarray, n_total: int, m: int, r0: int, r1: int) -> np.ndarray:
    """Rows [r0, r1) of a PLINK variant-major payload (r0 must be a multiple of 4)."""
    assert r0 % 4 == 0
    bpc = (n_total + 3) // 4
    cols = np.asarray(payload, dtype=np.uint8).reshape(m, bpc)
    out = cols[:, r0 // 4:(r1 + 3) // 4].copy()
    if (r1 - r0) % 4:
        out[:, -1] &= np.uint8((1 << (2 * ((r1 - r0) % 4))) - 1)   # pad fields of the local last byte -> code 00
    return np.ascontiguousarray(out).reshape(-1)


def global_col_stats(local_counts: np.ndarray, n_total: int, allreduce_sum: Optional[Callable] = None):
    """Column mean / population std over ALL rows from per-shard value counts."""
    from .api import stats_from_counts
    counts = np.ascontiguousarray(local_counts, dtype=np.int64)
    if allreduce_sum is not None:
        counts = allreduce_sum(counts)
    return stats_from_counts(counts, n_total)


def _gather_handles(world: int, mine: bytes, all_gather_bytes: Optional[Callable]):
    if all_gather_bytes is None:
        import torch.distributed as dist
        out = [None] * world
        dist.all_gather_object(out, mine)
        return [bytes(h) for h in out]
    return [bytes(h) for h in all_gather_bytes(mine)]


def connect_net(net, all_gather_bytes: Optional[Callable] = None):
    """Bulk peer-memory exchange of a net (grouped schedule: all-reduce of [gW | gb | rss] inside the library, parameter
    all-gather of Net.gradient): all-gather every rank's region handle and map the peers.  No-op on one rank."""
    world = net.ctx.world
    if world == 1:
        return
    net.comm_connect(_gather_handles(world, net.comm_handle(), all_gather_bytes))


def connect_ranks(ctx, all_gather_bytes: Optional[Callable] = None):
    """Peer-memory exchange set-up: all-gather every rank's inbox handle and map the peers.
    `all_gather_bytes(b) -> [bytes of rank 0, ..., bytes of rank world-1]`; the default uses
    torch.distributed.all_gather_object on the default process group."""
    if ctx.world == 1:
        return
    mine = ctx.comm_handle()
    if all_gather_bytes is None:
        import torch.distributed as dist
        out = [None] * ctx.world
        dist.all_gather_object(out, mine)
        handles = out
    else:
        handles = all_gather_bytes(mine)
    ctx.comm_connect([bytes(h) for h in handles])
