"""World-size-2 `gloo` tests (CPU) of the row-sharded path's host logic: shard boundaries, global
column statistics from all-reduced counts, and the all-reduce of the per-step [gW | gb | rss] sums
(here produced by the oracle, on the GPU box by K1): the reduced sums must equal the unsharded ones."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bed as obed
from oracle.branch import Branch, make_cfg


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, m, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rs_bann_b200 as rb
    g = obed.random_genotypes(n, m, seed=seed)
    payload = obed.pack_columns(g)
    r0, r1 = rb.row_shard(n, rank, world)
    local = rb.shard_payload(payload, n, m, r0, r1)
    assert np.array_equal(obed.decode_columns(local, r1 - r0, range(m)), g[r0:r1].astype(np.float32))
    dec = g[r0:r1]
    counts = np.stack([(dec == v).sum(axis=0) for v in range(3)], axis=1).astype(np.int64)

    def allreduce(c):
        t = torch.from_numpy(c.copy())
        dist.all_reduce(t)
        return t.numpy()

    mu, sd = rb.global_col_stats(counts, n, allreduce)
    # per-step sums of one branch on this shard (oracle stands in for K1), then the collective
    rng = np.random.default_rng(seed)
    cfg = make_cfg("ridge_ard", m, [5], 5, rng=rng)
    y = rng.normal(size=n).astype(np.float32)
    x = obed.submatrix_standardized(local, r1 - r0, range(m), mu, sd, np.float64)
    rss, gW, gb = Branch(cfg, np.float64).backpropagate(x, y[r0:r1].astype(np.float64))
    part = np.concatenate([Branch.join_vec(gW, gb), [rss]])
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    if rank == 0:
        np.savez(os.path.join(out_dir, "res.npz"), mu=mu, sd=sd, red=t.numpy(), shard=np.array([r0, r1]))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1030])
def test_two_rank_gloo_allreduce_matches_unsharded(tmp_path, n):
    m, seed, world = 17, 7, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, m, seed, str(tmp_path)), nprocs=world, join=True)
    res = np.load(tmp_path / "res.npz")
    g = obed.random_genotypes(n, m, seed=seed)
    payload = obed.pack_columns(g)
    mu, sd = obed.col_stats(payload, n, m)
    assert np.array_equal(res["mu"], mu) and np.allclose(res["sd"], sd, rtol=2e-5)
    rng = np.random.default_rng(seed)
    cfg = make_cfg("ridge_ard", m, [5], 5, rng=rng)
    y = rng.normal(size=n).astype(np.float32)
    x = obed.submatrix_standardized(payload, n, range(m), res["mu"], res["sd"], np.float64)
    rss, gW, gb = Branch(cfg, np.float64).backpropagate(x, y.astype(np.float64))
    full = np.concatenate([Branch.join_vec(gW, gb), [rss]])
    assert np.allclose(res["red"], full, rtol=1e-10, atol=1e-10)
    assert res["shard"][0] == 0 and res["shard"][1] % 128 == 0


def test_row_shards_cover_all_rows():
    import rs_bann_b200 as rb
    for n in (1, 127, 128, 129, 1000, 100000):
        for world in (1, 2, 4, 8):
            tiles = (n + 127) // 128
            tpr = (tiles + world - 1) // world
            filled = (tiles + tpr - 1) // tpr
            prev = 0
            for r in range(world):
                if r >= filled:            # an empty shard is refused with a clear message (not a zero-block launch later)
                    with pytest.raises(ValueError):
                        rb.row_shard(n, r, world)
                    continue
                r0, r1 = rb.row_shard(n, r, world)
                assert r0 == prev and r0 % 128 == 0 and r1 > r0
                prev = r1
            assert prev == n
