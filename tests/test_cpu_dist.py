"""world_size-2 gloo test of the N>1 host logic (sharding + the single all-reduce of scalars)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from enoki_b200.dist import shard_range, allreduce_scalars
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 1_000_003
    x = np.sin(np.arange(n, dtype=np.float64) * 1e-3)
    lo, hi = shard_range(n, rank, world)
    # per-rank partial "loss" and "gradient of a size-1 leaf": sum and dot over the local slice
    part = np.array([x[lo:hi].sum(), (x[lo:hi] * x[lo:hi]).sum(), float(hi - lo)])
    tot = allreduce_scalars(part)
    q.put((rank, lo, hi, tot.tolist()))
    dist.destroy_process_group()


def test_sharded_scalars_allreduce_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29577
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs: p.join(timeout=60)
    res.sort()
    n = 1_000_003
    x = np.sin(np.arange(n, dtype=np.float64) * 1e-3)
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n      # contiguous cover
    for r in res:
        assert abs(r[3][0] - x.sum()) < 1e-6 and abs(r[3][1] - (x * x).sum()) < 1e-6 and r[3][2] == n


def test_shard_range_covers():
    from enoki_b200.dist import shard_range
    for n in (1, 7, 64, 1 << 26):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
