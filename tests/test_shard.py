"""CPU tests of the multi-GPU host logic: deterministic sharding and the world-size-2 count
all-reduce over gloo (the GPU path uses the same code over NCCL)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from machineboss_b200 import shard


def test_lpt_assign_balances_and_partitions():
    rng = np.random.default_rng(1)
    li, lo = rng.integers(50, 500, 1000), rng.integers(50, 500, 1000)
    costs = (li + 1.0) * (lo + 1.0)
    x_off, y_off = np.concatenate([[0], np.cumsum(li)]), np.concatenate([[0], np.cumsum(lo)])
    for world in (1, 2, 4, 8):
        bins = shard.lpt_assign(x_off, y_off, world)
        allk = np.sort(np.concatenate(bins))
        assert np.array_equal(allk, np.arange(1000))
        loads = np.array([costs[b].sum() for b in bins])
        assert loads.max() / loads.mean() < 1.01
    assert all(np.array_equal(a, b) for a, b in zip(shard.lpt_assign(x_off, y_off, 4), shard.lpt_assign(x_off, y_off, 4)))
    # the longest pair opens shard 0, the next ones shards 1, 2, 3: longest-processing-time first
    order = np.argsort(-costs, kind="stable")
    first = [int(np.flatnonzero([k in b for b in shard.lpt_assign(x_off, y_off, 4)])[0]) for k in order[:4]]
    assert first == [0, 1, 2, 3]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = np.arange(0, 10, dtype=np.int64)      # pair k costs (k + 1) * 1 cells
    x_off, y_off = np.concatenate([[0], np.cumsum(lens)]), np.zeros(11, dtype=np.int64)
    mine = shard.lpt_assign(x_off, y_off, world)[rank]
    # stand-in for per-shard E-step results: counts proportional to the pair index, ll = -index
    counts = np.zeros(5)
    for k in mine:
        counts += np.arange(5) * (k + 1)
    c, ll = shard.allreduce_counts(counts, -float(mine.sum()))
    per_pair = shard.gather_by_pair(mine.astype(np.float64) * 2, mine, 10)
    if rank == 0:
        out.put((c.tolist(), ll, per_pair.tolist()))
    dist.destroy_process_group()


def test_count_allreduce_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    c, ll, per_pair = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert c == (np.arange(5) * 55.0).tolist()
    assert ll == -45.0
    assert per_pair == (np.arange(10) * 2.0).tolist()
