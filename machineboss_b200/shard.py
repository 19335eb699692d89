"""Sharding of a SeqPairList over the GPUs of one box, and the one exchange step of the path.

Every sequence pair is an independent unit: Forward, Viterbi and alignment
need no communication.  The E-step of Baum-Welch (MachineCounts over the list, src/counts.cpp:37-43)
sums per-pair counts, so after each rank has processed its shard the count vector and the total
log-likelihood are combined with ONE all-reduce of nTransitions + 1 doubles -- NCCL over NVLink on
the GPUs, gloo in the CPU tests.

Two launch models share the deal of pairs to devices (mb_shard_pairs in the library):
  * one process, several GPUs: the C ABI's mb_group_* entry points (a host thread per device, ncclAllReduce of
    the counts) -- what the host mirror's MachineCounts / MachineFitter and the boss_b200 CLI use;
  * one process per GPU under torchrun (bench.py --gpus N): this module, the exchange step through torch.distributed.
"""
from __future__ import annotations

import numpy as np


def lpt_assign(x_off, y_off, world: int):
    """The library's deal of pairs to `world` shards (mb_shard_pairs: longest-processing-time first by cell count
    (Li+1)(Lo+1), each pair to the least loaded shard, deterministic so every rank computes the same plan);
    returns the list of index arrays, one per rank.  Ranks of a torchrun job use it to pick their pairs; inside one
    process the same deal is what mb_group_batch_create applies to the devices of a group."""
    from . import capi
    shard_of = capi.shard_pairs(x_off, y_off, world)
    return [np.flatnonzero(shard_of == r).astype(np.int64) for r in range(world)]


def allreduce_counts(counts: np.ndarray, loglike: float, device=None):
    """Sum the per-rank count vectors and log-likelihoods over the process group (no-op if not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts, loglike
    t = torch.from_numpy(np.concatenate([np.asarray(counts, dtype=np.float64), [float(loglike)]]))
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    t = t.cpu().numpy()
    return t[:-1].copy(), float(t[-1])


def gather_by_pair(values: np.ndarray, mine: np.ndarray, n_total: int, device=None) -> np.ndarray:
    """Place this rank's per-pair results at their pair indices and combine over ranks."""
    import torch
    import torch.distributed as dist
    out = np.zeros(n_total, dtype=np.float64)
    out[mine] = values
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return out
    t = torch.from_numpy(out)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
